"""vg_connected_components (the B200 equivalent of the reference's sam2._C extension) vs the oracle, bit for bit, and vs the reference's own
kernel when oracle/_ref holds the extension built from the unmodified source (oracle/build_ref.py)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import cc as O  # noqa: E402  (checker only)
from videoglamm_b200 import ops  # noqa: E402
from videoglamm_b200.postprocess import fill_holes_in_mask_scores, get_connected_components  # noqa: E402


def _check(mask_np):
    lab, cnt = ops.connected_components(torch.from_numpy(mask_np).cuda())
    wl, wc = O.connected_components(mask_np)
    assert lab.dtype == torch.int32 and cnt.dtype == torch.int32 and tuple(lab.shape) == mask_np.shape
    assert np.array_equal(lab.cpu().numpy(), wl)
    assert np.array_equal(cnt.cpu().numpy(), wc)


@pytest.mark.parametrize("shape,density", [((1, 1, 2, 2), 0.5), ((3, 1, 16, 16), 0.3), ((2, 1, 64, 94), 0.5), ((2, 1, 64, 94), 0.62),
                                           ((1, 1, 480, 854), 0.55), ((2, 1, 1024, 1024), 0.58), ((1, 1, 720, 1280), 0.9)])
def test_random_masks_bit_exact(shape, density):
    rng = np.random.default_rng(hash((shape, density)) % (2 ** 31))
    _check((rng.random(shape) < density).astype(np.uint8))


def test_structured_masks():
    h, w = 32, 48
    yy, xx = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    for m in (np.zeros((h, w)), np.ones((h, w)), (yy + xx) % 2 == 0, (yy % 2 == 0) & (xx % 2 == 0), (yy == xx) | (yy + xx == h - 1), yy % 4 == 1,
              ((yy // 3 + xx // 5) % 2 == 0)):
        _check(m.astype(np.uint8)[None, None])
    spiral = np.zeros((64, 64), np.uint8)     # one long snake: worst case for pointer chasing
    for r in range(0, 64, 4):
        spiral[r, :] = 1
        spiral[r + 1: r + 4, 63 if (r // 4) % 2 == 0 else 0] = 1
    _check(spiral[None, None])


def test_errors_and_helpers():
    with pytest.raises(RuntimeError):
        ops.connected_components(torch.zeros(1, 1, 5, 4, dtype=torch.uint8, device="cuda"))      # odd height (reference: AT_ASSERTM)
    with pytest.raises(RuntimeError):
        ops.connected_components(torch.zeros(1, 1, 4, 4, dtype=torch.float32, device="cuda"))
    with pytest.raises(RuntimeError):
        ops.connected_components(torch.zeros(1, 1, 4, 4, dtype=torch.uint8))
    torch.manual_seed(0)
    scores = torch.randn(2, 1, 64, 64, device="cuda")
    lab, area = get_connected_components(scores <= 0)
    wl, wc = O.connected_components((scores <= 0).cpu().numpy().astype(np.uint8))
    assert np.array_equal(lab.cpu().numpy(), wl) and np.array_equal(area.cpu().numpy(), wc)
    filled = fill_holes_in_mask_scores(scores, 8)                                             # misc.py:216-227
    want = torch.where(torch.from_numpy((wl > 0) & (wc <= 8)).cuda(), 0.1, scores)
    assert torch.equal(filled, want)


def test_against_the_reference_extension():
    from oracle import build_ref
    ref = build_ref.load_ref()
    if ref is None:
        pytest.skip("oracle/_ref not built (python oracle/build_ref.py needs /root/reference)")
    rng = np.random.default_rng(7)
    for shape, density in (((2, 1, 64, 94), 0.5), ((1, 1, 480, 854), 0.57), ((3, 1, 256, 256), 0.62), ((1, 1, 1024, 1024), 0.3)):
        m = torch.from_numpy((rng.random(shape) < density).astype(np.uint8)).cuda()
        want_l, want_c = ref(m)
        got_l, got_c = ops.connected_components(m)
        assert torch.equal(got_l, want_l) and torch.equal(got_c, want_c)
