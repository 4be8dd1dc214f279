"""GPU pre-processing of the SAM2 frames: uint8 frames in, [T,3,1024,1024] fp32 out, vs the reference's CPU path
(PIL bilinear ResizeLongestSide -> normalise -> second bilinear squash; utils/sam_transforms.py:26-65 mirrored by preprocess.sam_preprocess,
which tests/test_preprocess_cpu.py checks against the reference's own function)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from videoglamm_b200 import ops, preprocess as P  # noqa: E402


@pytest.mark.parametrize("H,W", [(336, 336), (480, 854), (720, 1280), (53, 91)])
def test_gpu_resize_is_bit_exact_with_pil_and_squash_matches(H, W):
    from PIL import Image
    rng = np.random.default_rng(H * 7 + W)
    frames = rng.integers(0, 256, (3, H, W, 3), dtype=np.uint8)
    pre = P.SAM_v2_PreprocessGPU()
    (newh, neww), tabs = pre.tables(H, W)
    got = ops.resize_u8_pil(torch.from_numpy(frames).cuda(), newh, neww, tabs).cpu().numpy()
    for t in range(3):
        ref = np.array(Image.fromarray(frames[t]).resize((neww, newh), Image.BILINEAR))
        assert np.array_equal(got[t], ref)                                   # integer work: bit-exact
    out, shape = pre.preprocess_video(frames)
    assert out.shape == (3, 3, 1024, 1024) and out.dtype == torch.float32 and shape == (newh, neww)
    for t in range(3):
        want, rs = P.sam_preprocess(frames[t])
        assert tuple(rs) == shape
        err = (out[t].cpu() - want).abs().max().item()
        assert err < 2e-5, err                                               # fp32: same formula, different association (values are O(2))


def test_inference_accepts_raw_uint8_sam_frames():
    """`images_for_sam` may be the raw uint8 frames [T,H,W,3] (host or device): the engine then pre-processes them on the GPU; masks equal the
    ones obtained from the CPU-preprocessed fp32 tensor."""
    from oracle import lmm as M, sam2 as O, weights as Wt
    from videoglamm_b200.config import Sam2Config, make_lmm_config
    from videoglamm_b200.model import VideoGLaMMForCausalLM
    seg = 300
    ocfg = M.make_cfg(**M.TINY_LMM, seg_token_idx=seg)
    tiny1024 = {k: v for k, v in O.TINY.items() if k != "image_size"}           # the SAM2 input of sam_preprocess is always 1024 px
    sd = Wt.make_videoglamm_state_dict(ocfg, O.Sam2Cfg(**tiny1024), seed=7)
    model = VideoGLaMMForCausalLM(sd, make_lmm_config(**M.TINY_LMM, seg_token_idx=seg), Sam2Config(**tiny1024))
    rng = np.random.default_rng(3)
    frames = rng.integers(0, 256, (2, 60, 100, 3), dtype=np.uint8)
    T = 4
    torch.manual_seed(1)
    images, ctx = torch.randn(T, 3, 56, 56), torch.randn(T, 3, 56, 56)
    ids = torch.cat([torch.randint(3, 290, (5,)), torch.full((T,), -200), torch.randint(3, 290, (7,))])[None].cuda()
    sam_cpu = torch.stack([P.sam_preprocess(f)[0] for f in frames])
    a = model.inference([images.cuda()], [ctx.cuda()], [sam_cpu.cuda()], ids, None, [(60, 100)], max_new_tokens=3, use_sam2_video_branch=True, forced_tokens={1: seg})
    b = model.inference([images.cuda()], [ctx.cuda()], [torch.from_numpy(frames)], ids, None, [(60, 100)], max_new_tokens=3, use_sam2_video_branch=True,
                        forced_tokens={1: seg})
    assert torch.equal(a[0], b[0])
    for t in a[1][0]:
        assert (a[1][0][t][0] == b[1][0][t][0]).mean() > 0.999
