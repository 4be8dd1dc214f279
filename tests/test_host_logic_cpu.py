"""Host-side logic that needs no GPU: prompt-length buckets of the prefill graphs, HF shortest-edge sizes, bench decode-group sizing."""
import torch

from videoglamm_b200 import preprocess as P
from videoglamm_b200.lmm_engine import LmmEngine


def test_prefill_bucket_and_padding():
    eng = object.__new__(LmmEngine)
    assert [eng._bucket(n) for n in (1, 63, 64, 65, 1696, 1728, 3360)] == [64, 64, 64, 128, 1728, 1728, 3392]
    x = torch.arange(12.0).view(3, 4)
    y = LmmEngine._pad_rows(x, 5)
    assert y.shape == (5, 4) and torch.equal(y[:3], x) and float(y[3:].abs().sum()) == 0.0
    assert LmmEngine._pad_rows(x, 3) is x


def test_clip_shortest_edge_sizes_match_hf():
    from transformers.image_transforms import get_resize_output_image_size
    import numpy as np
    for h, w in [(336, 336), (480, 854), (720, 1280), (1280, 720), (100, 60), (37, 53), (1080, 1920)]:
        want = get_resize_output_image_size(np.zeros((h, w, 3), np.uint8), size=336, default_to_square=False)
        assert tuple(want) == P.clip_resize_shape(h, w, 336)
