"""Result writers: file layout, pixel values and the small pieces of arithmetic the eval scripts apply before writing."""
import json
import os

import numpy as np
import torch
from PIL import Image

from videoglamm_b200 import writers as W


def _segments(T=3, n_obj=2, hw=(6, 8), seed=0):
    rng = np.random.default_rng(seed)
    return {t: {o: rng.random(hw) > 0.5 for o in range(n_obj)} for t in range(T)}


def test_binary_mask_pngs(tmp_path):
    seg = _segments()
    paths = W.save_binary_mask_pngs(seg, str(tmp_path / "vid" / "0"))
    assert [os.path.basename(p) for p in paths] == ["00000.png", "00001.png", "00002.png"]       # eval_mevis.py:195
    for t, p in enumerate(paths):
        im = np.array(Image.open(p))
        assert im.dtype == np.uint8 and set(np.unique(im)) <= {0, 255}
        assert np.array_equal(im == 255, seg[t][0])                                                 # first object only (:203)
    paths = W.save_binary_mask_pngs(seg, str(tmp_path / "ytvos"), frame_names=["00010", "00015", "00020"])
    assert [os.path.basename(p) for p in paths] == ["00010.png", "00015.png", "00020.png"]       # :192-193


def test_davis_merge_and_palette(tmp_path):
    rng = np.random.default_rng(1)
    obj = (rng.random((3, 2, 5, 7)) > 0.6).astype(np.float32)
    got = W.merge_object_masks(obj)
    # the script's own torch expressions (eval_referdavis_infer.py:247-255)
    a = torch.from_numpy(obj.copy())
    a[a < 0.5] = 0.0
    want = torch.argmax(torch.cat([0.1 * torch.ones(1, 2, 5, 7), a], 0), 0).numpy().astype(np.uint8)
    assert np.array_equal(got, want) and got.max() <= 3
    palette = [v for i in range(256) for v in (i, (i * 7) % 256, (i * 13) % 256)]
    paths = W.save_palette_pngs(got, str(tmp_path / "anno_0" / "bear"), palette)
    assert [os.path.basename(p) for p in paths] == ["00000.png", "00001.png"]
    im = Image.open(paths[0])
    assert im.mode == "P" and np.array_equal(np.array(im), got[0]) and im.getpalette()[:6] == palette[:6]


def test_remove_small_blobs_uses_4_connectivity():
    m = np.zeros((8, 8), np.uint8)
    m[0, 0] = m[1, 1] = m[2, 2] = 1          # diagonal chain: three separate 1-pixel objects under 4-connectivity
    m[5:7, 3:6] = 1                          # 6-pixel block
    out = W.remove_small_blobs(m, min_size=3)
    assert out.dtype == m.dtype and out[:3, :3].sum() == 0 and out[5:7, 3:6].sum() == 6
    assert np.array_equal(W.remove_small_blobs(m, min_size=0), m)
    assert W.remove_small_blobs(m, min_size=7).sum() == 0     # "fewer than min_size" pixels are removed: the 6-pixel block goes at 7, stays at 6
    assert W.remove_small_blobs(m, min_size=6)[5:7, 3:6].sum() == 6


def test_gcg_result(tmp_path):
    seg = _segments(T=2, n_obj=2, hw=(12, 12), seed=3)
    seg[0][1][:] = False
    seg[0][1][0, 0] = True                     # a 1-pixel blob: removed by min_size=20
    out = W.save_gcg_result(str(tmp_path / "v"), seg, {"pred_text": "a [SEG] b", "gt_text": "g"})
    res = json.load(open(tmp_path / "v" / "res.json"))
    assert list(res) == ["gt_text", "gt_text_cleaned", "gt_phrases", "pred_text", "pred_text_cleaned", "pred_phrases"] and res["pred_text"] == "a [SEG] b"
    assert sorted(os.listdir(tmp_path / "v")) == ["pred_masks_0", "pred_masks_1", "res.json"]
    assert out[1][0].sum() == 0
    im = np.array(Image.open(tmp_path / "v" / "pred_masks_0" / "mask_1.png"))
    assert np.array_equal(im == 255, out[0][1])
