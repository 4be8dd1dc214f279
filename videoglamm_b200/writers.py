"""Result writers: the on-disk formats the reference's eval scripts produce from `inference()`'s output. Host code,
downstream of the hot path; each function restates the lines of the script it replaces so that the official evaluation tools read the files
unchanged. `video_segments` is `inference()`'s per-sample dict {frame_idx: {obj_id: bool ndarray [H, W]}}."""
from __future__ import annotations

import json
import os

import numpy as np


def save_binary_mask_pngs(video_segments: dict, out_dir: str, frame_names: list | None = None) -> list[str]:
    """MeViS / Refer-YouTube-VOS / Refer-DAVIS submission layout (eval_mevis.py:182-203): one 8-bit PNG per frame holding the FIRST object's mask
    scaled to {0, 255} (`break` after the first object, :203), named `{t:05d}.png` (MeViS) or `{frame_names[t]}.png` (YTVOS / DAVIS, :192-193),
    inside `out_dir` = .../{video_name}/{exp_id}. Returns the paths written."""
    from PIL import Image

    os.makedirs(out_dir, exist_ok=True)
    paths = []
    for t, per_obj in video_segments.items():
        for _, m in per_obj.items():
            arr = ((np.asarray(m) > 0) * 255).astype(np.uint8)
            name = f"{frame_names[t]}.png" if frame_names is not None else f"{t:05d}.png"
            path = os.path.join(out_dir, name)
            Image.fromarray(arr).save(path)
            paths.append(path)
            break
    return paths


def merge_object_masks(obj_masks: np.ndarray) -> np.ndarray:
    """Refer-DAVIS annotator merge (eval_referdavis_infer.py:247-255): obj_masks float/bool [num_obj, T, H, W] -> uint8 [T, H, W] with 0 = background and
    i + 1 = object i: values < 0.5 are zeroed, a constant 0.1 background plane is prepended and the argmax over planes is taken (first maximum wins,
    so overlapping objects resolve to the lowest object index)."""
    m = np.asarray(obj_masks, dtype=np.float32).copy()
    m[m < 0.5] = 0.0
    planes = np.concatenate([np.full((1,) + m.shape[1:], 0.1, np.float32), m], axis=0)
    return np.argmax(planes, axis=0).astype(np.uint8)


def save_palette_pngs(index_masks: np.ndarray, out_dir: str, palette) -> list[str]:
    """Refer-DAVIS output (eval_referdavis_infer.py:257-264): uint8 [T, H, W] object-index maps -> palettised PNGs `{f:05d}.png`; `palette` is the flat
    768-entry list the script takes from a DAVIS annotation (`Image.getpalette()`)."""
    from PIL import Image

    os.makedirs(out_dir, exist_ok=True)
    paths = []
    for f in range(index_masks.shape[0]):
        im = Image.fromarray(np.asarray(index_masks[f], dtype=np.uint8))
        im.putpalette(palette)
        path = os.path.join(out_dir, "{:05d}.png".format(f))
        im.save(path)
        paths.append(path)
    return paths


def remove_small_blobs(binary_mask: np.ndarray, min_size: int = 0) -> np.ndarray:
    """eval_gcg_infer.py:20-29 -> skimage.morphology.remove_small_objects(mask.astype(bool), min_size): connected components with FEWER than
    `min_size` pixels are cleared; skimage's default connectivity for a 2-D array is 1, i.e. the 4-neighbourhood (diagonal pixels are separate
    objects) — restated with scipy.ndimage because skimage is not a dependency here. dtype is preserved like the reference does."""
    if min_size <= 0:
        return binary_mask
    from scipy import ndimage

    dtype = binary_mask.dtype
    fg = binary_mask.astype(bool)
    lab, n = ndimage.label(fg)                       # default structure: 4-connectivity
    if n:
        sizes = np.bincount(lab.ravel())
        small = sizes < min_size
        small[0] = False
        fg = fg & ~small[lab]
    return fg.astype(dtype)


def save_gcg_result(save_dir: str, video_segments: dict, texts: dict, min_size: int = 20) -> dict:
    """GCG output (eval_gcg_infer.py:172-219): `res.json` with the six text fields the script stores and, per object, `pred_masks_{obj}/mask_{i}.png`
    (frame order) after `remove_small_blobs(mask > 0, min_size=20)`; masks are written as 8-bit {0, 255} images, which is what
    `skimage.img_as_ubyte(bool)` yields. Returns {obj_id: bool ndarray [T, H, W]} (the cleaned masks)."""
    from PIL import Image

    os.makedirs(save_dir, exist_ok=True)
    keys = ("gt_text", "gt_text_cleaned", "gt_phrases", "pred_text", "pred_text_cleaned", "pred_phrases")
    with open(os.path.join(save_dir, "res.json"), "w") as f:
        json.dump({k: texts.get(k) for k in keys}, f)
    per_obj: dict = {}
    for t in video_segments:
        for obj_id, m in video_segments[t].items():
            per_obj.setdefault(obj_id, []).append(remove_small_blobs(np.asarray(m) > 0, min_size=min_size))
    out = {}
    for obj_id, masks in per_obj.items():
        d = os.path.join(save_dir, f"pred_masks_{obj_id}")
        os.makedirs(d, exist_ok=True)
        stacked = np.stack(masks)
        for i, m in enumerate(stacked):
            Image.fromarray((m * 255).astype(np.uint8)).save(os.path.join(d, f"mask_{i}.png"))
        out[obj_id] = stacked
    return out
