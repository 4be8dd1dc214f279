"""TEST INFRASTRUCTURE (oracle) — CPU restatement of the reference's connected-components extension
(model/segment_anything_2/sam2/csrc/connected_components.cu:213-289, bound by sam2/utils/misc.py:47-63).

The reference is a CUDA extension (torch C++ API): it cannot run in a container without a GPU, so this restates its OUTPUT CONTRACT, derived
from the kernels: `merge` (:75-118) unions 2x2 blocks whose foreground pixels are 8-adjacent (pixels of one block always are), `union_` (:42-59)
keeps the minimum block index as root, `final_labeling` (:129-166) writes root + 1 to every foreground pixel of a block, `init_counting` /
`final_counting` (:168-201) give every foreground pixel the number of foreground pixels carrying its label. Block index = linear index
row * W + col of the block's top-left pixel (even row, even col), whether or not that pixel is foreground.
Pinning: hand-computed known answers in tests/test_cc_cpu.py; on the GPU box the test also runs the reference kernel itself when
oracle/_ref/ holds the extension built by oracle/build_ref.py from the unmodified source."""
from __future__ import annotations

import numpy as np


def connected_components(mask: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """mask [N,1,H,W] (nonzero = foreground), H and W even -> (labels int32, counts int32), both [N,1,H,W]."""
    from scipy import ndimage

    n, c, h, w = mask.shape
    assert c == 1 and h % 2 == 0 and w % 2 == 0
    rows, cols = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    block_idx = (rows & ~1) * w + (cols & ~1)
    labels = np.zeros((n, 1, h, w), np.int32)
    counts = np.zeros((n, 1, h, w), np.int32)
    for i in range(n):
        fg = mask[i, 0] != 0
        comp, k = ndimage.label(fg, structure=np.ones((3, 3), int))
        if k == 0:
            continue
        idx = np.arange(1, k + 1)
        root = ndimage.minimum(block_idx, comp, idx).astype(np.int64)
        area = ndimage.sum(fg, comp, idx).astype(np.int64)
        lut_l = np.concatenate([[0], root + 1]).astype(np.int32)
        lut_c = np.concatenate([[0], area]).astype(np.int32)
        labels[i, 0] = lut_l[comp]
        counts[i, 0] = lut_c[comp]
    return labels, counts


def connected_components_slow(mask: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """Pure-python flood fill (small cases only): independent of scipy, used to cross-check `connected_components`."""
    n, _, h, w = mask.shape
    labels = np.zeros((n, 1, h, w), np.int32)
    counts = np.zeros((n, 1, h, w), np.int32)
    for i in range(n):
        fg = mask[i, 0] != 0
        seen = np.zeros((h, w), bool)
        for r in range(h):
            for c in range(w):
                if not fg[r, c] or seen[r, c]:
                    continue
                stack, comp = [(r, c)], []
                seen[r, c] = True
                while stack:
                    y, x = stack.pop()
                    comp.append((y, x))
                    for dy in (-1, 0, 1):
                        for dx in (-1, 0, 1):
                            yy, xx = y + dy, x + dx
                            if 0 <= yy < h and 0 <= xx < w and fg[yy, xx] and not seen[yy, xx]:
                                seen[yy, xx] = True
                                stack.append((yy, xx))
                root = min((y & ~1) * w + (x & ~1) for y, x in comp)
                for y, x in comp:
                    labels[i, 0, y, x] = root + 1
                    counts[i, 0, y, x] = len(comp)
    return labels, counts
