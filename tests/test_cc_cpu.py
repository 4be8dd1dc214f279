"""Oracle for the connected-components operator (oracle/cc.py) against hand-computed known answers and an independent flood fill."""
import numpy as np

from oracle import cc as O


def test_known_answers():
    # 4 x 6 mask, W = 6. Blocks (top-left linear index): row 0-1: 0, 2, 4; row 2-3: 12, 14, 16.
    m = np.array([[1, 0, 0, 0, 0, 1],
                  [0, 1, 0, 0, 0, 0],
                  [0, 0, 1, 0, 0, 0],
                  [0, 0, 0, 0, 1, 1]], np.uint8)[None, None]
    lab, cnt = O.connected_components(m)
    # diagonal chain (0,0)-(1,1)-(2,2) is ONE 8-connected component whose smallest block is 0 -> label 1, area 3;
    # (0,5) alone lies in block 4 -> label 5, area 1; (3,4),(3,5) lie in block 16 -> label 17, area 2.
    want_l = np.array([[1, 0, 0, 0, 0, 5], [0, 1, 0, 0, 0, 0], [0, 0, 1, 0, 0, 0], [0, 0, 0, 0, 17, 17]], np.int32)
    want_c = np.array([[3, 0, 0, 0, 0, 1], [0, 3, 0, 0, 0, 0], [0, 0, 3, 0, 0, 0], [0, 0, 0, 0, 2, 2]], np.int32)
    assert np.array_equal(lab[0, 0], want_l) and np.array_equal(cnt[0, 0], want_c)
    # the label is the block's top-left index even when that pixel is background: a single pixel at (1,1) -> block 0 -> label 1
    m2 = np.zeros((1, 1, 2, 2), np.uint8)
    m2[0, 0, 1, 1] = 1
    lab, cnt = O.connected_components(m2)
    assert lab[0, 0].tolist() == [[0, 0], [0, 1]] and cnt[0, 0].tolist() == [[0, 0], [0, 1]]
    # empty and full
    z = np.zeros((2, 1, 4, 4), np.uint8)
    lab, cnt = O.connected_components(z)
    assert not lab.any() and not cnt.any()
    lab, cnt = O.connected_components(z + 1)
    assert (lab == 1).all() and (cnt == 16).all()


def test_scipy_restatement_equals_flood_fill():
    rng = np.random.default_rng(0)
    for density in (0.2, 0.45, 0.6, 0.9):
        m = (rng.random((3, 1, 18, 26)) < density).astype(np.uint8)
        a, b = O.connected_components(m), O.connected_components_slow(m)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
