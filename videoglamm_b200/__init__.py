"""videoglamm_b200 — B200-native (sm_100a) engine for VideoGLaMM's grounded-generation forward pass.

Only the hot path lives here: `csrc/` (hand-written CUDA + C ABI, built into `lib/libvglamm.so`), `_lib.py`
(ctypes binding), `ops.py` (tensor-level wrappers) and the host-side mirror of the reference's `model/` surface.
"""
__version__ = "0.1.0"
