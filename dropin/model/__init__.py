"""Drop-in `model` package: put this directory (`<repo>/dropin`) in front of the reference checkout on PYTHONPATH and the reference's UNMODIFIED
`chat.py` / `eval_*_infer.py` run on the B200 engine:

    PYTHONPATH=<repo>/dropin:<repo>:<reference>/VideoGLaMM python <reference>/VideoGLaMM/chat.py ...

`from model.VideoGLaMM import VideoGLaMMForCausalLM` (chat.py:278) then resolves to the engine-backed class, while every other `model.*` import the
reference's host code makes (model.videogpt_plus.constants / conversation / mm_utils, model.segment_anything.utils.transforms, model.chatunivi.*:
pre-processing and prompt building, utils/*.py) still resolves to the reference's own package: its `model/` directory is appended to this package's
search path. The reference root is taken from $VIDEOGLAMM_REFERENCE_ROOT or found on sys.path."""
import os
import sys


def _reference_model_dir():
    cands = [os.environ.get("VIDEOGLAMM_REFERENCE_ROOT", "")] + list(sys.path)
    here = os.path.dirname(os.path.abspath(__file__))
    for c in cands:
        d = os.path.join(c or ".", "model")
        if os.path.isdir(os.path.join(d, "videogpt_plus")) and os.path.abspath(d) != here:
            return os.path.abspath(d)
    return None


_ref = _reference_model_dir()
if _ref is not None and _ref not in __path__:
    __path__.append(_ref)
