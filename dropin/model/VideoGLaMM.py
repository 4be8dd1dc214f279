"""`model.VideoGLaMM` of the drop-in package: the reference's class name backed by the B200 engine (videoglamm_b200/model/VideoGLaMM.py)."""
from videoglamm_b200.model.VideoGLaMM import VideoGLaMMForCausalLM  # noqa: F401
