"""Mirror of the reference's `model/` package surface for the grounded-generation path."""
from .VideoGLaMM import VideoGLaMMForCausalLM  # noqa: F401
