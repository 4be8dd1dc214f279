"""The entry points the reference's eval scripts import from chat.py exist with the same names, parameters and early
error behaviour. No GPU, no weights: only the surface is checked here."""
import ast
import inspect
import os
from types import SimpleNamespace

import pytest

REF_CHAT = "/root/reference/VideoGLaMM/chat.py"


def test_surface():
    from videoglamm_b200 import chat
    assert chat.IMAGE_TOKEN_INDEX == -200
    assert list(inspect.signature(chat.initialize_model_videogptplus).parameters) == [
        "model_base", "precision", "local_rank", "load_in_8bit", "load_in_4bit", "use_sam2_video_branch", "base_llm_type"]
    assert list(inspect.signature(chat.preprocess_vision).parameters)[:6] == [
        "np_images", "type", "enc_preprocessor", "sam_preprocessor", "conv_generator", "precision"]
    with pytest.raises(ValueError):                                   # chat.py:266-267
        chat.initialize_model_videogptplus("x", "bf16", 0, False, False, True, "llama3_1")
    with pytest.raises(NotImplementedError):
        chat.initialize_model_videogptplus("x", "bf16", 0, True, False, True, "phi3")


@pytest.mark.reference
@pytest.mark.skipif(not os.path.exists(REF_CHAT), reason="reference tree not present")
def test_parameter_names_equal_the_reference():
    from videoglamm_b200 import chat
    tree = ast.parse(open(REF_CHAT).read())
    ref = {n.name: [a.arg for a in n.args.args] for n in ast.walk(tree) if isinstance(n, ast.FunctionDef)}
    assert ref["initialize_model_videogptplus"] == list(inspect.signature(chat.initialize_model_videogptplus).parameters)
    ours = list(inspect.signature(chat.preprocess_vision).parameters)
    assert ref["preprocess_vision"] == ours[:len(ref["preprocess_vision"])]


def test_eos_ids_normalisation():
    from videoglamm_b200.model import VideoGLaMMForCausalLM
    f = VideoGLaMMForCausalLM._eos_ids
    assert f(SimpleNamespace(config=SimpleNamespace(eos_token_id=[2, 32007]))) == (2, 32007)
    assert f(SimpleNamespace(config=SimpleNamespace(eos_token_id=2))) == (2,)
    assert f(SimpleNamespace(config=SimpleNamespace(eos_token_id=None))) == ()
    assert f(SimpleNamespace(config=SimpleNamespace())) == ()
