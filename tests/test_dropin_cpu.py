"""The reference's UNMODIFIED chat.initialize_model_videogptplus runs against the drop-in `model` package (dropin/) and a synthetic exported
checkpoint: from_pretrained kwargs, tokenizer / [SEG] resize, config assignments, initialize_vision_modules, tower handles, dtype / device calls."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.reference
@pytest.mark.skipif(not os.path.exists("/root/reference/VideoGLaMM/chat.py"), reason="reference tree not present")
def test_reference_chat_initialize_runs_on_the_dropin_package():
    env = dict(os.environ, NUM_FRAMES="4")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "dropin_chat_script.py")], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0 and "DROPIN_OK videoglamm_b200.model.VideoGLaMM" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
