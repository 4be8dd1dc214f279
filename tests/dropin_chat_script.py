"""Run in a SUBPROCESS by tests/test_dropin_cpu.py (a clean interpreter: `model` must resolve to the drop-in package, not to an already imported
reference package). Executes the reference's UNMODIFIED chat.initialize_model_videogptplus against a synthetic exported checkpoint directory."""
import json
import os
import sys
import tempfile
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("VG_REFERENCE_ROOT", "/root/reference")
sys.path.insert(0, ROOT)
from oracle import lmm as M, ref_harness as H, sam2 as O, weights as Wt  # noqa: E402

H.install_shims()                                   # stubs for decord / matplotlib / hydra / timm; puts <reference>/VideoGLaMM on sys.path
sys.modules["decord"].VideoReader = sys.modules["decord"].cpu = object      # chat.py imports these names at module level (video decoding is not exercised)
sys.path.insert(0, os.path.join(ROOT, "dropin"))    # ... and the drop-in `model` package in front of it
os.environ["VIDEOGLAMM_REFERENCE_ROOT"] = os.path.join(REF, "VideoGLaMM")

work = tempfile.mkdtemp()
os.chdir(work)                                      # the reference resolves "openai/clip-vit-large-patch14-336" relative to the cwd when it is a directory
seg = 300
arch = dict(M.TINY_LMM)
ocfg, scfg = M.make_cfg(**arch, seg_token_idx=seg), O.Sam2Cfg(**O.TINY)
sd = Wt.make_videoglamm_state_dict(ocfg, scfg, seed=7)
ckpt = os.path.join(work, "ckpt")
clip_dir = os.path.join(work, "openai", "clip-vit-large-patch14-336")
os.makedirs(ckpt)
os.makedirs(clip_dir)
clip_pre, iv2_pre = "model.image_vision_tower.vision_tower.", "model.vision_tower."
from safetensors.torch import save_file  # noqa: E402
save_file({k[len(clip_pre):]: v.contiguous() for k, v in sd.items() if k.startswith(clip_pre)}, os.path.join(clip_dir, "model.safetensors"))
from transformers import CLIPImageProcessor  # noqa: E402
CLIPImageProcessor(size={"shortest_edge": 336}, crop_size={"height": 336, "width": 336}).save_pretrained(clip_dir)
iv2_pt = os.path.join(work, "iv2.pt")
torch.save({k[len(iv2_pre):]: v for k, v in sd.items() if k.startswith(iv2_pre)}, iv2_pt)
torch.save({k: v for k, v in sd.items() if not k.startswith(clip_pre) and not k.startswith(iv2_pre)}, os.path.join(ckpt, "pytorch_model.bin"))
p = ocfg.phi
json.dump(dict(vocab_size=p.vocab, hidden_size=p.hidden, num_hidden_layers=p.layers, num_attention_heads=p.heads, intermediate_size=p.inter,
               rope_theta=p.rope_theta, rms_norm_eps=p.eps, sliding_window=p.sliding_window, seg_token_idx=seg, out_dim=256, use_sam2=True,
               mm_vision_tower=iv2_pt, image_mm_vision_tower="openai/clip-vit-large-patch14-336", eos_token_id=2, max_sequence_length=4096,
               vg_arch_overrides={k: v for k, v in arch.items() if k != "phi"}, vg_sam2_overrides={k: (list(v) if isinstance(v, tuple) else v) for k, v in O.TINY.items()}),
          open(os.path.join(ckpt, "config.json"), "w"))
json.dump(dict(eos_token_id=[2, 7, 9]), open(os.path.join(ckpt, "generation_config.json"), "w"))

import chat  # noqa: E402  the reference's own chat.py, unmodified
assert chat.__file__.startswith(REF), chat.__file__


class Tok:
    """Minimal tokenizer standing in for AutoTokenizer.from_pretrained (no Phi-3 tokenizer files exist offline)."""
    unk_token, eos_token, pad_token = "<unk>", "</s>", None
    eos_token_id, bos_token_id, pad_token_id = 2, 1, 0

    def __init__(self):
        self.vocab = {f"t{i}": i for i in range(seg)}

    def add_tokens(self, toks, special_tokens=False):
        toks = [toks] if isinstance(toks, str) else toks
        n = 0
        for t in toks:
            if t not in self.vocab:
                self.vocab[t] = len(self.vocab)
                n += 1
        return n

    def convert_tokens_to_ids(self, t):
        return self.vocab[t]

    def __len__(self):
        return len(self.vocab)


chat.AutoTokenizer = types.SimpleNamespace(from_pretrained=lambda *a, **k: Tok())
import model.VideoGLaMM as MV  # noqa: E402
from videoglamm_b200.model.VideoGLaMM import VideoGLaMMForCausalLM as Ours  # noqa: E402
assert MV.VideoGLaMMForCausalLM is Ours, "model.VideoGLaMM did not resolve to the drop-in package"
import model.videogpt_plus.constants as ref_constants  # noqa: E402  (reference host code still resolves through the merged package path)
assert ref_constants.__file__.startswith(REF)
if not torch.cuda.is_available():
    Ours.cuda = lambda self, *a, **k: self          # no GPU in this container: the engines are not built; everything else runs as written

model_, tokenizer, enc_pre, conv_gen, sam_pre = chat.initialize_model_videogptplus(ckpt, "bf16", 0, False, False, True, "phi3")
assert isinstance(model_, Ours)
assert model_.config.seg_token_idx == seg + 0 and tokenizer.convert_tokens_to_ids("[SEG]") == seg
assert model_.lmm_cfg.phi.vocab == len(tokenizer)                               # resize_token_embeddings(len(tokenizer)) grew 320 -> 301? no: shrank/grew to the tokenizer
assert model_._sd is None or model_._sd["lm_head.weight"].shape[0] == len(tokenizer)
assert model_.config.eos_token_id == 2 and model_._eos_ids() == (2, 7, 9)       # generation_config.json ids still stop generation
assert model_.config.max_sequence_length == 4096
assert any(k.startswith(clip_pre) for k in (model_._sd or {})) or model_.lmm is not None
assert type(enc_pre).__module__ == "utils.enc_preprocessors" and type(sam_pre).__name__ == "SAM_v2_Preprocess"
print("DROPIN_OK", type(model_).__module__, len(tokenizer))
if torch.cuda.is_available():                                                   # GPU box: the reference is absent there, so this branch only runs here with a GPU
    ids = torch.cat([torch.randint(3, 290, (5,)), torch.full((4,), -200), torch.randint(3, 290, (7,))])[None].cuda()
    out = model_.inference([torch.randn(4, 3, 56, 56).cuda()], [torch.randn(4, 3, 56, 56).cuda()], [torch.randn(2, 3, 256, 256).cuda()], ids, None, [(48, 64)],
                           max_new_tokens=3, use_sam2_video_branch=True)
    print("INFERENCE_OK", out[0].shape)
