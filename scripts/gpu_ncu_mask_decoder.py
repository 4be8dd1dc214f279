"""Launches the fused SAM2 mask decoder at full size (HW = 4096 -> 128 row blocks, B = 1 object, 9 tokens): tensor-core (split-bf16) variant, then the
fp32-FMA variant. For `ncu -k regex:mask_decoder_kernel --set full` (only these two launches are profiled)."""
import sys, torch
sys.path.insert(0, ".")
from videoglamm_b200.config import Sam2Config
from videoglamm_b200.sam2_engine import Sam2Engine
from videoglamm_b200.statedict import random_state_dict, sam2_param_shapes
dev = "cuda"
cfg = Sam2Config()
eng = Sam2Engine(random_state_dict(sam2_param_shapes(cfg), dev, seed=1), cfg, dev)
HW = eng.HW
pix = torch.randn(HW, 256, device=dev); s0 = torch.randn(16 * HW, 32, device=dev); s1 = torch.randn(4 * HW, 64, device=dev); sp = torch.randn(1, 3, 256, device=dev)
torch.cuda.synchronize()
eng.mask_decoder(pix, s0, s1, sp)
torch.cuda.synchronize()
eng.fused_exact = True
eng.mask_decoder(pix, s0, s1, sp)
torch.cuda.synchronize()
