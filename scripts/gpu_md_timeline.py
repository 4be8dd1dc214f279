"""Stage timeline of the fused mask-decoder kernel at full size (HW = 4096, 128 row blocks): %globaltimer of block (0,0) after every grid barrier
(vg_mask_decoder_set_timeline). Run on the GPU box:  python scripts/gpu_md_timeline.py [B]"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from videoglamm_b200 import _lib as L, build, ops  # noqa: E402
from videoglamm_b200.config import Sam2Config  # noqa: E402
from videoglamm_b200.sam2_engine import Sam2Engine  # noqa: E402
from videoglamm_b200.statedict import random_state_dict, sam2_param_shapes  # noqa: E402

NAMES = []
for l in (0, 1):
    NAMES += [f"L{l} T1 self-attn qkv" + (" (+I2 of L0)" if l else " (+load rows)"), f"L{l} T2 attn core + out proj", f"L{l} T3 norm1 + t2i q + K,V proj (2 img GEMMs)",
              f"L{l} I1 t2i partials", f"L{l} T4a merge", f"L{l} T4b out proj", f"L{l} T5 norm2 + fc1", f"L{l} T6 fc2 (4 chunks)", f"L{l} T7 norm3 + i2t k,v + img q GEMM"]
NAMES += ["F  I2 of L1 + final q + K,V proj", "F  partials", "F  merge", "F  out proj", "H1 norm_final + hidden 1", "H2 hidden 2", "H3 last layers",
          "UP upscaling (up1 GEMM, LN, 4 x up2 GEMM) + masks"]


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    build.build()
    cfg = Sam2Config()
    dev = "cuda"
    sd = random_state_dict(sam2_param_shapes(cfg), dev, seed=1)
    eng = Sam2Engine(sd, cfg, dev)
    HW, Cc = eng.HW, 256
    g = torch.Generator().manual_seed(0)
    pix = torch.randn(B * HW, Cc, generator=g).to(dev)
    s0, s1 = torch.randn(16 * HW, 32, generator=g).to(dev), torch.randn(4 * HW, 64, generator=g).to(dev)
    sparse = torch.randn(B, 3, Cc, generator=g).to(dev)
    tl = torch.zeros(64, dtype=torch.int64, device=dev)
    lib = L.lib()
    lib.vg_mask_decoder_set_timeline.restype = None
    for it in range(5):
        lib.vg_mask_decoder_set_timeline(C.c_void_p(tl.data_ptr()) if it == 4 else None)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.mask_decoder(pix, s0, s1, sparse)
        e1.record()
        torch.cuda.synchronize()
    lib.vg_mask_decoder_set_timeline(None)
    t = tl.cpu().tolist()
    nbar = t[63]
    print(f"B={B}: kernel {e0.elapsed_time(e1) * 1e3:.1f} us (event, incl. memset node); {nbar} barriers; block(0,0) span {(t[nbar + 1] - t[0]) / 1e3:.1f} us")
    for i in range(nbar + 1):
        print(f"  {(t[i + 1] - t[i]) / 1e3:7.1f} us  {NAMES[i] if i < len(NAMES) else '?'}")
    sub = ["T2 start", "T2 qkv rows loaded", "T2 attention core done", "T2 out proj issued -> barrier", "T3 start", "T3 LN done", "T3 q proj done", "T3 K GEMM done", "T3 V GEMM done"]
    st = t[32:32 + 30]
    print("  intra-stage stamps of layer 0 (relative to the previous stamp):")
    for i in range(1, 9):
        if st[i]:
            print(f"  {(st[i] - st[i - 1]) / 1e3:7.2f} us  -> {sub[i]}")


if __name__ == "__main__":
    main()
