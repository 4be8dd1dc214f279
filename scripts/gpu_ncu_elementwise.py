"""One launch of each elementwise / layout kernel at its config-2 shape for `ncu --set full` (roofline evidence for the glue).
Order: LayerNorm Hiera stage 3 (32768 x 576 fp32 -> bf16), LayerNorm stage 2 (131072 x 288), 2x2 max pool of the stage-2 -> 3 transition q (bf16),
depthwise 7x7 of the memory fuser (64 x 64 x 256 fp32), table RoPE on packed q|k (4096 x 512 of 768), residual-free add (pos embed, 524288 x 144),
im2col of the 7x7/4 patch embedding (8 frames 1024^2), bit-packed final upsampling (8 planes 256^2 -> 336^2)."""
import sys, torch
sys.path.insert(0, ".")
from videoglamm_b200 import ops
dev = "cuda"
BF = torch.bfloat16
x = torch.randn(32768, 576, device=dev); w, b = torch.randn(576, device=dev), torch.randn(576, device=dev)
ops.norm(x, w, b, eps=1e-6, out_dtype=BF); torch.cuda.synchronize()
x = torch.randn(131072, 288, device=dev); w, b = torch.randn(288, device=dev), torch.randn(288, device=dev)
ops.norm(x, w, b, eps=1e-6, out_dtype=BF); torch.cuda.synchronize()
q = torch.randn(8 * 128 * 128, 3 * 576, device=dev).bfloat16()[:, :576]
ops.pool2x2(q, 8, 128, 128, is_max=True); torch.cuda.synchronize()
x = torch.randn(64 * 64, 256, device=dev); dw, db = torch.randn(256, 7, 7, device=dev), torch.randn(256, device=dev)
ops.dwconv(x, dw, db, 1, 64, 64, 7); torch.cuda.synchronize()
table = ops.rope2d_table(256, 64, 64, dev); torch.cuda.synchronize()
qkv = torch.randn(4096, 768, device=dev).bfloat16()
ops.rope2d_(qkv[:, :512], 256, 64, 64, table=table, nblk=2); torch.cuda.synchronize()
a, p = torch.randn(8 * 65536, 144, device=dev), torch.randn(65536, 144, device=dev)
ops.add(a, p); torch.cuda.synchronize()
img = torch.randn(8, 3, 1024, 1024, device=dev)
ops.im2col(img, 7, 4, 3, kpad=152, out_dtype=BF); torch.cuda.synchronize()
low = torch.randn(8, 256, 256, device=dev)
ops.bilinear(low, 336, 336, mode=4); torch.cuda.synchronize()
