"""Bring-up: SAM2 engine (CUDA) vs oracle (CPU fp32) stage by stage on identical inputs."""
import sys, time
import torch
sys.path.insert(0, ".")
from oracle import sam2 as O, weights as Wt
from videoglamm_b200.sam2_engine import Sam2Engine
from videoglamm_b200.config import Sam2Config
from videoglamm_b200 import ops

tiny = "--full" not in sys.argv
kw = O.TINY if tiny else {}
ocfg = O.Sam2Cfg(**kw)
ecfg = Sam2Config(**{k: v for k, v in kw.items()})
sd = Wt.make_sam2_state_dict(ocfg, 1234)
sd = {k: v.bfloat16().float() if v.dim() >= 2 else v for k, v in sd.items()}  # both sides see bf16-representable matrices
eng = Sam2Engine(sd, ecfg)
dev = "cuda"
rel = lambda a, b: ((a.float().cpu() - b.float().cpu()).norm() / b.float().cpu().norm()).item()
torch.manual_seed(5678)
S = ocfg.image_size
T, nobj = 3, 2
images = torch.randn(T, 3, S, S)
text = torch.randn(nobj, 1, 256)
# 1. image encoder
t0 = time.time(); fpn, pos = O.forward_image(sd, ocfg, images[:1]); print("oracle forward_image s", time.time() - t0)
f = eng.forward_image(images[:1].cuda()); torch.cuda.synchronize()
tok = lambda x: x[0].permute(1, 2, 0).reshape(-1, x.shape[1])
print("s0", rel(f["s0"], tok(fpn[0])), "s1", rel(f["s1"], tok(fpn[1])), "top", rel(f["top"], tok(fpn[2])))
# 2. mask decoder on identical inputs (oracle's features)
pix = fpn[2] + sd["no_mem_embed"].reshape(1, -1, 1, 1)
sparse, dense = O.prompt_encode(sd, ocfg, 1, text[:1])
om = O.mask_decoder(sd, ocfg, pix, O.dense_pe(sd, ocfg), sparse, dense, fpn[:2], True)
em = eng.mask_decoder(tok(pix).cuda().contiguous(), tok(fpn[0]).cuda().bfloat16().contiguous(), tok(fpn[1]).cuda().bfloat16().contiguous(), sparse.cuda())
print("dec masks", rel(em[0].view(1, 4, 4 * ocfg.feat_size, -1), om[4]), "iou", rel(em[1], om[5]), "obj", em[3].item(), om[3].item(), "dense_pe", rel(eng.dense_pe, tok(O.dense_pe(sd, ocfg))))
# 3. sam heads
oh = O.forward_sam_heads(sd, ocfg, pix, fpn[:2], text_inputs=text[:1])
eh = eng.forward_sam_heads(tok(pix).cuda().contiguous(), tok(fpn[0]).cuda().bfloat16().contiguous(), tok(fpn[1]).cuda().bfloat16().contiguous(), text[:1].cuda(), 1)
print("heads low", rel(eh["low_res_masks"], oh["low_res_masks"][:, 0]), "ptr", rel(eh["obj_ptr"], oh["obj_ptr"]))
# 4. memory encoder
low = oh["low_res_masks"]
high = torch.nn.functional.interpolate(low, size=(S, S), mode="bilinear", align_corners=False)
for binz in (True, False):
    mf, mp = O.encode_new_memory(sd, ocfg, fpn[2], high, is_mask_from_pts=binz)
    emf = eng.encode_memory(tok(fpn[2]).cuda().contiguous(), low[:, 0].cuda().contiguous(), binarize=binz)
    print("memenc bin", binz, rel(emf[0], tok(mf)), "pos", rel(eng.pos_mem, tok(mp)))
# 5. memory attention
Nk = 2 * eng.HW + 8
mem = torch.randn(Nk, nobj, 64).bfloat16().float(); mpos = torch.randn(Nk, 1, 64).expand(-1, nobj, -1)
curr = O._flat(fpn[2]).expand(-1, nobj, -1); cpos = O._flat(pos[2]).expand(-1, nobj, -1)
oma = O.memory_attention(sd, ocfg, curr, cpos, mem, mpos, 8)
ema = eng.memory_attention(tok(fpn[2]).cuda().contiguous(), nobj, mem.permute(1, 0, 2).contiguous().cuda().bfloat16(), mpos[:, 0].contiguous().cuda(), 8)
print("memattn", rel(ema.view(nobj, eng.HW, 256), oma.permute(1, 0, 2)), "pos_top", rel(eng.pos_top, tok(pos[2])))
# 6. full video
t0 = time.time(); oc = {}; oo = O.run_video(sd, ocfg, images, text, (60, 100), collect=oc); print("oracle video s", time.time() - t0)
ec = {}
eo = eng.run_video(images.cuda(), text.cuda(), (60, 100), collect=ec); torch.cuda.synchronize()
for t in range(T):
    a, b = eo[t].cpu(), oo[t][:, 0]
    print(f"frame {t}: rel {rel(a, b):.4f} maxabs {(a - b).abs().max():.3f} refmax {b.abs().max():.1f} mask agreement {((a > 0) == (b > 0)).float().mean():.5f}")
ops.launch_count_reset(); t0 = time.time()
eo = eng.run_video(images.cuda(), text.cuda(), (60, 100)); torch.cuda.synchronize()
print("engine video s", time.time() - t0, "launches", ops.launch_count())
