import sys, torch
sys.path.insert(0, ".")
import torch.nn.functional as F
from oracle import sam2 as O, weights as Wt
from videoglamm_b200.sam2_engine import Sam2Engine
from videoglamm_b200.config import Sam2Config
ocfg = O.Sam2Cfg(**O.TINY); ecfg = Sam2Config(**O.TINY)
sd = Wt.make_sam2_state_dict(ocfg, 1234)
sd = {k: v.bfloat16().float() if v.dim() >= 2 else v for k, v in sd.items()}
eng = Sam2Engine(sd, ecfg)
rel = lambda a, b: ((a.float().cpu() - b.float().cpu()).norm() / b.float().cpu().norm()).item()
torch.manual_seed(5678)
images = torch.randn(1, 3, 256, 256); text = torch.randn(1, 1, 256)
fpn, pos = O.forward_image(sd, ocfg, images)
tok = lambda x: x[0].permute(1, 2, 0).reshape(-1, x.shape[1])
for scale in (1.0, 0.2):
    pix = (fpn[2] + sd["no_mem_embed"].reshape(1, -1, 1, 1)) * scale
    sparse, dense = O.prompt_encode(sd, ocfg, 1, text * scale)
    p = "sam_mask_decoder"
    out_tokens = torch.cat([sd[p + ".obj_score_token.weight"], sd[p + ".iou_token.weight"], sd[p + ".mask_tokens.weight"]], 0)
    tokens = torch.cat([out_tokens[None], sparse], 1)
    hs, src = O.two_way_transformer(sd, p + ".transformer", ocfg, pix + dense, O.dense_pe(sd, ocfg), tokens)
    om = O.mask_decoder(sd, ocfg, pix, O.dense_pe(sd, ocfg), sparse, dense, fpn[:2], True)
    dbg = {}
    em = eng.mask_decoder(tok(pix).cuda().contiguous(), tok(fpn[0]).cuda().bfloat16().contiguous(), tok(fpn[1]).cuda().bfloat16().contiguous(), sparse.cuda(), debug=dbg)
    print("scale", scale, "hs", rel(dbg["hs"], hs), "keys", rel(dbg["keys"], src[0]), "masks", rel(em[0].view(1, 4, 64, 64), om[4]), "iou", rel(em[1], om[5]))
    print("  per-token hs err", [round(rel(dbg["hs"][0, i], hs[0, i]), 4) for i in range(hs.shape[1])])
    # attention logit scale in oracle, layer 0 t2i
    lp = p + ".transformer.layers.0.cross_attn_token_to_image"
    q = O.linear(sd, lp + ".q_proj", tokens); k = O.linear(sd, lp + ".k_proj", (pix + dense).flatten(2).permute(0, 2, 1))
    lg = (q.view(1, -1, 8, 16).transpose(1, 2) @ k.view(1, -1, 8, 16).transpose(1, 2).transpose(-1, -2)) / 4
    print("  t2i logits std", lg.std().item(), "max softmax prob mean", lg.softmax(-1).max(-1).values.mean().item())
