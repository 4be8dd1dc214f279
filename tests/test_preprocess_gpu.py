"""GPU pre-processing of the SAM2 frames: uint8 frames in, [T,3,1024,1024] fp32 out, vs the reference's CPU path
(PIL bilinear ResizeLongestSide -> normalise -> second bilinear squash; utils/sam_transforms.py:26-65 mirrored by preprocess.sam_preprocess,
which tests/test_preprocess_cpu.py checks against the reference's own function)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from videoglamm_b200 import ops, preprocess as P  # noqa: E402


@pytest.mark.parametrize("H,W", [(336, 336), (480, 854), (720, 1280), (53, 91)])
def test_gpu_resize_is_bit_exact_with_pil_and_squash_matches(H, W):
    from PIL import Image
    rng = np.random.default_rng(H * 7 + W)
    frames = rng.integers(0, 256, (3, H, W, 3), dtype=np.uint8)
    pre = P.SAM_v2_PreprocessGPU()
    (newh, neww), tabs = pre.tables(H, W)
    got = ops.resize_u8_pil(torch.from_numpy(frames).cuda(), newh, neww, tabs).cpu().numpy()
    for t in range(3):
        ref = np.array(Image.fromarray(frames[t]).resize((neww, newh), Image.BILINEAR))
        assert np.array_equal(got[t], ref)                                   # integer work: bit-exact
    out, shape = pre.preprocess_video(frames)
    assert out.shape == (3, 3, 1024, 1024) and out.dtype == torch.float32 and shape == (newh, neww)
    for t in range(3):
        want, rs = P.sam_preprocess(frames[t])
        assert tuple(rs) == shape
        err = (out[t].cpu() - want).abs().max().item()
        assert err < 2e-5, err                                               # fp32: same formula, different association (values are O(2))


def test_inference_accepts_raw_uint8_sam_frames():
    """`images_for_sam` may be the raw uint8 frames [T,H,W,3] (host or device): the engine then pre-processes them on the GPU; masks equal the
    ones obtained from the CPU-preprocessed fp32 tensor."""
    from oracle import lmm as M, sam2 as O, weights as Wt
    from videoglamm_b200.config import Sam2Config, make_lmm_config
    from videoglamm_b200.model import VideoGLaMMForCausalLM
    seg = 300
    ocfg = M.make_cfg(**M.TINY_LMM, seg_token_idx=seg)
    tiny1024 = {k: v for k, v in O.TINY.items() if k != "image_size"}           # the SAM2 input of sam_preprocess is always 1024 px
    sd = Wt.make_videoglamm_state_dict(ocfg, O.Sam2Cfg(**tiny1024), seed=7)
    model = VideoGLaMMForCausalLM(sd, make_lmm_config(**M.TINY_LMM, seg_token_idx=seg), Sam2Config(**tiny1024))
    rng = np.random.default_rng(3)
    frames = rng.integers(0, 256, (2, 60, 100, 3), dtype=np.uint8)
    T = 4
    torch.manual_seed(1)
    images, ctx = torch.randn(T, 3, 56, 56), torch.randn(T, 3, 56, 56)
    ids = torch.cat([torch.randint(3, 290, (5,)), torch.full((T,), -200), torch.randint(3, 290, (7,))])[None].cuda()
    sam_cpu = torch.stack([P.sam_preprocess(f)[0] for f in frames])
    a = model.inference([images.cuda()], [ctx.cuda()], [sam_cpu.cuda()], ids, None, [(60, 100)], max_new_tokens=3, use_sam2_video_branch=True, forced_tokens={1: seg})
    b = model.inference([images.cuda()], [ctx.cuda()], [torch.from_numpy(frames)], ids, None, [(60, 100)], max_new_tokens=3, use_sam2_video_branch=True,
                        forced_tokens={1: seg})
    assert torch.equal(a[0], b[0])
    for t in a[1][0]:
        assert (a[1][0][t][0] == b[1][0][t][0]).mean() > 0.999


@pytest.mark.parametrize("H,W", [(336, 336), (480, 854), (720, 1280), (100, 60)])
def test_gpu_encoder_preprocessing_matches_the_reference_processors(H, W):
    """EncPreprocessGPU (raw uint8 frames -> IV2 frames + CLIP context images on the GPU) against the reference's CPU processors
    (preprocess.EncPreprocessor_VideoGPTPlus = cv2.resize + float64 normalise / HF CLIPImageProcessor): the integer resampling bit-exact, the
    normalised tensors to float32 rounding."""
    import cv2
    from PIL import Image
    rng = np.random.default_rng(H + 11 * W)
    frames = rng.integers(0, 256, (4, H, W, 3), dtype=np.uint8)
    pre = P.EncPreprocessGPU()
    cv_tabs, (ch, cw), pil = pre.tables(H, W)
    small = ops.resize_u8_cv(torch.from_numpy(frames).cuda(), 224, 224, cv_tabs).cpu().numpy()
    for t in range(4):
        assert np.array_equal(small[t], cv2.resize(frames[t], (224, 224)))
    if pil is not None:
        big = ops.resize_u8_pil(torch.from_numpy(frames).cuda(), ch, cw, pil).cpu().numpy()
        for t in range(4):
            assert np.array_equal(big[t], np.array(Image.fromarray(frames[t]).resize((cw, ch), Image.BICUBIC)))
    images, context = pre.preprocess_video(frames)
    ref = P.EncPreprocessor_VideoGPTPlus(num_frames=4).preprocess([frames[t] for t in range(4)])
    want_im = torch.stack(ref["images"]).float()
    want_ctx = torch.stack(ref["context_images"]).float()
    assert images.shape == (4, 3, 224, 224) and context.shape == (4, 3, 336, 336)
    assert torch.equal(images.cpu(), want_im)                                   # float64 arithmetic rounded once: identical
    d = (context.cpu() - want_ctx).abs()                                         # installed HF 5.x resizes through torchvision, not PIL (<= 2 levels on < 5 %)
    assert d.max().item() < 2.02 / 255 / min(P.CLIP_STD) and (d > 2e-6).float().mean().item() < 0.05
    r = frames if pil is None else big                                          # ... and exactly the PIL-based arithmetic of transformers 4.41
    top, left = (ch - 336) // 2, (cw - 336) // 2
    v = (r[:, top:top + 336, left:left + 336].astype(np.float64) * (1 / 255)).astype(np.float32)
    want41 = ((v - np.array(P.CLIP_MEAN, np.float32)) / np.array(P.CLIP_STD, np.float32)).transpose(0, 3, 1, 2)
    assert np.array_equal(context.cpu().numpy(), want41)


def test_inference_accepts_raw_uint8_frames_for_all_three_inputs():
    """`images` / `context_images` may be the raw uint8 frames too: the engine pre-processes the IV2, CLIP and SAM2 inputs on the GPU; ids and masks
    equal the ones obtained from the CPU-preprocessed tensors (cv2 / HF processors of preprocess.EncPreprocessor_VideoGPTPlus at the tiny sizes)."""
    from oracle import lmm as M, sam2 as O, weights as Wt
    from videoglamm_b200.config import Sam2Config, make_lmm_config
    from videoglamm_b200.model import VideoGLaMMForCausalLM
    seg = 300
    ocfg = M.make_cfg(**M.TINY_LMM, seg_token_idx=seg)
    tiny1024 = {k: v for k, v in O.TINY.items() if k != "image_size"}
    sd = Wt.make_videoglamm_state_dict(ocfg, O.Sam2Cfg(**tiny1024), seed=7)
    model = VideoGLaMMForCausalLM(sd, make_lmm_config(**M.TINY_LMM, seg_token_idx=seg), Sam2Config(**tiny1024))
    model._ensure_engines()
    size_iv2, size_clip = model.lmm.cfg.iv2.image_size, model.lmm.cfg.clip.image_size
    rng = np.random.default_rng(5)
    frames = rng.integers(0, 256, (4, 60, 100, 3), dtype=np.uint8)
    pre = P.EncPreprocessGPU("cuda", size_iv2, size_clip)
    images, ctx = pre.preprocess_video(frames)
    assert images.shape == (4, 3, size_iv2, size_iv2) and ctx.shape == (4, 3, size_clip, size_clip)
    ids = torch.cat([torch.randint(3, 290, (5,)), torch.full((4,), -200), torch.randint(3, 290, (7,))])[None].cuda()
    raw = torch.from_numpy(frames)
    a = model.inference([images], [ctx], [raw[:2]], ids, None, [(60, 100)], max_new_tokens=3, use_sam2_video_branch=True, forced_tokens={1: seg})
    b = model.inference([raw], [raw], [raw[:2]], ids, None, [(60, 100)], max_new_tokens=3, use_sam2_video_branch=True, forced_tokens={1: seg})
    assert torch.equal(a[0], b[0])
    for t in a[1][0]:
        assert np.array_equal(a[1][0][t][0], b[1][0][t][0])
