"""Host-side pre-processing / prompt building vs the reference's own functions (when the tree is present)
and self-consistency checks that run anywhere."""
import numpy as np
import pytest
import torch

from oracle import ref_harness as H
from videoglamm_b200 import preprocess as P


class FakeTok:
    """whitespace tokenizer with a BOS, enough to exercise tokenizer_image_token's chunk/sentinel/BOS logic"""
    bos_token_id = 1

    def __call__(self, text):
        class R:
            pass
        r = R()
        r.input_ids = [1] + [3 + (sum(map(ord, w)) % 500) for w in text.split()]
        return r


def test_tokenizer_image_token_and_prompt_template():
    tok = FakeTok()
    ids = P.ConvGenerator_VideoGPTPlus(num_frames=8).apply_for_chat("where is the cat ?", "video", tok, device="cpu")
    assert ids.shape[0] == 1 and ids.dtype == torch.long
    flat = ids[0].tolist()
    assert flat[0] == 1 and flat.count(1) == 1                       # exactly one BOS
    run = [i for i, t in enumerate(flat) if t == P.IMAGE_TOKEN_INDEX]
    assert len(run) == 8 and run == list(range(run[0], run[0] + 8))  # NUM_FRAMES contiguous sentinels
    prompt = P.phi3_prompt([(P.PHI3_ROLES[0], "hi"), (P.PHI3_ROLES[1], "")])
    assert prompt == "<|system|>\nYou are a helpful AI assistant.<|end|>\n<|user|>\nhi<|end|>\n<|assistant|>\n"


def test_sam_preprocess_shapes_and_pad_by_repeat():
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (120, 214, 3), dtype=np.uint8)
    t, rs = P.sam_preprocess(img)
    assert t.shape == (3, 1024, 1024) and t.dtype == torch.float32 and rs == (574, 1024)
    enc = P.EncPreprocessor_VideoGPTPlus(num_frames=4)
    from PIL import Image
    d = enc.preprocess([Image.fromarray(img)] * 2)                  # 2 frames padded to 4 by repeating the last (enc_preprocessors.py:145-149)
    assert len(d["images"]) == 4 and d["images"][0].shape == (3, 224, 224) and d["images"][0].dtype == torch.float64
    assert len(d["context_images"]) == 4 and d["context_images"][0].shape == (3, 336, 336)
    assert torch.equal(d["images"][1], d["images"][3])


@pytest.mark.reference
@pytest.mark.skipif(not H.available(), reason="reference tree not present")
def test_matches_reference_functions():
    H.install_shims()
    from model.videogpt_plus.mm_utils import tokenizer_image_token as ref_tok
    from model.videogpt_plus.model.internvideo.utils import VideoTrainProcessor
    from utils.sam_transforms import sam_preprocess as ref_sam
    rng = np.random.default_rng(1)
    frames = [rng.integers(0, 256, (96, 170, 3), dtype=np.uint8) for _ in range(3)]
    mine, rs = P.sam_preprocess(frames[0])
    ref, rrs = ref_sam(frames[0], model_type="sam2")
    assert tuple(rs) == tuple(rrs) and torch.equal(mine, ref)
    ref_v = VideoTrainProcessor(num_frames=3).preprocess([f.copy() for f in frames])["pixel_values"]
    for a, b in zip(P.iv2_frames_to_tensors(frames), ref_v):
        assert torch.equal(a, b)
    tok = FakeTok()
    prompt = P.phi3_prompt([(P.PHI3_ROLES[0], "<image>" * 4 + "\nfind the dog"), (P.PHI3_ROLES[1], "")])
    assert P.tokenizer_image_token(prompt, tok) == ref_tok(prompt, tok)
    from model.videogpt_plus import conversation as C
    conv = C.conv_phi3_instruct.copy()
    conv.messages = []
    conv.append_message(conv.roles[0], "hello")
    conv.append_message(conv.roles[1], "")
    assert conv.get_prompt() == P.phi3_prompt([(P.PHI3_ROLES[0], "hello"), (P.PHI3_ROLES[1], "")])


@pytest.mark.parametrize("H_,W_,oh,ow", [(336, 336, 1024, 1024), (480, 854, 576, 1024), (720, 1280, 576, 1024), (53, 91, 596, 1024), (1500, 1100, 1024, 751)])
def test_pil_bilinear_restatement_is_bit_exact(H_, W_, oh, ow):
    """The fixed-point two-pass resampling that vg_resize_u8_pil executes (coefficient tables from pil_bilinear_coeffs) equals
    PIL.Image.resize(BILINEAR) bit for bit, up- and down-scaling (antialiased support)."""
    from PIL import Image
    img = np.random.default_rng(H_ + W_).integers(0, 256, (H_, W_, 3), dtype=np.uint8)
    ref = np.array(Image.fromarray(img).resize((ow, oh), Image.BILINEAR))
    assert np.array_equal(P.pil_bilinear_resize_numpy(img, oh, ow), ref)
    assert (oh, ow) == P.get_preprocess_shape(H_, W_, 1024) or (H_, W_) == (1500, 1100)


@pytest.mark.parametrize("H_,W_", [(336, 336), (480, 854), (720, 1280), (100, 60), (448, 448), (37, 53), (1080, 1920)])
def test_encoder_preprocessing_restatements_are_bit_exact(H_, W_):
    """What EncPreprocessGPU executes, restated in numpy and pinned against the libraries the reference calls (utils/enc_preprocessors.py:108-166):
    cv2.resize INTER_LINEAR (IV2 frames, internvideo/utils.py:120-131) and PIL BICUBIC shortest-edge resize (HF CLIPImageProcessor), bit for bit,
    down- and up-scaling incl. the exact 2x case."""
    import cv2
    from PIL import Image
    img = np.random.default_rng(H_ * 3 + W_).integers(0, 256, (H_, W_, 3), dtype=np.uint8)
    assert np.array_equal(P.cv_resize_linear_numpy(img, 224, 224), cv2.resize(img, (224, 224)))
    ch, cw = P.clip_resize_shape(H_, W_, 336)
    assert min(ch, cw) == 336
    ref = np.array(Image.fromarray(img).resize((cw, ch), Image.BICUBIC))
    assert np.array_equal(P.pil_resize_numpy(img, ch, cw, "bicubic"), ref)
    # the whole CLIP branch against the HF processor (resize -> centre crop -> rescale -> normalize)
    want = P.clip_image_processor().preprocess(Image.fromarray(img), return_tensors="pt")["pixel_values"][0].numpy()
    r = P.pil_resize_numpy(img, ch, cw, "bicubic") if (ch, cw) != (H_, W_) else img
    top, left = (ch - 336) // 2, (cw - 336) // 2
    v = (r[top:top + 336, left:left + 336].astype(np.float64) * (1 / 255)).astype(np.float32)
    got = ((v - np.array(P.CLIP_MEAN, np.float32)) / np.array(P.CLIP_STD, np.float32)).transpose(2, 0, 1)
    # transformers 4.41 (pinned by the reference) resizes through PIL — the arithmetic restated above; the installed 5.x processor resizes through
    # torchvision and lands one (two when up-scaling) 8-bit levels away on a few percent of the pixels at most: everything else agrees to float32 rounding
    d = np.abs(want - got)
    assert want.shape == got.shape == (3, 336, 336) and d.max() < 2.02 / 255 / min(P.CLIP_STD) and (d > 2e-6).mean() < 0.05
