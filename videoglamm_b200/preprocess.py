"""Host-side pre-processing and prompt building with the reference's semantics. These are CPU
image/string transforms in the reference as well (cv2 / PIL / HF CLIPImageProcessor / tokenizer); the arithmetic-heavy path starts at
the tensors they produce. Paths cited are relative to /root/reference/VideoGLaMM/."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

IMAGE_TOKEN_INDEX = -200                 # model/videogpt_plus/constants.py:12
DEFAULT_IMAGE_TOKEN, DEFAULT_VIDEO_TOKEN = "<image>", "<video>"
CHUNK_SIZE = 4

IMAGENET_MEAN, IMAGENET_STD = np.array([0.485, 0.456, 0.406]).reshape(1, 1, 3), np.array([0.229, 0.224, 0.225]).reshape(1, 1, 3)
CLIP_MEAN, CLIP_STD = [0.48145466, 0.4578275, 0.40821073], [0.26862954, 0.26130258, 0.27577711]
SAM_MEAN, SAM_STD = torch.tensor([123.675, 116.28, 103.53]).view(-1, 1, 1), torch.tensor([58.395, 57.12, 57.375]).view(-1, 1, 1)


def iv2_frames_to_tensors(frames: list, size: int = 224) -> list[torch.Tensor]:
    """VideoTrainProcessor.preprocess / frames2tensor (model/videogpt_plus/model/internvideo/utils.py:104-143): cv2 bilinear resize,
    (x/255 - mean)/std in float64, HWC -> CHW."""
    import cv2
    out = []
    for f in frames:
        x = cv2.resize(np.asarray(f), (size, size))
        x = (x / 255.0 - IMAGENET_MEAN) / IMAGENET_STD
        out.append(torch.from_numpy(np.transpose(x, (2, 0, 1))))
    return out


def clip_image_processor(size: int = 336):
    """HF CLIPImageProcessor with openai/clip-vit-large-patch14-336's preprocessor_config (built locally: no hub access needed)."""
    from transformers import CLIPImageProcessor
    return CLIPImageProcessor(do_resize=True, size={"shortest_edge": size}, resample=3, do_center_crop=True, crop_size={"height": size, "width": size},
                              do_rescale=True, rescale_factor=1 / 255, do_normalize=True, image_mean=CLIP_MEAN, image_std=CLIP_STD, do_convert_rgb=True)


class EncPreprocessor_VideoGPTPlus:
    """utils/enc_preprocessors.py:108-166 (video branch): trim / pad-by-repeat to `num_frames`, IV2 frames + CLIP context images."""

    def __init__(self, num_frames: int = 16, image_processor=None):
        self.num_frames = num_frames
        self.image_processor = image_processor or clip_image_processor()

    def preprocess(self, pil_images):
        if not isinstance(pil_images, list):
            return {"images": self.image_processor.preprocess(pil_images, return_tensors="pt")["pixel_values"][0], "context_images": None}
        pil_images = list(pil_images[: self.num_frames])
        while len(pil_images) < self.num_frames:
            pil_images.append(pil_images[-1])
        video_frames = iv2_frames_to_tensors(pil_images)
        context = [self.image_processor.preprocess(im, return_tensors="pt")["pixel_values"][0] for im in pil_images]
        return {"images": video_frames, "context_images": context}


def get_preprocess_shape(oldh: int, oldw: int, long_side: int) -> tuple[int, int]:
    """ResizeLongestSide.get_preprocess_shape (model/segment_anything/utils/transforms.py:103-113)"""
    scale = long_side * 1.0 / max(oldh, oldw)
    return int(oldh * scale + 0.5), int(oldw * scale + 0.5)


def sam_preprocess(x: np.ndarray, img_size: int = 1024):
    """sam_preprocess(model_type="sam2") utils/sam_transforms.py:26-65: PIL bilinear ResizeLongestSide(1024) on uint8, normalise, then a
    SECOND bilinear squash to 1024x1024 (F.interpolate, align_corners=False). Returns ([3,1024,1024] fp32, resize_shape)."""
    from PIL import Image
    newh, neww = get_preprocess_shape(x.shape[0], x.shape[1], img_size)
    r = np.array(Image.fromarray(x).resize((neww, newh), Image.BILINEAR))
    t = torch.from_numpy(r).permute(2, 0, 1).contiguous()
    t = (t - SAM_MEAN) / SAM_STD
    t = F.interpolate(t.unsqueeze(0), (img_size, img_size), mode="bilinear").squeeze(0)
    return t, (newh, neww)


class SAM_v2_Preprocess:
    def preprocess(self, x: np.ndarray):
        return sam_preprocess(x)


def _pil_filter(name: str):
    """Pillow's resampling kernels (src/libImaging/Resample.c): (function on |x| arrays, support)."""
    if name == "bilinear":
        return (lambda a: np.where(a < 1.0, 1.0 - a, 0.0)), 1.0
    if name == "bicubic":                                        # a = -0.5
        def f(a, A=-0.5):
            return np.where(a < 1.0, ((A + 2.0) * a - (A + 3.0)) * a * a + 1.0, np.where(a < 2.0, (((a - 5.0) * a + 8.0) * a - 4.0) * A, 0.0))
        return f, 2.0
    raise ValueError(name)


def pil_coeffs(in_size: int, out_size: int, filt: str = "bilinear") -> tuple[np.ndarray, np.ndarray]:
    """Pillow's precompute_coeffs + normalize_coeffs_8bpc (src/libImaging/Resample.c) for the BILINEAR (support 1) or BICUBIC (support 2, a = -0.5)
    filter, support scaled by the down-sampling factor: int32 coefficients [out, ksize] with 22 fractional bits and bounds [out, 2] = (first input
    index, tap count). Evaluated in float64 exactly as Pillow does, so vg_resize_u8_pil reproduces Image.resize bit for bit (tests/test_preprocess_cpu.py)."""
    import math
    fn, sup = _pil_filter(filt)
    scale = in_size / out_size
    fs = max(scale, 1.0)
    support = sup * fs
    ksize = int(math.ceil(support)) * 2 + 1
    kk = np.zeros((out_size, ksize), dtype=np.float64)
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    inv = 1.0 / fs
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = fn(np.abs((np.arange(xmax) + xmin - center + 0.5) * inv))
        ww = w.sum()
        kk[xx, :xmax] = w / ww if ww != 0 else w
        bounds[xx] = (xmin, xmax)
    ki = np.where(kk < 0, (-0.5 + kk * (1 << 22)).astype(np.int64), (0.5 + kk * (1 << 22)).astype(np.int64)).astype(np.int32)
    return ki, bounds


def pil_bilinear_coeffs(in_size: int, out_size: int) -> tuple[np.ndarray, np.ndarray]:
    return pil_coeffs(in_size, out_size, "bilinear")


def pil_resize_numpy(img: np.ndarray, oh: int, ow: int, filt: str = "bilinear") -> np.ndarray:
    """Host restatement of Pillow's 8-bit two-pass resampling (the arithmetic vg_resize_u8_pil executes); used by the CPU tests."""
    H, W, C = img.shape
    kh, bh = pil_coeffs(W, ow, filt)
    tmp = np.zeros((H, ow, C), np.uint8)
    for x in range(ow):
        lo, n = bh[x]
        acc = (1 << 21) + (img[:, lo:lo + n, :].astype(np.int64) * kh[x, :n][None, :, None]).sum(1)
        tmp[:, x] = np.clip(acc >> 22, 0, 255)
    kv, bv = pil_coeffs(H, oh, filt)
    out = np.zeros((oh, ow, C), np.uint8)
    for y in range(oh):
        lo, n = bv[y]
        acc = (1 << 21) + (tmp[lo:lo + n].astype(np.int64) * kv[y, :n][:, None, None]).sum(0)
        out[y] = np.clip(acc >> 22, 0, 255)
    return out


def pil_bilinear_resize_numpy(img: np.ndarray, oh: int, ow: int) -> np.ndarray:
    return pil_resize_numpy(img, oh, ow, "bilinear")


def cv_linear_tables(in_size: int, out_size: int, vertical: bool) -> tuple[np.ndarray, np.ndarray]:
    """OpenCV `cv::resize(..., INTER_LINEAR)` on 8-bit images (modules/imgproc/src/resize.cpp): per output index the left source index and the two
    11-bit fixed-point coefficients (INTER_RESIZE_COEF_BITS). Horizontally the index is clamped and the fraction reset at both borders;
    vertically the fraction is kept and the two ROW indices are clipped instead (so a border row is a two-term sum of the same row)."""
    scale = in_size / out_size
    idx, coef = np.zeros(out_size, np.int32), np.zeros((out_size, 2), np.int32)
    for d in range(out_size):
        f = np.float32((d + 0.5) * scale - 0.5)
        s = int(np.floor(f))
        f = np.float32(f - s)
        if not vertical:
            if s < 0:
                s, f = 0, np.float32(0)
            if s >= in_size - 1:
                s, f = in_size - 1, np.float32(0)
        idx[d] = s
        coef[d] = (int(np.rint(np.float32((np.float32(1.0) - f) * 2048))), int(np.rint(np.float32(f * 2048))))
    return idx, coef


def cv_resize_linear_numpy(img: np.ndarray, oh: int, ow: int) -> np.ndarray:
    """Host restatement of cv2.resize(img, (ow, oh)) for uint8 HWC images, bit for bit (tests/test_preprocess_cpu.py): the arithmetic of vg_resize_u8_cv."""
    H, W, _ = img.shape
    xi, xc = cv_linear_tables(W, ow, False)
    yi, yc = cv_linear_tables(H, oh, True)
    x1 = np.minimum(xi + 1, W - 1)
    rows = img[:, xi, :].astype(np.int64) * xc[:, 0][None, :, None] + img[:, x1, :].astype(np.int64) * xc[:, 1][None, :, None]
    y0, y1 = np.clip(yi, 0, H - 1), np.clip(yi + 1, 0, H - 1)
    b0, b1 = yc[:, 0][:, None, None], yc[:, 1][:, None, None]
    out = (((b0 * (rows[y0] >> 4)) >> 16) + ((b1 * (rows[y1] >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def clip_resize_shape(h: int, w: int, size: int = 336) -> tuple[int, int]:
    """HF get_resize_output_image_size for {"shortest_edge": size}: the short side becomes `size`, the long side int(size * long / short)."""
    short, long = (h, w) if h <= w else (w, h)
    new_short, new_long = size, int(size * long / short)
    return (new_short, new_long) if h <= w else (new_long, new_short)


class EncPreprocessGPU:
    """EncPreprocessor_VideoGPTPlus.preprocess (utils/enc_preprocessors.py:108-166) for a whole clip ON THE GPU from the raw
    uint8 frames [T,H,W,3]:
      IV2 frames    cv2.resize(x, (224, 224)) (INTER_LINEAR, 11-bit fixed point, bit-exact) -> (x / 255 - mean) / std in float64 -> fp32 [T,3,224,224]
      CLIP context  PIL BICUBIC resize of the short side to 336 (Pillow's two-pass fixed-point resampling, bit-exact) -> centre crop 336x336 ->
                    float32(x * (1/255)) -> (v - mean) / std in float32 (HF CLIPImageProcessor's arithmetic) -> fp32 [T,3,336,336]
    Frame trimming / padding to NUM_FRAMES stays with the caller (it is list bookkeeping)."""

    def __init__(self, device="cuda", iv2_size: int = 224, clip_size: int = 336):
        self.device, self.iv2_size, self.clip_size, self._tables = torch.device(device), iv2_size, clip_size, {}

    def tables(self, h: int, w: int):
        key = (h, w)
        if key not in self._tables:
            if len(self._tables) > 32:
                self._tables.clear()
            dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(self.device)
            xi, xc = cv_linear_tables(w, self.iv2_size, False)
            yi, yc = cv_linear_tables(h, self.iv2_size, True)
            ch, cw = clip_resize_shape(h, w, self.clip_size)
            pil = None
            if (ch, cw) != (h, w):                      # PIL returns a copy when the size does not change
                kh, bh = pil_coeffs(w, cw, "bicubic")
                kv, bv = pil_coeffs(h, ch, "bicubic")
                pil = tuple(dev(a) for a in (kh, bh, kv, bv))
            self._tables[key] = (tuple(dev(a) for a in (xi, xc, yi, yc)), (ch, cw), pil)
        return self._tables[key]

    def preprocess_video(self, frames) -> tuple[torch.Tensor, torch.Tensor]:
        """frames: uint8 [T,H,W,3] (numpy, or a torch tensor on the host / device) -> (images fp32 [T,3,224,224], context_images fp32 [T,3,336,336])."""
        from . import ops
        t = torch.from_numpy(np.ascontiguousarray(frames)) if isinstance(frames, np.ndarray) else frames
        assert t.dtype == torch.uint8 and t.dim() == 4 and t.shape[3] == 3
        t = t.to(self.device, non_blocking=True).contiguous()
        cv_tabs, (ch, cw), pil = self.tables(t.shape[1], t.shape[2])
        small = ops.resize_u8_cv(t, self.iv2_size, self.iv2_size, cv_tabs)
        images = ops.u8_normalize_chw(small, self.iv2_size, self.iv2_size, 0, 0, tuple(IMAGENET_MEAN.flatten().tolist()), tuple(IMAGENET_STD.flatten().tolist()), mode=1)
        r = t if pil is None else ops.resize_u8_pil(t, ch, cw, pil)
        S = self.clip_size
        context = ops.u8_normalize_chw(r, S, S, (ch - S) // 2, (cw - S) // 2, tuple(CLIP_MEAN), tuple(CLIP_STD), mode=0)
        return images, context


class SAM_v2_PreprocessGPU:
    """sam_preprocess(model_type="sam2") for a whole clip ON THE GPU: the raw uint8 frames [T,H,W,3] are uploaded (H*W*3 bytes per
    frame instead of 12.6 MB of fp32) and resized (PIL-bilinear, bit-exact), normalised and squashed to 1024x1024 by two libvglamm kernels.
    Coefficient tables are cached per frame size."""

    def __init__(self, device="cuda", img_size: int = 1024):
        self.device, self.img_size, self._tables = torch.device(device), img_size, {}

    def tables(self, h: int, w: int):
        key = (h, w)
        if key not in self._tables:
            if len(self._tables) > 32:
                self._tables.clear()
            newh, neww = get_preprocess_shape(h, w, self.img_size)
            kh, bh = pil_bilinear_coeffs(w, neww)
            kv, bv = pil_bilinear_coeffs(h, newh)
            self._tables[key] = ((newh, neww), tuple(torch.from_numpy(np.ascontiguousarray(a)).to(self.device) for a in (kh, bh, kv, bv)))
        return self._tables[key]

    def preprocess_video(self, frames) -> tuple[torch.Tensor, tuple]:
        """frames: uint8 [T,H,W,3] (numpy, or a torch tensor on the host / device) -> (fp32 [T,3,1024,1024] on the device, resize_shape)."""
        from . import ops
        t = torch.from_numpy(np.ascontiguousarray(frames)) if isinstance(frames, np.ndarray) else frames
        assert t.dtype == torch.uint8 and t.dim() == 4 and t.shape[3] == 3
        t = t.to(self.device, non_blocking=True).contiguous()
        (newh, neww), tabs = self.tables(t.shape[1], t.shape[2])
        r = ops.resize_u8_pil(t, newh, neww, tabs)
        return ops.sam_normalize_squash(r, self.img_size, tuple(SAM_MEAN.flatten().tolist()), tuple(SAM_STD.flatten().tolist())), (newh, neww)

    def preprocess(self, x: np.ndarray):
        out, shape = self.preprocess_video(x[None])
        return out[0], shape


def tokenizer_image_token(prompt: str, tokenizer, image_token_index: int = IMAGE_TOKEN_INDEX, return_tensors=None):
    """model/videogpt_plus/mm_utils.py:17-37: tokenise the text between <image> markers, join with the -200 sentinel, keep one BOS."""
    chunks = [tokenizer(c).input_ids for c in prompt.split(DEFAULT_IMAGE_TOKEN)]
    ids, offset = [], 0
    if len(chunks) > 0 and len(chunks[0]) > 0 and chunks[0][0] == tokenizer.bos_token_id:
        offset = 1
        ids.append(chunks[0][0])
    sep = [image_token_index] * (offset + 1)
    inter = [e for pair in zip(chunks, [sep] * len(chunks)) for e in pair][:-1]
    for x in inter:
        ids.extend(x[offset:])
    if return_tensors == "pt":
        return torch.tensor(ids, dtype=torch.long)
    if return_tensors is not None:
        raise ValueError(f"Unsupported tensor type: {return_tensors}")
    return ids


PHI3_SYSTEM, PHI3_ROLES, PHI3_SEP = "<|system|>\nYou are a helpful AI assistant.", ("\n<|user|>\n", "\n<|assistant|>\n"), "<|end|>"


def phi3_prompt(messages: list[tuple[str, str]]) -> str:
    """conv_phi3_instruct with SeparatorStyle.MPT (model/videogpt_plus/conversation.py:60-68, 124-132)."""
    ret = PHI3_SYSTEM + PHI3_SEP
    for role, message in messages:
        ret += role + message + PHI3_SEP if message else role
    return ret


class ConvGenerator_VideoGPTPlus:
    """utils/conv_generator.py:86-131 (chat path, mm_use_im_start_end False)."""

    def __init__(self, num_frames: int = 16):
        self.NUM_FRAMES = num_frames

    def apply_for_chat(self, prompt_text: str, type: str = "video", tokenizer=None, device="cuda"):
        if type == "video":
            prompt = (DEFAULT_VIDEO_TOKEN + "\n" + prompt_text).replace(DEFAULT_VIDEO_TOKEN, DEFAULT_IMAGE_TOKEN * self.NUM_FRAMES)
        elif type == "image":
            prompt = DEFAULT_IMAGE_TOKEN + "\n" + prompt_text
        else:
            raise ValueError(type)
        prompt = phi3_prompt([(PHI3_ROLES[0], prompt), (PHI3_ROLES[1], "")])
        return tokenizer_image_token(prompt, tokenizer, return_tensors="pt").unsqueeze(0).to(device)


def preprocess_vision(np_images, type="video", enc_preprocessor=None, sam_preprocessor=None, conv_generator=None, precision="fp32", device="cuda"):
    """chat.py:402-456 (video): sub-sample encoder frames with np.linspace, encode ALL frames for SAM2.
    Returns (enc_image, enc_context_image, image_sam, original_size_list, resize_list), each a list of length 1."""
    from PIL import Image
    if type != "video":
        raise NotImplementedError("still-image inputs are not on the video path")
    assert len(np_images) == 1, "Batch size must be 1"
    frames = np_images[0]
    n = conv_generator.NUM_FRAMES
    enc_frames = [frames[i] for i in np.linspace(0, len(frames) - 1, n, dtype=int)] if len(frames) > n else list(frames)
    d = enc_preprocessor.preprocess([Image.fromarray(f) for f in enc_frames])
    dt = {"bf16": torch.bfloat16, "fp16": torch.float16}.get(precision, torch.float32)
    enc_image = [torch.stack(d["images"], 0).to(dt).to(device, non_blocking=True)]
    enc_ctx = [torch.stack(d["context_images"], 0).to(dt).to(device, non_blocking=True)]
    pre = [sam_preprocessor.preprocess(f) for f in frames]
    image_sam = [torch.stack([p[0] for p in pre], 0).to(dt).to(device, non_blocking=True)]
    return enc_image, enc_ctx, image_sam, [frames[0].shape[:2]], [pre[0][1]]
