"""Multi-GPU plumbing for the path: independent (clip, query) pairs shard across ranks with no data-path exchange;
the only collective step is ONE gather per batch of bit-packed final masks (+ token ids) to rank 0. One process per GPU,
torch.distributed (NCCL on GPUs, gloo in the CPU tests).

Wire format of a mask stack: uint8 [T, n_obj, ceil(H*W/8)] — every [H,W] plane thresholded (> 0) and bit-packed MSB first
(numpy.packbits order), padded to whole bytes. On the GPU it is produced directly by the final upsampling kernel (vg_bilinear mode 4),
so masks never exist as one byte (or one fp32) per pixel outside the kernel."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch
import torch.distributed as dist

HDR = 8   # int64 fields per clip in the fixed-size header: clip index, #ids, #packed bytes, T, n_obj, H, W, valid flag


@dataclass
class PackedMasks:
    """Bit-packed masks of one clip: `bits` uint8 [T, n_obj, ceil(H*W/8)] (any device), shape = (T, n_obj, H, W)."""
    bits: torch.Tensor
    shape: tuple

    def to_segments(self) -> dict:
        return unpack_masks(self.bits, self.shape)


def shard_clips(n_clips: int, rank: int, world: int) -> list[int]:
    """Clip i -> rank i mod world (round-robin keeps data-dependent decode lengths balanced on average)."""
    return list(range(rank, n_clips, world))


def pack_masks(segments: dict) -> tuple[torch.Tensor, tuple]:
    """{frame: {obj: bool [H,W]}} -> (uint8 [T, n_obj, ceil(H*W/8)] tensor, (T, n_obj, H, W)) on the host (CPU callers / tests; the engine
    packs on the device)."""
    frames = sorted(segments)
    if not frames:
        return torch.zeros((0, 0, 0), dtype=torch.uint8), (0, 0, 0, 0)
    objs = sorted(segments[frames[0]])
    m = np.stack([np.stack([segments[t][o] for o in objs]) for t in frames])
    T, n_obj, H, W = m.shape
    return torch.from_numpy(np.packbits(m.reshape(T, n_obj, H * W), axis=-1)), (T, n_obj, H, W)


def unpack_masks(packed: torch.Tensor, shape: tuple) -> dict:
    T, n_obj, H, W = shape
    if T == 0:
        return {}
    bits = np.unpackbits(packed.cpu().numpy().reshape(T, n_obj, -1), axis=-1, count=H * W).reshape(T, n_obj, H, W).view(np.bool_)
    return {t: {o: bits[t, o] for o in range(n_obj)} for t in range(T)}


def gather_results(local: list[tuple], device: torch.device | str = "cpu", dst: int = 0, max_local: int | None = None, unpack: bool = True):
    """local: [(clip_index, output_ids int64 [1,L], masks)] produced by this rank for ONE batch; `masks` is a PackedMasks (device-packed
    bits, the engine's form) or a {frame: {obj: bool [H,W]}} dict. Returns on rank `dst` {clip_index: (output_ids, segments | PackedMasks)}
    (`unpack=False` keeps the bit-packed form); None elsewhere.
    Exactly two tensor collectives per BATCH, no pickling: an all_gather of a fixed-size int64 header ([max_local, 8] per rank) and a padded
    gather of the byte payload (ids + packed bits, O(MB): latency-bound). `max_local` = the largest number of clips any rank
    holds (default: len(local), i.e. equal shares)."""
    world, rank = dist.get_world_size(), dist.get_rank()
    K = max_local if max_local is not None else len(local)
    assert len(local) <= K
    hdr = torch.zeros((K, HDR), dtype=torch.int64)
    parts = []
    for j, (idx, ids, masks) in enumerate(local):
        if not isinstance(masks, PackedMasks):
            masks = PackedMasks(*pack_masks(masks))
        ids_b = ids.reshape(-1).to(torch.int64)
        bits = masks.bits.reshape(-1)
        hdr[j] = torch.tensor([idx, ids_b.numel(), bits.numel(), *masks.shape, 1])
        parts += [ids_b.to(device).view(torch.uint8), bits.to(device)]
    payload = torch.cat(parts) if parts else torch.zeros(0, dtype=torch.uint8, device=device)
    hdr = hdr.to(device)
    all_hdr = [torch.empty_like(hdr) for _ in range(world)]
    dist.all_gather(all_hdr, hdr)
    all_hdr = [h.cpu() for h in all_hdr]
    max_len = max(int((h[:, 1] * 8 + h[:, 2]).sum()) for h in all_hdr)
    buf = torch.zeros(max(max_len, 1), dtype=torch.uint8, device=device)
    buf[: payload.numel()] = payload
    outs = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
    dist.gather(buf, outs, dst=dst)
    if rank != dst:
        return None
    merged = {}
    for r in range(world):
        data, off = outs[r].cpu(), 0
        for idx, n_ids, n_packed, T, n_obj, H, W, valid in all_hdr[r].tolist():
            if not valid:
                continue
            ids = data[off: off + 8 * n_ids].clone().view(torch.int64)[None]
            off += 8 * n_ids
            pm = PackedMasks(data[off: off + n_packed].reshape(T, n_obj, -1) if T else data[off: off], (T, n_obj, H, W))
            merged[idx] = (ids, pm.to_segments() if unpack else pm)
            off += n_packed
    return merged
