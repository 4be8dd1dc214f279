"""Multi-GPU plumbing for the path: independent (clip, query) pairs shard across ranks with no data-path exchange;
the only collective is a gather of bit-packed final masks (+ token ids) to rank 0. One process per GPU, torch.distributed (NCCL on
GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_clips(n_clips: int, rank: int, world: int) -> list[int]:
    """Clip i -> rank i mod world (round-robin keeps data-dependent decode lengths balanced on average)."""
    return list(range(rank, n_clips, world))


def pack_masks(segments: dict) -> tuple[torch.Tensor, tuple]:
    """{frame: {obj: bool [H,W]}} -> (uint8 bit-packed 1-D tensor, (T, n_obj, H, W))."""
    frames = sorted(segments)
    if not frames:
        return torch.zeros(0, dtype=torch.uint8), (0, 0, 0, 0)
    objs = sorted(segments[frames[0]])
    m = np.stack([np.stack([segments[t][o] for o in objs]) for t in frames])
    return torch.from_numpy(np.packbits(m.reshape(-1))), tuple(m.shape)


def unpack_masks(packed: torch.Tensor, shape: tuple) -> dict:
    T, n_obj, H, W = shape
    if T == 0:
        return {}
    bits = np.unpackbits(packed.cpu().numpy())[: T * n_obj * H * W].reshape(T, n_obj, H, W).astype(np.bool_)
    return {t: {o: bits[t, o] for o in range(n_obj)} for t in range(T)}


def gather_results(local: list[tuple[int, torch.Tensor, dict]], device: torch.device | str = "cpu", dst: int = 0):
    """local: [(clip_index, output_ids int64 [1,L], segments)] produced by this rank. Returns on rank `dst` the merged
    {clip_index: (output_ids, segments)}; None elsewhere. Two collectives: an all_gather of the (variable) payload sizes, then a padded
    gather of the bit-packed payload — messages are O(MB), latency-bound."""
    world, rank = dist.get_world_size(), dist.get_rank()
    metas, payloads = [], []
    for idx, ids, segs in local:
        packed, shape = pack_masks(segs)
        ids_b = ids.reshape(-1).to(torch.int64).cpu()
        metas.append((idx, ids_b.numel(), packed.numel(), *shape))
        payloads += [ids_b.view(torch.uint8), packed]
    payload = torch.cat(payloads) if payloads else torch.zeros(0, dtype=torch.uint8)
    all_meta: list = [None] * world
    dist.all_gather_object(all_meta, (metas, payload.numel()))
    max_len = max(m[1] for m in all_meta)
    buf = torch.zeros(max(max_len, 1), dtype=torch.uint8, device=device)
    buf[: payload.numel()] = payload.to(device)
    outs = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
    dist.gather(buf, outs, dst=dst)
    if rank != dst:
        return None
    merged = {}
    for r in range(world):
        data, off = outs[r].cpu(), 0
        for idx, n_ids, n_packed, T, n_obj, H, W in all_meta[r][0]:
            ids = data[off: off + 8 * n_ids].clone().view(torch.int64)[None]
            off += 8 * n_ids
            merged[idx] = (ids, unpack_masks(data[off: off + n_packed], (T, n_obj, H, W)))
            off += n_packed
    return merged
