"""world_size-2 gloo test of the N>1 path's host logic: clip sharding and the gather of bit-packed masks to rank 0."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from videoglamm_b200 import parallel as P


def _fake_result(idx):
    rng = np.random.default_rng(idx)
    T, n_obj = 3 + idx % 2, 1 + idx % 3
    segs = {t: {o: rng.random((37, 53)) > 0.5 for o in range(n_obj)} for t in range(T)} if idx != 4 else {}
    ids = torch.arange(10 + idx, dtype=torch.int64)[None] * (idx + 1) - 200
    return ids, segs


def _worker(rank, world, port, n_clips, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = P.shard_clips(n_clips, rank, world)
    local = [(i, *_fake_result(i)) for i in mine]
    merged = P.gather_results(local, max_local=(n_clips + world - 1) // world)   # ONE header all_gather + ONE payload gather per batch
    if rank == 0:
        ok = sorted(merged) == list(range(n_clips))
        for i in range(n_clips):
            ids, segs = _fake_result(i)
            ok &= torch.equal(merged[i][0], ids) and sorted(merged[i][1]) == sorted(segs)
            for t in segs:
                for o in segs[t]:
                    ok &= bool(np.array_equal(merged[i][1][t][o], segs[t][o]))
        q.put(ok)
    else:
        assert merged is None
    dist.destroy_process_group()


def test_shard_and_gather_world2():
    assert P.shard_clips(7, 0, 2) == [0, 2, 4, 6] and P.shard_clips(7, 1, 2) == [1, 3, 5]
    assert sorted(sum((P.shard_clips(32, r, 8) for r in range(8)), [])) == list(range(32))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, 29611, 7, q)) for r in range(2)]
    [p.start() for p in procs]
    [p.join(120) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert q.get(timeout=5) is True


def test_pack_roundtrip_ragged_sizes():
    for H, W in ((1, 1), (7, 9), (336, 336), (480, 854)):
        segs = {t: {0: np.random.default_rng(t).random((H, W)) > 0.3} for t in range(2)}
        packed, shape = P.pack_masks(segs)
        assert packed.numel() == 2 * ((H * W + 7) // 8)               # every plane padded to whole bytes (the device kernel's format)
        back = P.unpack_masks(packed, shape)
        assert all(np.array_equal(back[t][0], segs[t][0]) for t in segs)
    assert P.unpack_masks(*P.pack_masks({})) == {}
