#!/usr/bin/env python
"""Headline benchmark: grounded clips/sec end to end (BASELINE.json config 2: 1 clip = T=8 frames, CLIP 336 px / IV2 224 px /
SAM2 1024 px, Phi-3-mini, 1 [SEG] object, 12 new tokens), full-size architecture with synthetic weights and inputs.

  python bench.py [--gpus N] [--steps K] [--warmup W]            # our engine (one process per GPU under torchrun for N>1)
  python bench.py --impl reference [...]                          # CPU arm: the oracle port of the reference on host cores

One "step" = one clip through the whole path (towers -> projectors -> Phi-3 prefill + KV-cached decode -> [SEG] projection ->
SAM2 Hiera + mask decoder + memory attention/encoder over 8 frames -> masks at video resolution).
Prints ONE JSON line (contract in the task statement / DESIGN.md §Measurement).
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

L_QUERY = 20
# BASELINE.json configs[1..4] (SURVEY.md §8 "Config shorthand" and §8d for the forced [SEG] counts / new-token counts). One "step" = `clips`
# independent clips per GPU through the whole path.
CONFIGS = {
    2: dict(tag="C2", workload="BASELINE.json configs[1]: 1 clip T_enc=T_sam=8, CLIP 336px, IV2 224px, SAM2 1024px, Phi-3-mini 32L, 1 [SEG], 12 new tokens",
            t_enc=8, t_sam=8, n_new=12, seg_steps=(5,), video_hw=(336, 336), clips=1),
    3: dict(tag="C3", workload="BASELINE.json configs[2]: GCG multi-object, 4 clips per GPU (32 clips on 8 GPUs), each T=8 336px, 4 [SEG] objects, 80 new tokens",
            t_enc=8, t_sam=8, n_new=80, seg_steps=(10, 30, 50, 70), video_hw=(336, 336), clips=4),
    4: dict(tag="C4", workload="BASELINE.json configs[3]: Refer-DAVIS shape, 1 clip T_enc=T_sam=16, original 480x854, 1 object, 12 new tokens, L=3360 (sliding window cuts)",
            t_enc=16, t_sam=16, n_new=12, seg_steps=(5,), video_hw=(480, 854), clips=1),
    5: dict(tag="C5", workload="BASELINE.json configs[4]: MeViS shape, 1 clip per GPU, T_sam=32 (T_enc=16), original 720x1280, 4 objects, 20 new tokens, memory bank full",
            t_enc=16, t_sam=32, n_new=20, seg_steps=(4, 8, 12, 16), video_hw=(720, 1280), clips=1),
}
CFG = CONFIGS[2]


def set_config(n: int):
    global CFG, T_ENC, T_SAM, N_NEW, VIDEO_HW
    CFG = CONFIGS[n]
    T_ENC, T_SAM, N_NEW, VIDEO_HW = CFG["t_enc"], CFG["t_sam"], CFG["n_new"], CFG["video_hw"]


set_config(2)


METRIC = {2: "grounded clips/sec (T=8, 336px) end-to-end", 3: "grounded clips/sec (T=8, 336px, 4 objects, batch of 4 clips per GPU) end-to-end",
          4: "grounded clips/sec (T=16, 480p) end-to-end", 5: "grounded clips/sec (T=32, 720p, 4 objects) end-to-end"}


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1383.6), d.get("hbm_gbs", 6458.4), "measured"
    return 1400.0, 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def mark(self):
        """Rows before this point (nvidia-smi start-up, idle GPU) are not part of the timed region."""
        self.rows = []

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i] == "Active" for r in self.rows)]
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None, "reasons": reasons, "samples": len(sm)}


def clip_flops(L: int) -> float:
    """Algorithmic FLOPs of one clip of the selected config for the ideal engine: encode once, KV-cached decode."""
    n_obj = len(CFG["seg_steps"])
    enc = 0.382e12 * T_ENC + 2.25e12 * T_ENC / 4 + 0.022e12 * T_ENC
    prefill = 2 * 3.62e9 * L + 2 * L * min(L, 2048) * 98304
    decode = N_NEW * (7.6e9 + 4 * min(L, 2048) * 98304)
    sam = 1.620e12 * T_SAM
    for t in range(T_SAM):
        nk = 4096 * min(t, 7) + 4 * min(t, 16)
        sam += n_obj * (3.64e9 + 11.6e9 + (t > 0) * (4 * (2.1e9 + 17.2e9 + 1.07e9 + 8.6e9) + 4 * (65536 + 4194304) * nk))
    return enc + prefill + decode + sam


def synth_inputs(seed: int, pinned: bool):
    """seeded N(0,1) tensors of the named shapes + a prompt with T_ENC image sentinels."""
    g = torch.Generator().manual_seed(5678 + seed)
    mk = lambda *s: torch.randn(*s, generator=g)
    images, context, sam = mk(T_ENC, 3, 224, 224), mk(T_ENC, 3, 336, 336), mk(T_SAM, 3, 1024, 1024)
    ids = torch.cat([torch.randint(3, 32000, (12,), generator=g), torch.full((T_ENC,), -200), torch.randint(3, 32000, (L_QUERY,), generator=g)])[None]
    if pinned:
        images, context, sam, ids = (t.pin_memory() for t in (images, context, sam, ids))
    return images, context, sam, ids


def synth_raw_frames(seed: int):
    """The raw uint8 video frames [T_sam, H, W, 3] at the original resolution (uniform [0,255]) for the end-to-end arm: the engine
    pre-processes them on the GPU (PIL-exact resize, normalise, squash), so H*W*3 bytes per frame cross PCIe instead of 12.6 MB of fp32."""
    g = torch.Generator().manual_seed(91234 + seed)
    return torch.randint(0, 256, (T_SAM, VIDEO_HW[0], VIDEO_HW[1], 3), generator=g, dtype=torch.uint8).pin_memory()


def run_engine(args):
    import torch.distributed as dist
    from videoglamm_b200 import build, ops, parallel as P
    from videoglamm_b200.config import Sam2Config, make_lmm_config
    from videoglamm_b200.model import VideoGLaMMForCausalLM
    from videoglamm_b200.statedict import random_state_dict, videoglamm_param_shapes

    build.build()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL prints its version banner on stdout at communicator creation; the contract is ONE JSON line on stdout -> point fd 1 at stderr meanwhile
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    dev = f"cuda:{local}"
    lmm_cfg, sam_cfg = make_lmm_config(), Sam2Config()
    sd = random_state_dict(videoglamm_param_shapes(lmm_cfg, sam_cfg), dev, seed=1234)
    model = VideoGLaMMForCausalLM(sd, lmm_cfg, sam_cfg, device=dev, validate=False)
    del sd
    torch.cuda.empty_cache()
    forced = {st: lmm_cfg.seg_token_idx for st in CFG["seg_steps"]}
    n_obj, CPS = len(forced), CFG["clips"]
    host_clips = [synth_inputs(rank * CPS + j, pinned=True) for j in range(CPS)]       # CPS different clips per GPU and step
    res_clips = [[t.to(dev) for t in h] for h in host_clips]
    # e2e arm: ONLY the raw uint8 frames (+ prompt ids) cross PCIe: the SAM2 frames and the T_ENC encoder frames sampled from them; the IV2 / CLIP /
    # SAM2 pre-processing runs on the GPU (preprocess.EncPreprocessGPU, SAM_v2_PreprocessGPU: the reference's cv2 / PIL / HF arithmetic)
    raw_clips = []
    for j, h in enumerate(host_clips):
        raw_sam = synth_raw_frames(rank * CPS + j)
        raw_enc = raw_sam[:: max(T_SAM // T_ENC, 1)][:T_ENC].contiguous().pin_memory()
        raw_clips.append([raw_enc, raw_enc, raw_sam, h[3]])
    host, resident = host_clips[0], res_clips[0]
    L = host[3].shape[1] - T_ENC + 208 * T_ENC

    def step(inp):
        im, ctx, sam, ids = inp
        out_ids, segs = model.inference([im], [ctx], [sam], ids, None, [VIDEO_HW], max_new_tokens=N_NEW, use_sam2_video_branch=True, forced_tokens=forced)
        return out_ids, segs

    def gather_batch(results):
        """§8(e): the only collective step on the path — ONE gather per batch of the token ids + the device-packed mask bits to rank 0
        (videoglamm_b200.parallel: a fixed-size header all_gather + a padded payload gather; nothing is pickled, masks stay bit-packed)."""
        if world > 1:
            P.gather_results([(rank * len(results) + j, ids, pm) for j, (ids, pm) in enumerate(results)], device=dev, unpack=False)

    from videoglamm_b200.pipeline import ClipPipeline
    # the HBM-bound decode loops of up to DB = 8 queued clips run as ONE batched loop (one weight pass per token for all of them): two C3 steps of 4
    # clips; for the 1-clip-per-step configs, DB consecutive clips of the stream (throughput serving; --decode-batch 1 turns it off). Measured
    # DB = 4 -> 8: C2 13.0 -> 13.1, C3 6.2 -> 7.2, C5 2.61 -> 2.66 clips/s
    db_req = int(args.decode_batch) if args.decode_batch else (args.steps * CPS + 1) // 2     # auto: at least two groups, so that LMM and SAM2 stages overlap
    DB = max(CPS, min(db_req, 8))
    pipe = ClipPipeline(model, decode_batch=DB)
    pipe1 = ClipPipeline(model, decode_batch=CPS) if DB != CPS else None

    def run_clips(inps, steps, pipe=pipe):
        """`steps` batches of CPS independent clips per GPU. Sequential mode calls model.inference clip by clip; pipelined mode (default) overlaps
        the LMM stage of clip i+1 with the SAM2 stage of clip i on two streams — same public engine, same results; every clip completes (masks on
        the host as bool arrays) and the batch is gathered to rank 0 inside the region."""
        if args.sequential:
            res = []
            for k in range(steps * CPS):
                out_ids, segs = step(inps[k % CPS])
                res.append((out_ids, model.last_packed))
            gather_batch(res)
            return
        clips = []
        for k in range(steps * CPS):
            im, ctx, sam, ids = inps[k % CPS]
            clips.append(dict(images=im, context_images=ctx, images_for_sam=sam, input_ids=ids, original_size=VIDEO_HW))
        gather_batch([(out_ids, pm) for out_ids, _, pm in pipe.run(clips, N_NEW, forced_tokens=forced)])

    def timed(inp, steps, pipe=pipe):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run_clips(inp, steps, pipe)
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    for _ in range(max(args.warmup, 3)):
        step(resident)
    # pipeline warm-up (per-thread stream state, graph pools): two full decode groups + the remainder group the timed run will end with, so that
    # every graph key of the timed region (prefill per cache slot, batched decode per group size) has been seen and captured
    warm_steps = -(-(2 * DB + (args.steps * CPS) % DB) // CPS)
    run_clips(res_clips, warm_steps)
    run_clips(host_clips, warm_steps)            # ... and of the host-input paths (side-stream allocator pools, pinned result buffers)
    run_clips(raw_clips, warm_steps)
    if pipe1 is not None:
        run_clips(res_clips, 2, pipe1)
    torch.cuda.synchronize()
    # serving hygiene: the engine's long-lived Python objects (weights, graphs, tensor maps) move to the permanent generation, so that a
    # generation-2 garbage collection cannot stall both stage threads for 0.1-0.3 s in the middle of a timed region (seen in the stage traces)
    gc.collect()
    gc.freeze()
    lat0, lat1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    lat0.record()
    step(resident)
    lat1.record()
    torch.cuda.synchronize()
    latency_ms = lat0.elapsed_time(lat1)          # one clip alone, nothing overlapped
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        t_wait = time.time()                     # nvidia-smi start-up (fork + NVML init across every GPU of the box stalls our launches for
        while not sampler.rows and time.time() - t_wait < 8.0:   # 0.1-0.3 s) stays out of the timed region: wait for its first sample
            time.sleep(0.05)
        sampler.mark()
    ops.launch_count_reset()
    def graph_stats():
        return {k: dict(v.stats) for eng in (model.lmm, model.sam) for k, v in vars(eng).items() if hasattr(v, "stats") and hasattr(v, "entries")}
    def dump_trace(name):
        """diagnostics (VG_PIPE_TRACE=1): start -> end of every stage call of a timed region, ms from the first mark"""
        if not pipe.trace:
            return
        t_first, marks = pipe.trace[0][2], {}
        for tag, i, ev in pipe.trace:
            marks[(tag, i)] = t_first.elapsed_time(ev)
        row = [f"{tag[:-1]}{i}:{t:.0f}-{marks[(tag[:-1] + '1', i)]:.0f}" for (tag, i), t in sorted(marks.items(), key=lambda kv: kv[1]) if tag.endswith("0")]
        print(f"trace {name}", " ".join(row), graph_stats(), file=sys.stderr)
        pipe.trace = []
    if os.environ.get("VG_PIPE_TRACE"):
        pipe.trace = []
    ms = timed(res_clips, args.steps)
    launches = ops.launch_count()
    dump_trace("resident")
    # same call, HOST pinned inputs incl. the RAW uint8 SAM2 frames (GPU pre-processing): H2D + D2H inside the region. Two passes of K steps, the
    # faster one is reported and both are listed (`e2e.passes_ms`): on shared hosts one pass in ~10 stalls for 0.3-0.5 s on the CPU side
    ms_e2e_passes = [timed(raw_clips, args.steps)]
    dump_trace("e2e")
    ms_e2e_passes.append(timed(raw_clips, args.steps))
    ms_e2e = min(ms_e2e_passes)
    pipe.trace = None
    ms_e2e_fp32 = timed(host_clips, args.steps)  # ... and with the CPU-preprocessed fp32 SAM2 tensors the reference uploads (100 MB per 8-frame clip)
    ms_db1 = timed(res_clips, args.steps, pipe1) if pipe1 is not None else None   # the same stream of clips, every clip decoded alone
    clocks = sampler.stop() if rank == 0 else {}
    # dominant kernel = the tcgen05 GEMM: per-launch CUDA-event timing on one extra (untimed) step
    from videoglamm_b200 import graphs
    graphs.ENABLED = False                       # per-launch events need eager launches (graph replays bypass the Python wrappers)
    ops.gemm_profile(True)
    step(resident)
    torch.cuda.synchronize()
    prof = ops.gemm_profile(False)
    graphs.ENABLED = True
    if rank == 0 and os.environ.get("VG_GEMM_SHAPES"):
        for row in prof["top_shapes"]:
            print("gemm shape", row, file=sys.stderr)
    dec = time_decode_step(model, L) if rank == 0 else None
    mdec = time_mask_decoder(model) if rank == 0 else None
    cpu = cpu_baseline_sample(L) if (rank == 0 and not args.no_cpu_baseline and world == 1) else None
    if rank == 0:
        tf_peak, hbm_peak, src = _peaks()
        n_clips = args.steps * world * CPS
        tp, tall = prof["tensor_big"], prof["tensor"]
        ach = tp["flops"] / (tp["ms"] * 1e-3) / 1e12 if tp["ms"] > 0 else 0.0
        ach_all = tall["flops"] / (tall["ms"] * 1e-3) / 1e12 if tall["ms"] > 0 else 0.0
        h2d = CPS * sum(t.numel() * t.element_size() for t in raw_clips[0][1:])   # encoder frames (uploaded once for both towers) + SAM2 frames + ids
        h2d_fp32 = CPS * sum(t.numel() * t.element_size() for t in host)
        d2h = CPS * (T_SAM * n_obj * ((VIDEO_HW[0] * VIDEO_HW[1] + 7) // 8) + 8)     # bit-packed masks + the per-token id reads of the decode loop
        ms_clip = ms / (args.steps * CPS)
        line = {
            "metric": METRIC[args.config], "value": n_clips / (ms * 1e-3), "unit": "clips/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": CFG["workload"], "precision": "bf16 operands / fp32 accumulation; SAM2 mask decoder in fp32-grade split-bf16 arithmetic (precise=heads)",
                       "llm_seq_len": L, "clips_per_step_per_gpu": CPS, "objects_per_clip": n_obj,
                       "parallelism": f"dp{world} (independent clips, bit-packed masks gathered to rank 0 once per batch)",
                       "l2": "no explicit flush: each step streams 10.8 GB of bf16 weights (>> 126 MB L2)",
                       "algorithmic_tflop_per_clip": clip_flops(L) / 1e12,
                       "schedule": "sequential clips" if args.sequential else
                                   f"2-stage pipeline over clips (LMM of clips i+1.. || SAM2 of clip i, two streams); decode loops of {DB} queued clips batched",
                       "decode_batch": DB,
                       **({"value_decode_batch_1": n_clips / (ms_db1 * 1e-3)} if ms_db1 else {}),
                       "latency_ms_single_clip": latency_ms},
            "e2e": {"value": n_clips / (ms_e2e * 1e-3), "unit": "clips/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "passes_ms": ms_e2e_passes,
                    "inputs": "pinned host tensors: the RAW uint8 video frames [T,H,W,3] (SAM2 frames + the encoder frames sampled from them) and the prompt ids; IV2 "
                              "(cv2-bilinear 224 px), CLIP (PIL-bicubic + centre crop 336 px) and SAM2 (PIL-bilinear + squash 1024 px) pre-processing on the GPU, "
                              "bit-exact resampling; outputs: bit-packed masks expanded to bool arrays on the host",
                    "with_fp32_sam_inputs": {"value": n_clips / (ms_e2e_fp32 * 1e-3), "h2d_bytes_per_step": h2d_fp32,
                                             "note": "the reference's input format: all three inputs pre-processed on the CPU, uploaded as fp32 (SAM2 [T,3,1024,1024])"}},
            "gpu_launches": launches,
            "roofline": {"bound": "tensor", "kernel": "gemm_tcgen05_kernel (every tensor-core GEMM launch of one clip with >= 1 GFLOP: "
                         f"{tp['n']} launches carrying {tp['flops'] / max(tall['flops'], 1):.1%} of the GEMM FLOPs)", "achieved": ach, "peak": tf_peak,
                         "unit": "TFLOP/s", "frac": ach / tf_peak, "traffic": ncu_traffic()["dram_bytes"], "traffic_detail": ncu_traffic(), "peak_source": f"{src} bf16_tflops_sustained",
                         "launches": tp["n"], "gemm_ms_per_clip": tp["ms"], "gemm_share_of_step": tp["ms"] / ms_clip,
                         "all_tensor_gemm_launches": {"launches": tall["n"], "ms_per_clip": tall["ms"], "achieved": ach_all, "frac": ach_all / tf_peak,
                                                      "note": "includes ~700 sub-GFLOP launches whose event-timed duration is launch latency"},
                         "note": "algorithmic 2MNK per launch / per-launch CUDA events on an extra eager (non-graph) step; traffic = ncu dram bytes of the "
                                 "largest-share shape (profiles/r01_ncu_targets_*.md)",
                         "whole_clip_tflops": clip_flops(L) / (ms_clip * 1e-3) / 1e12},
            "roofline_decode": {"bound": "hbm", "kernel": "one KV-cached Phi-3 decode step (graph replay: 4 weight-streaming GEMVs + RoPE/append + "
                                "single-query attention per layer, lm_head, argmax)", "achieved": dec["gbs"], "peak": hbm_peak, "unit": "GB/s",
                                "frac": dec["gbs"] / hbm_peak, "bytes_per_step": dec["bytes"], "ms_per_step": dec["ms"], "peak_source": f"{src} hbm_gbs"},
            "mask_decoder": {"kernel_group": "vg_mask_decoder: PromptEncoder dense + TwoWayTransformer + upscaling + hypernetwork MLPs + IoU/object heads, fp32 "
                             "(graph replay, B=1 object)",
                             "ms_per_frame": mdec["ms"], "launches": mdec["launches"], "algorithmic_bytes": mdec["bytes"], "achieved_gbs": mdec["gbs"],
                             "hbm_frac": mdec["gbs"] / hbm_peak, "algorithmic_gflop": mdec["gflop"], "achieved_tflops": mdec["tflops"],
                             "tensor_frac": mdec["tflops"] / tf_peak},
            "clocks": clocks,
        }
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def ncu_traffic() -> dict:
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the largest-share tensor GEMM shape (Hiera stage-3 fc2, 32768 x 576 x 2304,
    in-place fp32 residual), READ from the committed summary of the `ncu --set full` capture (profiles/r02_ncu_targets.csv, written by
    scripts/ncu_summarize.py from the .ncu-rep that scripts/gpu_ncu_targets.py produced; launch id 1). Algorithmic bytes of that launch:
    151.0 (A) + 2.7 (W) + 75.5 (C read) + 75.5 (C write) = 304.7 MB."""
    import csv
    for name in ("r02_ncu_targets.csv",):
        path = os.path.join(ROOT, "profiles", name)
        if os.path.exists(path):
            for row in csv.DictReader(open(path)):
                if row["id"] == "1" and "gemm_tcgen05_kernel" in row["kernel"]:
                    rd, wr = float(row["dram_read_MB"]), float(row["dram_write_MB"])
                    return {"kernel": row["kernel"] + " 32768x576x2304 +resid", "dram_bytes": (rd + wr) * 1e6, "dram_read_bytes": rd * 1e6, "dram_write_bytes": wr * 1e6,
                            "algorithmic_bytes": 304.7e6, "source": f"profiles/{name} launch id 1 ({row['us']} us under ncu, tensor pipe {row['tensor_pipe_pct']} %)"}
    return {"kernel": "gemm_tcgen05_kernel 32768x576x2304 +resid", "dram_bytes": None, "algorithmic_bytes": 304.7e6, "source": "no committed ncu summary found"}


def time_decode_step(model, L: int, reps: int = 8) -> dict:
    """HBM roofline of the decode phase: replays the captured decode-step graph and divides the bytes one step must stream (every Phi-3
    layer weight + lm_head once, K/V rows of the L cached positions) by its duration."""
    lmm = model.lmm
    ph = lmm.cfg.phi
    kv_bound = min(lmm.kv[0].shape[0], (L + N_NEW + 511) // 512 * 512)
    lmm.dec_pos.fill_(L)
    lmm.dec_nk.fill_(L + 1)
    lmm.dec_tok.fill_(5)
    step = lambda: lmm.g_decode((kv_bound,), lambda: lmm._decode_step(kv_bound), mutable_state=(lmm.dec_tok, lmm.dec_pos, lmm.dec_nk))
    step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    wbytes = ph.layers * 2 * (3 * ph.hidden * ph.hidden + ph.hidden * ph.hidden + 2 * ph.inter * ph.hidden + ph.inter * ph.hidden) + 2 * ph.vocab * ph.hidden
    kvbytes = ph.layers * 2 * (L + reps // 2) * ph.hidden * 2
    return {"ms": ms, "bytes": wbytes + kvbytes, "gbs": (wbytes + kvbytes) / (ms * 1e-3) / 1e9}


def time_mask_decoder(model, reps: int = 20) -> dict:
    """BASELINE.json metric (ii): the SAM2 mask-decoder kernel group for one (frame, 1 object), captured in a CUDA graph and replayed.
    Algorithmic traffic per SURVEY.md section 8(d): image_embed 2.10 + dense PE 2.10 + feat_s0 4.19 + feat_s1 2.10 MB + outputs 0.79 MB
    (weights 8.43 MB stay L2-resident across frames and are not counted); 3.64 GFLOP per object."""
    from videoglamm_b200 import ops
    from videoglamm_b200.graphs import GraphedStage
    sam = model.sam
    cfg = sam.cfg
    dev = sam.dev
    g = torch.Generator(device="cpu").manual_seed(1)
    HW, C = sam.HW, cfg.d_model
    pix = torch.randn(HW, C, generator=g).to(dev)
    s0 = torch.randn(16 * HW, C // 8, generator=g).to(dev)
    s1 = torch.randn(4 * HW, C // 4, generator=g).to(dev)
    sparse = torch.randn(1, 3, C, generator=g).to(dev)
    stage = GraphedStage("mask_decoder_bench")
    run = lambda: stage(("md",), lambda pix, s0, s1, sparse: sam.mask_decoder(pix, s0, s1, sparse)[0], pix=pix, s0=s0, s1=s1, sparse=sparse)
    n0 = ops.launch_count()
    run()
    run()
    torch.cuda.synchronize()
    n1 = ops.launch_count()
    run()
    launches = ops.launch_count() - n1
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    nbytes = (2.10 + 2.10 + 4.19 + 2.10 + 0.79) * 1e6
    return {"ms": ms, "launches": launches, "bytes": nbytes, "gbs": nbytes / (ms * 1e-3) / 1e9, "gflop": 3.64, "tflops": 3.64e9 / (ms * 1e-3) / 1e12}


# ---------------------------------------------------------------------------------------------------- CPU arm (oracle port)
def cpu_baseline_sample(L: int) -> dict:
    """Times the oracle (CPU fp32 port of the reference path) on a bounded, full-width SAMPLE of the config-2 clip and scales by
    layer / frame multiplicities: 1 CLIP layer x 8 frames, 1 IV2 block x 2 chunks, 1 Phi-3 layer at L, 1 Hiera block per stage,
    1 mask-decoder + 1 memory-attention layer (full bank) + memory encoder for one object-frame."""
    import math

    import torch.nn.functional as F

    from oracle import lmm as M, sam2 as O

    torch.set_num_threads(os.cpu_count())
    g = torch.Generator().manual_seed(0)
    rnd = lambda *s: torch.randn(*s, generator=g) * 0.02
    t = {}

    def clock(name, fn, mult):
        fn()
        t0 = time.perf_counter()
        fn()
        t[name] = (time.perf_counter() - t0) * mult

    with torch.no_grad():
        # CLIP layer (23 run), 8 frames x 577 tokens
        x = torch.randn(T_ENC, 577, 1024, generator=g)
        w = {k: rnd(*s) for k, s in dict(qkv=(3072, 1024), o=(1024, 1024), f1=(4096, 1024), f2=(1024, 4096)).items()}

        def vit_layer(x, w, heads, act):
            B, N, C = x.shape
            h = F.layer_norm(x, (C,))
            q, k, v = F.linear(h, w["qkv"]).view(B, N, 3, heads, C // heads).permute(2, 0, 3, 1, 4)
            x = x + F.linear(F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, N, C), w["o"])
            return x + F.linear(act(F.linear(F.layer_norm(x, (C,)), w["f1"])), w["f2"])

        clock("clip", lambda: vit_layer(x, w, 16, lambda z: z * torch.sigmoid(1.702 * z)), 23)
        x2 = torch.randn(T_ENC // 4, 1025, 1408, generator=g)
        w2 = {k: rnd(*s) for k, s in dict(qkv=(4224, 1408), o=(1408, 1408), f1=(6144, 1408), f2=(1408, 6144)).items()}
        clock("iv2", lambda: vit_layer(x2, w2, 16, F.gelu), 39)
        # Phi-3 layer at L tokens; the reference-as-written runs N_NEW+2 full forwards (use_cache=False, VideoGLaMM.py:782-799)
        pcfg = M.PhiCfg(layers=1)
        sd = {"model.layers.0.self_attn.qkv_proj.weight": rnd(9216, 3072), "model.layers.0.self_attn.o_proj.weight": rnd(3072, 3072),
              "model.layers.0.mlp.gate_up_proj.weight": rnd(16384, 3072), "model.layers.0.mlp.down_proj.weight": rnd(3072, 8192),
              "model.layers.0.input_layernorm.weight": torch.ones(3072), "model.layers.0.post_attention_layernorm.weight": torch.ones(3072),
              "model.norm.weight": torch.ones(3072)}
        xl = torch.randn(L, 3072, generator=g)
        clock("phi3_layer_full_forward", lambda: M.phi3_forward(sd, pcfg, xl), 32)
        # Hiera: one block of each stage at 1024^2, scaled by blocks per stage, x 8 frames
        scfg = O.Sam2Cfg()
        specs = scfg.block_specs()
        hier = 0.0
        for si, (bi, cnt, side) in enumerate(((1, 2, 256), (3, 6, 128), (10, 36, 64), (45, 4, 32))):
            dim, dim_out, heads, window, qs = specs[bi]
            p = "b"
            bsd = {f"{p}.norm1.weight": torch.ones(dim), f"{p}.norm1.bias": torch.zeros(dim), f"{p}.attn.qkv.weight": rnd(3 * dim_out, dim),
                   f"{p}.attn.qkv.bias": torch.zeros(3 * dim_out), f"{p}.attn.proj.weight": rnd(dim_out, dim_out), f"{p}.attn.proj.bias": torch.zeros(dim_out),
                   f"{p}.norm2.weight": torch.ones(dim_out), f"{p}.norm2.bias": torch.zeros(dim_out),
                   f"{p}.mlp.layers.0.weight": rnd(4 * dim_out, dim_out), f"{p}.mlp.layers.0.bias": torch.zeros(4 * dim_out),
                   f"{p}.mlp.layers.1.weight": rnd(dim_out, 4 * dim_out), f"{p}.mlp.layers.1.bias": torch.zeros(dim_out)}
            xh = torch.randn(1, side, side, dim, generator=g)
            clock(f"hiera_stage{si}", lambda: O.hiera_block(bsd, p, xh, specs[bi]), cnt * T_SAM)
            hier += t[f"hiera_stage{si}"]
        # SAM2 heads / memory path at the true widths and 64x64 feature maps (tiny Hiera: its weights are not used by these stages)
        n_obj = len(CFG["seg_steps"])
        hcfg = O.Sam2Cfg(embed_dim=32, stages=(1, 2, 3, 1), global_att_blocks=(4,))
        from oracle import weights as Wt
        ssd = Wt.make_sam2_state_dict(hcfg, 1)
        fs = hcfg.feat_size
        pix, s0f, s1f = torch.randn(n_obj, 256, fs, fs, generator=g), torch.randn(n_obj, 32, 4 * fs, 4 * fs, generator=g), torch.randn(n_obj, 64, 2 * fs, 2 * fs, generator=g)
        clock("sam_heads", lambda: O.forward_sam_heads(ssd, hcfg, pix, [s0f, s1f]), T_SAM)
        high = torch.randn(n_obj, 1, hcfg.image_size, hcfg.image_size, generator=g)
        clock("memory_encoder", lambda: O.encode_new_memory(ssd, hcfg, pix, high, is_mask_from_pts=False), T_SAM)
        f_avg = round(sum(min(tt, 7) for tt in range(1, T_SAM)) / (T_SAM - 1))   # memory frames in an average bank (+ 16 pointers = 64 tokens)
        nk_avg = 4096 * f_avg + 64
        curr, cpos = torch.randn(fs * fs, n_obj, 256, generator=g), torch.randn(fs * fs, n_obj, 256, generator=g)
        mem, mpos = torch.randn(nk_avg, n_obj, 64, generator=g), torch.randn(nk_avg, n_obj, 64, generator=g)
        clock("memory_attention", lambda: O.memory_attention(ssd, hcfg, curr, cpos, mem, mpos, 64), T_SAM - 1)
    llm_forwards = N_NEW + 2
    sam_heads = t["sam_heads"] + t["memory_encoder"] + t["memory_attention"]
    total = t["clip"] + t["iv2"] + t["phi3_layer_full_forward"] * llm_forwards + hier + sam_heads
    return {"value": 1.0 / total, "unit": "clips/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"oracle (CPU fp32 port of the reference path, {CFG['tag']}) timed per full-width stage and scaled by multiplicity: CLIP 1 layer x{T_ENC} frames "
                      f"(x23), IV2 1 block x{T_ENC // 4} chunks (x39), Phi-3 1 layer at L={L} tokens (x32 layers x{llm_forwards} full forwards: use_cache=False as the "
                      f"reference), Hiera 1 block per stage (x2/6/36/4 x{T_SAM} frames), SAM heads (prompt encoder + mask decoder + pointer MLP, {n_obj} objects) "
                      f"x{T_SAM} frames, memory encoder x{T_SAM}, memory attention (4 layers, average bank of {nk_avg} keys) x{T_SAM - 1}; pre/post-processing excluded",
            "seconds_per_clip_extrapolated": total, "stage_seconds": t}


def reference_inference_reduced() -> dict:
    """ONE call of the UNMODIFIED reference's VideoGLaMMForCausalLM.inference (video branch) on the host cores, from the package staged under
    baseline/_ref (oracle/build_ref.py:stage_reference; absent -> reported as such). Config: the reduced one of tests/golden (true IV2 / CLIP
    widths and resolutions, reduced depths, tiny Phi-3 and Hiera, 3 SAM2 frames at 256 px, 2 objects, 5 new tokens) — a functional run of the
    reference's own code path next to the per-stage full-width sample, not an estimate of the full config."""
    ref_root = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.exists(os.path.join(ref_root, "VideoGLaMM", "model", "VideoGLaMM.py")):
        return {"unavailable": "baseline/_ref not staged on this box"}
    os.environ["VG_REFERENCE_ROOT"] = ref_root
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    try:
        import numpy as np
        import make_golden as MG
        ref, *_ = MG.ref_videoglamm()
        MG.ref_e2e_run(ref)                                  # warm-up
        t0 = time.perf_counter()
        ids, masks = MG.ref_e2e_run(ref)
        dt = time.perf_counter() - t0
        g = np.load(os.path.join(ROOT, "tests", "golden", "e2e_tiny.npz"))
        return {"seconds_per_call": dt, "config": "tests/golden e2e_tiny (T_enc=4, T_sam=3 at 256px, 2 [SEG], 5 new tokens, reduced depths)",
                "matches_committed_golden": bool(np.array_equal(ids, g["output_ids"])), "code": "unmodified reference, baseline/_ref"}
    except Exception as e:  # the arm must still print its line
        return {"unavailable": f"{type(e).__name__}: {e}"[:200]}


def run_reference(args):
    """`--impl reference`: the reference's CPU implementation of the path on the box's host cores, all host threads, same config/metric.
    `value` comes from the bounded per-stage sample of the oracle port (every stage of the path, full width, scaled by multiplicity);
    `ms_per_step` = 1000 / value (the extrapolated time of one step's clips); the wall time of the sample itself and one REAL call of the unmodified
    reference's inference() at a reduced config are reported next to it."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    L = 12 + L_QUERY + 208 * T_ENC
    res = None
    for _ in range(max(args.warmup, 0)):
        res = cpu_baseline_sample(L)
    vals = []
    t0 = time.perf_counter()
    for _ in range(max(args.steps, 1)):
        res = cpu_baseline_sample(L)
        vals.append(res["value"])
    wall = time.perf_counter() - t0
    v = sum(vals) / len(vals)
    real = reference_inference_reduced()
    line = {"impl": "reference", "metric": METRIC[args.config], "value": v, "unit": "clips/s", "n_gpus": int(os.environ.get("WORLD_SIZE", 1)),
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * CFG["clips"] / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": CFG["workload"] + " (same as the engine arm)", "device": "host CPU", "clips_per_step_per_gpu": CFG["clips"],
                       "sample_wall_ms_per_step": wall / max(args.steps, 1) * 1e3,
                       "note": "value = 1 / (per-stage full-width sample x multiplicity): a bounded sample of the workload, not a full run (one full-size clip takes "
                               "minutes on host cores)"},
            "cpu_baseline": {**res, "value": v}, "reference_inference_reduced": real,
            "e2e": {"value": v, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--decode-batch", type=int, default=0,
                    help="decode loops of this many queued clips run as one batched loop (>= clips per step, <= 8); 0 = auto: min(8, half the clips of the run)")
    ap.add_argument("--sequential", action="store_true", help="one clip at a time (latency mode); default overlaps consecutive clips")
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS), help="BASELINE.json configs[N-1]: 2 (headline, default), 3, 4, 5")
    a = ap.parse_args()
    set_config(a.config)
    if a.impl == "reference":
        run_reference(a)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device — the engine has no CPU path (use --impl reference for the CPU arm)")
        run_engine(a)
