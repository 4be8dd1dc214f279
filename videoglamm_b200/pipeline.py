"""Two-stage software pipeline over independent clips for throughput serving: the LMM stage (towers, prefill, latency-bound decode) of clip
i+1 runs on one CUDA stream while the SAM2 stage (tensor-bound Hiera / memory attention) of clip i runs on another. Each stage has its own
host thread (the decode loop syncs with the host once per token) and its own CUDA-graph pools. Results are identical to
`VideoGLaMMForCausalLM.inference` called clip by clip — only the interleaving on the device differs."""
from __future__ import annotations

import queue
import threading

import numpy as np
import torch


class ClipPipeline:
    def __init__(self, model, depth: int = 2, decode_batch: int = 1):
        """decode_batch > 1: the LMM stage takes that many consecutive clips at a time and runs their HBM-bound decode loops as ONE batched loop
        (LmmEngine.generate_batch: every weight matrix is streamed once per step for all of them) — BASELINE.json configs[2], 4 clips per GPU."""
        self.model = model
        model._ensure_engines()
        self.depth = max(depth, decode_batch)
        self.decode_batch = max(1, min(int(decode_batch), 8))
        self.s_lmm = torch.cuda.Stream(device=model.device)
        self.s_sam = torch.cuda.Stream(device=model.device)
        self.s_copy = torch.cuda.Stream(device=model.device)   # host -> device prefetch of the SAM2 frames (100 MB per clip at 1024 px)
        self.trace = None                                      # set to a list to collect (tag, index, cuda event) marks of both stages (diagnostics)

    def _mark(self, tag: str, i: int, stream):
        if self.trace is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(stream)
            self.trace.append((tag, i, ev))

    @torch.no_grad()
    def run(self, clips: list[dict], max_new_tokens: int, forced_tokens: dict | None = None) -> list[tuple]:
        """clips: [{"images", "context_images", "images_for_sam", "input_ids", "original_size"}] (host or device tensors).
        Returns [(output_ids, segments, packed)] in order; segments as in `inference` ({frame: {obj: bool ndarray}}), packed = the same masks
        bit-packed and still on the device (parallel.PackedMasks; None for a clip without [SEG]) for the multi-GPU gather."""
        m = self.model
        dev = torch.device("cuda", self.model.device.index if self.model.device.index is not None else torch.cuda.current_device())
        q: queue.Queue = queue.Queue(maxsize=self.depth)
        results: list = [None] * len(clips)
        errors: list = []

        def lmm_stage():
            try:
                torch.cuda.set_device(dev)
                with torch.cuda.stream(self.s_lmm):
                    S = self.decode_batch
                    for i0 in range(0, len(clips), S):
                        group, reqs, sams = clips[i0:i0 + S], [], []
                        for c in group:
                            # the SAM2 frames of a clip are not needed before its LMM stage is over: copy them on their own stream now, so
                            # the transfer runs under the LMM kernels instead of in front of the SAM2 stage
                            sam_dev, ev_copy = c["images_for_sam"], None
                            if not sam_dev.is_cuda:
                                with torch.cuda.stream(self.s_copy):
                                    sam_dev = sam_dev.to(dev, non_blocking=True)
                                    ev_copy = torch.cuda.Event()
                                    ev_copy.record(self.s_copy)
                            sams.append((sam_dev, ev_copy))
                            enc_images, enc_ctx = m.encoder_inputs(c["images"], c.get("context_images"))   # fp32 tensors, or raw uint8 frames -> GPU pre-processing
                            reqs.append((c["input_ids"].to(dev, non_blocking=True), enc_images, enc_ctx))
                        self._mark("lmm0", i0, self.s_lmm)
                        outs = m.lmm.generate_batch(reqs, max_new_tokens, eos_ids=m._eos_ids(), forced=forced_tokens)
                        self._mark("lmm1", i0, self.s_lmm)
                        for j, ((out_ids, emb), (sam_dev, ev_copy)) in enumerate(zip(outs, sams)):
                            emb = emb.clone()
                            emb.record_stream(self.s_sam)      # consumed on the other stream: keep the allocator from recycling it early
                            ev = torch.cuda.Event()
                            ev.record(self.s_lmm)
                            q.put((i0 + j, out_ids, emb, ev, sam_dev, ev_copy))
            except Exception as e:  # surface, don't swallow (the reference's eval loops silently skip failures)
                errors.append(e)
            finally:
                q.put(None)

        def sam_stage():
            try:
                torch.cuda.set_device(dev)
                with torch.cuda.stream(self.s_sam):
                    while True:
                        item = q.get()
                        if item is None:
                            return
                        i, out_ids, emb, ev, sam_dev, ev_copy = item
                        self.s_sam.wait_event(ev)
                        if ev_copy is not None:
                            self.s_sam.wait_event(ev_copy)
                            sam_dev.record_stream(self.s_sam)
                        if emb.shape[0] == 0:
                            results[i] = (out_ids, {}, None)
                            continue
                        H, W = clips[i]["original_size"]
                        self._mark("sam0", i, self.s_sam)
                        if sam_dev.dtype == torch.uint8:       # raw frames [T,H,W,3]: GPU pre-processing (resize / normalise / squash)
                            sam_dev = m.sam_preprocessor_gpu().preprocess_video(sam_dev)[0]
                        frames = m.sam.run_video(sam_dev, emb[:, None, :], (int(H), int(W)), binary="packed")
                        segs, packed = m.masks_to_host(frames, int(H), int(W))
                        self._mark("sam1", i, self.s_sam)
                        results[i] = (out_ids, segs, packed)
            except Exception as e:
                errors.append(e)
                while q.get() is not None:
                    pass

        cur = torch.cuda.current_stream(dev)
        self.s_lmm.wait_stream(cur)
        self.s_sam.wait_stream(cur)
        self.s_copy.wait_stream(cur)
        ta, tb = threading.Thread(target=lmm_stage), threading.Thread(target=sam_stage)
        ta.start(); tb.start()
        ta.join(); tb.join()
        cur.wait_stream(self.s_lmm)
        cur.wait_stream(self.s_sam)
        cur.wait_stream(self.s_copy)
        if errors:
            raise errors[0]
        return results
