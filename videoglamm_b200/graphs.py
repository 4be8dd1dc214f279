"""CUDA-graph capture of launch-bound stages. The per-launch host cost of the Python/ctypes binding (~20 us) dominates
stages made of small kernels (LLM decode: ~260 launches per token; SAM2 per-frame heads: hundreds); every stage whose launch
sequence is a pure function of static shapes is captured per shape key with `torch.cuda.CUDAGraph` (our kernels are
launched on torch's current stream, so they are captured like any other node) and replayed afterwards.

Cache policy (real eval sets vary in frame count, resolution, [SEG] count and prompt length on nearly every sample):
  - the FIRST call of a key runs eagerly (it is the warm-up: lazy function attributes, tensor-map cache, allocator pools) and returns that
    result; the key is captured on its second call, so a shape seen once costs one eager run, not warm-up + capture + replay;
  - entries are kept in LRU order and evicted beyond `max_entries` (VG_GRAPH_CACHE, default 6 per stage); all entries of a stage share ONE
    private memory pool, so evicting a graph returns its activations to the pool the next capture draws from instead of growing HBM use
    (safe because a stage's outputs are consumed before that stage is replayed again, and a stage never replays concurrently with itself);
  - captures are serialised process-wide and use `capture_error_mode="thread_local"` with a stream (not device) synchronise, so that the other
    stage thread of the clip pipeline may keep launching and allocating while one stage captures a cold shape."""
from __future__ import annotations

import collections
import gc
import os
import threading

import torch

from . import ops

ENABLED = os.environ.get("VG_NO_GRAPHS", "0") != "1"
MAX_ENTRIES = int(os.environ.get("VG_GRAPH_CACHE", "6"))
_CAPTURE_LOCK = threading.Lock()


class GraphedStage:
    """fn(**static_inputs) -> tensor | tuple | list | dict of tensors, captured per key. Inputs are copied into static buffers before each
    replay; outputs are views of graph-owned memory, valid until the next call of the same stage."""

    def __init__(self, name: str, max_entries: int | None = None):
        self.name = name
        self.max_entries = max_entries or MAX_ENTRIES
        self.entries: "collections.OrderedDict" = collections.OrderedDict()
        self.seen: "collections.OrderedDict" = collections.OrderedDict()       # keys that ran eagerly once (bounded)
        self.pool = None
        self.stats = {"eager": 0, "captures": 0, "replays": 0, "evictions": 0}

    def clear(self):
        self.entries.clear()
        self.seen.clear()

    def __call__(self, key, fn, mutable_state: tuple = (), **inputs: torch.Tensor):
        """`mutable_state`: tensors fn updates in place (device-side counters); they are restored after the capture so that
        only the replay takes effect."""
        if not ENABLED:
            return fn(**inputs)
        ent = self.entries.get(key)
        if ent is None:
            if key not in self.seen:                           # first sight: eager run (= the warm-up), no capture yet
                self.seen[key] = True
                while len(self.seen) > 64:
                    self.seen.popitem(last=False)
                self.stats["eager"] += 1
                return fn(**inputs)
            ent = self._capture(key, fn, mutable_state, inputs)
        else:
            self.entries.move_to_end(key)
        g, static, out, n_nodes = ent
        for k, v in inputs.items():
            if static[k].data_ptr() != v.data_ptr():
                static[k].copy_(v, non_blocking=True)
        g.replay()
        self.stats["replays"] += 1
        ops.launch_count_add(n_nodes)
        return out

    def _capture(self, key, fn, mutable_state, inputs):
        with _CAPTURE_LOCK:
            while len(self.entries) >= self.max_entries:       # LRU eviction: the graph and its static buffers go back to the shared pool
                self.entries.popitem(last=False)
                self.stats["evictions"] += 1
            cur = torch.cuda.current_stream()
            static = {k: v.clone() for k, v in inputs.items()}
            saved = [t.clone() for t in mutable_state]
            cur.synchronize()                                  # this stream only: the other pipeline stage keeps running
            if self.pool is None:
                self.pool = torch.cuda.graph_pool_handle()
            g = torch.cuda.CUDAGraph()
            n0 = ops.launch_count()
            # A garbage collection inside the capture may destroy an unreachable CUDAGraph of an earlier engine (cudaGraphExecDestroy is not
            # permitted while this thread captures and invalidates the capture): collect before, keep the collector off until the capture ends.
            gc_on = gc.isenabled()
            gc.collect()
            gc.disable()
            try:
                with torch.cuda.graph(g, pool=self.pool, capture_error_mode="thread_local"):
                    out = fn(**static)
            finally:
                if gc_on:
                    gc.enable()
            for t, sv in zip(mutable_state, saved):
                t.copy_(sv)
            ent = (g, static, out, ops.launch_count() - n0)
            self.entries[key] = ent
            self.stats["captures"] += 1
        return ent
