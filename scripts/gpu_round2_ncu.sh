#!/bin/bash
# ncu --set full of each hot kernel at its config-2 shape incl. the batched-decode GEMVs and the fused mask decoder (one GPU; no bench number is
# taken from this run). Only the summaries travel back.
mkdir -p gpurun_out
timeout 1200 ncu -k regex:"gemm_tcgen05|gemv_|flash_attn|window_attn|attn_decode|mask_decoder" --set full --clock-control none -o /tmp/r02_targets -f python scripts/gpu_ncu_targets.py > gpurun_out/r02_ncu_targets.log 2>&1
tail -2 gpurun_out/r02_ncu_targets.log
ncu -i /tmp/r02_targets.ncu-rep --page raw --csv > /tmp/r02_targets_raw.csv 2>/dev/null
python scripts/ncu_summarize.py /tmp/r02_targets_raw.csv gpurun_out/r02_ncu_targets labelled | cut -c1-250
