#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu -k regex:"norm_vec|pool2x2|dwconv|rope2d_cached|add_|im2col|bilinear" --set full --clock-control none -o /tmp/r02_elem -f python scripts/gpu_ncu_elementwise.py > gpurun_out/r02_ncu_elem.log 2>&1
tail -2 gpurun_out/r02_ncu_elem.log
ncu -i /tmp/r02_elem.ncu-rep --page raw --csv > /tmp/r02_elem_raw.csv 2>/dev/null
python scripts/ncu_summarize.py /tmp/r02_elem_raw.csv gpurun_out/r02_ncu_elementwise | cut -c1-220
