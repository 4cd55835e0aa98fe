#!/bin/bash
# one full step's KPConv tensor-core launches under ncu (run the same bench command plain first)
set -e
python bench.py --no-alt --no-cpu-baseline --steps 1 --warmup 3 > gpurun_out/plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_kpconv_tc --launch-skip 30 -c 10 \
    -f -o gpurun_out/kpconv_tc_step python bench.py --no-alt --no-cpu-baseline --steps 1 --warmup 3 > gpurun_out/ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep
