#!/bin/bash
# every KPConv launch of one bench step under ncu --set full (the same command is run plain first)
set -e
python bench.py --no-alt --no-cpu-baseline --steps 1 --warmup 3 > gpurun_out/plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_kpconv --launch-skip 33 -c 11 \
    -f -o gpurun_out/kpconv_step python bench.py --no-alt --no-cpu-baseline --steps 1 --warmup 3 > gpurun_out/ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep
