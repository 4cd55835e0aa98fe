#!/bin/bash
# ncu --set full of one launch of a kernel of the bench step:  tools/ncu_kernel.sh <regex> <launch-skip> <count> <out>
set -e
python bench.py --no-alt --no-cpu-baseline --steps 1 --warmup 3 > gpurun_out/plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:$1 --launch-skip $2 -c $3 \
    -f -o gpurun_out/$4 python bench.py --no-alt --no-cpu-baseline --steps 1 --warmup 3 > gpurun_out/ncu.log 2>&1
ls -la gpurun_out/$4.ncu-rep
