#!/bin/bash
# every KPConv launch of one bench step under ncu --set full (the same command is run plain first); the report stays on
# the box (gpurun merges at most 64 MiB back): only the condensed CSV and the traffic JSON travel
set -e
python bench.py --no-alt --no-cpu-baseline --steps 1 --warmup 3 > gpurun_out/plain.log 2>&1
ncu --set full --clock-control none -k regex:k_kpconv --launch-skip 33 -c 11 \
    -f -o /tmp/kpconv_step python bench.py --no-alt --no-cpu-baseline --steps 1 --warmup 3 > gpurun_out/ncu.log 2>&1
python tools/make_traffic.py /tmp/kpconv_step.ncu-rep gpurun_out/kpconv_step_ncu_summary.csv > gpurun_out/kpconv_traffic.log 2>&1
cp profiles/kpconv_traffic.json gpurun_out/kpconv_traffic.json
tail -12 gpurun_out/kpconv_traffic.log
