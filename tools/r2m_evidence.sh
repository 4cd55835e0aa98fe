#!/bin/bash
# Round-2 (second half) profiling evidence in one call: KPConv step capture (traffic + summary), launch list of the bench
# command, full captures of the stem and the tcgen05 attention kernel.  Only small files travel back.
python bench.py --no-alt --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/r2m_plain.log 2>&1; echo "plain rc=$?"
ncu --set full --clock-control none -k regex:k_kpconv --launch-skip 33 -c 11 \
    -f -o /tmp/kpconv_step python bench.py --no-alt --no-cpu-baseline --steps 1 --warmup 3 > gpurun_out/r2m_ncu_kpconv.log 2>&1; echo "ncu kpconv rc=$?"
python tools/make_traffic.py /tmp/kpconv_step.ncu-rep gpurun_out/r2m_kpconv_step_ncu_summary.csv > gpurun_out/r2m_kpconv_traffic.log 2>&1
cp profiles/kpconv_traffic.json gpurun_out/r2m_kpconv_traffic.json
tail -8 gpurun_out/r2m_kpconv_traffic.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2m_launches.csv \
    python bench.py --no-alt --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/r2m_ncu_list.log 2>&1; echo "ncu list rc=$?"
python tools/launch_summary.py gpurun_out/r2m_launches.csv gpurun_out/r2m_launches_summary.csv
rm -f gpurun_out/r2m_launches.csv
ncu --set full --clock-control none --import-source on -k regex:"k_attention_tc|k_kpconv_cin1_t" --launch-skip 39 -c 3 \
    -f -o gpurun_out/r2m_attn_stem python bench.py --no-alt --no-cpu-baseline --steps 1 --warmup 3 > gpurun_out/r2m_ncu_attn.log 2>&1; echo "ncu attn rc=$?"
python tools/ncu_summary.py gpurun_out/r2m_attn_stem.ncu-rep gpurun_out/r2m_attention_tc_stem_ncu_summary.csv
ls -la gpurun_out/r2m_*
