set -x
ncu --set full --clock-control none --import-source on -k regex:k_kpconv_tc -c 3 -f -o gpurun_out/kpconv_tc_src python tools/kpconv_gen_bench.py --pairs 8 --reps 1 > gpurun_out/kpconv_tc_src.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep
