set -x
ncu --set full --clock-control none --import-source on -k regex:k_kpconv_s -c 3 -f -o gpurun_out/kpconv_s_src python tools/kpconv_gen_bench.py --pairs 8 --reps 1 --gens 3 > gpurun_out/kpconv_s_src.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep
