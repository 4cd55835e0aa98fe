timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"k_kpconv_s<.int.256" -c 1 -f -o gpurun_out/r2n_c256 python tools/kpconv_gen_bench.py --pairs 32 --reps 1 --gens 3 > gpurun_out/r2n_c256.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r2n_c256.log
ls -la gpurun_out/r2n_c256.ncu-rep
