for m in 0 1; do echo "SPR_S_STREAM=$m"; SPR_S_STREAM=$m timeout 300 python tools/kpconv_gen_bench.py --pairs 8 --gens 1,3 2>&1 | tail -9; done > gpurun_out/kpconv_s_ab.log 2>&1
cat gpurun_out/kpconv_s_ab.log
