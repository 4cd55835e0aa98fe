timeout 600 python tools/kpconv_gen_bench.py --pairs 32 --gens 1,3 > gpurun_out/kpconv_s_genbench32.log 2>&1; echo "genbench rc=$?"
tail -12 gpurun_out/kpconv_s_genbench32.log
