set -x
timeout 600 python -m pytest tests/test_gpu_kpconv_staged.py -x -q -s > gpurun_out/kpconv_s_tests.log 2>&1; echo "staged rc=$?"
tail -25 gpurun_out/kpconv_s_tests.log
timeout 300 python tools/kpconv_gen_bench.py --pairs 8 --gens 1,3 > gpurun_out/kpconv_s_genbench.log 2>&1; echo "genbench rc=$?"
tail -12 gpurun_out/kpconv_s_genbench.log
