set -x
timeout 60 ./tools/umma_mn_test > gpurun_out/kpconv_g_umma_mn.log 2>&1; echo "umma_mn rc=$?"; cat gpurun_out/kpconv_g_umma_mn.log
timeout 400 python -m pytest tests/test_gpu_kpconv_gather.py -x -q -s > gpurun_out/kpconv_g_gather.log 2>&1; echo "gather rc=$?"
tail -30 gpurun_out/kpconv_g_gather.log
timeout 300 python tools/kpconv_gen_bench.py --pairs 8 > gpurun_out/kpconv_g_genbench.log 2>&1; echo "genbench rc=$?"
cat gpurun_out/kpconv_g_genbench.log | tail -15
