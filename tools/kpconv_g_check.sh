set -x
timeout 400 python -m pytest tests/test_gpu_kpconv_gather.py -x -q -s > gpurun_out/kpconv_g_gather.log 2>&1; echo "gather rc=$?"
tail -30 gpurun_out/kpconv_g_gather.log
timeout 300 python tools/kpconv_gen_bench.py --pairs 8 > gpurun_out/kpconv_g_genbench.log 2>&1; echo "genbench rc=$?"
cat gpurun_out/kpconv_g_genbench.log | tail -15
timeout 300 bash tools/kpconv_g_trace.sh > gpurun_out/kpconv_g_trace.log 2>&1; echo "trace rc=$?"
cat gpurun_out/kpconv_g_trace.log | grep -v "^+" | tail -50
