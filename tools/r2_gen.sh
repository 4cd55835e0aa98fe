timeout 300 python tools/kpconv_gen_bench.py --pairs 32 --gens 3 > gpurun_out/r2s_gen54.log 2>&1; echo "rc=$?"
grep -v Warn gpurun_out/r2s_gen54.log | tail -8
SPR_KPCONV_TQCAP=64 timeout 300 python tools/kpconv_gen_bench.py --pairs 32 --gens 3 > gpurun_out/r2s_gen64.log 2>&1; grep -v Warn gpurun_out/r2s_gen64.log | tail -8
SPR_KPCONV_TQCAP=36 timeout 300 python tools/kpconv_gen_bench.py --pairs 32 --gens 3 > gpurun_out/r2s_gen36.log 2>&1; grep -v Warn gpurun_out/r2s_gen36.log | tail -8
timeout 600 python -m pytest tests/test_gpu_kpconv_staged.py -x -q > gpurun_out/r2s_pytest_s.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2s_pytest_s.log
