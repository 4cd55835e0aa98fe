timeout 300 python tools/kpconv_gen_bench.py --pairs 32 --gens 1,3 > gpurun_out/r2r_gen.log 2>&1; echo "rc=$?"
grep -v Warn gpurun_out/r2r_gen.log | tail -9
SPR_KPCONV_STASH=0 timeout 300 python tools/kpconv_gen_bench.py --pairs 32 --gens 3 > gpurun_out/r2r_gen0.log 2>&1; grep -v Warn gpurun_out/r2r_gen0.log | tail -8
timeout 600 python -m pytest tests/test_gpu_kpconv_staged.py tests/test_gpu_kpconv.py -x -q > gpurun_out/r2r_pytest_s.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2r_pytest_s.log
