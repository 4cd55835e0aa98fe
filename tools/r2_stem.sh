timeout 300 python tools/stem_bench.py > gpurun_out/r2f_stem.log 2>&1; echo "rc=$?"
grep -v Warn gpurun_out/r2f_stem.log | tail -3
timeout 600 python -m pytest tests/test_gpu_kpconv.py -x -q -k "stem or golden or producer_block" > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2f_pytest.log
