timeout 300 python -m pytest tests/test_gpu_kpconv.py -x -q -k "producer_block" > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2g_pytest.log
timeout 300 python tools/host_profile.py > gpurun_out/r2g_host.log 2>&1; echo "rc=$?"; grep -v Warn gpurun_out/r2g_host.log | head -70
