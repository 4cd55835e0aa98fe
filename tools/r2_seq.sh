timeout 600 python -m pytest tests/test_gpu_preprocess.py tests/test_gpu_matching.py -x -q > gpurun_out/r2p_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2p_pytest.log
timeout 300 python tools/host_profile.py --top 3 > gpurun_out/r2p_host.log 2>&1; echo "rc=$?"; grep -v Warn gpurun_out/r2p_host.log | head -2
