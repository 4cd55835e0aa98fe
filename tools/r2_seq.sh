timeout 300 python -m pytest tests/test_gpu_transformer.py -x -q > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2o_pytest.log
timeout 300 python tools/host_profile.py --top 12 > gpurun_out/r2o_host.log 2>&1; echo "rc=$?"; grep -v Warn gpurun_out/r2o_host.log | head -4
timeout 300 python tools/host_profile.py --pairs 4 --points 5000 --top 3 > gpurun_out/r2o_host4.log 2>&1; grep -v Warn gpurun_out/r2o_host4.log | head -2
