timeout 300 python tools/kernel_times.py --pairs 32 --top 12 > gpurun_out/r2j_kt.log 2>&1; echo "rc=$?"; grep -v Warn gpurun_out/r2j_kt.log | grep -E "total device|k_attention"
timeout 300 python tools/kernel_times.py --kind modelnet --pairs 64 --top 6 > gpurun_out/r2j_kt_m.log 2>&1; echo "rc=$?"; grep -v Warn gpurun_out/r2j_kt_m.log | grep -E "total device|k_attention"
timeout 300 python tools/kernel_times.py --kind kitti --pairs 8 --top 6 > gpurun_out/r2j_kt_k.log 2>&1; echo "rc=$?"; grep -v Warn gpurun_out/r2j_kt_k.log | grep -E "total device|k_attention"
timeout 300 python -m pytest tests/test_gpu_transformer.py -x -q > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2j_pytest.log
