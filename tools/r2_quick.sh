set -x
timeout 600 python -m pytest tests/test_gpu_kpconv.py tests/test_gpu_transformer.py -x -q > gpurun_out/quick_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/quick_pytest.log
timeout 300 python tools/kernel_times.py --pairs 32 --arch 4stage --top 14 > gpurun_out/quick_kt.log 2>&1; echo "kt rc=$?"; grep -v Warn gpurun_out/quick_kt.log | head -18
