timeout 600 python -m pytest tests/test_gpu_kpconv.py tests/test_gpu_matching.py -x -q > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2q_pytest.log
for k in "3dmatch 32" "kitti 8"; do set -- $k
timeout 300 python tools/kernel_times.py --kind $1 --pairs $2 --top 12 > gpurun_out/r2q_kt_$1.log 2>&1; echo "rc=$?"; grep -v Warn gpurun_out/r2q_kt_$1.log | grep -E "total device|k_max_pool"
done
