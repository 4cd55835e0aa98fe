timeout 600 python -m pytest tests/test_gpu_kpconv.py tests/test_gpu_matching.py -x -q > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2q_pytest.log
timeout 300 python tools/kernel_times.py --kind 3dmatch --pairs 32 --top 14 > gpurun_out/r2q_kt_3dmatch.log 2>&1; echo "rc=$?"; grep -v Warn gpurun_out/r2q_kt_3dmatch.log | grep -E "total device|cin1|k_max_pool"
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-alt > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/r2q_bench.json
