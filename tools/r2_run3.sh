set -x
timeout 60 ./tools/umma_mn_test > gpurun_out/r2c_umma_mn.log 2>&1; echo "umma_mn rc=$?"; cat gpurun_out/r2c_umma_mn.log
timeout 300 python -m pytest tests/test_gpu_kpconv_gather.py -x -q -s > gpurun_out/r2c_gather.log 2>&1; echo "gather rc=$?"
tail -25 gpurun_out/r2c_gather.log
python -m pytest tests -m gpu -q --deselect tests/test_gpu_kpconv_gather.py > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?"
tail -40 gpurun_out/r2c_pytest.log
SPR_GEMM_DETAIL=1 SPR_IN_DETAIL=1 SPR_GAPS=1 python tools/kernel_times.py --pairs 32 --arch 4stage --top 40 > gpurun_out/r2c_kernel_times.log 2>&1; echo "kt rc=$?"
head -45 gpurun_out/r2c_kernel_times.log
python tools/profile_step.py > gpurun_out/r2c_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2c_launches.csv python tools/profile_step.py > gpurun_out/r2c_ncu_list.log 2>&1; echo "list rc=$?"
ncu --set full --clock-control none --profile-from-start off -k 'regex:k_gemm_tc|k_in_apply|k_in_stats|k_in_partial|k_max_pool|k_attention' -f -o /tmp/r2c_rest_step python tools/profile_step.py > gpurun_out/r2c_ncu_rest.log 2>&1; echo "ncu rest rc=$?"
python tools/ncu_summary.py /tmp/r2c_rest_step.ncu-rep gpurun_out/r2c_rest_step_ncu_summary.csv
du -sh gpurun_out
