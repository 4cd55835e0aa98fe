set -x
python -m pytest tests -m gpu -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r2b_pytest.log
SPR_GEMM_DETAIL=1 SPR_IN_DETAIL=1 SPR_GAPS=1 python tools/kernel_times.py --pairs 32 --arch 4stage --top 40 > gpurun_out/r2b_kernel_times.log 2>&1; echo "kt rc=$?"
head -50 gpurun_out/r2b_kernel_times.log
python tools/profile_step.py > gpurun_out/r2b_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2b_launches.csv python tools/profile_step.py > gpurun_out/r2b_ncu_list.log 2>&1; echo "list rc=$?"
ncu --set full --clock-control none --import-source on --profile-from-start off -k 'regex:k_gemm_tc|k_in_apply|k_in_stats|k_in_partial|k_max_pool|k_attention' -f -o gpurun_out/r2b_rest_step python tools/profile_step.py > gpurun_out/r2b_ncu_rest.log 2>&1; echo "ncu rest rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -3
