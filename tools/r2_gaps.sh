SPR_GAPS=1 timeout 300 python tools/kernel_times.py --pairs 32 --top 3 > gpurun_out/r2k_gaps.log 2>&1; echo "rc=$?"; grep -v Warn gpurun_out/r2k_gaps.log | tail -18
