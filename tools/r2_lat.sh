SPR_GAPS=1 timeout 300 python tools/kernel_times.py --pairs 1 --points 5000 --arch 3stage --top 8 > gpurun_out/r2o_lat.log 2>&1; echo "rc=$?"; grep -v Warn gpurun_out/r2o_lat.log | tail -28
