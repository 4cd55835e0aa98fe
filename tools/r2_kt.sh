timeout 300 python tools/kernel_times.py --kind kitti --pairs 8 --top 25 > gpurun_out/r2g_kt_kitti.log 2>&1; echo "rc=$?"; grep -v Warn gpurun_out/r2g_kt_kitti.log | head -30
timeout 300 python tools/kernel_times.py --kind modelnet --pairs 64 --top 16 > gpurun_out/r2g_kt_modelnet.log 2>&1; echo "rc=$?"; grep -v Warn gpurun_out/r2g_kt_modelnet.log | head -20
timeout 300 python tools/kernel_times.py --pairs 32 --top 22 > gpurun_out/r2g_kt_3dmatch.log 2>&1; echo "rc=$?"; grep -v Warn gpurun_out/r2g_kt_3dmatch.log | head -26
