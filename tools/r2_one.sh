python -m pytest tests/test_gpu_matching.py -q -x -s -k "full_forward" > gpurun_out/one.log 2>&1; echo "rc=$?"; grep -v "^$" gpurun_out/one.log | tail -25
