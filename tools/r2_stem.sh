timeout 300 python tools/stem_bench.py > gpurun_out/r2f_stem.log 2>&1; echo "rc=$?"
grep -v Warn gpurun_out/r2f_stem.log | tail -3
