python bench.py --no-alt --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_order.json
SPR_NO_ORDER=1 python bench.py --no-alt --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_noorder.json
