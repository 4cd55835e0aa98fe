#!/bin/bash
# Round-end evidence: GPU tests, smoke, both bench arms, the launch list and the KPConv capture.
python -m pytest tests -m gpu -x -q > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?"
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_ref.err; echo "ref rc=$?"
python bench.py --no-alt --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/plain2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/final_launches.csv \
    python bench.py --no-alt --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/ncu_list.log 2>&1; echo "launch list rc=$?"
bash tools/ncu_kpconv.sh; echo "kpconv ncu rc=$?"
