python -m pytest tests -m gpu -q > gpurun_out/mid_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/mid_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/mid_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/mid_smoke.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/mid_bench.json 2> gpurun_out/mid_bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/mid_bench.json
