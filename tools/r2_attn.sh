timeout 300 python -m pytest tests/test_gpu_transformer.py -x -q -s -k "attention_varlen" > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "gen 2|passed|failed|Error|error" gpurun_out/r2l_pytest.log | head -12
for k in "3dmatch 32" "modelnet 64" "kitti 8"; do set -- $k
timeout 300 python tools/kernel_times.py --kind $1 --pairs $2 --top 12 > gpurun_out/r2l_kt_$1.log 2>&1; echo "rc=$?"; grep -v Warn gpurun_out/r2l_kt_$1.log | grep -E "total device|k_attention"
done
