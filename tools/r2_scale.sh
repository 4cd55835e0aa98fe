N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2m_bench_n$N.json 2> gpurun_out/r2m_bench_n$N.err; echo "rc=$?"; cut -c1-300 gpurun_out/r2m_bench_n$N.json
