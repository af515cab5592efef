mkdir -p gpurun_out
N=${N:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 2 --warmup 3 > gpurun_out/bench_${N}gpu.log 2>&1; tail -1 gpurun_out/bench_${N}gpu.log | cut -c1-400
