mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1 --warmup 1 --ddpm-steps 20 > gpurun_out/bench_2gpu_dbg.log 2>&1; echo "2gpu rc=$?"
tail -n 3 gpurun_out/bench_2gpu_dbg.log | cut -c1-600
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/bench_2gpu_ref.log 2>&1; echo "2gpu ref rc=$?"
tail -n 1 gpurun_out/bench_2gpu_ref.log | cut -c1-300
