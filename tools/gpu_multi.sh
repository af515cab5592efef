# 2-GPU check of the sharded pipeline + NCCL all-gather (run with: gpurun --gpus 2 -- 'bash tools/gpu_multi.sh')
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_2gpu.log 2>&1; echo "2gpu rc=$?"
tail -n 1 gpurun_out/bench_2gpu.log | cut -c1-700
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/bench_2gpu_ref.log 2>&1; echo "2gpu ref rc=$?"
tail -n 1 gpurun_out/bench_2gpu_ref.log | cut -c1-300
