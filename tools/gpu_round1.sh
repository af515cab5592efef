mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
python -m pytest tests/test_gpu_index_ops.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/t_index.log 2>&1; echo "index rc=$?" >> gpurun_out/summary.txt
python -m pytest tests/test_gpu_program.py -m gpu -q --timeout 900 -p no:cacheprovider -k "simt" > gpurun_out/t_prog_simt.log 2>&1; echo "prog_simt rc=$?" >> gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_program.py -m gpu -q --timeout 600 -p no:cacheprovider -k "not simt" > gpurun_out/t_prog_auto.log 2>&1; echo "prog_auto rc=$?" >> gpurun_out/summary.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --ddpm-steps 20 --steps 1 --warmup 1 --backend simt > gpurun_out/bench_dbg_simt.log 2>&1; echo "bench_simt rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --ddpm-steps 20 --steps 1 --warmup 1 --backend auto > gpurun_out/bench_dbg_auto.log 2>&1; echo "bench_auto rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -5 gpurun_out/t_index.log gpurun_out/t_prog_simt.log gpurun_out/t_prog_auto.log gpurun_out/smoke.log gpurun_out/bench_dbg_simt.log gpurun_out/bench_dbg_auto.log
