mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
python -m pytest tests/test_gpu_index_ops.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/t_index.log 2>&1; echo "index rc=$?" >> gpurun_out/summary.txt
timeout 1200 python -m pytest tests/test_gpu_program.py -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/t_prog.log 2>&1; echo "prog rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --ddpm-steps 20 --steps 1 --warmup 1 --backend simt > gpurun_out/bench_dbg_simt.log 2>&1; echo "bench_simt rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --ddpm-steps 20 --steps 1 --warmup 1 --backend auto > gpurun_out/bench_dbg_auto.log 2>&1; echo "bench_auto rc=$?" >> gpurun_out/summary.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1.csv python bench.py --ddpm-steps 2 --steps 1 --warmup 1 --backend auto > gpurun_out/ncu_launch.log 2>&1; echo "ncu rc=$?" >> gpurun_out/summary.txt
timeout 1500 python bench.py --steps 1 --warmup 1 --backend auto > gpurun_out/bench_full_auto.log 2>&1; echo "bench_full rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
for f in t_index t_prog bench_dbg_simt bench_dbg_auto bench_full_auto; do echo "== $f"; tail -n 4 gpurun_out/$f.log | cut -c1-600; done
