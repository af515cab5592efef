mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
python -m pytest tests/test_gpu_index_ops.py tests/test_gpu_dropin_modules.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/t_index.log 2>&1; echo "index+dropin rc=$?" >> gpurun_out/summary.txt
timeout 1500 python -m pytest tests/test_gpu_program.py -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/t_prog.log 2>&1; echo "prog rc=$?" >> gpurun_out/summary.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/summary.txt
python tools/profile_records.py lat 256 auto > gpurun_out/prof_lat_auto.txt 2>&1
python tools/profile_records.py pos 256 auto > gpurun_out/prof_pos_auto.txt 2>&1
timeout 900 python bench.py --steps 1 --warmup 1 > gpurun_out/bench_full_auto.log 2>&1; echo "bench rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -n 4 gpurun_out/t_index.log | cut -c1-300; tail -n 6 gpurun_out/t_prog.log | cut -c1-300; tail -n 2 gpurun_out/smoke.log
head -1 gpurun_out/prof_lat_auto.txt gpurun_out/prof_pos_auto.txt
tail -n 1 gpurun_out/bench_full_auto.log | cut -c1-200
