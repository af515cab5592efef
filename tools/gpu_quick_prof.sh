mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_program.py -m gpu -q --timeout 900 -p no:cacheprovider -k "golden" > gpurun_out/t_prog.log 2>&1; echo "prog rc=$?" > gpurun_out/summary.txt
python tools/profile_records.py lat 256 auto > gpurun_out/prof_lat_auto.txt 2>&1
python tools/profile_records.py pos 256 auto > gpurun_out/prof_pos_auto.txt 2>&1
cat gpurun_out/summary.txt; tail -n 2 gpurun_out/t_prog.log | cut -c1-200
head -1 gpurun_out/prof_lat_auto.txt gpurun_out/prof_pos_auto.txt
grep -E "SA1" gpurun_out/prof_lat_auto.txt | cut -c1-110
