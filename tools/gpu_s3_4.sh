mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_program.py -x -q -m gpu -p no:cacheprovider --timeout 600 -k "teacher_forced" > gpurun_out/t_prog.log 2>&1; echo "pytest teacher rc=$?" >> gpurun_out/summary.txt
for v in 2 1; do
  SLIDE_TC_PERSIST=$v timeout 300 python tools/profile_records.py lat 256 auto > gpurun_out/s4_lat_p$v.txt 2>&1
  SLIDE_TC_PERSIST=$v timeout 300 python tools/profile_records.py pos 256 auto > gpurun_out/s4_pos_p$v.txt 2>&1
  echo "persist=$v $(head -1 gpurun_out/s4_lat_p$v.txt)" >> gpurun_out/summary.txt
  echo "persist=$v $(head -1 gpurun_out/s4_pos_p$v.txt)" >> gpurun_out/summary.txt
done
cat gpurun_out/summary.txt; tail -n 6 gpurun_out/t_prog.log | cut -c1-400
