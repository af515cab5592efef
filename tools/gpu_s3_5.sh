mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_program.py -x -q -m gpu -p no:cacheprovider --timeout 600 -k "teacher_forced" > gpurun_out/t_prog.log 2>&1; echo "pytest teacher rc=$?" >> gpurun_out/summary.txt
for mt in 296 148; do
  SLIDE_TC_PERSIST_MIN_TILES=$mt timeout 300 python tools/profile_records.py lat 256 auto > gpurun_out/s5_lat_mt$mt.txt 2>&1
  SLIDE_TC_PERSIST_MIN_TILES=$mt timeout 300 python tools/profile_records.py pos 256 auto > gpurun_out/s5_pos_mt$mt.txt 2>&1
  echo "min_tiles=$mt $(head -1 gpurun_out/s5_lat_mt$mt.txt)" >> gpurun_out/summary.txt
  echo "min_tiles=$mt $(head -1 gpurun_out/s5_pos_mt$mt.txt)" >> gpurun_out/summary.txt
done
cat gpurun_out/summary.txt; tail -n 6 gpurun_out/t_prog.log | cut -c1-400
