mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_program.py -m gpu -q --timeout 900 -p no:cacheprovider -k "golden and lat" > gpurun_out/t_prog.log 2>&1; echo "prog rc=$?" > gpurun_out/summary.txt
for d in 0 1 2 4 8 3 7 15; do
  SLIDE_TC_DEBUG=$d python tools/profile_records.py lat 256 auto > gpurun_out/prof_lat_dbg$d.txt 2>&1
done
cat gpurun_out/summary.txt
for d in 0 1 2 4 8 3 7 15; do echo "dbg=$d"; grep -E "^lat|SA1.att.v |SA1.att.w2 |SA1.mlp.res |SA1.att.w1q |SA0.mlp.conv1 |FP1.mlp2.conv1 " gpurun_out/prof_lat_dbg$d.txt | cut -c1-120; done
