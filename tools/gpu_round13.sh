mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 1500 python -m pytest tests/test_gpu_program.py tests/test_gpu_pipeline.py -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/t_prog.log 2>&1; echo "prog rc=$?" >> gpurun_out/summary.txt
for v in 1 0; do
    SLIDE_FACTOR_GROUP=$v python tools/profile_records.py lat 256 auto > gpurun_out/ab_lat_fac$v.txt 2>&1
    SLIDE_FACTOR_GROUP=$v python tools/profile_records.py pos 256 auto > gpurun_out/ab_pos_fac$v.txt 2>&1
    echo "factor=$v: $(head -1 gpurun_out/ab_lat_fac$v.txt)" >> gpurun_out/summary.txt
    echo "factor=$v: $(head -1 gpurun_out/ab_pos_fac$v.txt)" >> gpurun_out/summary.txt
done
timeout 900 python bench.py --steps 1 --warmup 1 > gpurun_out/bench_full_auto.log 2>&1; echo "bench rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -n 8 gpurun_out/t_prog.log | cut -c1-400
tail -n 1 gpurun_out/bench_full_auto.log | cut -c1-200
grep -E "PAIR|\.U " gpurun_out/ab_lat_fac1.txt | cut -c1-100
