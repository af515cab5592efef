# session-3 first call: validate HEAD (full GPU suite + smoke), A/B the persistent GEMM, bench line
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 1200 python -m pytest tests -x -q -m gpu -p no:cacheprovider --timeout 900 > gpurun_out/t_all.log 2>&1; echo "pytest -m gpu rc=$?" >> gpurun_out/summary.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/summary.txt
for v in 1 0; do
    SLIDE_TC_PERSIST=$v timeout 300 python tools/profile_records.py lat 256 auto > gpurun_out/ab_lat_p$v.txt 2>&1
    SLIDE_TC_PERSIST=$v timeout 300 python tools/profile_records.py pos 256 auto > gpurun_out/ab_pos_p$v.txt 2>&1
    echo "persist=$v: $(head -1 gpurun_out/ab_lat_p$v.txt)" >> gpurun_out/summary.txt
    echo "persist=$v: $(head -1 gpurun_out/ab_pos_p$v.txt)" >> gpurun_out/summary.txt
done
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_s3.log 2>&1; echo "bench rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -n 5 gpurun_out/t_all.log | cut -c1-300; tail -n 1 gpurun_out/smoke.log
tail -n 1 gpurun_out/bench_s3.log | cut -c1-400
