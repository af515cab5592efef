# session-3 call 2: parity of the edited kernels, per-record timings, ncu of small position-DDPM records
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_program.py tests/test_gpu_pipeline.py -x -q -m gpu -p no:cacheprovider --timeout 600 > gpurun_out/t_prog.log 2>&1; echo "pytest prog rc=$?" >> gpurun_out/summary.txt
timeout 300 python tools/profile_records.py lat 256 auto > gpurun_out/s3_lat.txt 2>&1
timeout 300 python tools/profile_records.py pos 256 auto > gpurun_out/s3_pos.txt 2>&1
head -1 gpurun_out/s3_lat.txt >> gpurun_out/summary.txt
head -1 gpurun_out/s3_pos.txt >> gpurun_out/summary.txt
PROFILE_ONLY=net.SA0.mlp.conv1,net.FP1.mlp2.conv1,net.SA0.att.q,net.SA1.att.w1k,net.SA0.group timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/s3_pos_small python tools/profile_records.py pos 256 auto > gpurun_out/ncu_pos_small.log 2>&1
echo "ncu rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -n 4 gpurun_out/t_prog.log | cut -c1-300
grep -E "PAIR|softmax|\.res " gpurun_out/s3_lat.txt | cut -c1-100
