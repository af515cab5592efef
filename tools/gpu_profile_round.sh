# Produces the round's committed evidence: bench line, ncu launch list of the same command (short), ncu --set full of
# the dominant kernel, per-record timing tables, and (when the instrumented library was built: python tools/tc_timeline.py
# --build) the phase timeline of the tcgen05 GEMM CTAs.
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 > gpurun_out/bench_r1.log 2>&1
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_r1_reference.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r1.csv python tools/one_step.py > gpurun_out/ncu_launch.log 2>&1
python tools/profile_records.py lat 256 auto > gpurun_out/prof_lat_auto.txt 2>&1
python tools/profile_records.py pos 256 auto > gpurun_out/prof_pos_auto.txt 2>&1
PROFILE_ONLY=net.SA1.att.v,net.SA1.att.w2+softmax,net.SA1.mlp.res,net.SA1.mlp.conv1 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/prof_gemm_tc_r1 python tools/profile_records.py lat 256 auto > gpurun_out/ncu_full.log 2>&1
if [ -f slide_b200/libslide_b200_tl.so ]; then bash tools/gpu_timeline.sh > /dev/null 2>&1; fi
tail -n 1 gpurun_out/bench_r1.log | cut -c1-300
tail -n 1 gpurun_out/bench_r1_reference.log | cut -c1-300
