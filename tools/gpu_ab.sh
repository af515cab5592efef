mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_program.py -m gpu -q --timeout 600 -p no:cacheprovider -k "golden or (teacher_forced and auto)" > gpurun_out/t_prog.log 2>&1; echo "prog rc=$?"; tail -n 3 gpurun_out/t_prog.log | cut -c1-300
for v in 1 0; do
    SLIDE_TC_PERSIST=$v timeout 300 python tools/profile_records.py lat 256 auto > gpurun_out/ab_lat_p$v.txt 2>&1
    SLIDE_TC_PERSIST=$v timeout 300 python tools/profile_records.py pos 256 auto > gpurun_out/ab_pos_p$v.txt 2>&1
    echo "persist=$v: $(head -1 gpurun_out/ab_lat_p$v.txt)"
    echo "persist=$v: $(head -1 gpurun_out/ab_pos_p$v.txt)"
    grep -E "att.v |att.q |mlp.conv0 |\.res " gpurun_out/ab_lat_p$v.txt | cut -c1-100
done
