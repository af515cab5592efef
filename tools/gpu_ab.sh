mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_program.py -m gpu -q --timeout 900 -p no:cacheprovider -k "golden or (teacher_forced and auto)" > gpurun_out/t_prog.log 2>&1; echo "prog rc=$?"; tail -n 3 gpurun_out/t_prog.log | cut -c1-300
for v in 1 0; do
    SLIDE_TC_TMA=$v python tools/profile_records.py lat 256 auto > gpurun_out/ab_lat_tma$v.txt 2>&1
    SLIDE_TC_TMA=$v python tools/profile_records.py pos 256 auto > gpurun_out/ab_pos_tma$v.txt 2>&1
    echo "tma=$v: $(head -1 gpurun_out/ab_lat_tma$v.txt)"
    echo "tma=$v: $(head -1 gpurun_out/ab_pos_tma$v.txt)"
    grep -E "SA1.att.v |SA1.att.k |SA1.mlp.res |SA1.mlp.conv0 |SA1.att.q " gpurun_out/ab_lat_tma$v.txt | cut -c1-100
done
