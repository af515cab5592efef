# A/B of a tcgen05-GEMM tuning knob on the per-record timings of one DDPM step (batch 256), e.g.
#   gpurun -- 'KNOB=SLIDE_TC_PERSIST VALUES="2 1 0" bash tools/gpu_ab.sh'
# knobs (read per launch by gemm_tc.cu): SLIDE_TC_PERSIST (0 never, 1 TMA-fed only, 2 all), SLIDE_TC_PERSIST_MIN_TILES,
# SLIDE_TC_PERSIST_MIN_K, SLIDE_TC_PERSIST_MIN_BN, SLIDE_TC_PREPASS, SLIDE_PAIR_MIN_CTAS, SLIDE_PAIR_MIN_ROWS, SLIDE_TC_TMA, SLIDE_SIDE_BRANCH, SLIDE_FACTOR_GROUP.
mkdir -p gpurun_out
for v in ${VALUES:-1 0}; do
    env ${KNOB:-SLIDE_TC_PERSIST}=$v timeout 300 python tools/profile_records.py lat 256 auto > gpurun_out/ab_lat_$v.txt 2>&1
    env ${KNOB:-SLIDE_TC_PERSIST}=$v timeout 300 python tools/profile_records.py pos 256 auto > gpurun_out/ab_pos_$v.txt 2>&1
    echo "${KNOB:-SLIDE_TC_PERSIST}=$v: $(head -1 gpurun_out/ab_lat_$v.txt)"
    echo "${KNOB:-SLIDE_TC_PERSIST}=$v: $(head -1 gpurun_out/ab_pos_$v.txt)"
done
