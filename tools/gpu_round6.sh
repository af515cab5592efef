mkdir -p gpurun_out
PROFILE_ONLY=27,24,21,44,19 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/prof_gemm_tc_r1b python tools/profile_records.py lat 256 auto > gpurun_out/ncu_full.log 2>&1
tail -n 8 gpurun_out/ncu_full.log
