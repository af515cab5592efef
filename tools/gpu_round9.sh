mkdir -p gpurun_out
PROFILE_ONLY=27,26 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/prof_gemm_tc_r1c python tools/profile_records.py lat 256 auto > gpurun_out/ncu_full.log 2>&1
SLIDE_TC_DEBUG=15 PROFILE_ONLY=27 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/prof_gemm_tc_r1c_dbg15 python tools/profile_records.py lat 256 auto > gpurun_out/ncu_full2.log 2>&1
tail -n 3 gpurun_out/ncu_full.log gpurun_out/ncu_full2.log
