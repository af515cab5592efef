mkdir -p gpurun_out
python tools/profile_records.py lat 256 auto > gpurun_out/prof_lat_auto.txt 2>&1
python tools/profile_records.py pos 256 auto > gpurun_out/prof_pos_auto.txt 2>&1
python tools/profile_records.py lat 256 simt > gpurun_out/prof_lat_simt.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 40 -c 3 -o gpurun_out/prof_gemm_tc_r1 python tools/profile_records.py lat 256 auto > gpurun_out/ncu_full.log 2>&1
head -3 gpurun_out/prof_lat_auto.txt gpurun_out/prof_pos_auto.txt gpurun_out/prof_lat_simt.txt
