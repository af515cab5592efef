mkdir -p gpurun_out
PROFILE_ONLY=net.SA1.mlp.conv2 timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/s8_conv2 python tools/profile_records.py lat 256 auto > gpurun_out/ncu_s8.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_s8.log
