mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 1500 python -m pytest tests/test_gpu_program.py -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/t_prog.log 2>&1; echo "prog rc=$?" >> gpurun_out/summary.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r1.csv python tools/one_step.py > gpurun_out/ncu_launch.log 2>&1; echo "launchlist rc=$?" >> gpurun_out/summary.txt
PROFILE_ONLY=net.SA1.att.v,net.SA1.att.w2+softmax,net.SA1.mlp.res,net.SA1.mlp.conv1 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/prof_gemm_tc_r1 python tools/profile_records.py lat 256 auto > gpurun_out/ncu_full.log 2>&1; echo "ncufull rc=$?" >> gpurun_out/summary.txt
timeout 1200 python tools/bench_autoencoder.py > gpurun_out/bench_autoencoder.log 2>&1; echo "ae rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -n 6 gpurun_out/t_prog.log | cut -c1-300; tail -n 4 gpurun_out/bench_autoencoder.log | cut -c1-400
