mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_program.py -x -q -m gpu -p no:cacheprovider --timeout 600 > gpurun_out/t_prog.log 2>&1; echo "pytest prog rc=$?"; tail -n 4 gpurun_out/t_prog.log | cut -c1-300
KNOB=SLIDE_TC_PREPASS VALUES="1 0" bash tools/gpu_ab.sh
paste <(cut -c1-62 gpurun_out/ab_lat_1.txt) <(cut -c50-62 gpurun_out/ab_lat_0.txt) | grep "M=  4096"
