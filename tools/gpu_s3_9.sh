mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_generation.py tests/test_gpu_pipeline.py -x -q -m gpu -p no:cacheprovider --timeout 600 > gpurun_out/t_gen.log 2>&1; echo "pytest generation rc=$?"; tail -n 15 gpurun_out/t_gen.log | cut -c1-300
