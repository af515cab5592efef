# Final-state check of the GPU suite minus the SAP / mesh files (those ran last against the shipped .so); bounded so that it
# fits the GPU minutes left.  Verbose log is flushed per test, so a cut-off run still shows how far it got.
mkdir -p gpurun_out
timeout 128 python -u -m pytest tests/test_gpu_pair.py tests/test_gpu_baseline_sizes.py tests/test_gpu_program.py \
  tests/test_gpu_knn_warp.py tests/test_gpu_chain.py tests/test_gpu_resident.py tests/test_gpu_overlap.py \
  tests/test_gpu_pipeline.py tests/test_gpu_rng.py tests/test_gpu_generation.py tests/test_gpu_index_ops.py \
  tests/test_dropin_golden.py tests/test_gpu_dropin_modules.py tests/test_gpu_reference_unmodified.py \
  -x -v -m gpu -p no:cacheprovider --durations=15 > gpurun_out/final_subset.log 2>&1
echo "rc=$?" >> gpurun_out/final_subset.log
grep -c PASSED gpurun_out/final_subset.log; tail -n 3 gpurun_out/final_subset.log | cut -c1-200
