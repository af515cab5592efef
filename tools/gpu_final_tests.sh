mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
python -m pytest tests -x -q -m gpu -p no:cacheprovider --timeout 900 > gpurun_out/t_all.log 2>&1; echo "pytest -m gpu rc=$?" >> gpurun_out/summary.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -n 5 gpurun_out/t_all.log | cut -c1-300; tail -n 1 gpurun_out/smoke.log
