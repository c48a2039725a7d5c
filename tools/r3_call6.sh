set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tensorcore.py tests/test_gpu_pinned_paths.py -m gpu -q -x > gpurun_out/r3_tests6.log 2>&1; echo "tests rc=$?"
tail -4 gpurun_out/r3_tests6.log
TC_BENCH_ENGINES=1,2 TC_BENCH_LANES=0 timeout 600 python tools/tc_bench.py 888 > gpurun_out/r3_tc_bench6.jsonl 2>&1; echo "tcbench rc=$?"
cat gpurun_out/r3_tc_bench6.jsonl | cut -c1-1500
