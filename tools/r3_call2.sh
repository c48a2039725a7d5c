set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_hostpath.py tests/test_gpu_fit.py tests/test_gpu_fit_dp.py -m gpu -q > gpurun_out/r3_tests2.log 2>&1; echo "tests rc=$?"
tail -8 gpurun_out/r3_tests2.log
timeout 600 python tools/fit_slow_probe.py > gpurun_out/r3_fit_slow.jsonl 2> gpurun_out/r3_fit_slow.err; echo "probe rc=$?"
cat gpurun_out/r3_fit_slow.jsonl; tail -3 gpurun_out/r3_fit_slow.err
TC_BENCH_ENGINES=2 TC_BENCH_LANES=0 timeout 600 python tools/tc_bench.py 888 > gpurun_out/r3_tc_bench.jsonl 2>&1; echo "tcbench rc=$?"
cat gpurun_out/r3_tc_bench.jsonl
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_fit.py --deselect tests/test_gpu_hostpath.py --deselect tests/test_gpu_fit_dp.py > gpurun_out/r3_tests_rest.log 2>&1; echo "tests(rest) rc=$?"
tail -5 gpurun_out/r3_tests_rest.log
