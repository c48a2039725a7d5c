# round 2, session 3, call 1: GPU tests of the new paths, fit probe, bench line
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_fit.py > gpurun_out/r3_tests_main.log 2>&1; echo "tests(main) rc=$?"
tail -5 gpurun_out/r3_tests_main.log
timeout 900 python -m pytest tests/test_gpu_fit.py -m gpu -q > gpurun_out/r3_tests_fit.log 2>&1; echo "tests(fit) rc=$?"
tail -25 gpurun_out/r3_tests_fit.log
timeout 600 python tools/fit_probe.py 60 > gpurun_out/r3_fit_probe.jsonl 2> gpurun_out/r3_fit_probe.err; echo "fit probe rc=$?"
cat gpurun_out/r3_fit_probe.jsonl; tail -3 gpurun_out/r3_fit_probe.err
timeout 900 python bench.py > gpurun_out/r3_bench.json 2> gpurun_out/r3_bench.err; echo "bench rc=$?"
tail -c 6000 gpurun_out/r3_bench.json; tail -5 gpurun_out/r3_bench.err
