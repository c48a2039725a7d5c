set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r3_tests_all_2gpu.log 2>&1; echo "tests rc=$?"
tail -6 gpurun_out/r3_tests_all_2gpu.log
