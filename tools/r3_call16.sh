set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cluster.py tests/test_gpu_api.py -m gpu -q > gpurun_out/r3_tests16.log 2>&1; echo "tests rc=$?"
tail -8 gpurun_out/r3_tests16.log
