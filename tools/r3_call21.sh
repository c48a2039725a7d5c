set -x
timeout 900 python -m pytest tests/test_gpu_fit.py tests/test_gpu_fit_dp.py -m gpu -q 2>&1 | tail -3
timeout 300 python tools/fit_probe.py 60 2>&1 | grep '"batch": 32, "barrier": "cg", "gemm_tiles": "128"\|"batch": 10\|"depth": 2' | cut -c1-260
