set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_cluster.py -m gpu -q -x > gpurun_out/r3_tests12.log 2>&1; echo "cluster tests rc=$?"
tail -12 gpurun_out/r3_tests12.log
timeout 300 python tools/latency_probe.py > gpurun_out/r3_latency.txt 2>&1; echo "latency rc=$?"
cat gpurun_out/r3_latency.txt | tail -30
timeout 120 python tools/prof_run.py 1 200 2 0 0 2>&1 | tail -3
