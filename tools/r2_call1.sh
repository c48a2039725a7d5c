#!/bin/bash
# round 2, GPU call 1: tcgen05 probes + the pin tests + a bench line of the round-1 kernel on this box
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2c1_gpu.txt 2>&1
nproc >> gpurun_out/r2c1_gpu.txt; lscpu | grep "Model name" >> gpurun_out/r2c1_gpu.txt
timeout 60 tools/next/tc_probe.bin > gpurun_out/r2_tc_probe_tf32.txt 2>&1; echo "tc_probe rc=$?" >> gpurun_out/r2_tc_probe_tf32.txt
timeout 180 tools/tc/tc_gemm_probe.bin > gpurun_out/r2_tc_gemm_probe.txt 2>&1; echo "tc_gemm_probe rc=$?" >> gpurun_out/r2_tc_gemm_probe.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c1_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c1_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2c1_bench.json 2> gpurun_out/r2c1_bench.err; echo "bench rc=$?" >> gpurun_out/r2c1_bench.err
tail -5 gpurun_out/r2_tc_probe_tf32.txt; tail -40 gpurun_out/r2_tc_gemm_probe.txt; tail -15 gpurun_out/r2c1_tests.log; tail -c 600 gpurun_out/r2c1_bench.json
