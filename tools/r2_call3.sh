#!/bin/bash
mkdir -p gpurun_out
TC_BENCH_LANES=6 timeout 600 python tools/tc_bench.py 888 > gpurun_out/r2c3_tcbench48.txt 2>&1; echo "rc=$?" >> gpurun_out/r2c3_tcbench48.txt
UISRNN_B200_TC_N=32 TC_BENCH_LANES=4 timeout 600 python tools/tc_bench.py 592 > gpurun_out/r2c3_tcbench32.txt 2>&1; echo "rc=$?" >> gpurun_out/r2c3_tcbench32.txt
tail -3 gpurun_out/r2c3_tcbench48.txt; tail -3 gpurun_out/r2c3_tcbench32.txt
