#!/bin/bash
mkdir -p gpurun_out
for mode in 0 4 2; do
  UIS_DBG_MODE=$mode TC_BENCH_ENGINES=2 TC_BENCH_LANES=6 timeout 300 python tools/tc_bench.py 888 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
  try: d = json.loads(ln)
  except Exception: print(ln.strip()); continue
  print('mode $mode', {k: d[k] for k in ('beam_ms','frames_per_s','mismatching_frames','phase_us_per_cta_step','mma_issuer_us_per_pass')})
" >> gpurun_out/r2c5_modes.txt
done
cat gpurun_out/r2c5_modes.txt
