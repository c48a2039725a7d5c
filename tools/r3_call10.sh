set -x
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py --smoke > gpurun_out/r3_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r3_smoke.log | cut -c1-300
for one in 1 0; do
  UISRNN_B200_ONE_COPY_STREAM=$one timeout 600 python bench.py --no-secondary --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/r3_bench_copy$one.json 2> gpurun_out/r3_bench_copy$one.err; echo "bench rc=$?"
  python - <<PY
import json
d = json.loads(open('gpurun_out/r3_bench_copy$one.json').read().strip().splitlines()[-1])
print('one_copy_stream=$one', 'value', round(d['value']), 'e2e', round(d['e2e']['value']), d['e2e']['breakdown_ms_rank0_last_step'])
PY
done
timeout 300 python -m pytest tests/test_gpu_hostpath.py -m gpu -q 2>&1 | tail -2
