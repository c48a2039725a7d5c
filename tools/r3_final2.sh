# last run of the round: GPU tests, smoke and bench line of the final tree
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r3g_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r3g_tests.log
timeout 600 python __graft_entry__.py --smoke > gpurun_out/r3g_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py > gpurun_out/r3g_bench.json 2> gpurun_out/r3g_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r3g_bench.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'e2e', round(d['e2e']['value']), d['e2e']['breakdown_ms_rank0_last_step'])
print({k: (v.get('frames_per_s') or v.get('ms_per_iteration')) for k, v in d['secondary'].items()})
PY
