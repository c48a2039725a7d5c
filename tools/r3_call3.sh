set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fit.py tests/test_gpu_fit_dp.py -m gpu -q > gpurun_out/r3_tests3.log 2>&1; echo "tests rc=$?"
tail -4 gpurun_out/r3_tests3.log
timeout 600 python tools/fit_probe.py 60 > gpurun_out/r3_fit_probe2.jsonl 2> gpurun_out/r3_fit_probe2.err; echo "fit probe rc=$?"
cat gpurun_out/r3_fit_probe2.jsonl; tail -3 gpurun_out/r3_fit_probe2.err
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r3_bench2.json 2> gpurun_out/r3_bench2.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r3_bench2.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], d['e2e'].get('breakdown_ms_rank0_last_step'))
print(json.dumps(d.get('secondary'), indent=1)[:3000])
PY
tail -5 gpurun_out/r3_bench2.err
