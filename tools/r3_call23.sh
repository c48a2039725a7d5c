set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r3_bench_n8.json 2> gpurun_out/r3_bench_n8.err; echo "bench N=8 rc=$?"
python - <<'PY'
import json
line = [l for l in open('gpurun_out/r3_bench_n8.json').read().strip().splitlines() if l.startswith('{')][-1]
d = json.loads(line)
print('N', d['n_gpus'], 'value', round(d['value']), 'e2e', round(d['e2e']['value']), d['e2e'].get('breakdown_ms_rank0_last_step'))
print('parity', d['parity']['identical'], '/', d['parity']['reference_golden_utterances_checked'])
s = d.get('secondary', {})
print({k: (v.get('e2e_ms'), v.get('frames_per_s'), v.get('label_checksum'), v.get('ms_per_iteration'), v.get('error')) for k, v in s.items()})
PY
tail -3 gpurun_out/r3_bench_n8.err
