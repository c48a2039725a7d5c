set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r3_bench_n4.json 2> gpurun_out/r3_bench_n4.err; echo "bench N=4 rc=$?"
python - <<'PY'
import json
line = [l for l in open('gpurun_out/r3_bench_n4.json').read().strip().splitlines() if l.startswith('{')][-1]
d = json.loads(line)
print('N', d['n_gpus'], 'value', round(d['value']), 'e2e', round(d['e2e']['value']), d['e2e'].get('breakdown_ms_rank0_last_step'))
print('parity', d['parity']['identical'], '/', d['parity']['reference_golden_utterances_checked'])
s = d.get('secondary', {})
print({k: (v.get('e2e_ms'), v.get('frames_per_s'), v.get('label_checksum'), v.get('ms_per_iteration'), v.get('error')) for k, v in s.items()})
PY
tail -3 gpurun_out/r3_bench_n4.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --impl reference --gpus 4 --steps 1 --warmup 0 > gpurun_out/r3_ref_n4.json 2> gpurun_out/r3_ref_n4.err; echo "ref arm under torchrun rc=$?"
tail -c 600 gpurun_out/r3_ref_n4.json
