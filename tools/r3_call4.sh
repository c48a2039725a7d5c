# 2-GPU box: the tests that need two devices, the N = 2 bench line (partitioned list, gather to rank 0, strong-scaling
# list, data-parallel fit), NCCL topology line
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest "tests/test_gpu_api.py::test_parallel_predict_thread_branch_on_two_devices" "tests/test_gpu_api.py::test_parallel_predict_on_cuda" "tests/test_gpu_fit_dp.py" -m gpu -q > gpurun_out/r3_tests_2gpu.log 2>&1; echo "tests rc=$?"
tail -6 gpurun_out/r3_tests_2gpu.log
NCCL_DEBUG=INFO timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r3_bench_n2.json 2> gpurun_out/r3_bench_n2.err; echo "bench N=2 rc=$?"
grep -i "NVLS\|via P2P\|Channel 00/0\|Connected all" gpurun_out/r3_bench_n2.err | head -8
python - <<'PY'
import json
line = [l for l in open('gpurun_out/r3_bench_n2.json').read().strip().splitlines() if l.startswith('{')][-1]
d = json.loads(line)
print('N', d['n_gpus'], 'value', d['value'], 'e2e', d['e2e']['value'], d['e2e'].get('breakdown_ms_rank0_last_step'))
print('parity', d['parity'])
print(json.dumps(d.get('secondary'), indent=1)[:2500])
PY
tail -3 gpurun_out/r3_bench_n2.err
timeout 600 python bench.py --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/r3_bench_n1_same_box.json 2> gpurun_out/r3_bench_n1_same_box.err; echo "bench N=1 rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r3_bench_n1_same_box.json').read().strip().splitlines()[-1])
print('N', d['n_gpus'], 'value', d['value'], 'e2e', d['e2e']['value'])
s = d['secondary']
print({k: s[k] for k in ('strong_scaling_fixed_list', 'fit_data_parallel_batch32')})
PY
