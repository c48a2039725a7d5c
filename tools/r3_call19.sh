set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_hostpath.py tests/test_gpu_api.py -m gpu -q 2>&1 | tail -3
for t in 4 8 16; do
UISRNN_B200_COPY_THREADS=$t timeout 600 python bench.py --pageable --no-secondary --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/r3_bench_pageable_t$t.json 2>gpurun_out/r3_bench_pageable.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r3_bench_pageable_t$t.json").read().strip().splitlines()[-1])
print("threads $t pageable e2e", round(d["e2e"]["value"]), d["e2e"]["breakdown_ms_rank0_last_step"])
PY
done
timeout 600 python bench.py --no-secondary --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/r3_bench_pinned_check.json 2>/dev/null
python - <<PY
import json
d=json.loads(open("gpurun_out/r3_bench_pinned_check.json").read().strip().splitlines()[-1])
print("pinned e2e", round(d["e2e"]["value"]), d["e2e"]["breakdown_ms_rank0_last_step"])
PY
