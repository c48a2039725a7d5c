#!/bin/bash
# round 2, GPU call 2: first run of the tensor-core engine
mkdir -p gpurun_out
timeout 300 python - > gpurun_out/r2c2_sanity.txt 2>&1 <<'PY'
import numpy as np, sys, os
sys.path.insert(0, 'tests')
from helpers import load_weights, toy_utterances
from uisrnn_b200 import native
native.load_library()
m = native.NativeModel(load_weights('model_toy100.npz'))
xs, labs = toy_utterances()
a = m.predict(xs[:3], engine=1)
print('ffma ok', m.stats()['beam_ms'])
b = m.predict(xs[:3], engine=2)
st = m.stats()
print('tc stats', st)
for i, (u, v, w) in enumerate(zip(a, b, labs)):
  print(i, 'tc==ffma', u.tolist() == v.tolist(), 'tc==ref', v.tolist() == w.tolist(), 'mismatch', int((u != v).sum()), 'of', len(u))
PY
echo "sanity rc=$?" >> gpurun_out/r2c2_sanity.txt
tail -12 gpurun_out/r2c2_sanity.txt
timeout 900 python -m pytest tests/test_gpu_tensorcore.py -x -q > gpurun_out/r2c2_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c2_tests.log
tail -25 gpurun_out/r2c2_tests.log
timeout 600 python tools/tc_bench.py 296 888 > gpurun_out/r2c2_tcbench.txt 2>&1; echo "tcbench rc=$?" >> gpurun_out/r2c2_tcbench.txt
cat gpurun_out/r2c2_tcbench.txt | tail -12
