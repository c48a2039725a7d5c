"""Tensor-core engine: 6 lanes x 148 CTAs (888 utterances, kcap 16) against 7 lanes (1036 utterances, kcap 12)."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from uisrnn_b200 import native
from uisrnn_b200.synth import synth_utt
model = native.NativeModel(dict(np.load(os.path.join(ROOT, 'tests', 'golden', 'model_toy100.npz'))))
N = 500
U = 1036
x = torch.from_numpy(np.concatenate([synth_utt(100000 + u, n_frames=N)[0] for u in range(U)]).astype(np.float32)).cuda()
ref = None
for lanes, kcap, u in ((6, 16, 888), (6, 12, 888), (7, 12, 1036), (7, 12, 888)):
  off = np.arange(u + 1, dtype=np.int64) * N
  lab = torch.empty(u * N, dtype=torch.int32, device='cuda')
  try:
    for _ in range(2):
      model.predict_device(x.data_ptr(), off, lab.data_ptr(), engine=2, lanes=lanes, kcap=kcap)
      st = model.stats()
  except native.NativeError as err:
    print(json.dumps({'lanes': lanes, 'kcap': kcap, 'U': u, 'error': str(err)[:160]})); continue
  got = lab.cpu().numpy()[:888 * N].copy()
  if ref is None: ref = got
  print(json.dumps({'lanes_req': lanes, 'lanes': st['lanes'], 'kcap': kcap, 'U': u, 'ctas': st['ctas'], 'beam_ms': round(st['beam_ms'], 2),
                    'frames_per_s': round(u * N / ((st['beam_ms'] + st['prepass_ms']) / 1e3)), 'passes_per_step': round(st['weight_passes'] * st['lanes'] / max(1, st['beam_steps']), 3),
                    'same_labels_first_888': bool(np.array_equal(ref, got))}), flush=True)
