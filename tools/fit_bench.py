"""BASELINE config 4 probe: fit() on 50k concatenated synthetic frames, D=256, H=512, batch_size=32."""
import os, sys, time, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import uisrnn
from uisrnn_b200 import utils, native
from uisrnn_b200.synth import synth_training_set
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 50
np.random.seed(0); random.seed(0); torch.manual_seed(0)
m, t, _ = uisrnn.parse_arguments([])
m.verbosity = 0
t.batch_size, t.train_iteration = 32, iters
seqs, ids = synth_training_set(2000, 500, n_frames=100, dim=256, n_spk=3)
t0 = time.perf_counter()
x, y = utils.concatenate_training_data(seqs, ids, True, True)
subs, lens = utils.resize_sequence(x, np.array(y), t.num_permutations)
print('host prep (concatenate + resize_sequence) on %d frames: %.2f s, %d sub-sequences' % (len(x), time.perf_counter() - t0, len(subs)))
for mode in ('native', 'torch'):
  os.environ['UISRNN_B200_TORCH_FIT'] = '1' if mode == 'torch' else '0'
  np.random.seed(1); torch.manual_seed(1)
  model = uisrnn.UISRNN(m)
  t.train_iteration = 5
  model.fit_concatenated(x, np.array(y), t)      # warm-up (includes resize_sequence)
  torch.cuda.synchronize()
  # time the iteration loop only
  rows = 0
  if mode == 'native':
    state = {k: v.detach().cpu().numpy() for k, v in model.rnn_model.state_dict().items()}
    params = {k: state[k] for k in native.PARAM_ORDER[:8]}
    params['rnn_init_hidden'] = model.rnn_init_hidden.detach().cpu().numpy().reshape(-1)
    params['sigma2'] = model.sigma2.detach().cpu().numpy()
    hp = {'learning_rate': 1e-3, 'sigma_alpha': 1.0, 'sigma_beta': 1.0, 'regularization_weight': 1e-5, 'grad_max_norm': 5.0, 'train_sigma2': True}
    tr = native.NativeTrainer(params, hp)
    np.random.seed(1)
    index_lists, lens_i = utils.resize_indices(np.array(y), t.num_permutations)
    tr.set_corpus(x, index_lists)
    sampler = utils.BatchSampler(lens_i, 32)
    t0 = time.perf_counter()
    for _ in range(iters):
      chosen, li = sampler.draw()
      rows += int(li.sum())
      tr.step_corpus(chosen)
    tr.losses(1)
    dt = time.perf_counter() - t0
  else:
    t.train_iteration = iters
    t0 = time.perf_counter()
    model.fit_concatenated(x, np.array(y), t)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
  print('%s: %d iterations in %.3f s -> %.1f it/s (%.2f ms/it)%s' % (mode, iters, dt, iters / dt, 1e3 * dt / iters,
        (', %.0f packed rows/s' % (rows / dt)) if rows else ' (includes resize_sequence)'))
