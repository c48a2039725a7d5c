import os, sys, time, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from uisrnn_b200 import utils, native
from uisrnn_b200.synth import synth_training_set
np.random.seed(0); random.seed(0)
D, H = 256, 512
seqs, ids = synth_training_set(2000, 60, n_frames=100, dim=D, n_spk=3)
x, y = utils.concatenate_training_data(seqs, ids, True, True)
subs, lens = utils.resize_sequence(x, np.array(y), 10)
rng = np.random.default_rng(0)
params = {'gru.weight_ih_l0': rng.standard_normal((3*H, D))*0.05, 'gru.weight_hh_l0': rng.standard_normal((3*H, H))*0.05,
          'gru.bias_ih_l0': np.zeros(3*H), 'gru.bias_hh_l0': np.zeros(3*H), 'linear_mean1.weight': rng.standard_normal((H, H))*0.05,
          'linear_mean1.bias': np.zeros(H), 'linear_mean2.weight': rng.standard_normal((D, H))*0.05, 'linear_mean2.bias': np.zeros(D),
          'rnn_init_hidden': np.zeros(H), 'sigma2': np.full(D, 0.1)}
hp = {'learning_rate': 1e-3, 'sigma_alpha': 1.0, 'sigma_beta': 1.0, 'regularization_weight': 1e-5, 'grad_max_norm': 5.0, 'train_sigma2': True}
tr = native.NativeTrainer(params, hp)
tp = 0.0
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
  t0 = time.perf_counter()
  xi, li = utils.pack_batch(subs, lens, 32, D)
  xi = xi.astype(np.float32)
  tp += time.perf_counter() - t0
  print('L', xi.shape[0], 'rows', int(li.sum()), tr.step(xi, li))
print('host pack time per iter %.2f ms' % (1e3 * tp / 3))
