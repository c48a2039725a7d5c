"""Shared by the CPU and GPU fit() trajectory tests: replays a case of tests/golden/fit_traj.npz (written by
oracle/make_golden.py from the unmodified reference's fit()) through this repo's UISRNN.fit()."""
import random

import numpy as np
import torch

from helpers import GOLDEN

CASES = ('d1_b16', 'd1_b48', 'd2_b16')


def run_case(name, enable_cuda):
  """Returns (losses [20, 3] of this repo's fit(), reference losses [20, 3], trained model, reference final weights)."""
  import uisrnn
  from uisrnn_b200 import loss_func
  from uisrnn_b200.synth import synth_training_set
  g = np.load(GOLDEN + '/fit_traj.npz')
  seed, depth, batch = (int(v) for v in g[name + '_args'])
  margs, targs, _ = uisrnn.parse_arguments([])
  margs.observation_dim, margs.rnn_hidden_size, margs.rnn_depth, margs.rnn_dropout = 64, 128, depth, 0.0 if depth > 1 else 0.2
  margs.enable_cuda, margs.verbosity = enable_cuda, 0
  targs.train_iteration, targs.learning_rate, targs.num_permutations, targs.batch_size = 20, 1e-3, 4, batch
  seqs, ids = synth_training_set(7100 + seed, 40, n_frames=60, dim=64, n_spk=3, noise=0.08)
  model = uisrnn.UISRNN(margs)
  init = {k[len(name) + 6:]: g[k] for k in g.files if k.startswith(name + '_init_')}
  sd = {'linear_mean1.weight': init['w1'], 'linear_mean1.bias': init['b1'],
        'linear_mean2.weight': init['w2'], 'linear_mean2.bias': init['b2']}
  for l in range(depth):
    for nm in ('weight_ih', 'weight_hh', 'bias_ih', 'bias_hh'):
      sd['gru.{}_l{}'.format(nm, l)] = init['{}_l{}'.format(nm, l)]
  model.rnn_model.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd.items()})
  with torch.no_grad():
    model.rnn_init_hidden.data.copy_(torch.from_numpy(np.array(init['h0'])))
    model.sigma2.data.copy_(torch.from_numpy(np.array(init['sigma2'])))
  # the torch (CPU / autograd) path logs nothing per iteration: record the three terms as the fixture did
  rec = {'l1': [], 'l2': [], 'l3': []}
  orig = (loss_func.weighted_mse_loss, loss_func.sigma2_prior_loss, loss_func.regularization_loss)

  def wrap(fn, key):
    def inner(*a, **k):
      v = fn(*a, **k)
      rec[key].append(float(v.detach()))
      return v
    return inner
  loss_func.weighted_mse_loss = wrap(orig[0], 'l1')
  loss_func.sigma2_prior_loss = wrap(orig[1], 'l2')
  loss_func.regularization_loss = wrap(orig[2], 'l3')
  try:
    np.random.seed(seed + 1000); random.seed(seed + 1000); torch.manual_seed(seed + 1000)
    model.fit(seqs, ids, targs)
  finally:
    loss_func.weighted_mse_loss, loss_func.sigma2_prior_loss, loss_func.regularization_loss = orig
  if rec['l1']:
    losses = np.array([rec['l1'], rec['l2'], rec['l3']]).T
  else:  # device trainer: the per-iteration losses live on the device and are collected by fit()
    losses = np.asarray(model.last_training_loss_terms)
  final = {k[len(name) + 7:]: g[k] for k in g.files if k.startswith(name + '_final_')}
  return losses, g[name + '_losses'], model, final
