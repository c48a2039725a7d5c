"""Shared helpers for the test-suite (loading golden fixtures, the oracle, comparisons)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import uis_oracle  # noqa: E402  (test infrastructure)


def load_weights(name):
  return dict(np.load(os.path.join(GOLDEN, name)))


def oracle_model(name):
  return uis_oracle.OracleModel(load_weights(name))


def toy_utterances():
  g = np.load(os.path.join(GOLDEN, 'toy_test.npz'))
  off = np.concatenate([[0], np.cumsum(g['lengths'])])
  xs = [g['x'][off[i]:off[i + 1]].astype(np.float64) for i in range(int(g['n_utt']))]
  labs = [g['labels'][off[i]:off[i + 1]] for i in range(int(g['n_utt']))]
  return xs, labs


def depth2_cases():
  return small_cases('depth2_cases.npz')


def small_cases(fixture='small_cases.npz'):
  g = np.load(os.path.join(GOLDEN, fixture))
  out = []
  for name in g['names']:
    name = str(name)
    b, la, t = [int(v) for v in g[name + '_args']]
    out.append(dict(name=name, x=g[name + '_x'].astype(np.float64), beam_size=b, look_ahead=la,
                    test_iteration=t,
                    **{k: g['{}_{}'.format(name, k)] for k in
                       ('labels', 'win', 'score', 'off', 'nfinite', 'final_scores', 'final_mean',
                        'final_hidden', 'final_blocks', 'full_trace')}))
  return out


def rel_err(a, b):
  a = np.asarray(a, dtype=np.float64)
  b = np.asarray(b, dtype=np.float64)
  return float(np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b)))) if a.size else 0.0


def compare_trace(win, score, off, gwin, gscore, goff, rtol=1e-5):
  """Compares two beam-search traces (per-step ranked winners + scores).

  Ranked score lists must agree within `rtol`; the hypotheses (identified by their full label
  history) must be the same and in the same order, EXCEPT that hypotheses whose float32 scores
  are within `rtol` of each other may swap ranks, or swap in/out at the beam cut-off (the
  reference ranks float32 scores with an unstable sort, uisrnn.py:546-549, so order inside a
  near-tie is not defined by the reference either).  Returns the number of near-tie swaps."""
  assert np.array_equal(off, goff), 'different beam sizes per step'
  assert rel_err(score, gscore) < rtol
  prev_a, prev_b = [()], [()]
  swaps = 0
  for s in range(len(off) - 1):
    lo, hi = int(off[s]), int(off[s + 1])
    a = [prev_a[int(win[r, 0])] + tuple(int(v) for v in win[r, 1:] if v >= 0) for r in range(lo, hi)]
    b = [prev_b[int(gwin[r, 0])] + tuple(int(v) for v in gwin[r, 1:] if v >= 0) for r in range(lo, hi)]
    if a != b:
      sb = {h: float(v) for h, v in zip(b, gscore[lo:hi])}
      cutoff = float(gscore[hi - 1])
      for r, h in enumerate(a):
        v = float(score[lo + r])
        tol = rtol * max(1.0, abs(v))
        if h in sb:
          assert abs(sb[h] - v) <= tol, 'step %d: same hypothesis, different score' % s
          if b[r] != h:  # rank moved: must be inside a near-tie
            assert abs(float(gscore[lo + r]) - sb[h]) <= tol, 'step %d: rank moved across a real gap' % s
            swaps += 1
        else:  # only allowed for a tie at the cut-off
          assert abs(v - cutoff) <= tol, 'step %d: different hypothesis set beyond a cut-off tie' % s
          swaps += 1
    prev_a, prev_b = a, b
  return swaps


def uisrnn_from_weights(weights, enable_cuda=False, verbosity=0):
  """Builds a `uisrnn.UISRNN` (this repo's drop-in package) holding golden-fixture weights."""
  import torch
  import uisrnn
  model_args, _, _ = uisrnn.parse_arguments([])
  model_args.observation_dim = int(weights['w2'].shape[0])
  model_args.rnn_hidden_size = int(weights['w1'].shape[0])
  model_args.rnn_depth = int(weights['depth'])
  model_args.enable_cuda = enable_cuda
  model_args.verbosity = verbosity
  model_args.transition_bias = float(weights['transition_bias'])
  model_args.crp_alpha = float(weights['crp_alpha'])
  model = uisrnn.UISRNN(model_args)
  sd = {'linear_mean1.weight': weights['w1'], 'linear_mean1.bias': weights['b1'],
        'linear_mean2.weight': weights['w2'], 'linear_mean2.bias': weights['b2']}
  for layer in range(model_args.rnn_depth):
    for name in ('weight_ih', 'weight_hh', 'bias_ih', 'bias_hh'):
      sd['gru.{}_l{}'.format(name, layer)] = weights['{}_l{}'.format(name, layer)]
  model.rnn_model.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd.items()})
  model.rnn_init_hidden = torch.nn.Parameter(torch.from_numpy(np.array(weights['h0'])).to(model.device))
  model.sigma2 = torch.nn.Parameter(torch.from_numpy(np.array(weights['sigma2'])).to(model.device))
  return model


def inference_args(beam_size=10, look_ahead=1, test_iteration=2):
  import uisrnn
  _, _, args = uisrnn.parse_arguments([])
  args.beam_size, args.look_ahead, args.test_iteration = beam_size, look_ahead, test_iteration
  return args
