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


def small_cases():
  g = np.load(os.path.join(GOLDEN, 'small_cases.npz'))
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
