"""Synthetic d-vector utterances (BASELINE.md §3 / SURVEY.md §8d "Config 2" generator).

The statistics mirror the reference's toy data (`/root/reference/data/toy_testing_data.npz`):
unit-norm rows, near-orthogonal speaker centroids, within-speaker per-dim std ~0.043,
mean speaker run ~15 frames.
"""
import numpy as np


def synth_utt(seed, n_frames=500, dim=256, n_spk=4, mean_run=15, noise=0.059):
  """Returns (x float64 [n_frames, dim], labels int64 [n_frames])."""
  rng = np.random.default_rng(seed)
  centres = rng.standard_normal((n_spk, dim))
  centres /= np.linalg.norm(centres, axis=1, keepdims=True)
  labels = np.empty(n_frames, dtype=np.int64)
  pos = 0
  spk = int(rng.integers(n_spk))
  while pos < n_frames:
    run = 1 + int(rng.geometric(1.0 / mean_run))
    labels[pos:pos + run] = spk
    pos += run
    if n_spk > 1:
      nxt = int(rng.integers(n_spk - 1))
      spk = nxt if nxt < spk else nxt + 1
  x = centres[labels] + noise * rng.standard_normal((n_frames, dim))
  x /= np.linalg.norm(x, axis=1, keepdims=True)
  return x.astype(np.float64), labels


def synth_training_set(first_seed, n_utt, n_frames=100, dim=256, n_spk=3, **kw):
  """A list of utterances and string cluster ids (`"<u>_<spk>"`), SURVEY.md §8d config 4."""
  seqs, ids = [], []
  for u in range(n_utt):
    x, lab = synth_utt(first_seed + u, n_frames=n_frames, dim=dim, n_spk=n_spk, **kw)
    seqs.append(x)
    ids.append(np.array(['{}_{}'.format(u, int(s)) for s in lab]))
  return seqs, ids
