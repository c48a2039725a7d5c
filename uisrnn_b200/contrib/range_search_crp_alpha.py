"""Grid search for `crp_alpha` maximising the ddCRP likelihood of the training label sequences
(mirrors `/root/reference/uisrnn/contrib/range_search_crp_alpha.py`; offline helper, not used by
fit()/predict()).

Labels look like `"<utterance>_<speaker>"`; consecutive labels with the same utterance prefix form
one sequence.  For a sequence with K speakers whose speaker changes happen at positions i, the
likelihood is  alpha^(K-1) / prod_i ( sum_{k != z_{i-1}} N_{k,i-1} + alpha )  (Eq. 8 of the paper),
N_{k,t} being the number of blocks of speaker k up to t.
"""
import numpy as np


def estimate_crp_alpha(train_cluster_id, search_range=1, search_step=0.01):
  """Returns the alpha on the grid `search_step * {1, 2, ...}` (< search_range) with the largest
  log-likelihood (first maximum wins); nan if the grid is empty."""
  best_alpha, best_value = np.nan, -np.inf
  for step in range(1, int(np.ceil(search_range / search_step))):
    alpha = step * search_step
    value = _get_cdf(train_cluster_id, alpha)
    if value > best_value:
      best_alpha, best_value = alpha, value
  return best_alpha


def _get_cdf(train_cluster_id, alpha):
  """Log-likelihood of all label sequences for one alpha."""
  total = 0
  for sequence in _get_cluster_id_single(train_cluster_id):
    total += np.log(_get_cdf_single(sequence, alpha))
  return total


def _get_cdf_single(cluster_id_single, alpha):
  """Likelihood of one normalised label sequence."""
  speakers_so_far = _get_k_t(cluster_id_single)
  blocks = _get_n_kt(cluster_id_single)
  numerator = alpha ** (len(set(cluster_id_single)) - 1)
  denominator = 1
  for i in range(1, len(cluster_id_single)):
    previous = cluster_id_single[i - 1]
    if cluster_id_single[i] != previous:
      others = sum(blocks[i - 1, k] for k in range(speakers_so_far[i - 1]) if k != previous)
      denominator *= others + alpha
  return numerator / denominator


def _get_k_t(cluster_id_single):
  """K_t: number of distinct speakers among the first t+1 labels."""
  seen, counts = set(), []
  for label in cluster_id_single:
    seen.add(label)
    counts.append(len(seen))
  return np.array(counts)


def _get_n_kt(cluster_id_single):
  """N_{k,t}: blocks of speaker k up to t (row 0 is all zeros, as in the reference)."""
  n_speakers = len(set(cluster_id_single))
  table = np.zeros((len(cluster_id_single), n_speakers))
  running = np.zeros(n_speakers)
  current = None
  for t, speaker in enumerate(cluster_id_single):
    if t == 0 or speaker != current:
      current = speaker
      running[speaker] += 1
    if t > 0:
      table[t] = running
  return table


def _get_cluster_id_single(train_cluster_id):
  """Yields the normalised label sequence of each utterance.  Like the reference, a sequence is
  emitted when the prefix changes or at the last element, and the slice excludes the element at
  which it is emitted (so the very last label is dropped)."""
  start, prefix = 0, train_cluster_id[0].split('_')[0]
  last = len(train_cluster_id) - 1
  for i, label in enumerate(train_cluster_id):
    current = label.split('_')[0]
    if current != prefix or i == last:
      yield _get_normalized_id(train_cluster_id[start:i])
      start, prefix = i, current


def _get_normalized_id(cluster_id_single):
  """Relabels speakers 0, 1, 2, ... in order of first appearance."""
  speakers = [int(label.split('_')[1]) for label in cluster_id_single]
  first_seen = {}
  for speaker in speakers:
    first_seen.setdefault(speaker, len(first_seen))
  return np.array([first_seen[speaker] for speaker in speakers])
