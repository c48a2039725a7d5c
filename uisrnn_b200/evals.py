"""Evaluation helper: accuracy of a predicted label sequence under the best one-to-one label
matching (Hungarian algorithm), as `/root/reference/uisrnn/evals.py:40-73`."""
import numpy as np
from scipy import optimize


def get_list_inverse_index(unique_ids):
  """Maps each element of a list of unique ids to its position (evals.py:21-37)."""
  if not isinstance(unique_ids, list):
    raise TypeError('unique_ids must be a list')
  return {value: position for position, value in enumerate(unique_ids)}


def compute_sequence_match_accuracy(sequence1, sequence2):
  """Fraction of positions that agree under the optimal matching of the two label sets.

  Raises TypeError for non-list inputs and ValueError for empty or unequal lengths, like the
  reference.  The co-occurrence matrix is accumulated with one vectorised scatter-add.
  """
  if not isinstance(sequence1, list) or not isinstance(sequence2, list):
    raise TypeError('sequence1 and sequence2 must be lists')
  if not sequence1 or len(sequence1) != len(sequence2):
    raise ValueError('sequence1 and sequence2 must have the same non-zero length')
  index1 = get_list_inverse_index(sorted(set(sequence1)))
  index2 = get_list_inverse_index(sorted(set(sequence2)))
  rows = np.fromiter((index1[s] for s in sequence1), dtype=np.int64, count=len(sequence1))
  cols = np.fromiter((index2[s] for s in sequence2), dtype=np.int64, count=len(sequence2))
  counts = np.zeros((len(index1), len(index2)))
  np.add.at(counts, (rows, cols), 1.0)
  matched_rows, matched_cols = optimize.linear_sum_assignment(-counts)
  return counts[matched_rows, matched_cols].sum() / len(sequence1)
