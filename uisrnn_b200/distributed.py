"""Multi-process (one process per GPU) utterance sharding for predict().

The path shards naturally: every test utterance is an independent beam search (the reference
maps `predict_single` over the list, `/root/reference/uisrnn/uisrnn.py:587-589, 619-621`), so the
only cross-rank traffic is the gather of the label lists -- no data-path collective.  Works with
any `torch.distributed` backend (NCCL on the GPU box, gloo in the CPU tests).
"""
import numpy as np
import torch
import torch.distributed as dist

from .uisrnn import shard_by_frames


def predict_sharded(model, test_sequences, args, group=None, lengths=None, root=None, as_arrays=False):
  """Every rank passes the same list; rank r decodes the r-th shard (longest-first partition by
  frame count) with `model.predict`, and every rank returns the complete, ordered result.

  `lengths` (optional): the frame counts of ALL utterances.  With it a rank only needs to hold the
  utterances of its own shard -- `test_sequences[i]` may be None (or a zero-argument callable that
  produces the array) for every other i -- so a large list is never materialised on every rank
  (`my_shard(lengths)` tells a rank which entries it owns).

  The labels travel as ONE int32 tensor per rank (all_gather / gather over NCCL when the model lives on a
  CUDA device, gloo otherwise), not as pickled Python lists.  `root=r`: only rank r receives the merged
  result (what the reference's `parallel_predict` caller gets, uisrnn.py:619-623); the other ranks return
  their own shard in place and None elsewhere.  `as_arrays=True`: entries are numpy int32 arrays instead of
  lists of Python ints (building Python ints costs ~10 ns per label in one thread, which at several million
  frames per second per GPU is the slowest stage of an 8-GPU job)."""
  if not isinstance(test_sequences, list):
    raise TypeError('test_sequences must be a list.')
  if lengths is not None and len(lengths) != len(test_sequences):
    raise ValueError('lengths must have one entry per test sequence.')
  if not (dist.is_available() and dist.is_initialized()):
    if as_arrays:
      return _predict_arrays(model, [_materialise(s) for s in test_sequences], args)
    return model.predict([_materialise(s) for s in test_sequences], args)
  world = dist.get_world_size(group)
  rank = dist.get_rank(group)
  lengths = [len(s) for s in test_sequences] if lengths is None else [int(n) for n in lengths]
  shards = shard_by_frames(lengths, world)
  mine = _predict_arrays(model, [_materialise(test_sequences[i]) for i in shards[rank]], args) if shards[rank] else []
  counts = [sum(lengths[i] for i in shard) for shard in shards]
  use_cuda = getattr(model, 'device', None) is not None and model.device.type == 'cuda' and \
      dist.get_backend(group) == 'nccl'
  device = model.device if use_cuda else torch.device('cpu')
  flat = np.concatenate(mine).astype(np.int32, copy=False) if mine else np.zeros(0, np.int32)
  assert flat.size == counts[rank]
  width = max(max(counts), 1)
  send = torch.zeros(width, dtype=torch.int32, device=device)   # equal-sized pieces: pad to the largest shard
  send[:flat.size] = torch.from_numpy(flat).to(device)
  receives = rank == root or root is None
  if root is None:
    recv = [torch.empty(width, dtype=torch.int32, device=device) for _ in range(world)]
    dist.all_gather(recv, send, group=group)
  else:
    recv = [torch.empty(width, dtype=torch.int32, device=device) for _ in range(world)] if rank == root else None
    dist.gather(send, recv, dst=root, group=group)
  merged = [None] * len(test_sequences)
  if receives:
    for shard, piece, count in zip(shards, recv, counts):
      flat_r = piece[:count].cpu().numpy()
      pos = 0
      for i in shard:
        merged[i] = flat_r[pos:pos + lengths[i]]
        pos += lengths[i]
  else:
    for i, lab in zip(shards[rank], mine):
      merged[i] = lab
  if not as_arrays:
    merged = [m.tolist() if m is not None else None for m in merged]
  return merged


def _predict_arrays(model, sequences, args):
  """model.predict, but int32 arrays straight from the device when the model has the native path."""
  if getattr(model, 'device', None) is not None and model.device.type == 'cuda' and hasattr(model, '_predict_cuda'):
    from .uisrnn import _check_test_sequence
    for sequence in sequences:
      _check_test_sequence(sequence, model.observation_dim)  # the reference's TypeError / ValueError sites
    return model._predict_cuda(sequences, args, as_arrays=True)  # pylint: disable=protected-access
  return [np.asarray(o, dtype=np.int32) for o in model.predict(sequences, args)]


def my_shard(lengths, group=None):
  """Indices of the utterances this rank decodes in `predict_sharded(..., lengths=lengths)`."""
  if not (dist.is_available() and dist.is_initialized()):
    return list(range(len(lengths)))
  return shard_by_frames(list(lengths), dist.get_world_size(group))[dist.get_rank(group)]


def _materialise(entry):
  if callable(entry):
    entry = entry()
  if entry is None:
    raise ValueError('predict_sharded: an utterance of this rank\'s shard is missing (None).')
  return entry
