"""Multi-process (one process per GPU) utterance sharding for predict().

The path shards naturally: every test utterance is an independent beam search (the reference
maps `predict_single` over the list, `/root/reference/uisrnn/uisrnn.py:587-589, 619-621`), so the
only cross-rank traffic is the gather of the label lists -- no data-path collective.  Works with
any `torch.distributed` backend (NCCL on the GPU box, gloo in the CPU tests).
"""
import torch.distributed as dist

from .uisrnn import shard_by_frames


def predict_sharded(model, test_sequences, args, group=None, lengths=None):
  """Every rank passes the same list; rank r decodes the r-th shard (longest-first partition by
  frame count) with `model.predict`, and every rank returns the complete, ordered result.

  `lengths` (optional): the frame counts of ALL utterances.  With it a rank only needs to hold the
  utterances of its own shard -- `test_sequences[i]` may be None (or a zero-argument callable that
  produces the array) for every other i -- so a large list is never materialised on every rank
  (`my_shard(lengths)` tells a rank which entries it owns)."""
  if not isinstance(test_sequences, list):
    raise TypeError('test_sequences must be a list.')
  if lengths is not None and len(lengths) != len(test_sequences):
    raise ValueError('lengths must have one entry per test sequence.')
  if not (dist.is_available() and dist.is_initialized()):
    return model.predict([_materialise(s) for s in test_sequences], args)
  world = dist.get_world_size(group)
  rank = dist.get_rank(group)
  shards = shard_by_frames([len(s) for s in test_sequences] if lengths is None else list(lengths), world)
  mine = model.predict([_materialise(test_sequences[i]) for i in shards[rank]], args) if shards[rank] else []
  gathered = [None] * world
  dist.all_gather_object(gathered, mine, group=group)
  merged = [None] * len(test_sequences)
  for shard, labels in zip(shards, gathered):
    for i, lab in zip(shard, labels):
      merged[i] = lab
  return merged


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
