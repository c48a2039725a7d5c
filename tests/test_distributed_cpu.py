"""N > 1 path on CPU: utterance sharding over ranks (gloo, world_size 2) gives exactly the
single-process result; the partition is balanced and order-preserving."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

from helpers import ROOT, inference_args, load_weights, uisrnn_from_weights


def test_shard_by_frames_is_a_balanced_partition():
  from uisrnn_b200.uisrnn import shard_by_frames
  rng = np.random.default_rng(0)
  lengths = rng.integers(1, 500, size=101).tolist()
  for n in (1, 2, 3, 8):
    shards = shard_by_frames(lengths, n)
    assert sorted(i for s in shards for i in s) == list(range(101))
    loads = [sum(lengths[i] for i in s) for s in shards]
    assert max(loads) - min(loads) <= max(lengths)
    assert all(s == sorted(s) for s in shards)
  assert shard_by_frames([], 4) == [[], [], [], []]


def _worker(rank, world, port, out_dir):
  sys.path.insert(0, ROOT)
  sys.path.insert(0, os.path.join(ROOT, 'tests'))
  import torch.distributed as dist
  from uisrnn_b200.distributed import predict_sharded
  from uisrnn_b200.synth import synth_utt
  dist.init_process_group('gloo', init_method='tcp://127.0.0.1:%d' % port, rank=rank, world_size=world)
  model = uisrnn_from_weights(load_weights('model_small.npz'))
  seqs = [synth_utt(700 + i, n_frames=10 + 3 * i, dim=64, n_spk=2, noise=0.08)[0] for i in range(5)]
  merged = predict_sharded(model, seqs, inference_args(beam_size=4))
  # lazy form: a rank holds only its own shard, the others are None / produced on demand
  from uisrnn_b200.distributed import my_shard
  lengths = [len(s) for s in seqs]
  own = set(my_shard(lengths))
  assert 0 < len(own) < len(seqs)
  lazy = [(seqs[i] if i % 2 else (lambda i=i: seqs[i])) if i in own else None for i in range(len(seqs))]
  assert predict_sharded(model, lazy, inference_args(beam_size=4), lengths=lengths) == merged
  # root = 0: only rank 0 receives the merged result; the others keep their own shard in place.  as_arrays: int32 arrays
  rooted = predict_sharded(model, lazy, inference_args(beam_size=4), lengths=lengths, root=0, as_arrays=True)
  if rank == 0:
    assert [r.tolist() for r in rooted] == merged and all(r.dtype == np.int32 for r in rooted)
  else:
    assert all((rooted[i] is not None and rooted[i].tolist() == merged[i]) if i in own else rooted[i] is None
               for i in range(len(seqs)))
  np.save(os.path.join(out_dir, 'rank%d.npy' % rank), np.array(merged, dtype=object), allow_pickle=True)
  dist.destroy_process_group()


def test_predict_sharded_gloo_world2(tmp_path):
  with socket.socket() as s:
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
  mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
  from uisrnn_b200.synth import synth_utt
  model = uisrnn_from_weights(load_weights('model_small.npz'))
  seqs = [synth_utt(700 + i, n_frames=10 + 3 * i, dim=64, n_spk=2, noise=0.08)[0] for i in range(5)]
  want = model.predict(seqs, inference_args(beam_size=4))
  for rank in (0, 1):
    got = np.load(str(tmp_path / ('rank%d.npy' % rank)), allow_pickle=True).tolist()
    assert [list(g) for g in got] == want


def test_shard_columns_partition_keeps_length_order():
  from uisrnn_b200.uisrnn import shard_columns
  lengths = np.sort(np.random.default_rng(3).integers(2, 90, size=32))[::-1]
  for world in (1, 2, 3, 8, 40):
    parts = [shard_columns(32, r, world) for r in range(world)]
    assert sorted(int(c) for p in parts for c in p) == list(range(32))
    for p in parts:
      assert np.all(np.diff(lengths[p]) <= 0)          # every shard is still sorted by decreasing length
    sizes = [len(p) for p in parts]
    assert max(sizes) - min(sizes) <= 1
