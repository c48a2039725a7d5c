"""One process per GPU (the launch shape of `bench.py --gpus N`): `predict_sharded` partitions the utterance list by frame
count, every rank decodes its shard through uis_predict on its own device, the labels travel as one int32 tensor per
rank over NCCL -- gathered to rank 0 or to every rank.  Needs two devices (skipped elsewhere)."""
import os
import socket
import sys

import numpy as np
import pytest

from helpers import ROOT, inference_args, load_weights, toy_utterances, uisrnn_from_weights

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out_dir):
  sys.path.insert(0, ROOT)
  sys.path.insert(0, os.path.join(ROOT, 'tests'))
  import torch
  import torch.distributed as dist
  from uisrnn_b200.distributed import my_shard, predict_sharded
  torch.cuda.set_device(rank)
  dist.init_process_group('nccl', init_method='tcp://127.0.0.1:%d' % port, rank=rank, world_size=world,
                          device_id=torch.device('cuda', rank))
  model = uisrnn_from_weights(load_weights('model_toy100.npz'), enable_cuda=True)
  assert model.device.index == rank
  xs, _ = toy_utterances()
  lengths = [len(x) for x in xs]
  own = set(my_shard(lengths))
  lazy = [xs[i] if i in own else None for i in range(len(xs))]          # a rank holds only its own shard
  everyone = predict_sharded(model, lazy, inference_args(), lengths=lengths)
  rooted = predict_sharded(model, lazy, inference_args(), lengths=lengths, root=0, as_arrays=True)
  if rank == 0:
    assert [r.tolist() for r in rooted] == everyone
  else:
    assert all((rooted[i] is not None) == (i in own) for i in range(len(xs)))
  np.save(os.path.join(out_dir, 'rank%d.npy' % rank), np.array(everyone, dtype=object), allow_pickle=True)
  dist.destroy_process_group()


def test_predict_sharded_nccl_world2(tmp_path):
  import torch
  import torch.multiprocessing as mp
  if torch.cuda.device_count() < 2:
    pytest.skip('needs 2 GPUs')
  with socket.socket() as s:
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
  mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
  _, labs = toy_utterances()
  for rank in (0, 1):
    got = np.load(str(tmp_path / ('rank%d.npy' % rank)), allow_pickle=True).tolist()
    assert [list(g) for g in got] == [l.tolist() for l in labs]     # the reference's labels, on every rank
