"""Data-parallel fit() timing (torchrun, one rank per GPU): BASELINE config 4 shapes (D=256, H=512,
batch 32 sharded over the ranks) through uisrnn.UISRNN.fit_concatenated.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/fit_dp_bench.py 200
"""
import os, sys, time, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.distributed as dist
import uisrnn
from uisrnn_b200 import utils
from uisrnn_b200.synth import synth_training_set
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rank = int(os.environ.get('LOCAL_RANK', 0)); world = int(os.environ.get('WORLD_SIZE', 1))
torch.cuda.set_device(rank)
if world > 1:
  dist.init_process_group('nccl', device_id=torch.device('cuda', rank))
np.random.seed(0); random.seed(0); torch.manual_seed(0)
m, t, _ = uisrnn.parse_arguments([])
m.verbosity = 0
t.batch_size = 32
seqs, ids = synth_training_set(2000, 200, n_frames=100, dim=256, n_spk=3)
x, y = utils.concatenate_training_data(seqs, ids, True, True)
y = np.array(y)
model = uisrnn.UISRNN(m)
real_resize = utils.resize_sequence
cache = {}
def cached_resize(**kw):          # time the iteration loop, not the host-side permutations
  if 'v' not in cache: cache['v'] = real_resize(**kw)
  return cache['v']
utils.resize_sequence = cached_resize
t.train_iteration = 10
model.fit_concatenated(x, y, t)
torch.cuda.synchronize()
if world > 1: dist.barrier()
t.train_iteration = iters
t0 = time.perf_counter()
model.fit_concatenated(x, y, t)
torch.cuda.synchronize()
if world > 1: dist.barrier()
dt = time.perf_counter() - t0
if rank == 0:
  print('world %d: %d iterations in %.3f s -> %.2f ms/iteration (final loss1 %.4f)' % (world, iters, dt, 1e3 * dt / iters, model.last_training_losses[-1]))
if world > 1: dist.destroy_process_group()
