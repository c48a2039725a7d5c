#!/usr/bin/env python3
"""bench.py -- UIS-RNN predict() throughput on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
  python bench.py --impl reference --gpus N --steps K ...  # reference CPU arm (host cores)

A "step" is one predict() pass over one batch of synthetic utterances (BASELINE config 2:
500-frame 256-d utterances, hidden 512, beam_size 10, look_ahead 1, test_iteration 2).
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the definitions.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

METRIC = 'predict_frames_per_sec_beam10_256d'
UNIT = 'frames/s'
N_FRAMES, DIM, HIDDEN, BEAM, LOOK_AHEAD, TEST_ITER = 500, 256, 512, 10, 1, 2
WORKLOAD = ('configs[1]: predict() synthetic 256-d d-vectors, 500-frame utterances, hidden=512, '
            'beam_size=10, look_ahead=1, test_iteration=2')
MODEL_FIXTURE = os.path.join(ROOT, 'tests', 'golden', 'model_toy100.npz')


def synth_batch(first_seed, n_utt, pinned=False):
  from uisrnn_b200.synth import synth_utt
  seqs = []
  for u in range(n_utt):
    x = synth_utt(first_seed + u, n_frames=N_FRAMES, dim=DIM)[0]
    if pinned:
      import torch
      t = torch.from_numpy(x).pin_memory()
      seqs.append((t.numpy(), t))  # keep the pinned tensor alive next to its numpy view
    else:
      seqs.append((x, None))
  return seqs


class ClockSampler:
  """Samples nvidia-smi SM clocks / throttle reasons while the timed region runs."""
  Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
       'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
       'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

  def __init__(self, gpu_index=0):
    self.gpu = gpu_index
    self.lines = []
    self.proc = None

  def start(self):
    try:
      self.proc = subprocess.Popen(
          ['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits',
           '-lms', '100'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
      self.thr = threading.Thread(target=self._pump, daemon=True)
      self.thr.start()
    except Exception:  # pylint: disable=broad-except
      self.proc = None

  def _pump(self):
    for line in self.proc.stdout:
      self.lines.append(line.strip())

  def stop(self):
    if not self.proc:
      return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
    self.proc.terminate()
    try:
      self.proc.wait(timeout=2)
    except Exception:  # pylint: disable=broad-except
      self.proc.kill()
    sm, smax, reasons, power = [], [], set(), []
    for ln in self.lines:
      f = [s.strip() for s in ln.split(',')]
      if len(f) < 9:
        continue
      try:
        sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
      except ValueError:
        continue
      for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[5:9]):
        if v.lower().startswith('active'):
          reasons.add(name)
    if not sm:
      return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no samples']}
    busy = [c for c, p in zip(sm, power) if p > 0.5 * max(power)] or sm
    return {'sm_mhz': float(np.median(busy)), 'sm_max_mhz': float(max(smax)), 'reasons': sorted(reasons),
            'samples': len(sm), 'power_w_max': float(max(power))}


def load_peaks():
  path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
  if os.path.exists(path):
    with open(path) as f:
      d = json.load(f)
    return float(d['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)', float(d.get('sm_max_mhz', 1965.0))
  return 6650.0, 'fallback (B200_PROFILING.md)', 1965.0


def measured_traffic(utts):
  """DRAM bytes per launch of the beam kernel from the committed ncu capture (profiles/r1_traffic.json);
  only valid for the workload it was captured on."""
  path = os.path.join(ROOT, 'profiles', 'r1_traffic.json')
  try:
    with open(path) as f:
      d = json.load(f)
    if d['utterances'] == utts and d['frames_per_utterance'] == N_FRAMES:
      return d['dram_bytes_read'] + d['dram_bytes_write']
  except Exception:  # pylint: disable=broad-except
    pass
  return None


def secondary_metrics(model, torch):
  """Best-effort extra numbers for the other BASELINE configs (never allowed to break the headline line)."""
  out = {}
  try:  # config 3: beam_size=30, look_ahead=2 (wide-beam stress), device-resident, 148 x 100 frames
    from uisrnn_b200.synth import synth_utt
    U3, N3 = 148, 100
    x3 = torch.from_numpy(np.concatenate([synth_utt(1000 + u, n_frames=N3, dim=DIM)[0] for u in range(U3)]).astype(np.float32)).cuda()
    lab3 = torch.empty(U3 * N3, dtype=torch.int32, device='cuda')
    off3 = np.arange(U3 + 1, dtype=np.int64) * N3
    for _ in range(2):
      model.predict_device(x3.data_ptr(), off3, lab3.data_ptr(), beam_size=30, look_ahead=2, test_iteration=TEST_ITER)
      st3 = model.stats()
    out['config3_beam30_lookahead2'] = {'frames_per_s': U3 * N3 / (st3['beam_ms'] / 1e3), 'kernel_ms': st3['beam_ms'],
                                        'gru_columns_per_step': st3['gru_columns'] / max(1, st3['beam_steps'])}
  except Exception as err:  # pylint: disable=broad-except
    out['config3_beam30_lookahead2'] = {'error': str(err)[:200]}
  try:  # latency mode (U=1) and a small batch (U=64), device-resident, same workload
    from uisrnn_b200.synth import synth_utt
    for U1 in (1, 64):
      xs = torch.from_numpy(np.concatenate([synth_utt(1000 + u, n_frames=N_FRAMES, dim=DIM)[0] for u in range(U1)]).astype(np.float32)).cuda()
      lab = torch.empty(U1 * N_FRAMES, dtype=torch.int32, device='cuda')
      offs = np.arange(U1 + 1, dtype=np.int64) * N_FRAMES
      for _ in range(2):
        model.predict_device(xs.data_ptr(), offs, lab.data_ptr(), beam_size=BEAM, look_ahead=LOOK_AHEAD, test_iteration=TEST_ITER)
        stu = model.stats()
      out['config2_U%d' % U1] = {'frames_per_s': U1 * N_FRAMES / ((stu['beam_ms'] + stu['prepass_ms']) / 1e3),
                                 'ms': stu['beam_ms'] + stu['prepass_ms'], 'ctas': stu['ctas'], 'lanes': stu['lanes'],
                                 'cluster': stu['cluster']}   # CTAs per utterance (latency mode)
  except Exception as err:  # pylint: disable=broad-except
    out['config2_small_batches'] = {'error': str(err)[:200]}
  try:  # config 4: fit() iteration on 50k concatenated frames, batch_size=32 (device trainer, csrc/uis_train.cu)
    import random
    from uisrnn_b200 import native, utils
    from uisrnn_b200.synth import synth_training_set
    np.random.seed(0); random.seed(0)
    seqs, ids = synth_training_set(2000, 500, n_frames=100, dim=DIM, n_spk=3)
    xcat, ycat = utils.concatenate_training_data(seqs, ids, True, True)
    index_lists, lens = utils.resize_indices(np.array(ycat), 10)
    w = dict(np.load(MODEL_FIXTURE))
    params = {'gru.weight_ih_l0': w['weight_ih_l0'], 'gru.weight_hh_l0': w['weight_hh_l0'], 'gru.bias_ih_l0': w['bias_ih_l0'],
              'gru.bias_hh_l0': w['bias_hh_l0'], 'linear_mean1.weight': w['w1'], 'linear_mean1.bias': w['b1'],
              'linear_mean2.weight': w['w2'], 'linear_mean2.bias': w['b2'], 'rnn_init_hidden': w['h0'].reshape(-1),
              'sigma2': w['sigma2']}
    hp = {'learning_rate': 1e-3, 'sigma_alpha': 1.0, 'sigma_beta': 1.0, 'regularization_weight': 1e-5,
          'grad_max_norm': 5.0, 'train_sigma2': True}
    tr = native.NativeTrainer(params, hp)
    tr.set_corpus(xcat, index_lists)             # as UISRNN.fit does: training set resident on the device
    sampler = utils.BatchSampler(lens, 32)
    iters, rows = 100, 0
    for i in range(5 + iters):
      if i == 5:
        tr.losses(1); t0 = time.perf_counter()
      chosen, li = sampler.draw()
      if i >= 5:
        rows += int(li.sum())
      tr.step_corpus(chosen)                     # asynchronous; the batch is gathered on the device
    tr.losses(1)                                 # synchronises
    dt = time.perf_counter() - t0
    out['config4_fit_batch32'] = {'ms_per_iteration': 1e3 * dt / iters, 'packed_rows_per_s': rows / dt,
                                  'includes': 'batch draw (host RNG) + device gather + forward/backward/clip/Adam kernels'}
    tr.close()
  except Exception as err:  # pylint: disable=broad-except
    out['config4_fit_batch32'] = {'error': str(err)[:200]}
  return out


def cpu_baseline_port(n_utts=3):
  """Times the CPU oracle port (numpy, 1 thread of control) on a bounded sample."""
  sys.path.insert(0, os.path.join(ROOT, 'oracle'))
  import uis_oracle  # bench.py's cpu_baseline leg is allowed to run the oracle (as a baseline)
  om = uis_oracle.OracleModel.load(MODEL_FIXTURE)
  seqs = [s for s, _ in synth_batch(1000, n_utts)]
  t0 = time.perf_counter()
  for s in seqs:
    uis_oracle.predict_single(om, s, beam_size=BEAM, look_ahead=LOOK_AHEAD, test_iteration=TEST_ITER)
  dt = time.perf_counter() - t0
  return {'value': n_utts * N_FRAMES / dt, 'unit': UNIT, 'cores': 1, 'kind': 'port',
          'sample': '%d utterances x %d frames of the same workload (seeds 1000..), oracle/uis_oracle.py, '
                    '%.1f s' % (n_utts, N_FRAMES, dt)}


# --------------------------------------------------------------------------- reference arm

def _ref_worker(job):
  kind, weights_path, seed, n_frames = job
  import torch
  torch.set_num_threads(1)
  from uisrnn_b200.synth import synth_utt
  x = synth_utt(seed, n_frames=N_FRAMES, dim=DIM)[0][:n_frames]
  t0 = time.perf_counter()
  if kind == 'reference':
    sys.path[:0] = [os.path.join(ROOT, 'oracle', 'shims'), os.path.join(ROOT, 'baseline', '_ref')]
    import uisrnn as ref
    assert 'baseline' in ref.__file__
    argv, sys.argv = sys.argv, [sys.argv[0]]
    try:
      margs, _, iargs = ref.parse_arguments()
    finally:
      sys.argv = argv
    margs.enable_cuda = False
    margs.verbosity = 0
    w = dict(np.load(weights_path))
    margs.transition_bias = float(w['transition_bias'])
    margs.crp_alpha = float(w['crp_alpha'])
    model = ref.UISRNN(margs)
    sd = {'gru.weight_ih_l0': w['weight_ih_l0'], 'gru.weight_hh_l0': w['weight_hh_l0'],
          'gru.bias_ih_l0': w['bias_ih_l0'], 'gru.bias_hh_l0': w['bias_hh_l0'],
          'linear_mean1.weight': w['w1'], 'linear_mean1.bias': w['b1'],
          'linear_mean2.weight': w['w2'], 'linear_mean2.bias': w['b2']}
    model.rnn_model.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd.items()})
    model.rnn_init_hidden = torch.nn.Parameter(torch.from_numpy(np.array(w['h0'])))
    model.sigma2 = torch.nn.Parameter(torch.from_numpy(np.array(w['sigma2'])))
    t0 = time.perf_counter()
    model.predict(x, iargs)   # the reference's own public API, stock code path
  else:
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import uis_oracle
    om = uis_oracle.OracleModel.load(weights_path)
    t0 = time.perf_counter()
    uis_oracle.predict_single(om, x, beam_size=BEAM, look_ahead=LOOK_AHEAD, test_iteration=TEST_ITER)
  return time.perf_counter() - t0


def run_reference(args):
  rank = int(os.environ.get('RANK', '0'))
  if rank != 0:
    return
  import multiprocessing as mp
  have_ref = os.path.exists(os.path.join(ROOT, 'baseline', '_ref', 'uisrnn', 'uisrnn.py'))
  kind = 'reference' if have_ref else 'port'
  cores = os.cpu_count() or 1
  procs = max(1, min(cores, 128))
  # bounded sample: one utterance slice per process and step, sized for ~5-10 s per step
  n_frames = 24 if kind == 'reference' else N_FRAMES
  ctx = mp.get_context('spawn')
  times = []
  with ctx.Pool(procs) as pool:
    for step in range(args.warmup + args.steps):
      jobs = [(kind, MODEL_FIXTURE, 1000 + step * procs + i, n_frames) for i in range(procs)]
      t0 = time.perf_counter()
      pool.map(_ref_worker, jobs, chunksize=1)
      dt = time.perf_counter() - t0
      if step >= args.warmup:
        times.append(dt)
  frames = procs * n_frames
  total = sum(times)
  value = frames * len(times) / total
  sample = ('%d processes x 1 utterance slice of %d frames per step (same generator/seeds family as the GPU arm), '
            '%s, OMP threads=1 per process' % (procs, n_frames,
                                              'unmodified reference predict() from baseline/_ref' if have_ref
                                              else 'oracle/uis_oracle.py port'))
  out = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
         'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * total / len(times),
         'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
         'config': {'workload': WORKLOAD, 'sample_frames_per_step': frames},
         'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': procs, 'kind': kind, 'sample': sample},
         'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
  print(json.dumps(out), flush=True)


# --------------------------------------------------------------------------- this repo's arm

def run_b200(args):
  import torch
  import torch.distributed as dist
  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  local = int(os.environ.get('LOCAL_RANK', '0'))
  if not torch.cuda.is_available():
    raise SystemExit('bench.py: no CUDA device; the sm_100a path cannot run (no CPU fallback by design)')
  torch.cuda.set_device(local)
  if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
  import __graft_entry__ as ge
  ge.build()
  from uisrnn_b200 import native

  U = args.utts           # utterances per GPU per step (weak scaling: fixed per GPU)
  weights = dict(np.load(MODEL_FIXTURE))
  model = native.NativeModel(weights, device=local)
  batch = synth_batch(100000 + rank * U, U, pinned=True)
  seqs = [s for s, _ in batch]
  frames = U * N_FRAMES
  stream = torch.cuda.current_stream().cuda_stream

  # ---- device-resident leg (`value`): fp32 inputs already in HBM
  x_dev = torch.from_numpy(np.concatenate(seqs).astype(np.float32)).cuda()
  labels_dev = torch.empty(frames, dtype=torch.int32, device='cuda')
  off = np.arange(U + 1, dtype=np.int64) * N_FRAMES

  def step_dev():
    model.predict_device(x_dev.data_ptr(), off, labels_dev.data_ptr(), beam_size=BEAM, look_ahead=LOOK_AHEAD,
                         test_iteration=TEST_ITER, stream=stream)

  def barrier():
    torch.cuda.synchronize()
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  for _ in range(args.warmup):
    step_dev()
  barrier()
  sampler = ClockSampler(local)
  if rank == 0:
    sampler.start()
  ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  beam_ms, prepass_ms = [], []
  ev0.record()
  for _ in range(args.steps):
    step_dev()
    # stats() synchronises on the step: device-side counters + per-kernel CUDA-event times
    st = model.stats()
    beam_ms.append(st['beam_ms']); prepass_ms.append(st['prepass_ms'])
  ev1.record()
  barrier()
  dev_ms = ev0.elapsed_time(ev1)
  labels_first = labels_dev.cpu().numpy().copy()

  # ---- end-to-end leg (`e2e`): the public API a user calls -- uisrnn.UISRNN.predict(list of host
  #      float64 arrays) -> list of label lists.  Pinned host inputs; H2D, cast, GEMM, beam search,
  #      D2H and the Python list conversion are all inside the timed region.
  sys.path.insert(0, os.path.join(ROOT, 'tests'))
  import uisrnn
  margs, _, iargs = uisrnn.parse_arguments([])
  margs.verbosity, margs.transition_bias, margs.crp_alpha = 0, float(weights['transition_bias']), float(weights['crp_alpha'])
  api_model = uisrnn.UISRNN(margs)
  assert api_model.device.type == 'cuda'
  if local != 0:
    api_model.device = torch.device('cuda', local)
  sd = {'gru.weight_ih_l0': weights['weight_ih_l0'], 'gru.weight_hh_l0': weights['weight_hh_l0'],
        'gru.bias_ih_l0': weights['bias_ih_l0'], 'gru.bias_hh_l0': weights['bias_hh_l0'],
        'linear_mean1.weight': weights['w1'], 'linear_mean1.bias': weights['b1'],
        'linear_mean2.weight': weights['w2'], 'linear_mean2.bias': weights['b2']}
  api_model.rnn_model.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd.items()})
  api_model.rnn_init_hidden = torch.nn.Parameter(torch.from_numpy(np.array(weights['h0'])))
  api_model.sigma2 = torch.nn.Parameter(torch.from_numpy(np.array(weights['sigma2'])))
  iargs.beam_size, iargs.look_ahead, iargs.test_iteration = BEAM, LOOK_AHEAD, TEST_ITER

  def step_e2e():
    return api_model.predict(seqs, iargs)

  for _ in range(max(1, args.warmup // 2)):
    out = step_e2e()
  barrier()
  t0 = time.perf_counter()
  for _ in range(args.steps):
    out = step_e2e()
  torch.cuda.synchronize()
  e2e_s = time.perf_counter() - t0
  clocks = sampler.stop() if rank == 0 else None
  assert np.array_equal(np.concatenate([np.asarray(o, dtype=np.int32) for o in out]), labels_first), \
      'e2e and device-resident legs disagree'

  t = torch.tensor([dev_ms, e2e_s * 1e3], dtype=torch.float64, device='cuda')
  if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
  dev_ms, e2e_ms = float(t[0]), float(t[1])
  if rank != 0:
    if world > 1:
      dist.destroy_process_group()
    return

  total_frames = frames * world
  value = total_frames * args.steps / (dev_ms / 1e3)
  e2e_value = total_frames * args.steps / (e2e_ms / 1e3)
  H, D = HIDDEN, DIM
  wbytes_pass = 4 * (3 * H * H + H * H + H * D)             # W_hh, W1, W2 streamed once per pass
  io_bytes = frames * (4 * D + 4 * 3 * H * (1 + TEST_ITER) + 4)   # x read, gi write + T reads, labels
  alg_bytes = st['weight_passes'] * wbytes_pass + io_bytes
  beam_avg_ms = float(np.mean(beam_ms))
  peak, peak_src, sm_max = load_peaks()
  achieved = alg_bytes / (beam_avg_ms / 1e3) / 1e9
  flops = st['gru_columns'] * 2.0 * (3 * H * H + H * H + H * D)
  sm_mhz = (clocks or {}).get('sm_mhz') or sm_max
  fp32_peak = st['ctas'] * 128 * 2 * sm_mhz * 1e6 / 1e12
  out = {
      'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
      'ms_per_step': dev_ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
      'dtype': 'f32', 'data': 'synthetic',
      'config': {'workload': WORKLOAD, 'utterances_per_gpu_per_step': U, 'frames_per_gpu_per_step': frames,
                 'model': 'D=256 H=512 depth=1, weights = reference fit() 100 it on toy data (tests/golden/model_toy100.npz)',
                 'parallelism': 'utterance-sharded x%d, no collective' % world,
                 'l2': 'inputs larger than L2: x %.0f MB + gi %.0f MB rewritten every step' % (
                     frames * D * 4 / 1e6, frames * 3 * H * 4 / 1e6)},
      'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': frames * D * 8, 'd2h_bytes_per_step': frames * 4,
              'path': 'uisrnn.UISRNN.predict(list of pinned float64 ndarrays) -> uis_predict() C ABI: H2D, cast+GEMM+beam kernels, D2H int32 labels -> Python lists'},
      'gpu_launches': int(args.steps * 2),
      'clocks': clocks,
      'roofline': {'bound': 'hbm', 'kernel': 'uis_beam_kernel<512,256>', 'achieved': achieved, 'peak': peak,
                   'unit': 'GB/s', 'frac': achieved / peak, 'traffic': measured_traffic(U), 'peak_source': peak_src,
                   'traffic_source': 'profiles/r1_traffic.json (ncu dram__bytes_read+write, same workload)',
                   'algorithmic_bytes_per_launch': alg_bytes, 'kernel_ms': beam_avg_ms,
                   'note': 'algorithmic bytes = SURVEY 8(d): weights streamed once per beam-step pass (4.72 MB, L2-resident, '
                           'so this stream never reaches DRAM: traffic << algorithmic) + per-frame HBM I/O; the binding '
                           'unit is the fp32 FMA pipe, see fp32',
                   'fp32': {'achieved_tflops': flops / (beam_avg_ms / 1e3) / 1e12, 'peak_tflops': fp32_peak,
                            'frac': flops / (beam_avg_ms / 1e3) / 1e12 / fp32_peak, 'sm_mhz_used': sm_mhz}},
      'kernel_stats': {k: st[k] for k in ('beam_steps', 'gru_columns', 'weight_passes', 'candidates', 'max_k', 'ctas')},
      'prepass_ms': float(np.mean(prepass_ms)),
  }
  if world == 1 and not args.no_secondary:
    out['secondary'] = secondary_metrics(model, torch)
  if world == 1 and not args.no_cpu_baseline:  # reported at N = 1 only (a bounded CPU sample, ~35 s)
    out['cpu_baseline'] = cpu_baseline_port()
  print(json.dumps(out), flush=True)
  if world > 1:
    dist.destroy_process_group()


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=5)
  ap.add_argument('--warmup', type=int, default=3)
  ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
  ap.add_argument('--utts', type=int, default=296, help='utterances per GPU per step')
  ap.add_argument('--no-cpu-baseline', action='store_true')
  ap.add_argument('--no-secondary', action='store_true', help='skip the config-3 / config-4 side measurements')
  args = ap.parse_args()
  if args.impl == 'reference':
    run_reference(args)
  else:
    run_b200(args)


if __name__ == '__main__':
  main()
