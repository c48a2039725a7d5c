"""The UIS-RNN model class with the public surface of `/root/reference/uisrnn/uisrnn.py`
(`UISRNN(args)`, `.fit`, `.fit_concatenated`, `.predict`, `.predict_single`, `.save`, `.load`,
`parallel_predict`, and the module-level `CoreRNN` / `BeamState`).

Inference on a CUDA device runs the whole beam search inside libuisrnn_b200.so (hand-written
sm_100a kernels behind the C ABI in include/uisrnn_b200.h): `predict(list)` hands all utterances
to ONE native call, which shards them over persistent CTAs; nothing but the labels comes back.
There is no silent fallback: on a CUDA device an unsupported configuration raises.  On the CPU
device (explicit `--enable_cuda=False`, the reference's own device rule) the decoder in
`beam_cpu.py` is used.  Training uses PyTorch autograd on the model's device.
"""
import functools
import threading

import numpy as np
import torch
from torch import multiprocessing
from torch import nn
from torch import optim
import torch.nn.functional as F

from . import beam_cpu
from . import logger as logger_lib
from . import loss_func
from . import utils

_INITIAL_SIGMA2_VALUE = 0.1
_DEFAULT_KCAP = 0  # clusters per hypothesis held in device tables: 0 = the library's default (16 or 32, by kernel); grown on overflow


class CoreRNN(nn.Module):
  """GRU (+ dropout between layers when depth >= 2) followed by a two-layer MLP that predicts
  the mean of the next observation (uisrnn.py:32-52 of the reference; same parameter names, so
  `state_dict()`s are interchangeable)."""

  def __init__(self, input_dim, hidden_size, depth, observation_dim, dropout=0):
    super().__init__()
    self.hidden_size = hidden_size
    gru_kwargs = {'dropout': dropout} if depth >= 2 else {}
    self.gru = nn.GRU(input_dim, hidden_size, depth, **gru_kwargs)
    self.linear_mean1 = nn.Linear(hidden_size, hidden_size)
    self.linear_mean2 = nn.Linear(hidden_size, observation_dim)

  def forward(self, input_seq, hidden=None):
    output_seq, hidden = self.gru(input_seq, hidden)
    if isinstance(output_seq, nn.utils.rnn.PackedSequence):
      output_seq, _ = nn.utils.rnn.pad_packed_sequence(output_seq, batch_first=False)
    return self.linear_mean2(F.relu(self.linear_mean1(output_seq))), hidden


class BeamState:
  """Plain record of one beam-search hypothesis (uisrnn.py:55-77).  Kept for API compatibility;
  the CUDA path keeps the equivalent state on the device (slot pool + per-hypothesis tables)."""

  def __init__(self, source=None):
    if not source:
      self.mean_set, self.hidden_set, self.trace, self.block_counts = [], [], [], []
      self.neg_likelihood = 0
    else:
      self.mean_set = source.mean_set.copy()
      self.hidden_set = source.hidden_set.copy()
      self.trace = source.trace.copy()
      self.block_counts = source.block_counts.copy()
      self.neg_likelihood = source.neg_likelihood

  def append(self, mean, hidden, cluster):
    self.mean_set.append(mean.clone())
    self.hidden_set.append(hidden.clone())
    self.block_counts.append(1)
    self.trace.append(cluster)


def _check_test_sequence(test_sequence, observation_dim):
  """Input validation of predict_single (uisrnn.py:511-521): same exceptions, same order."""
  if not isinstance(test_sequence, np.ndarray) or test_sequence.dtype != float:
    raise TypeError('test_sequence should be a numpy array of float type.')
  if test_sequence.ndim != 2:
    raise ValueError('test_sequence must be 2-dim array.')
  if test_sequence.shape[1] != observation_dim:
    raise ValueError('test_sequence does not match the dimension specified by args.observation_dim.')


class UISRNN:
  """Unbounded Interleaved-State Recurrent Neural Network."""

  def __init__(self, args):
    self.observation_dim = args.observation_dim
    # uisrnn.py:97-98 pins 'cuda:0'; here it is the process's CURRENT device (0 unless the caller ran
    # torch.cuda.set_device, as one-process-per-GPU launches do)
    if torch.cuda.is_available() and args.enable_cuda:
      self.device = torch.device('cuda', torch.cuda.current_device())
    else:
      self.device = torch.device('cpu')
    self.rnn_model = CoreRNN(self.observation_dim, args.rnn_hidden_size, args.rnn_depth,
                             self.observation_dim, args.rnn_dropout).to(self.device)
    self.rnn_init_hidden = nn.Parameter(torch.zeros(args.rnn_depth, 1, args.rnn_hidden_size).to(self.device))
    self.estimate_sigma2 = (args.sigma2 is None)
    self.estimate_transition_bias = (args.transition_bias is None)
    sigma2 = _INITIAL_SIGMA2_VALUE if self.estimate_sigma2 else args.sigma2
    self.sigma2 = nn.Parameter(sigma2 * torch.ones(self.observation_dim).to(self.device))
    self.transition_bias = args.transition_bias
    self.transition_bias_denominator = 0.0
    self.crp_alpha = args.crp_alpha
    self.logger = logger_lib.Logger(args.verbosity)
    self._native = None          # (fingerprint, NativeModel) cache for the CUDA decoder
    self._native_lock = threading.Lock()
    if self.device.type == 'cuda':
      # say so NOW if the sm_100a kernels cannot hold this model (predict() / fit() would raise NativeError later;
      # there is no silent fallback): hidden <= 1024, dim <= 512, 1..4 GRU layers (one layer above hidden 512 / dim 256)
      too_big = args.rnn_hidden_size > 1024 or self.observation_dim > 512 or args.rnn_depth > 4 or (
          args.rnn_depth > 1 and (args.rnn_hidden_size > 512 or self.observation_dim > 256))
      if too_big:
        self.logger.print(
            1, 'Warning: the CUDA kernels of this build hold models up to rnn_hidden_size=1024, observation_dim=512, '
               'rnn_depth=4 (rnn_depth=1 above 512 / 256); predict() on this device will raise for this model. '
               'Use --enable_cuda=False for it.')

  def __getstate__(self):
    # the device-side twin and its lock are per-process; pickled copies (forkserver workers of
    # parallel_predict) rebuild them on demand
    state = dict(self.__dict__)
    state['_native'] = None
    state['_native_lock'] = None
    return state

  def __setstate__(self, state):
    self.__dict__.update(state)
    self._native_lock = threading.Lock()

  # ------------------------------------------------------------------ persistence
  def save(self, filepath):
    """Writes the checkpoint dictionary of the reference (uisrnn.py:135-147), so files are
    interchangeable between the two implementations."""
    torch.save({
        'rnn_state_dict': self.rnn_model.state_dict(),
        'rnn_init_hidden': self.rnn_init_hidden.detach().cpu().numpy(),
        'transition_bias': self.transition_bias,
        'transition_bias_denominator': self.transition_bias_denominator,
        'crp_alpha': self.crp_alpha,
        'sigma2': self.sigma2.detach().cpu().numpy()}, filepath)

  def load(self, filepath):
    """Restores a checkpoint written by `save` (of this package or of the reference).  The file
    holds numpy arrays, so it is read with `weights_only=False` (the reference's plain
    `torch.load`, uisrnn.py:155, fails on torch >= 2.6)."""
    var_dict = torch.load(filepath, map_location=self.device, weights_only=False)
    self.rnn_model.load_state_dict(var_dict['rnn_state_dict'])
    self.rnn_init_hidden = nn.Parameter(torch.from_numpy(var_dict['rnn_init_hidden']).to(self.device))
    self.transition_bias = float(var_dict['transition_bias'])
    self.transition_bias_denominator = float(var_dict['transition_bias_denominator'])
    self.crp_alpha = float(var_dict['crp_alpha'])
    self.sigma2 = nn.Parameter(torch.from_numpy(var_dict['sigma2']).to(self.device))
    self._native = None
    self.logger.print(
        3, 'Loaded model with transition_bias={}, crp_alpha={}, sigma2={}, rnn_init_hidden={}'.format(
            self.transition_bias, self.crp_alpha, var_dict['sigma2'], var_dict['rnn_init_hidden']))

  # ------------------------------------------------------------------ training
  def _get_optimizer(self, optimizer, learning_rate):
    groups = [{'params': self.rnn_model.parameters()}, {'params': self.rnn_init_hidden}]
    if self.estimate_sigma2:
      groups.append({'params': self.sigma2})
    assert optimizer == 'adam', 'Only adam optimizer is supported.'
    return optim.Adam(groups, lr=learning_rate)

  def fit_concatenated(self, train_sequence, train_cluster_id, args):
    """Trains on one concatenated sequence `train_sequence` [N, D] (float64) with string labels
    `train_cluster_id` [N] (uisrnn.py:172-313): per iteration a random batch of per-speaker
    sub-sequences, running-mean prediction, weighted-MSE + sigma^2 prior + norm regulariser,
    clipped Adam step, sigma^2 >= 1e-6."""
    if not isinstance(train_sequence, np.ndarray) or train_sequence.dtype != float:
      raise TypeError('train_sequence should be a numpy array of float type.')
    if isinstance(train_cluster_id, list):
      train_cluster_id = np.array(train_cluster_id)
    if (not isinstance(train_cluster_id, np.ndarray) or
        not train_cluster_id.dtype.name.startswith(('str', 'unicode'))):
      raise TypeError('train_cluster_id type be a numpy array of strings.')
    if train_sequence.ndim != 2:
      raise ValueError('train_sequence must be 2-dim array.')
    if train_cluster_id.ndim != 1:
      raise ValueError('train_cluster_id must be 1-dim array.')
    total_length, observation_dim = train_sequence.shape
    if observation_dim != self.observation_dim:
      raise ValueError('train_sequence does not match the dimension specified by args.observation_dim.')
    if total_length != len(train_cluster_id):
      raise ValueError('train_sequence length is not equal to train_cluster_id length.')

    self.rnn_model.train()
    optimizer = self._get_optimizer(optimizer=args.optimizer, learning_rate=args.learning_rate)
    if self._native_fit_supported(args):
      self._sync_replicas()
    if self._native_fit_supported(args):
      # device-resident training set: row indices per sub-sequence instead of num_permutations copies
      index_lists, seq_lengths = utils.resize_indices(train_cluster_id, args.num_permutations)
      self._fit_native(train_sequence, index_lists, seq_lengths, args)
      return
    self.last_fit_backend = 'torch'
    sub_sequences, seq_lengths = utils.resize_sequence(
        sequence=train_sequence, cluster_id=train_cluster_id, num_permutations=args.num_permutations)
    batch = None
    if args.batch_size is None:  # "batch learning": one fixed batch holding every sub-sequence
      batch = utils.pack_sequence(sub_sequences, seq_lengths, None, self.observation_dim, self.device)
    for num_iter in range(args.train_iteration):
      optimizer.zero_grad()
      if args.batch_size is not None:
        batch = utils.pack_sequence(sub_sequences, seq_lengths, args.batch_size, self.observation_dim,
                                    self.device)
      packed_input, rnn_truth = batch
      width = rnn_truth.size(1)
      mean, _ = self.rnn_model(packed_input, self.rnn_init_hidden.repeat(1, width, 1))
      # running average of the predictions over time (uisrnn.py:265-271 does it with a dense
      # diag(1/t) matrix product; the values are identical)
      steps = torch.arange(1, mean.size(0) + 1, device=self.device).float()
      mean = torch.cumsum(mean, dim=0) * (1.0 / steps).view(-1, 1, 1)
      mask = (rnn_truth != 0).float()
      weight = 1 / (2 * self.sigma2)
      loss1 = loss_func.weighted_mse_loss(input_tensor=mask * mean[:-1, :, :], target_tensor=rnn_truth,
                                          weight=weight)
      residual2 = ((mask * mean[:-1, :, :] - rnn_truth) ** 2).view(-1, observation_dim)
      num_non_zero = torch.sum((residual2 != 0).float(), dim=0).squeeze()
      loss2 = loss_func.sigma2_prior_loss(num_non_zero, args.sigma_alpha, args.sigma_beta, self.sigma2)
      loss3 = loss_func.regularization_loss(self.rnn_model.parameters(), args.regularization_weight)
      loss = loss1 + loss2 + loss3
      loss.backward()
      nn.utils.clip_grad_norm_(self.rnn_model.parameters(), args.grad_max_norm)
      optimizer.step()
      self.sigma2.data.clamp_(min=1e-6)
      if num_iter % 10 == 0 or num_iter == args.train_iteration - 1:
        self.logger.print(
            2, 'Iter: {:d}  \tTraining Loss: {:.4f}    \n    Negative Log Likelihood: {:.4f}\t'
               'Sigma2 Prior: {:.4f}\tRegularization: {:.4f}'.format(
                   num_iter, float(loss.data), float(loss1.data), float(loss2.data), float(loss3.data)))
    self._native = None
    self.logger.print(1, 'Done training with {} iterations'.format(args.train_iteration))

  def _sync_replicas(self):
    """Data-parallel fit(): inside an initialised `torch.distributed` job every rank adopts rank 0's
    parameters and host RNG state (numpy + `random`), so the shuffle of `concatenate_training_data`, the
    permutations of `resize_sequence` and every mini-batch draw are the same on all ranks.  No-op in a
    single process."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
      return
    with torch.no_grad():
      for tensor in list(self.rnn_model.state_dict().values()) + [self.rnn_init_hidden.data, self.sigma2.data]:
        dist.broadcast(tensor, src=0)
    import random
    rng = [np.random.get_state(), random.getstate()]  # utils.py draws from both generators
    dist.broadcast_object_list(rng, src=0)
    np.random.set_state(rng[0])
    random.setstate(rng[1])

  def _native_fit_supported(self, args):
    """On a CUDA device fit() runs on the hand-written training kernels (csrc/uis_train.cu): 1..4 stacked GRU
    layers (inter-layer dropout in train mode), any mini-batch width including batch_size=None (one batch of
    every sub-sequence).  UISRNN_B200_TORCH_FIT=1 selects PyTorch autograd instead (a developer switch for A/B
    checks); the CPU device always trains with PyTorch, as the reference does."""
    import os
    del args
    return (self.device.type == 'cuda' and 1 <= self.rnn_init_hidden.shape[0] <= 4 and
            os.environ.get('UISRNN_B200_TORCH_FIT', '0') != '1')

  def _fit_native(self, train_sequence, index_lists, seq_lengths, args):
    """fit_concatenated's iteration loop (uisrnn.py:252-311) on libuisrnn_b200.so: the training set,
    parameters, gradients and Adam state stay on the device; per iteration only the ids of the drawn
    sub-sequences go up (the batch is gathered on the device) and the loss scalars are read back
    when they are logged."""
    from . import native
    import torch.distributed as dist
    world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
    rank = dist.get_rank() if world > 1 else 0
    # world > 1: data-parallel fit (optional) -- every rank holds rank 0's parameters and draws
    # the SAME mini-batches (_sync_replicas); rank r owns batch columns r, r + world, ...
    state = {k: v.detach().cpu().numpy() for k, v in self.rnn_model.state_dict().items()}
    depth = int(self.rnn_init_hidden.shape[0])
    order = native.param_order(depth)
    params = {name: state[name] for name in order[:-2]}
    params['rnn_init_hidden'] = self.rnn_init_hidden.detach().cpu().numpy().reshape(-1)
    params['sigma2'] = self.sigma2.detach().cpu().numpy()
    # the dropout masks of the stacked GRU are seeded from torch's generator (so torch.manual_seed() makes a run
    # repeatable, as it does for the reference), one draw per fit_concatenated call (in a data-parallel job every
    # rank masks its own columns, so the ranks need not share the seed)
    dropout = float(self.rnn_model.gru.dropout) if depth > 1 else 0.0
    dropout_seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if dropout > 0 else 0
    hparams = {'learning_rate': args.learning_rate, 'sigma_alpha': args.sigma_alpha, 'sigma_beta': args.sigma_beta,
               'regularization_weight': args.regularization_weight, 'grad_max_norm': args.grad_max_norm,
               'train_sigma2': self.estimate_sigma2, 'rnn_depth': depth, 'rnn_dropout': dropout,
               'dropout_seed': dropout_seed}
    trainer = native.NativeTrainer(params, hparams, device=self.device.index or 0)
    self.last_training_losses = []
    self.last_training_loss_terms = []  # [iteration] -> (negative log likelihood, sigma2 prior, regularisation)
    self.last_fit_backend = 'native'
    comm = torch.zeros(trainer.comm_size(), dtype=torch.float32, device=self.device) if world > 1 else None
    sampler = utils.BatchSampler(seq_lengths, args.batch_size)
    try:
      trainer.set_corpus(train_sequence, index_lists)
      pending = 0  # steps enqueued since the losses were last read back
      for num_iter in range(args.train_iteration):
        chosen, _ = sampler.draw()  # same np.random.choice call as utils.pack_sequence
        if world == 1:
          trainer.step_corpus(chosen)  # asynchronous: the host runs ahead of the device
        else:
          # local forward/backward on this rank's columns -> ONE all-reduce(sum) of [gradients | loss
          # statistics] over NCCL -> identical normalise / clip / Adam step on every rank
          mine = shard_columns(len(chosen), rank, world)
          if len(mine):
            trainer.step_corpus(chosen[mine], mode=2)
            trainer.comm_export(comm.data_ptr())
          else:
            comm.zero_()
          dist.all_reduce(comm, op=dist.ReduceOp.SUM)
          trainer.comm_apply(comm.data_ptr())
        pending += 1
        log_now = num_iter % 10 == 0 or num_iter == args.train_iteration - 1
        if log_now or pending == 4096:
          recent = trainer.losses(pending)
          self.last_training_losses.extend(float(v) for v in recent[:, 0])
          self.last_training_loss_terms.extend(tuple(float(v) for v in row) for row in recent)
          pending = 0
          if log_now:
            loss1, loss2, loss3 = (float(v) for v in recent[-1])
            self.logger.print(
                2, 'Iter: {:d}  \tTraining Loss: {:.4f}    \n    Negative Log Likelihood: {:.4f}\t'
                   'Sigma2 Prior: {:.4f}\tRegularization: {:.4f}'.format(
                       num_iter, loss1 + loss2 + loss3, loss1, loss2, loss3))
      trained = trainer.parameters()
    finally:
      trainer.close()
    with torch.no_grad():
      self.rnn_model.load_state_dict({k: torch.from_numpy(trained[k].copy()) for k in order[:-2]})
      self.rnn_init_hidden.data.copy_(torch.from_numpy(trained['rnn_init_hidden'].reshape(depth, 1, -1)))
      self.sigma2.data.copy_(torch.from_numpy(trained['sigma2']))
    self._native = None
    self.logger.print(1, 'Done training with {} iterations'.format(args.train_iteration))

  def fit(self, train_sequences, train_cluster_ids, args):
    """Trains on a list of sequences (+ list of label sequences) or on one concatenated sequence
    (uisrnn.py:315-386).  Estimates / running-averages `transition_bias` unless it was given."""
    if isinstance(train_sequences, np.ndarray):
      if self.estimate_transition_bias:
        self.logger.print(
            2, 'Warning: transition_bias cannot be correctly estimated from a concatenated sequence; '
               'train_sequences will be treated as a single sequence. This can lead to inaccurate '
               'estimation of transition_bias. Please, consider estimating transition_bias before '
               'concatenating the sequences and passing it as argument.')
      train_sequences = [train_sequences]
      train_cluster_ids = [train_cluster_ids]
    elif not isinstance(train_sequences, list):
      raise TypeError('train_sequences must be a list or numpy.ndarray')
    if self._native_fit_supported(args):
      self._sync_replicas()  # before the shuffle inside concatenate_training_data
    if self.estimate_transition_bias:
      bias, denominator = utils.estimate_transition_bias(train_cluster_ids)
      if self.transition_bias is None:
        self.transition_bias = bias
        self.transition_bias_denominator = denominator
      else:  # weighted running average over successive fit() calls
        merged = self.transition_bias_denominator + denominator
        self.transition_bias = (self.transition_bias * self.transition_bias_denominator +
                                bias * denominator) / merged
        self.transition_bias_denominator = merged
    sequence, cluster_id = utils.concatenate_training_data(
        train_sequences, train_cluster_ids, args.enforce_cluster_id_uniqueness, True)
    self.fit_concatenated(sequence, cluster_id, args)

  # ------------------------------------------------------------------ inference
  def _fingerprint(self):
    """Identity of the parameter values the device-side twin was built from.  `_version` does not move on
    edits through `.data` (an idiom the reference's own tests use), so a cheap content checksum -- the L2 and L1
    norms of every tensor, two fused multi-tensor reductions -- is part of the key."""
    tensors = list(self.rnn_model.parameters()) + [self.rnn_init_hidden, self.sigma2]
    with torch.no_grad():
      by_device = {}  # the parameters need not share a device (callers assign rnn_init_hidden / sigma2 freely)
      for i, t in enumerate(tensors):
        by_device.setdefault(t.device, []).append((i, t.detach()))
      sums = [None] * len(tensors)
      for parts in by_device.values():  # two fused multi-tensor reductions and one device -> host copy per device
        ts = [t for _, t in parts]
        try:
          n2, n1 = torch._foreach_norm(ts, 2), torch._foreach_norm(ts, 1)  # pylint: disable=protected-access
          flat = torch.stack(list(n2) + list(n1)).double().cpu().tolist()
          pairs = list(zip(flat[:len(ts)], flat[len(ts):]))
        except (AttributeError, RuntimeError, TypeError):  # no multi-tensor kernels in this torch: one by one
          pairs = [tuple(torch.stack((t.double().norm(2), t.double().norm(1))).cpu().tolist()) for t in ts]
        for (i, _), pair in zip(parts, pairs):
          sums[i] = tuple(pair)
    return (tuple((t.data_ptr(), t._version) for t in tensors), tuple(sums),
            self.transition_bias, self.crp_alpha)

  def export_weights(self):
    """Weights as float32 numpy arrays in the layout libuisrnn_b200.so / the oracle expect."""
    state = {k: v.detach().cpu().numpy() for k, v in self.rnn_model.state_dict().items()}
    depth = self.rnn_init_hidden.shape[0]
    out = {'depth': depth,
           'w1': state['linear_mean1.weight'], 'b1': state['linear_mean1.bias'],
           'w2': state['linear_mean2.weight'], 'b2': state['linear_mean2.bias'],
           'h0': self.rnn_init_hidden.detach().cpu().numpy(),
           'sigma2': self.sigma2.detach().cpu().numpy(),
           'transition_bias': self.transition_bias, 'crp_alpha': self.crp_alpha}
    for layer in range(depth):
      for name in ('weight_ih', 'weight_hh', 'bias_ih', 'bias_hh'):
        out['{}_l{}'.format(name, layer)] = state['gru.{}_l{}'.format(name, layer)]
    return out

  def _native_model(self, device_index=None):
    """The device-side twin of this model (created lazily, rebuilt when any parameter changed)."""
    from . import native  # raises NativeError if the library has not been built
    index = (self.device.index or 0) if device_index is None else device_index
    with self._native_lock:
      key = (self._fingerprint(), index)
      if self._native is None or self._native[0] != key:
        if self.transition_bias is None:
          raise TypeError('transition_bias is not set: call fit() or pass --transition_bias.')
        self._native = (key, native.NativeModel(self.export_weights(), device=index))
      return self._native[1]

  def _predict_cuda(self, sequences, args, device_index=None, as_arrays=False):
    from . import native
    model = self._native_model(device_index)
    kcap = _DEFAULT_KCAP
    while True:
      try:
        with model.lock:  # a uis_model handle (one workspace) is not re-entrant
          labels = model.predict(sequences, beam_size=args.beam_size, look_ahead=args.look_ahead,
                                 test_iteration=args.test_iteration, kcap=kcap)
        return labels if as_arrays else [lab.tolist() for lab in labels]
      except native.NativeError as err:
        if err.code != native.UIS_ERR_OVERFLOW or kcap >= 1024:
          raise
        kcap = 32 if kcap == 0 else kcap * 2  # a hypothesis opened more clusters than the device tables hold: grow and retry

  def predict_single(self, test_sequence, args):
    """Labels (list of N ints) for one test sequence [N, D] float64 (uisrnn.py:479-562)."""
    _check_test_sequence(test_sequence, self.observation_dim)
    if self.device.type == 'cuda':
      return self._predict_cuda([test_sequence], args)[0]
    decoder = beam_cpu.CpuBeamSearch(self)
    return decoder.decode(test_sequence, args.beam_size, args.look_ahead, args.test_iteration)

  def predict(self, test_sequences, args):
    """Labels for one sequence (ndarray -> list of ints) or many (list -> list of lists)
    (uisrnn.py:564-590).  On CUDA a list is decoded by a single native call."""
    if isinstance(test_sequences, np.ndarray):
      return self.predict_single(test_sequences, args)
    if isinstance(test_sequences, list):
      if self.device.type == 'cuda':
        for sequence in test_sequences:
          _check_test_sequence(sequence, self.observation_dim)
        return self._predict_cuda(test_sequences, args)
      return [self.predict_single(sequence, args) for sequence in test_sequences]
    raise TypeError('test_sequences should be either a list or numpy array.')


def _predict_shard(model, args, device_index, sequences, out, position):
  out[position] = model._predict_cuda(sequences, args, device_index)  # pylint: disable=protected-access


def parallel_predict(model, test_sequences, args, num_processes=4):
  """Parallel prediction over a list of sequences (uisrnn.py:593-623).

  CPU model: a forkserver process pool, as the reference.  CUDA model: `num_processes` is the
  number of GPUs to use (capped by the visible devices); the list is split by total frame count
  and each shard is decoded by one native call on its own device, from its own host thread
  (the C ABI releases the GIL).  Utterances are independent, so there is no collective.
  """
  if not isinstance(test_sequences, list):
    raise TypeError('test_sequences must be a list.')
  if model.device.type == 'cuda':
    for sequence in test_sequences:
      _check_test_sequence(sequence, model.observation_dim)
    n_dev = max(1, min(int(num_processes), torch.cuda.device_count()))
    if n_dev == 1 or len(test_sequences) < 2:
      return model._predict_cuda(test_sequences, args)  # pylint: disable=protected-access
    shards = shard_by_frames([len(s) for s in test_sequences], n_dev)
    twins = [model] + [_clone_for_device(model, d) for d in range(1, n_dev)]
    results, threads = [None] * n_dev, []
    for d, shard in enumerate(shards):
      thread = threading.Thread(target=_predict_shard, args=(
          twins[d], args, d, [test_sequences[i] for i in shard], results, d))
      thread.start()
      threads.append(thread)
    for thread in threads:
      thread.join()
    merged = [None] * len(test_sequences)
    for shard, labels in zip(shards, results):
      if labels is None:
        raise RuntimeError('parallel_predict: a device shard failed')
      for i, lab in zip(shard, labels):
        merged[i] = lab
    return merged
  ctx = multiprocessing.get_context('forkserver')
  model.rnn_model.share_memory()
  with ctx.Pool(num_processes) as pool:
    return pool.map(functools.partial(model.predict_single, args=args), test_sequences)


class _DeviceTwin:
  """Just enough of a UISRNN to own a NativeModel on another device."""

  def __init__(self, weights):
    self._weights = weights
    self._models = {}
    self._lock = threading.Lock()

  def _predict_cuda(self, sequences, args, device_index):
    from . import native
    with self._lock:
      if device_index not in self._models:
        self._models[device_index] = native.NativeModel(self._weights, device=device_index)
      model = self._models[device_index]
    kcap = _DEFAULT_KCAP
    while True:
      try:
        with model.lock:
          labels = model.predict(sequences, beam_size=args.beam_size, look_ahead=args.look_ahead,
                                 test_iteration=args.test_iteration, kcap=kcap)
        return [lab.tolist() for lab in labels]
      except native.NativeError as err:
        if err.code != native.UIS_ERR_OVERFLOW or kcap >= 1024:
          raise
        kcap = 32 if kcap == 0 else kcap * 2


def _clone_for_device(model, device_index):
  del device_index
  return _DeviceTwin(model.export_weights())


def shard_columns(width, rank, world):
  """Columns of a (length-sorted) mini-batch owned by `rank` in data-parallel fit(): rank, rank + world,
  ... -- every shard stays sorted by decreasing length and the long sequences are spread evenly."""
  return np.arange(rank, width, world)


def shard_by_frames(lengths, n_shards):
  """Longest-processing-time-first partition of utterance indices into `n_shards` groups with
  near-equal total frame counts (cost of an utterance ~ its frame count).  Returns index lists."""
  order = sorted(range(len(lengths)), key=lambda i: -lengths[i])
  loads = [0] * n_shards
  shards = [[] for _ in range(n_shards)]
  for i in order:
    target = loads.index(min(loads))
    shards[target].append(i)
    loads[target] += lengths[i]
  for shard in shards:
    shard.sort()
  return shards
