"""fit() on the device (csrc/uis_train.cu through the C ABI) against PyTorch autograd on the same
batch: the three loss terms, every gradient, one Adam step, and a short training trajectory."""
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _model_and_data(D=64, H=128, seed=11, depth=1, dropout=0.0):
  import torch
  import uisrnn
  from uisrnn_b200.synth import synth_training_set
  # the autograd reference must be true fp32: cuDNN's RNN uses TF32 tensor cores by default
  torch.backends.cudnn.allow_tf32 = False
  torch.backends.cuda.matmul.allow_tf32 = False
  np.random.seed(seed); random.seed(seed); torch.manual_seed(seed)
  m, t, _ = uisrnn.parse_arguments([])
  m.observation_dim, m.rnn_hidden_size, m.verbosity = D, H, 0
  m.rnn_depth, m.rnn_dropout = depth, dropout
  t.batch_size, t.learning_rate, t.train_iteration = 12, 1e-3, 1
  model = uisrnn.UISRNN(m)
  assert model.device.type == 'cuda'
  with torch.no_grad():
    model.rnn_init_hidden.data.normal_(0, 0.1)
    model.sigma2.data.uniform_(0.05, 0.2)
  seqs, ids = synth_training_set(4000, 30, n_frames=50, dim=D, n_spk=3, noise=0.08)
  from uisrnn_b200 import utils
  x, y = utils.concatenate_training_data(seqs, ids, True, True)
  subs, lens = utils.resize_sequence(x, np.array(y), t.num_permutations)
  return model, t, subs, lens


def _torch_losses_and_grads(model, targs, rnn_input, lengths):
  import torch
  from torch import nn
  from uisrnn_b200 import loss_func
  dev = model.device
  x = torch.from_numpy(rnn_input).float().to(dev)
  packed = nn.utils.rnn.pack_padded_sequence(x, lengths, batch_first=False)
  truth = x[1:]
  for p in list(model.rnn_model.parameters()) + [model.rnn_init_hidden, model.sigma2]:
    p.grad = None
  mean, _ = model.rnn_model(packed, model.rnn_init_hidden.repeat(1, x.size(1), 1))
  steps = torch.arange(1, mean.size(0) + 1, device=dev).float()
  mean = torch.cumsum(mean, dim=0) * (1.0 / steps).view(-1, 1, 1)
  mask = (truth != 0).float()
  loss1 = loss_func.weighted_mse_loss(mask * mean[:-1], truth, 1 / (2 * model.sigma2))
  res = ((mask * mean[:-1] - truth) ** 2).view(-1, x.size(2))
  nnz = torch.sum((res != 0).float(), dim=0).squeeze()
  loss2 = loss_func.sigma2_prior_loss(nnz, targs.sigma_alpha, targs.sigma_beta, model.sigma2)
  loss3 = loss_func.regularization_loss(model.rnn_model.parameters(), targs.regularization_weight)
  (loss1 + loss2 + loss3).backward()
  grads = {n: p.grad.detach().cpu().numpy() for n, p in model.rnn_model.named_parameters()}
  grads['rnn_init_hidden'] = model.rnn_init_hidden.grad.detach().cpu().numpy().reshape(-1)
  grads['sigma2'] = model.sigma2.grad.detach().cpu().numpy()
  return (float(loss1), float(loss2), float(loss3)), grads


def _native_trainer(model, targs, dropout=0.0, dropout_seed=0):
  from uisrnn_b200 import native
  depth = int(model.rnn_init_hidden.shape[0])
  state = {k: v.detach().cpu().numpy() for k, v in model.rnn_model.state_dict().items()}
  params = {k: state[k] for k in native.param_order(depth)[:-2]}
  params['rnn_init_hidden'] = model.rnn_init_hidden.detach().cpu().numpy().reshape(-1)
  params['sigma2'] = model.sigma2.detach().cpu().numpy()
  hp = {'learning_rate': targs.learning_rate, 'sigma_alpha': targs.sigma_alpha, 'sigma_beta': targs.sigma_beta,
        'regularization_weight': targs.regularization_weight, 'grad_max_norm': targs.grad_max_norm,
        'train_sigma2': True, 'rnn_depth': depth, 'rnn_dropout': dropout, 'dropout_seed': dropout_seed}
  return native.NativeTrainer(params, hp)


def _rel(a, b):
  return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-20))


def test_losses_and_gradients_match_autograd():
  from uisrnn_b200 import utils
  model, targs, subs, lens = _model_and_data()
  model.rnn_model.train()
  rnn_input, lengths = utils.pack_batch(subs, lens, targs.batch_size, model.observation_dim)
  want_losses, want = _torch_losses_and_grads(model, targs, rnn_input, lengths)
  trainer = _native_trainer(model, targs)
  got_losses = trainer.step(rnn_input.astype(np.float32), lengths, grads_only=True)
  got = trainer.gradients()
  for a, b in zip(got_losses, want_losses):
    assert abs(a - b) <= 1e-4 * max(1.0, abs(b)), (got_losses, want_losses)
  for name in want:
    assert _rel(got[name].reshape(want[name].shape), want[name]) < 2e-3, name


def test_gradients_at_default_model_size_batch32():
  """BASELINE config 4 shape: D=256, H=512, batch_size=32."""
  from uisrnn_b200 import utils
  model, targs, subs, lens = _model_and_data(D=256, H=512, seed=21)
  targs.batch_size = 32
  model.rnn_model.train()
  rnn_input, lengths = utils.pack_batch(subs, lens, targs.batch_size, model.observation_dim)
  want_losses, want = _torch_losses_and_grads(model, targs, rnn_input, lengths)
  trainer = _native_trainer(model, targs)
  got_losses = trainer.step(rnn_input.astype(np.float32), lengths, grads_only=True)
  got = trainer.gradients()
  for a, b in zip(got_losses, want_losses):
    assert abs(a - b) <= 1e-4 * max(1.0, abs(b)), (got_losses, want_losses)
  for name in want:
    assert _rel(got[name].reshape(want[name].shape), want[name]) < 2e-3, name


def test_one_adam_step_matches_torch():
  import torch
  from torch import nn
  from uisrnn_b200 import utils
  model, targs, subs, lens = _model_and_data(seed=12)
  model.rnn_model.train()
  rnn_input, lengths = utils.pack_batch(subs, lens, targs.batch_size, model.observation_dim)
  trainer = _native_trainer(model, targs)
  trainer.step(rnn_input.astype(np.float32), lengths)
  got = trainer.parameters()
  optimizer = model._get_optimizer('adam', targs.learning_rate)   # pylint: disable=protected-access
  optimizer.zero_grad()
  _torch_losses_and_grads(model, targs, rnn_input, lengths)
  nn.utils.clip_grad_norm_(model.rnn_model.parameters(), targs.grad_max_norm)
  optimizer.step()
  model.sigma2.data.clamp_(min=1e-6)
  for name, p in model.rnn_model.named_parameters():
    assert np.max(np.abs(got[name] - p.detach().cpu().numpy())) < 2e-5, name   # lr = 1e-3: |update| <= 1e-3
  assert np.max(np.abs(got['sigma2'] - model.sigma2.detach().cpu().numpy())) < 2e-5
  assert np.max(np.abs(got['rnn_init_hidden'] - model.rnn_init_hidden.detach().cpu().numpy().reshape(-1))) < 2e-5


def test_training_trajectory_matches_torch_path(monkeypatch):
  """20 iterations with the same RNG stream: native loss1 trajectory within 1e-3 of the autograd path
  (SURVEY.md section 8(d), config 4 criterion)."""
  import torch
  from uisrnn_b200 import utils

  def run(native_path):
    model, targs, subs, lens = _model_and_data(seed=13)
    np.random.seed(99)
    losses = []
    if native_path:
      trainer = _native_trainer(model, targs)
      for _ in range(20):
        x, l = utils.pack_batch(subs, lens, targs.batch_size, model.observation_dim)
        losses.append(trainer.step(x.astype(np.float32), l)[0])
    else:
      model.rnn_model.train()
      opt = model._get_optimizer('adam', targs.learning_rate)   # pylint: disable=protected-access
      for _ in range(20):
        x, l = utils.pack_batch(subs, lens, targs.batch_size, model.observation_dim)
        opt.zero_grad()
        (l1, _, _), _ = _torch_losses_and_grads(model, targs, x, l)
        torch.nn.utils.clip_grad_norm_(model.rnn_model.parameters(), targs.grad_max_norm)
        opt.step()
        model.sigma2.data.clamp_(min=1e-6)
        losses.append(l1)
    return np.array(losses)

  a, b = run(True), run(False)
  assert np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b))) < 1e-3, (a, b)


def test_fit_api_uses_native_trainer_and_learns():
  import torch
  import uisrnn
  from uisrnn_b200.synth import synth_training_set, synth_utt
  np.random.seed(5); random.seed(5); torch.manual_seed(5)
  m, t, i = uisrnn.parse_arguments([])
  m.observation_dim, m.rnn_hidden_size, m.verbosity = 64, 128, 0
  t.batch_size, t.learning_rate, t.train_iteration = 16, 2e-3, 150
  model = uisrnn.UISRNN(m)
  seqs, ids = synth_training_set(6000, 60, n_frames=60, dim=64, n_spk=3, noise=0.08)
  model.fit(seqs, ids, t)
  losses = model.last_training_losses
  assert len(losses) == 150 and np.all(np.isfinite(losses))   # likelihood term only, as uisrnn.py:311
  x, truth = synth_utt(6100, n_frames=80, dim=64, n_spk=3, noise=0.08)
  acc = uisrnn.compute_sequence_match_accuracy(model.predict(x, i), truth.tolist())
  assert acc > 0.9


def test_device_gathered_batch_equals_host_packed_batch():
  """uis_trainer_step_corpus (training set resident on the device, batch gathered there) is bit-identical
  to uis_trainer_step on the batch utils.pack_batch builds on the host."""
  from uisrnn_b200 import utils
  from uisrnn_b200.synth import synth_training_set
  model, targs, _, _ = _model_and_data(seed=17)
  seqs, ids = synth_training_set(4000, 30, n_frames=50, dim=64, n_spk=3, noise=0.08)
  x, y = utils.concatenate_training_data(seqs, ids, True, True)
  np.random.seed(23)
  subs, lens = utils.resize_sequence(x, np.array(y), targs.num_permutations)
  np.random.seed(23)
  index_lists, lens2 = utils.resize_indices(np.array(y), targs.num_permutations)
  assert lens == lens2
  host, dev = _native_trainer(model, targs), _native_trainer(model, targs)
  dev.set_corpus(x, index_lists)
  sampler = utils.BatchSampler(lens, targs.batch_size)
  for it in range(5):
    np.random.seed(100 + it)
    batch, lengths = utils.pack_batch(subs, lens, targs.batch_size, model.observation_dim)
    np.random.seed(100 + it)
    chosen, lengths2 = sampler.draw()
    assert np.array_equal(lengths, lengths2)
    want = host.step(batch.astype(np.float32), lengths)
    got = dev.step_corpus(chosen, want_losses=True)
    # same arithmetic on the same rows; the per-dimension loss sums use float atomics (order not fixed)
    assert np.allclose(got, want, rtol=2e-6, atol=0), (it, got, want)
  a, b = host.parameters(), dev.parameters()
  assert all(np.max(np.abs(a[k] - b[k])) < 1e-6 for k in a)
  with pytest.raises(Exception):
    dev.step_corpus(np.array([len(index_lists)]))   # id out of range
  host.close(); dev.close()


def test_persistent_recurrence_kernels_match_per_step_launches(monkeypatch):
  """The cooperative whole-sequence GRU kernels (one launch per direction) against the per-step launches
  (UISRNN_B200_TRAIN_STEPWISE=1): same losses and gradients up to fp32 summation order."""
  from uisrnn_b200 import utils
  model, targs, subs, lens = _model_and_data(seed=19)
  np.random.seed(7)
  batch, lengths = utils.pack_batch(subs, lens, targs.batch_size, model.observation_dim)
  persistent = _native_trainer(model, targs)
  want_losses = persistent.step(batch.astype(np.float32), lengths, grads_only=True)
  want = persistent.gradients()
  monkeypatch.setenv('UISRNN_B200_TRAIN_STEPWISE', '1')
  stepwise = _native_trainer(model, targs)
  got_losses = stepwise.step(batch.astype(np.float32), lengths, grads_only=True)
  got = stepwise.gradients()
  assert np.allclose(got_losses, want_losses, rtol=1e-5)
  for name in want:
    assert _rel(got[name], want[name]) < 2e-5, name
  persistent.close(); stepwise.close()


def _check_against_autograd(model, targs, rnn_input, lengths, tol=2e-3):
  want_losses, want = _torch_losses_and_grads(model, targs, rnn_input, lengths)
  trainer = _native_trainer(model, targs)
  got_losses = trainer.step(rnn_input.astype(np.float32), lengths, grads_only=True)
  got = trainer.gradients()
  trainer.close()
  for a, b in zip(got_losses, want_losses):
    assert abs(a - b) <= 1e-4 * max(1.0, abs(b)), (got_losses, want_losses)
  assert set(want) == set(got)
  for name in want:
    assert _rel(got[name].reshape(want[name].shape), want[name]) < tol, name


@pytest.mark.parametrize('D,H,depth,batch', [(64, 128, 1, 48), (64, 128, 2, 12), (64, 128, 3, 40), (256, 512, 2, 70),
                                             (40, 96, 2, 37)])
def test_stacked_layers_and_wide_batches_match_autograd(D, H, depth, batch):
  """rnn_depth 2..3 (nn.GRU(num_layers), uisrnn.py:39-41; dropout 0 so that autograd is deterministic) and
  mini-batches wider than one 32-column group (any --batch_size, arguments.py:126-131): losses and every gradient
  against torch autograd.  H = 96 is not a multiple of 128: the per-step launches serve it instead of the
  persistent recurrence kernels."""
  from uisrnn_b200 import utils
  model, targs, subs, lens = _model_and_data(D=D, H=H, seed=31 + depth, depth=depth)
  targs.batch_size = batch
  model.rnn_model.train()
  rnn_input, lengths = utils.pack_batch(subs, lens, targs.batch_size, model.observation_dim)
  _check_against_autograd(model, targs, rnn_input, lengths)


def test_batch_learning_none_batch_size_matches_autograd():
  """--batch_size None = one batch holding every sub-sequence (utils.py:230-233)."""
  from uisrnn_b200 import utils
  model, targs, subs, lens = _model_and_data(seed=41)
  model.rnn_model.train()
  rnn_input, lengths = utils.pack_batch(subs, lens, None, model.observation_dim)
  assert rnn_input.shape[1] == len(subs) > 32
  _check_against_autograd(model, targs, rnn_input, lengths)


def test_inter_layer_dropout_gradients_match_autograd_with_the_same_masks():
  """Train-mode dropout between stacked GRU layers (nn.GRU(dropout=p), uisrnn.py:39-41): the device trainer's
  masks are a pure function of (seed, iteration, layer, element) -- native.dropout_keep_mask restates the hash --
  so the same masks are applied inside a layer-by-layer torch model that shares the parameter tensors, and all
  gradients are compared.  Two iterations: the masks change with the iteration number."""
  import torch
  from torch import nn
  from uisrnn_b200 import loss_func, native, utils
  p, seed = 0.3, 0x1234567890
  model, targs, subs, lens = _model_and_data(seed=43, depth=2, dropout=p)
  model.rnn_model.train()
  dev = model.device
  gru = model.rnn_model.gru
  layers = []
  for l in range(2):
    g = nn.GRU(model.observation_dim if l == 0 else gru.hidden_size, gru.hidden_size, 1).to(dev)
    for kind in ('weight_ih', 'weight_hh', 'bias_ih', 'bias_hh'):
      setattr(g, kind + '_l0', getattr(gru, '{}_l{}'.format(kind, l)))   # shared Parameter objects
    layers.append(g)
  trainer = _native_trainer(model, targs, dropout=p, dropout_seed=seed)
  for iteration in range(2):
    rnn_input, lengths = utils.pack_batch(subs, lens, targs.batch_size, model.observation_dim)
    x = torch.from_numpy(rnn_input).float().to(dev)
    L, B, H = x.size(0), x.size(1), gru.hidden_size
    for q in list(model.rnn_model.parameters()) + [model.rnn_init_hidden, model.sigma2]:
      q.grad = None
    packed = nn.utils.rnn.pack_padded_sequence(x, lengths, batch_first=False)
    out0, _ = layers[0](packed, model.rnn_init_hidden[0:1].repeat(1, B, 1))
    out0, _ = nn.utils.rnn.pad_packed_sequence(out0, batch_first=False)
    keep = native.dropout_keep_mask(seed, iteration, 0, L * B * H, p).reshape(L, B, H)
    assert 0.6 < keep.mean() < 0.8
    out0 = out0 * torch.from_numpy(keep.astype(np.float32) / np.float32(1 - p)).to(dev)
    out1, _ = layers[1](nn.utils.rnn.pack_padded_sequence(out0, lengths, batch_first=False),
                        model.rnn_init_hidden[1:2].repeat(1, B, 1))
    out1, _ = nn.utils.rnn.pad_packed_sequence(out1, batch_first=False)
    mean = model.rnn_model.linear_mean2(torch.relu(model.rnn_model.linear_mean1(out1)))
    steps = torch.arange(1, L + 1, device=dev).float()
    mean = torch.cumsum(mean, dim=0) * (1.0 / steps).view(-1, 1, 1)
    truth = x[1:]
    mask = (truth != 0).float()
    loss1 = loss_func.weighted_mse_loss(mask * mean[:-1], truth, 1 / (2 * model.sigma2))
    res = ((mask * mean[:-1] - truth) ** 2).view(-1, x.size(2))
    nnz = torch.sum((res != 0).float(), dim=0).squeeze()
    loss2 = loss_func.sigma2_prior_loss(nnz, targs.sigma_alpha, targs.sigma_beta, model.sigma2)
    loss3 = loss_func.regularization_loss(model.rnn_model.parameters(), targs.regularization_weight)
    (loss1 + loss2 + loss3).backward()
    want = {n: q.grad.detach().cpu().numpy() for n, q in model.rnn_model.named_parameters()}
    want['rnn_init_hidden'] = model.rnn_init_hidden.grad.detach().cpu().numpy().reshape(-1)
    want['sigma2'] = model.sigma2.grad.detach().cpu().numpy()
    got_losses = trainer.step(rnn_input.astype(np.float32), lengths, grads_only=True)   # iteration counter += 1
    got = trainer.gradients()
    assert abs(got_losses[0] - float(loss1)) <= 1e-4 * max(1.0, abs(float(loss1)))
    for name in want:
      assert _rel(got[name].reshape(want[name].shape), want[name]) < 2e-3, (iteration, name)
  trainer.close()


def test_fit_api_trains_depth2_with_dropout_and_wide_batch_on_the_device():
  """The public fit() on CUDA with --rnn_depth 2 (default --rnn_dropout 0.2) and --batch_size 40 runs on the
  device trainer (no PyTorch autograd on this path), is repeatable under torch.manual_seed, and learns."""
  import torch
  import uisrnn
  from uisrnn_b200.synth import synth_training_set, synth_utt

  def run():
    np.random.seed(5); random.seed(5); torch.manual_seed(5)
    m, t, i = uisrnn.parse_arguments([])
    m.observation_dim, m.rnn_hidden_size, m.rnn_depth, m.verbosity = 64, 128, 2, 0
    assert m.rnn_dropout == 0.2
    t.batch_size, t.learning_rate, t.train_iteration = 40, 2e-3, 150
    model = uisrnn.UISRNN(m)
    seqs, ids = synth_training_set(6000, 60, n_frames=60, dim=64, n_spk=3, noise=0.08)
    model.fit(seqs, ids, t)
    return model, i
  model, i = run()
  assert model.last_fit_backend == 'native'
  losses = np.array(model.last_training_losses)
  assert len(losses) == 150 and np.all(np.isfinite(losses))   # likelihood term only (it grows while sigma2 shrinks)
  again, _ = run()
  # same masks, same batches; the per-dimension loss sums use float atomics, so not bit-identical
  assert np.allclose(np.array(again.last_training_losses), losses, rtol=2e-3)
  x, truth = synth_utt(6100, n_frames=80, dim=64, n_spk=3, noise=0.08)
  acc = uisrnn.compute_sequence_match_accuracy(model.predict(x, i), truth.tolist())
  assert acc > 0.9


@pytest.mark.parametrize('name', ['d1_b16', 'd1_b48', 'd2_b16'])
def test_native_fit_follows_reference_trajectory(name):
  """SURVEY 8(d) config 4 criterion: the device trainer (csrc/uis_train.cu), started from the reference's initial
  parameters with the reference's RNG stream, follows the loss trajectory of the UNMODIFIED reference's fit()
  (tests/golden/fit_traj.npz): first 20 iterations, every loss term, relative error <= 1e-3; the parameters it
  ends at agree with the reference's to 1e-4."""
  from fit_traj import run_case
  losses, want, model, final = run_case(name, enable_cuda=True)
  assert model.last_fit_backend == 'native'
  assert losses.shape == want.shape == (20, 3)
  assert np.max(np.abs(losses - want) / np.maximum(1.0, np.abs(want))) < 1e-3
  sd = {k: v.cpu().numpy() for k, v in model.rnn_model.state_dict().items()}
  assert np.max(np.abs(sd['linear_mean2.weight'] - final['w2'])) < 1e-4
  assert np.max(np.abs(sd['gru.weight_hh_l0'] - final['weight_hh_l0'])) < 1e-4
  if 'weight_ih_l1' in final:
    assert np.max(np.abs(sd['gru.weight_ih_l1'] - final['weight_ih_l1'])) < 1e-4
  assert np.max(np.abs(model.sigma2.detach().cpu().numpy() - final['sigma2'])) < 1e-5
