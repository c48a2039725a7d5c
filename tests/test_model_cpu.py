"""UISRNN API on the CPU device: training/prediction plumbing, exceptions, checkpoints, and the
CPU decoder (uisrnn_b200/beam_cpu.py) against the reference's golden vectors.  CPU only.
Modelled on the reference's tests/uisrnn_test.py and tests/integration_test.py."""
import os
import random

import numpy as np
import pytest
import torch

import uisrnn
from helpers import depth2_cases, inference_args, load_weights, small_cases, toy_utterances, uisrnn_from_weights


def _tiny_args():
  m, t, i = uisrnn.parse_arguments([])
  m.enable_cuda, m.rnn_depth, m.rnn_hidden_size, m.observation_dim, m.verbosity = False, 1, 8, 16, 0
  t.learning_rate, t.train_iteration = 0.01, 50
  i.test_iteration = 1
  return m, t, i


@pytest.fixture(scope='module')
def single_label_model():
  np.random.seed(0); random.seed(0); torch.manual_seed(0)
  m, t, i = _tiny_args()
  model = uisrnn.UISRNN(m)
  model.fit(np.random.rand(1000, 16), np.array(['A'] * 1000), t)
  return model, i


def test_fit_concatenated_then_predict_single_label(single_label_model):
  model, iargs = single_label_model
  seq = np.random.rand(10, 16) / 10.0
  assert model.predict(seq, iargs) == [0] * 10
  out = model.predict([seq, seq[:4]], iargs)
  assert isinstance(out, list) and out == [[0] * 10, [0] * 4]
  assert 0 < model.transition_bias < 1 and model.transition_bias_denominator == 2 + 999


def test_parallel_predict_cpu(single_label_model):
  model, iargs = single_label_model
  seqs = [np.random.rand(6, 16) / 10.0, np.random.rand(5, 16) / 10.0]
  assert uisrnn.parallel_predict(model, seqs, iargs, num_processes=2) == [[0] * 6, [0] * 5]
  with pytest.raises(TypeError):
    uisrnn.parallel_predict(model, seqs[0], iargs)


def test_fit_list_input():
  np.random.seed(1); random.seed(1); torch.manual_seed(1)
  m, t, i = _tiny_args()
  model = uisrnn.UISRNN(m)
  model.fit([np.random.rand(100, 16), np.random.rand(200, 16)], [np.array(['A'] * 100), ['A'] * 200], t)
  assert model.predict(np.random.rand(8, 16) / 10.0, i) == [0] * 8


def test_type_and_shape_errors(single_label_model):
  model, iargs = single_label_model
  _, t, _ = _tiny_args()
  t.train_iteration = 1
  with pytest.raises(ValueError):
    model.fit(np.random.rand(50, 18), np.array(['A'] * 50), t)
  with pytest.raises(ValueError):
    model.fit_concatenated(np.random.rand(50, 16), np.array(['A'] * 49), t)
  with pytest.raises(TypeError):
    model.fit_concatenated(np.random.rand(50, 16).astype(np.float32), np.array(['A'] * 50), t)
  with pytest.raises(TypeError):
    model.fit_concatenated(np.random.rand(50, 16), np.arange(50), t)
  with pytest.raises(TypeError):
    model.fit('nope', ['A'], t)
  with pytest.raises(ValueError):
    model.predict(np.random.rand(10, 18), iargs)
  with pytest.raises(ValueError):
    model.predict(np.random.rand(16), iargs)
  with pytest.raises(TypeError):
    model.predict(np.random.rand(10, 16).astype(np.float32), iargs)
  with pytest.raises(TypeError):
    model.predict('nope', iargs)


def test_save_and_load_roundtrip(tmp_path):
  m, _, _ = uisrnn.parse_arguments([])
  m.enable_cuda, m.observation_dim, m.transition_bias, m.sigma2, m.verbosity = False, 16, 0.5, 0.05, 0
  model = uisrnn.UISRNN(m)
  path = str(tmp_path / 'model.uisrnn')
  model.save(path)
  other = uisrnn.UISRNN(m)
  other.load(path)
  assert other.transition_bias == 0.5 and other.crp_alpha == 1.0
  assert torch.equal(other.sigma2.data, model.sigma2.data)
  for a, b in zip(model.rnn_model.parameters(), other.rnn_model.parameters()):
    assert torch.equal(a.data, b.data)
  # checkpoint layout of the reference (uisrnn.py:135-147)
  blob = torch.load(path, weights_only=False)
  assert set(blob) == {'rnn_state_dict', 'rnn_init_hidden', 'transition_bias',
                       'transition_bias_denominator', 'crp_alpha', 'sigma2'}
  assert set(blob['rnn_state_dict']) == {
      'gru.weight_ih_l0', 'gru.weight_hh_l0', 'gru.bias_ih_l0', 'gru.bias_hh_l0',
      'linear_mean1.weight', 'linear_mean1.bias', 'linear_mean2.weight', 'linear_mean2.bias'}
  assert isinstance(blob['sigma2'], np.ndarray) and blob['rnn_init_hidden'].shape == (1, 1, 512)


def test_four_clusters_depth2_end_to_end(tmp_path):
  """The reference's integration scenario (tests/integration_test.py:56-154): depth-2 GRU, 4
  clusters on a unit square, list fit, save/load, 100 % accuracy, second fit moves transition_bias."""
  np.random.seed(1); random.seed(1); torch.manual_seed(1)
  centres = {'A': (0.0, 0.0), 'B': (0.0, 1.0), 'C': (1.0, 0.0), 'D': (1.0, 1.0)}

  def make(labels):
    return np.array([centres[c] for c in labels]) + np.random.rand(len(labels), 2) * 0.01

  train_ids = ['A'] * 400 + ['B'] * 300 + ['C'] * 200 + ['D'] * 100
  random.shuffle(train_ids)
  train = make(train_ids)
  cuts = [0, 100, 300, 600, 1000]
  test_ids = ['A'] * 10 + ['B'] * 20 + ['C'] * 30 + ['D'] * 40
  random.shuffle(test_ids)
  test = make(test_ids)
  m, t, i = uisrnn.parse_arguments([])
  m.enable_cuda, m.rnn_depth, m.rnn_hidden_size, m.observation_dim, m.verbosity = False, 2, 8, 2, 0
  t.learning_rate, t.train_iteration, t.enforce_cluster_id_uniqueness = 0.01, 200, False
  model = uisrnn.UISRNN(m)
  model.fit([train[a:b] for a, b in zip(cuts, cuts[1:])], [train_ids[a:b] for a, b in zip(cuts, cuts[1:])], t)
  path = str(tmp_path / 'm.uisrnn')
  model.save(path)
  assert uisrnn.compute_sequence_match_accuracy(model.predict(test, i), test_ids) == 1.0
  loaded = uisrnn.UISRNN(m)
  loaded.load(path)
  assert uisrnn.compute_sequence_match_accuracy(loaded.predict(test, i), test_ids) == 1.0
  before = model.transition_bias
  t.learning_rate, t.train_iteration = 0.001, 5
  model.fit(train[:100], train_ids[:100], t)
  assert model.transition_bias != before


@pytest.mark.parametrize('case', small_cases(), ids=lambda c: c['name'])
def test_cpu_decoder_matches_reference_golden(case):
  """beam_cpu.py (any look_ahead) against labels produced by the unmodified reference."""
  model = uisrnn_from_weights(load_weights('model_small.npz'))
  args = inference_args(case['beam_size'], case['look_ahead'], case['test_iteration'])
  assert model.predict(case['x'], args) == case['labels'].tolist()


@pytest.mark.parametrize('case', depth2_cases(), ids=lambda c: c['name'])
def test_cpu_decoder_depth2_matches_reference_golden(case):
  model = uisrnn_from_weights(load_weights('model_small_d2.npz'))
  args = inference_args(case['beam_size'], case['look_ahead'], case['test_iteration'])
  assert model.predict(case['x'], args) == case['labels'].tolist()


def test_cpu_decoder_matches_reference_on_toy_utterance():
  xs, labs = toy_utterances()
  model = uisrnn_from_weights(load_weights('model_toy100.npz'))
  assert model.predict(xs[3], inference_args()) == labs[3].tolist()


def test_module_surface():
  from uisrnn import uisrnn as mod
  assert hasattr(mod, 'CoreRNN') and hasattr(mod, 'BeamState') and hasattr(mod, 'UISRNN')
  import uisrnn.contrib.range_search_crp_alpha as rs
  ids = np.array(['0_0', '0_0', '0_1', '0_1', '0_1', '0_0', '0_0', '1_0', '1_0', '1_0', '1_1', '1_1', '1_1',
                  '1_0', '1_0', '1_0', '1_2', '1_2', '1_2'])
  assert rs.estimate_crp_alpha(ids, 1, 0.01) == 0.5
  state = mod.BeamState()
  state.append(torch.zeros(1, 1, 2), torch.zeros(1, 1, 3), 0)
  copy = mod.BeamState(state)
  assert copy.trace == [0] and copy.block_counts == [1] and copy.mean_set is not state.mean_set


def test_reads_reference_written_checkpoint_and_writes_the_same_format(tmp_path):
  """tests/golden/ref_checkpoint.uisrnn was written by the reference's own save()
  (oracle/make_golden.py); load() must reproduce the reference's predictions with it, and save() must
  write a file of the same structure (keys, types, shapes, dtypes)."""
  golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
  cases = np.load(os.path.join(golden, 'ref_checkpoint_cases.npz'))
  m, _, i = uisrnn.parse_arguments([])
  m.enable_cuda, m.observation_dim, m.rnn_hidden_size, m.verbosity = False, 8, 16, 0
  model = uisrnn.UISRNN(m)
  model.load(os.path.join(golden, 'ref_checkpoint.uisrnn'))
  assert model.transition_bias == float(cases['transition_bias'])
  assert model.transition_bias_denominator == float(cases['transition_bias_denominator'])
  i.beam_size, i.look_ahead, i.test_iteration = 5, 1, 2
  for k in (0, 1):
    assert model.predict(cases['x%d' % k], i) == cases['labels%d' % k].tolist()
  ours = str(tmp_path / 'ours.uisrnn')
  model.save(ours)
  a = torch.load(os.path.join(golden, 'ref_checkpoint.uisrnn'), weights_only=False)
  b = torch.load(ours, weights_only=False)

  def signature(v):
    if isinstance(v, dict):
      return {k: signature(x) for k, x in v.items()}
    if isinstance(v, torch.Tensor):
      return ('tensor', tuple(v.shape), str(v.dtype))
    if isinstance(v, np.ndarray):
      return ('ndarray', v.shape, str(v.dtype))
    return type(v).__name__
  sig_a, sig_b = signature(a), signature(b)
  # the reference's own load() casts the denominator to float (uisrnn.py:160-161), so a file saved after a
  # load holds a float where a file saved right after fit() holds the int count
  assert sig_a.pop('transition_bias_denominator') == 'int' and sig_b.pop('transition_bias_denominator') == 'float'
  assert sig_a == sig_b
  assert list(a) == list(b) and list(a['rnn_state_dict']) == list(b['rnn_state_dict'])   # same key order
  for k, v in a['rnn_state_dict'].items():
    assert torch.equal(v, b['rnn_state_dict'][k])
  assert np.array_equal(a['sigma2'], b['sigma2']) and np.array_equal(a['rnn_init_hidden'], b['rnn_init_hidden'])


@pytest.mark.parametrize('name', ['d1_b16', 'd1_b48', 'd2_b16'])
def test_fit_reproduces_reference_trajectory_cpu(name):
  """SURVEY 8(d) config 4 criterion on the CPU device: with the reference's RNG stream and initial parameters,
  fit() follows the reference's own loss trajectory (20 iterations, all three loss terms) and ends at the
  reference's parameters (tests/golden/fit_traj.npz, written by the unmodified reference)."""
  from fit_traj import run_case
  losses, want, model, final = run_case(name, enable_cuda=False)
  assert losses.shape == want.shape == (20, 3)
  assert np.max(np.abs(losses - want) / np.maximum(1.0, np.abs(want))) < 1e-4
  assert abs(model.transition_bias - float(final['transition_bias'])) < 1e-12
  sd = model.rnn_model.state_dict()
  assert np.max(np.abs(sd['linear_mean2.weight'].numpy() - final['w2'])) < 2e-5
  assert np.max(np.abs(sd['gru.weight_hh_l0'].numpy() - final['weight_hh_l0'])) < 2e-5
  assert np.max(np.abs(model.sigma2.detach().numpy() - final['sigma2'])) < 2e-6
