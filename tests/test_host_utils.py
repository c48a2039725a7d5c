"""Host-side mirror of the reference's utils / evals / arguments / loss_func (CPU only).
Modelled on the reference's tests/utils_test.py, evals_test.py (behaviour, not code)."""
import random

import numpy as np
import pytest
import torch

import uisrnn
from uisrnn import arguments, evals, loss_func, utils


# ---------------------------------------------------------------- arguments
def test_parse_arguments_defaults_match_reference():
  m, t, i = uisrnn.parse_arguments([])
  assert (m.observation_dim, m.rnn_hidden_size, m.rnn_depth, m.rnn_dropout) == (256, 512, 1, 0.2)
  assert m.transition_bias is None and m.sigma2 is None and m.crp_alpha == 1.0
  assert m.verbosity == 3 and m.enable_cuda is True
  assert (t.optimizer, t.learning_rate, t.train_iteration, t.batch_size) == ('adam', 1e-3, 20000, 10)
  assert (t.num_permutations, t.sigma_alpha, t.sigma_beta) == (10, 1.0, 1.0)
  assert (t.regularization_weight, t.grad_max_norm, t.enforce_cluster_id_uniqueness) == (1e-5, 5.0, True)
  assert (i.beam_size, i.look_ahead, i.test_iteration) == (10, 1, 2)


def test_parse_arguments_flags_and_short_options():
  m, t, i = uisrnn.parse_arguments(['--enable_cuda=False', '-l', '0.01', '-t', '7', '-b', '3', '-s', '4',
                                    '--look_ahead', '2', '-r', '0.5', '--sigma2', '0.3'])
  assert m.enable_cuda is False and m.sigma2 == 0.3
  assert (t.learning_rate, t.train_iteration, t.batch_size, t.regularization_weight) == (0.01, 7, 3, 0.5)
  assert (i.beam_size, i.look_ahead) == (4, 2)
  assert not hasattr(m, 'beam_size') and not hasattr(i, 'learning_rate')
  with pytest.raises(SystemExit):
    uisrnn.parse_arguments(['--no_such_flag', '1'])
  with pytest.raises(SystemExit):
    uisrnn.parse_arguments(['--enable_cuda=maybe'])
  assert arguments.str2bool('Yes') is True and arguments.str2bool('0') is False


# ---------------------------------------------------------------- utils
def test_enforce_cluster_id_uniqueness():
  ids = [['a', 'b'], np.array(['a', 'c'])]
  out = utils.enforce_cluster_id_uniqueness(ids)
  assert [len(x) for x in out] == [2, 2]
  assert out[0][0].endswith('_a') and out[1][1].endswith('_c')
  assert out[0][0].split('_')[0] == out[0][1].split('_')[0] != out[1][0].split('_')[0]
  assert len(out[0][0].split('_')[0]) == 6
  with pytest.raises(TypeError):
    utils.enforce_cluster_id_uniqueness('ab')
  with pytest.raises(TypeError):
    utils.enforce_cluster_id_uniqueness([('a',)])


def test_generate_random_string_consumes_one_choice_per_char():
  random.seed(3)
  a = utils.generate_random_string(6)
  random.seed(3)
  b = ''.join(random.choice('ABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789') for _ in range(6))
  assert a == b


def test_concatenate_training_data():
  seqs = [np.zeros((3, 2)), np.ones((2, 2))]
  ids = [['a', 'b', 'a'], np.array(['a', 'a'])]
  x, y = utils.concatenate_training_data(seqs, ids, False, False)
  assert x.shape == (5, 2) and y == ['a', 'b', 'a', 'a', 'a']
  x, y = utils.concatenate_training_data(seqs, ids, True, False)
  assert len(set(y)) == 3 and y[0] == y[2] != y[3]
  random.seed(0)
  x, y = utils.concatenate_training_data(seqs, ids, False, True)
  assert sorted(x.sum(axis=1).tolist()) == [0, 0, 0, 2, 2]
  with pytest.raises(TypeError):
    utils.concatenate_training_data(seqs, 'ab')
  with pytest.raises(ValueError):
    utils.concatenate_training_data(seqs, ids[:1])
  with pytest.raises(ValueError):
    utils.concatenate_training_data([np.zeros((3, 2)), np.zeros((2, 3))], ids)
  with pytest.raises(ValueError):
    utils.concatenate_training_data(seqs, [['a'], ['a', 'a']])


def test_sample_permuted_segments_preserves_blocks():
  np.random.seed(0)
  idx = np.array([1, 2, 6, 10, 11, 12])
  for s in utils.sample_permuted_segments(idx, 20):
    assert sorted(s.tolist()) == idx.tolist()
    text = ','.join(map(str, s.tolist()))
    assert '1,2' in text and '10,11,12' in text
  assert utils.sample_permuted_segments(np.array([5]), 2)[0].tolist() == [5]


def test_resize_sequence_exact():
  seq = np.arange(12, dtype=float).reshape(6, 2)
  ids = np.array(['b', 'a', 'b', 'b', 'a', 'c'])
  subs, lens = utils.resize_sequence(seq, ids)
  assert lens == [3, 4, 2]                      # np.unique order a, b, c; length + 1
  assert subs[0].tolist() == seq[[1, 4]].tolist() and subs[1].tolist() == seq[[0, 2, 3]].tolist()
  np.random.seed(1)
  subs, lens = utils.resize_sequence(seq, ids, num_permutations=3)
  assert lens == [3, 3, 3, 4, 4, 4, 2, 2, 2] and len(subs) == 9
  assert sorted(subs[3][:, 0].tolist()) == [0.0, 4.0, 6.0]


def test_resize_sequence_rng_stream_matches_per_cluster_permutations():
  """Same np.random call sequence as the reference: one permutation(#runs) per speaker per copy."""
  ids = np.array(list('aabbaaccbb'))
  seq = np.arange(10, dtype=float).reshape(10, 1)
  np.random.seed(5)
  subs, _ = utils.resize_sequence(seq, ids, num_permutations=2)
  np.random.seed(5)
  expected = []
  for runs in ([[0, 1], [4, 5]], [[2, 3], [8, 9]], [[6, 7]]):
    for _ in range(2):
      order = np.random.permutation(len(runs))
      expected.append(np.concatenate([runs[i] for i in order]).astype(float))
  assert [s[:, 0].tolist() for s in subs] == [e.tolist() for e in expected]


def test_pack_sequence_shapes_and_shift():
  np.random.seed(0)
  subs = [np.full((2, 3), 1.0), np.full((4, 3), 2.0), np.full((1, 3), 3.0)]
  lens = [3, 5, 2]
  packed, truth = utils.pack_sequence(subs, lens, 4, 3, 'cpu')
  padded, out_lens = torch.nn.utils.rnn.pad_packed_sequence(packed)
  assert padded.shape[1] == 4 and truth.shape == (padded.shape[0] - 1, 4, 3)
  assert torch.all(padded[0] == 0) and torch.equal(truth, padded[1:])
  assert out_lens.tolist() == sorted(out_lens.tolist(), reverse=True)


def test_estimate_transition_bias():
  bias, denom = utils.estimate_transition_bias([['a', 'a', 'b'], np.array(['c', 'd'])])
  assert denom == 2 + 3 and bias == (1 + 2) / 5
  bias, _ = utils.estimate_transition_bias([['a'] * 10])
  assert 0 < bias < 1
  bias, denom = utils.estimate_transition_bias([[]], smooth=2)
  assert bias == 0.5 and denom == 4


def test_output_result_appends_file(tmp_path, monkeypatch):
  monkeypatch.chdir(tmp_path)
  m, t, _ = uisrnn.parse_arguments([])
  text = uisrnn.output_result(m, t, [(1.0, 10), (0.5, 4)])
  assert 'averaged accuracy: 0.750000' in text and text.count('\n    ') == 2
  assert (tmp_path / 'layer_512_1_0.2_result.txt').read_text() == text


# ---------------------------------------------------------------- evals
def test_sequence_match_accuracy():
  assert evals.compute_sequence_match_accuracy([0, 0, 1, 2, 2], [3, 3, 4, 4, 1]) == 0.8
  assert evals.compute_sequence_match_accuracy(['a', 'b'], [7, 9]) == 1.0
  assert evals.compute_sequence_match_accuracy([0, 0, 0, 0], [0, 1, 0, 1]) == 0.5
  a, b = [0, 1, 1, 2, 0, 2], [1, 1, 0, 2, 2, 2]
  assert evals.compute_sequence_match_accuracy(a, b) == evals.compute_sequence_match_accuracy(b, a)
  with pytest.raises(TypeError):
    evals.compute_sequence_match_accuracy(np.array([0]), [0])
  with pytest.raises(ValueError):
    evals.compute_sequence_match_accuracy([], [])
  with pytest.raises(ValueError):
    evals.compute_sequence_match_accuracy([0], [0, 1])
  assert evals.get_list_inverse_index(['x', 'y']) == {'x': 0, 'y': 1}
  with pytest.raises(TypeError):
    evals.get_list_inverse_index(('x',))


# ---------------------------------------------------------------- loss_func
def test_weighted_mse_loss_matches_dense_formula():
  torch.manual_seed(0)
  a, b, w = torch.randn(5, 3, 4), torch.randn(5, 3, 4), torch.rand(4) + 0.1
  b[1, 2] = a[1, 2]  # a row whose first squared difference is zero is not counted
  got = loss_func.weighted_mse_loss(a, b, w)
  sq = ((a - b) ** 2).view(-1, 4)
  dense = torch.mean(torch.mm(sq, torch.diag(w))) * 4 * 15 / 14
  assert torch.allclose(got, dense, rtol=1e-6)
  one = loss_func.weighted_mse_loss(a[0, 0], b[0, 0], w)
  assert torch.allclose(one, (sq[0] * w).sum(), rtol=1e-6)


def test_sigma2_prior_and_regularization():
  n = torch.tensor([4.0, 8.0])
  s2 = torch.tensor([0.5, 2.0])
  want = ((2 * 1.0 + n + 2) / (2 * n) * torch.log(s2)).sum() + (3.0 / (s2 * n)).sum()
  assert torch.allclose(loss_func.sigma2_prior_loss(n, 1.0, 3.0, s2), want)
  ps = [torch.ones(3), torch.full((2, 2), 2.0)]
  assert torch.allclose(loss_func.regularization_loss(ps, 0.1), torch.tensor(0.1 * (3 ** 0.5 + 4.0)))


def test_resize_indices_and_batch_sampler_mirror_the_data_path():
  """The device-resident training path replaces sub-sequence copies by row indices and the packed batch
  by the ids of its columns: same RNG stream, same rows."""
  from uisrnn_b200 import utils
  rng = np.random.default_rng(0)
  ids = np.array(['%d' % v for v in np.repeat(rng.integers(0, 7, 60), rng.integers(1, 9, 60))])
  seq = rng.normal(size=(len(ids), 5))
  for perms in (None, 1, 4):
    np.random.seed(3)
    subs, lens_a = utils.resize_sequence(seq, ids, perms)
    np.random.seed(3)
    index_lists, lens_b = utils.resize_indices(ids, perms)
    assert lens_a == lens_b
    assert all(np.array_equal(x, seq[ix]) for x, ix in zip(subs, index_lists))
    np.random.seed(5)
    x, lengths = utils.pack_batch(subs, lens_a, 6, 5)
    state_after_pack = np.random.get_state()[1].copy()
    np.random.seed(5)
    chosen, lengths2 = utils.BatchSampler(lens_a, 6).draw()
    assert np.array_equal(state_after_pack, np.random.get_state()[1])   # consumed the same random numbers
    assert np.array_equal(lengths, lengths2)
    for column, k in enumerate(chosen):
      assert np.array_equal(x[1:lengths[column], column, :], subs[k])
      assert not x[0, column].any() and not x[lengths[column]:, column].any()


def test_param_order_and_dropout_mask_contract():
  """Host side of the device trainer's contracts: the tensor order uis_trainer_create takes for stacked layers, and
  the dropout keep mask (a pure function of seed, iteration, layer, element) restated in numpy."""
  from uisrnn_b200 import native
  assert native.param_order(1) == native.PARAM_ORDER and len(native.PARAM_ORDER) == 10
  order = native.param_order(3)
  assert len(order) == 4 * 3 + 6
  assert order[:4] == ('gru.weight_ih_l0', 'gru.weight_hh_l0', 'gru.bias_ih_l0', 'gru.bias_hh_l0')
  assert order[8:12] == ('gru.weight_ih_l2', 'gru.weight_hh_l2', 'gru.bias_ih_l2', 'gru.bias_hh_l2')
  assert order[-6:] == ('linear_mean1.weight', 'linear_mean1.bias', 'linear_mean2.weight', 'linear_mean2.bias',
                        'rnn_init_hidden', 'sigma2')
  a = native.dropout_keep_mask(0x1234567890, 3, 1, 200000, 0.2)
  b = native.dropout_keep_mask(0x1234567890, 3, 1, 200000, 0.2)
  assert a.dtype == bool and np.array_equal(a, b)                       # reproducible
  assert abs(a.mean() - 0.8) < 0.01                                     # keeps 1 - p of the elements
  for other in (native.dropout_keep_mask(0x1234567890, 4, 1, 200000, 0.2),    # another iteration,
                native.dropout_keep_mask(0x1234567890, 3, 0, 200000, 0.2),    # another layer,
                native.dropout_keep_mask(0x1234567891, 3, 1, 200000, 0.2)):   # another seed: another mask
    assert 0.3 < (a != other).mean() < 0.34                             # two independent masks differ in 2 p (1 - p)
  assert native.dropout_keep_mask(7, 0, 0, 1000, 0.0).all()
  # a prefix of a longer mask is the shorter mask (element i depends on i only)
  assert np.array_equal(native.dropout_keep_mask(5, 1, 0, 100, 0.5), native.dropout_keep_mask(5, 1, 0, 1000, 0.5)[:100])
