"""Pins the CPU oracle (oracle/uis_oracle.py) to the golden vectors produced by running the
unmodified reference (oracle/make_golden.py).  CPU only."""
import numpy as np
import pytest

from helpers import depth2_cases, load_weights, oracle_model, rel_err, small_cases, toy_utterances, uis_oracle, GOLDEN

SCORE_RTOL = 1e-5   # scores within 1e-5 relative
STATE_ATOL = 1e-5   # GRU hidden / mean within 1e-5 abs


@pytest.mark.parametrize('case', small_cases(), ids=lambda c: c['name'])
def test_small_cases_match_reference(case):
  m = oracle_model('model_small.npz')
  rec = {}
  lab = uis_oracle.predict_single(m, case['x'], beam_size=case['beam_size'],
                                  look_ahead=case['look_ahead'],
                                  test_iteration=case['test_iteration'], record=rec)
  assert lab == case['labels'].tolist()
  assert np.array_equal(rec['win'], case['win'])          # every per-step winner, in rank order
  assert np.array_equal(rec['off'], case['off'])
  assert np.array_equal(rec['nfinite'], case['nfinite'])
  assert rel_err(rec['score'], case['score']) < SCORE_RTOL
  assert rel_err(rec['final_scores'], case['final_scores']) < SCORE_RTOL
  assert np.max(np.abs(rec['final_hidden'] - case['final_hidden'])) < STATE_ATOL
  assert np.max(np.abs(rec['final_mean'] - case['final_mean'])) < STATE_ATOL
  assert np.array_equal(rec['final_blocks'], case['final_blocks'])
  assert np.array_equal(rec['full_trace'], case['full_trace'])


@pytest.mark.parametrize('case', depth2_cases(), ids=lambda c: c['name'])
def test_depth2_cases_match_reference(case):
  """Pins the oracle's stacked-GRU path (rnn_depth=2) to the unmodified reference."""
  m = oracle_model('model_small_d2.npz')
  rec = {}
  lab = uis_oracle.predict_single(m, case['x'], beam_size=case['beam_size'], look_ahead=case['look_ahead'],
                                  test_iteration=case['test_iteration'], record=rec)
  assert lab == case['labels'].tolist()
  assert np.array_equal(rec['win'], case['win'])
  assert rel_err(rec['score'], case['score']) < SCORE_RTOL
  assert np.max(np.abs(rec['final_hidden'] - case['final_hidden'])) < STATE_ATOL
  assert np.max(np.abs(rec['final_mean'] - case['final_mean'])) < STATE_ATOL


def test_per_model_constants_match_torch_free_formula():
  m = oracle_model('model_small.npz')
  mean0, hidden0 = uis_oracle.core_rnn(m, np.zeros(m.observation_dim, np.float32), m.h0)
  assert np.array_equal(mean0, m.mean0) and np.array_equal(hidden0, m.hidden0)


@pytest.mark.parametrize('idx', [0, 1, 2, 3, 4, 5])
def test_toy_utterances_match_reference(idx):
  """North-star fixture: data/toy_testing_data.npz labels (first 6 here; all 25 in the gpu suite)."""
  xs, labs = toy_utterances()
  m = oracle_model('model_toy100.npz')
  lab = uis_oracle.predict_single(m, xs[idx])
  assert lab == labs[idx].tolist()


@pytest.mark.parametrize('idx', [0, 1])
def test_toy_trace_matches_reference(idx):
  xs, _ = toy_utterances()
  g = np.load(GOLDEN + '/toy_trace.npz')
  m = oracle_model('model_toy100.npz')
  rec = {}
  uis_oracle.predict_single(m, xs[idx], record=rec)
  assert np.array_equal(rec['win'], g['u%d_win' % idx])
  assert rel_err(rec['score'], g['u%d_score' % idx]) < SCORE_RTOL
  assert np.max(np.abs(rec['final_hidden'] - g['u%d_final_hidden' % idx])) < STATE_ATOL
  assert np.max(np.abs(rec['final_mean'] - g['u%d_final_mean' % idx])) < STATE_ATOL


def test_input_validation_matches_reference():
  m = oracle_model('model_small.npz')
  with pytest.raises(TypeError):
    uis_oracle.predict_single(m, np.zeros((3, 64), np.float32))
  with pytest.raises(ValueError):
    uis_oracle.predict_single(m, np.zeros((3, 65)))
  with pytest.raises(ValueError):
    uis_oracle.predict_single(m, np.zeros(64))
  with pytest.raises(TypeError):
    uis_oracle.predict(m, 'nope')


def test_small500_matches_reference():
  """500-frame utterances (1000 beam steps) with the D=64 / H=128 model: the long-utterance pin of the oracle."""
  from uisrnn_b200.synth import synth_utt
  g = np.load(GOLDEN + '/small500.npz')
  m = oracle_model('model_small.npz')
  for seed, want in list(zip(g['seeds'], g['labels']))[:2]:
    x = synth_utt(int(seed), n_frames=500, dim=64, n_spk=4, noise=0.08)[0]
    assert uis_oracle.predict_single(m, x) == want.tolist()


def test_bench_utterance_matches_reference():
  """One utterance of bench.py's own workload (seed 100000), default shape, against the reference's labels."""
  from uisrnn_b200.synth import synth_utt
  g = np.load(GOLDEN + '/synth500_bench.npz')
  m = oracle_model('model_toy100.npz')
  assert int(g['seeds'][0]) == 100000
  assert uis_oracle.predict_single(m, synth_utt(100000)[0]) == g['labels'][0].tolist()
