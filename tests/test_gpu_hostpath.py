"""The host-buffer entry point uis_predict(): float64 rows travel in staging chunks on a copy stream while the cast
and the input projection of the previous chunk run; a list that does not fit the device at once is decoded in
groups.  Neither may change a label: both are compared with the one-chunk / one-group call and with the
reference's golden labels."""
import numpy as np
import pytest

from helpers import load_weights, toy_utterances

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def toy_native():
  from uisrnn_b200 import native
  return native.NativeModel(load_weights('model_toy100.npz'))


def _ragged():
  xs, labs = toy_utterances()
  xs = xs[:12]
  labs = [l for l in labs[:12]]
  # ragged list with an empty utterance and a one-frame utterance in the middle
  xs = xs[:5] + [np.zeros((0, 256))] + xs[5:9] + [xs[9][:1].copy()] + xs[9:]
  return xs, labs


def test_chunked_copies_do_not_change_labels(toy_native, monkeypatch):
  xs, labs = _ragged()
  want = toy_native.predict(xs)
  st = toy_native.stats()
  assert st['chunks'] == 1 and st['groups'] == 1 and st['h2d_ms'] > 0 and st['host_ms'] > 0
  golden = labs[:5] + [np.zeros(0, np.int32)] + labs[5:9] + [None] + labs[9:]
  for got, ref in zip(want, golden):
    if ref is not None:
      assert np.array_equal(got, ref)
  # 256-row chunks (the floor): utterances straddle chunk boundaries, the ring of 3 staging slots wraps many times
  monkeypatch.setenv('UISRNN_B200_CHUNK_MB', '0')
  got = toy_native.predict(xs)
  st = toy_native.stats()
  rows = sum(len(x) for x in xs)
  assert st['chunks'] == -(-rows // 256) and st['chunks'] > 3   # more chunks than staging slots: the ring wraps
  assert st['kernel_launches'] == 1 + 2 * st['chunks']
  for a, b in zip(got, want):
    assert np.array_equal(a, b)


def test_lists_larger_than_the_row_budget_are_decoded_in_groups(toy_native, monkeypatch):
  xs, _ = _ragged()
  want = toy_native.predict(xs)
  whole = toy_native.stats()
  longest = max(len(x) for x in xs)
  monkeypatch.setenv('UISRNN_B200_MAX_ROWS', str(2 * longest))
  got = toy_native.predict(xs)
  st = toy_native.stats()
  assert st['groups'] > 2 and st['utterances'] == len(xs) and st['frames'] == whole['frames']
  assert st['beam_steps'] == whole['beam_steps']
  for a, b in zip(got, want):
    assert np.array_equal(a, b)
  # a budget smaller than one utterance: every utterance becomes its own group
  monkeypatch.setenv('UISRNN_B200_MAX_ROWS', '1')
  got = toy_native.predict(xs)
  assert toy_native.stats()['groups'] == len(xs)
  for a, b in zip(got, want):
    assert np.array_equal(a, b)


def test_pageable_and_pinned_inputs_agree(toy_native, monkeypatch):
  """Pageable arrays (what numpy gives) travel through the library's pinned staging ring, filled by host threads;
  pinned arrays are copied directly.  Forced both ways, small chunks so that the ring wraps."""
  import torch
  xs, _ = _ragged()
  monkeypatch.setenv('UISRNN_B200_CHUNK_MB', '0')
  monkeypatch.setenv('UISRNN_B200_HOST_STAGING', '0')
  want = toy_native.predict(xs)
  assert toy_native.stats()['staged'] == 0
  monkeypatch.setenv('UISRNN_B200_HOST_STAGING', '1')
  for threads in ('1', '5'):
    monkeypatch.setenv('UISRNN_B200_COPY_THREADS', threads)   # (the pool is created once per model: the first value counts)
    got = toy_native.predict(xs)
    st = toy_native.stats()
    assert st['staged'] == 1 and st['chunks'] > 3
    for a, b in zip(got, want):
      assert np.array_equal(a, b)
  monkeypatch.delenv('UISRNN_B200_HOST_STAGING')
  pinned = [torch.from_numpy(x).pin_memory().numpy() if len(x) else x for x in xs]
  got = toy_native.predict(pinned)
  for a, b in zip(got, want):
    assert np.array_equal(a, b)
