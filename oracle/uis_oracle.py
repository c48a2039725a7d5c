"""CPU oracle for UIS-RNN's predict() hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package (uisrnn_b200/) never does.

What it is: a numpy restatement of the reference's beam-search inference
(/root/reference/uisrnn/uisrnn.py:388-453 `_update_beam_state`, :455-477 `_calculate_score`,
:479-562 `predict_single`, and /root/reference/uisrnn/loss_func.py:19-41 `weighted_mse_loss`),
with the redundant GRU evaluations removed (scores of the LAST look-ahead sub-step need no
GRU; winners are re-evaluated once) but with every numerically relevant quirk kept:

  * float32 GRU / MLP / MSE; log terms in float64, rounded into a float32 loss; the
    hypothesis score accumulates in float32 (uisrnn.py:411-420, 440-446, 452)
  * running mean  fl32(fl32(fl32(mu*(n-1)) + m) / n), n = visits BEFORE this one, true
    division (uisrnn.py:425-429)
  * ddCRP denominator = sum of ALL block counts + alpha (uisrnn.py:417-420, 444-446)
  * score table padded with +inf, ranked with the same numpy calls the reference uses
    (np.sort / np.trim_zeros / np.argsort on a float64 table, uisrnn.py:546-549)
  * weighted_mse_loss divides by the number of rows whose FIRST squared difference is
    non-zero (loss_func.py:36) -> inf/nan when mean[0] == x[0] exactly

PARITY PIN: tests/test_oracle_golden.py checks this module against tests/golden/*.npz, which
were produced by running the unmodified reference (oracle/make_golden.py): identical labels
and winners on every fixture, scores within 1e-5 relative, final hidden/mean within 1e-5 abs.
The GRU cell formula (PyTorch nn.GRU, gate order r,z,n; uisrnn.py:39-51) is the documented
public one:  r = s(W_ir x + b_ir + W_hr h + b_hr), z = s(W_iz x + b_iz + W_hz h + b_hz),
n = tanh(W_in x + b_in + r*(W_hn h + b_hn)), h' = (h - n)*z + n.
"""
import numpy as np

F32 = np.float32


class OracleModel:
  """Weights as float32 numpy arrays (layout = PyTorch state_dict, uisrnn.py:39-43)."""

  def __init__(self, d):
    self.depth = int(d['depth'])
    f = lambda k: np.ascontiguousarray(d[k], dtype=F32)
    self.w_ih = [f('weight_ih_l%d' % l) for l in range(self.depth)]
    self.w_hh = [f('weight_hh_l%d' % l) for l in range(self.depth)]
    self.b_ih = [f('bias_ih_l%d' % l) for l in range(self.depth)]
    self.b_hh = [f('bias_hh_l%d' % l) for l in range(self.depth)]
    self.w1, self.b1, self.w2, self.b2 = f('w1'), f('b1'), f('w2'), f('b2')
    self.h0 = f('h0').reshape(self.depth, -1)
    self.sigma2 = f('sigma2')
    self.transition_bias = float(d['transition_bias'])
    self.crp_alpha = float(d['crp_alpha'])
    self.hidden_size = self.w1.shape[0]
    self.observation_dim = self.w2.shape[0]
    # weight = 1 / (2 * sigma2): two float32 tensor ops (uisrnn.py:414, 443)
    self.w = (F32(1.0) / (F32(2.0) * self.sigma2)).astype(F32)
    # CoreRNN(zeros, rnn_init_hidden) is a per-model constant (uisrnn.py:435-439)
    self.mean0, self.hidden0 = core_rnn(self, np.zeros(self.observation_dim, F32), self.h0)

  @classmethod
  def load(cls, path):
    return cls(dict(np.load(path)))


def _sigmoid(v):
  return (F32(1.0) / (F32(1.0) + np.exp(-v, dtype=F32))).astype(F32)


def core_rnn(model, x, hidden):
  """CoreRNN.forward for one frame (uisrnn.py:45-52): x [D], hidden [depth,H]."""
  H = model.hidden_size
  inp = x.astype(F32)
  new_hidden = np.empty_like(hidden)
  for l in range(model.depth):
    gi = model.w_ih[l] @ inp + model.b_ih[l]
    gh = model.w_hh[l] @ hidden[l] + model.b_hh[l]
    r = _sigmoid(gi[:H] + gh[:H])
    z = _sigmoid(gi[H:2 * H] + gh[H:2 * H])
    n = np.tanh(gi[2 * H:] + r * gh[2 * H:], dtype=F32)
    new_hidden[l] = (hidden[l] - n) * z + n
    inp = new_hidden[l]
  act = np.maximum(model.w1 @ inp + model.b1, F32(0))
  mean = model.w2 @ act + model.b2
  return mean.astype(F32), new_hidden


def weighted_mse(mean, x, w):
  """loss_func.weighted_mse_loss for one row (loss_func.py:33-41)."""
  d2 = ((mean - x) ** 2).astype(F32)
  s = np.sum((d2 * w).astype(F32), dtype=F32)
  if d2[0] != 0:
    return F32(s)
  with np.errstate(divide='ignore', invalid='ignore'):
    return F32(s) / F32(0.0)  # zero "non-zero rows": inf (or nan if s == 0)


class Beam:
  """BeamState (uisrnn.py:55-77) with visit counters instead of trace rescans."""
  __slots__ = ('means', 'hiddens', 'visits', 'blocks', 'trace', 'nl')

  def __init__(self, src=None):
    if src is None:
      self.means, self.hiddens, self.visits, self.blocks, self.trace = [], [], [], [], []
      self.nl = 0
    else:  # shallow list copies, as the reference's copy-constructor (uisrnn.py:66-70)
      self.means, self.hiddens = list(src.means), list(src.hiddens)
      self.visits, self.blocks, self.trace = list(src.visits), list(src.blocks), list(src.trace)
      self.nl = src.nl


def _sub_step(model, beam, x, c, want_state):
  """One iteration of the loop body of _update_beam_state (uisrnn.py:405-452) on `beam`
  IN PLACE (beam must be a private copy).  Returns False for an invalid cluster index."""
  K = len(beam.means)
  p0 = model.transition_bias
  if c > K:  # invalid trace (uisrnn.py:406-408)
    beam.nl = float('inf')
    return False
  tot = sum(beam.blocks)
  if c < K:  # existing cluster (uisrnn.py:409-433)
    last = beam.trace[-1]
    loss = weighted_mse(beam.means[c], x, model.w)
    if c == last:
      loss = F32(np.float64(loss) - np.log(1 - p0))
    else:
      loss = F32(np.float64(loss) - (np.log(p0) + np.log(beam.blocks[c]) -
                                     np.log(tot + model.crp_alpha)))
    if want_state:
      m, h = core_rnn(model, x, beam.hiddens[c])
      n = beam.visits[c]
      beam.means[c] = ((beam.means[c] * F32(n - 1) + m) / F32(n)).astype(F32)
      beam.hiddens[c] = h
      beam.visits[c] = n + 1
      if c != last:
        beam.blocks[c] += 1
      beam.trace.append(c)
  else:  # new cluster (uisrnn.py:434-451)
    loss = weighted_mse(model.mean0, x, model.w)
    loss = F32(np.float64(loss) - (np.log(p0) + np.log(model.crp_alpha) -
                                   np.log(tot + model.crp_alpha)))
    if want_state:
      m, h = core_rnn(model, x, model.hidden0)
      beam.means.append(m)
      beam.hiddens.append(h)
      beam.visits.append(1)
      beam.blocks.append(1)
      beam.trace.append(c)
  # neg_likelihood += loss : int 0 at first, float32 afterwards (uisrnn.py:452)
  beam.nl = F32(loss) if isinstance(beam.nl, int) else F32(beam.nl + loss)
  return True


def update_beam_state(model, beam, chunk, cluster_seq):
  """_update_beam_state (uisrnn.py:388-453): full state update for one index tuple."""
  nb = Beam(beam)
  for i, c in enumerate(cluster_seq):
    if not _sub_step(model, nb, chunk[i], int(c), True):
      break
  return nb


def calculate_score(model, beam, chunk):
  """_calculate_score (uisrnn.py:455-477): scores of every index tuple; the state after a
  shared prefix is computed once, and the last sub-step is scored without its GRU."""
  la = chunk.shape[0]
  K = len(beam.means)
  table = np.full([K + 1 + i for i in range(la)], np.inf)

  def rec(state, depth, prefix):
    for c in range(table.shape[depth]):
      nb = Beam(state)
      last = depth == la - 1
      ok = _sub_step(model, nb, chunk[depth], c, not last)
      if not ok:
        continue  # whole sub-tree stays +inf
      if last:
        table[prefix + (c,)] = nb.nl
      else:
        rec(nb, depth + 1, prefix + (c,))

  rec(beam, 0, ())
  return table


def predict_single(model, seq, beam_size=10, look_ahead=1, test_iteration=2, record=None):
  """predict_single (uisrnn.py:479-562).  `record`, if a dict, receives the per-step
  winners / scores in the layout of oracle/make_golden.py::traced_predict."""
  if not isinstance(seq, np.ndarray) or seq.dtype != float:
    raise TypeError('test_sequence should be a numpy array of float type.')
  if seq.ndim != 2:
    raise ValueError('test_sequence must be 2-dim array.')
  n, dim = seq.shape
  if dim != model.observation_dim:
    raise ValueError('test_sequence does not match the dimension specified '
                     'by args.observation_dim.')
  tiled = np.tile(seq, (test_iteration, 1)).astype(F32)
  beams = [Beam()]
  win, sc, off, nfin = [], [], [0], []
  for t in range(0, test_iteration * n, look_ahead):
    chunk = tiled[t:t + look_ahead]
    la = chunk.shape[0]
    kmax = max(len(b.means) for b in beams)
    table = np.full([beam_size] + [kmax + 1 + i for i in range(la)], np.inf)
    for r, b in enumerate(beams):
      s = calculate_score(model, b, chunk)
      table[r] = np.pad(s, [(0, kmax - len(b.means))] * la, 'constant',
                        constant_values=np.inf)
    ranked = np.sort(table, axis=None)
    ranked[ranked == np.inf] = 0
    ranked = np.trim_zeros(ranked)
    order = np.argsort(table, axis=None)
    new_beams = []
    for r in range(min(len(ranked), beam_size)):
      idx = np.unravel_index(order[r], table.shape)
      nb = update_beam_state(model, beams[int(idx[0])], chunk, idx[1:])
      new_beams.append(nb)
      win.append([int(v) for v in idx] + [-1] * (look_ahead - la))
      sc.append(float(nb.nl))
    off.append(len(win))
    nfin.append(len(ranked))
    beams = new_beams
  best = beams[0]
  if record is not None:
    record.update(
        win=np.array(win, dtype=np.int32).reshape(-1, 1 + look_ahead),
        score=np.array(sc, dtype=np.float64), off=np.array(off, dtype=np.int64),
        nfinite=np.array(nfin, dtype=np.int64),
        final_scores=np.array([float(b.nl) for b in beams]),
        final_mean=np.stack(best.means), final_hidden=np.stack(best.hiddens),
        final_blocks=np.array(best.blocks, dtype=np.int64),
        full_trace=np.array(best.trace, dtype=np.int64))
  return [int(c) for c in best.trace[-n:]]


def predict(model, seqs, **kw):
  """predict (uisrnn.py:564-590)."""
  if isinstance(seqs, np.ndarray):
    return predict_single(model, seqs, **kw)
  if isinstance(seqs, list):
    return [predict_single(model, s, **kw) for s in seqs]
  raise TypeError('test_sequences should be either a list or numpy array.')
