"""Derivation check (CPU, torch): manual forward/backward of one fit_concatenated iteration in the
padded time-major layout planned for the CUDA trainer, against torch autograd on the reference
formulation (uisrnn_b200/uisrnn.py::fit_concatenated).  Not part of the product or the tests."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from uisrnn_b200 import loss_func
from uisrnn_b200.uisrnn import CoreRNN
torch.manual_seed(0); np.random.seed(0)
D, H, B = 6, 10, 5
lens = np.array([7, 6, 6, 4, 2])  # incl. the leading zero frame, sorted descending
L = lens[0]
x = np.zeros((L, B, D), np.float32)
for b, n in enumerate(lens):
  x[1:n, b] = np.random.randn(n - 1, D)
x = torch.from_numpy(x)
rnn = CoreRNN(D, H, 1, D)
h0 = torch.nn.Parameter(torch.randn(1, 1, H) * 0.1)
sigma2 = torch.nn.Parameter(0.1 * torch.ones(D) + 0.01 * torch.rand(D))
alpha_s, beta_s, reg = 1.0, 1.0, 1e-5

# ---------------- autograd reference
packed = torch.nn.utils.rnn.pack_padded_sequence(x, lens, batch_first=False)
truth = x[1:]
mean, _ = rnn(packed, h0.repeat(1, B, 1))
steps = torch.arange(1, mean.size(0) + 1).float()
mean = torch.cumsum(mean, dim=0) * (1.0 / steps).view(-1, 1, 1)
mask = (truth != 0).float()
loss1 = loss_func.weighted_mse_loss(mask * mean[:-1], truth, 1 / (2 * sigma2))
res2 = ((mask * mean[:-1] - truth) ** 2).view(-1, D)
nnz = torch.sum((res2 != 0).float(), dim=0).squeeze()
loss2 = loss_func.sigma2_prior_loss(nnz, alpha_s, beta_s, sigma2)
loss3 = loss_func.regularization_loss(rnn.parameters(), reg)
(loss1 + loss2 + loss3).backward()
ref = {n: p.grad.clone() for n, p in rnn.named_parameters()}
ref['h0'] = h0.grad.clone(); ref['sigma2'] = sigma2.grad.clone()

# ---------------- manual
with torch.no_grad():
  Wih, Whh = rnn.gru.weight_ih_l0, rnn.gru.weight_hh_l0
  bih, bhh = rnn.gru.bias_ih_l0, rnn.gru.bias_hh_l0
  W1, b1, W2, b2 = rnn.linear_mean1.weight, rnn.linear_mean1.bias, rnn.linear_mean2.weight, rnn.linear_mean2.bias
  lens_t = torch.from_numpy(lens)
  valid = (torch.arange(L).view(-1, 1) < lens_t.view(1, -1)).float()          # [L,B]
  gi = x @ Wih.t() + bih                                                       # [L,B,3H]
  hs = torch.zeros(L + 1, B, H); hs[0] = h0.view(1, H).repeat(B, 1)            # hs[t+1] = h_t
  R = torch.zeros(L, B, H); Z = torch.zeros(L, B, H); Nn = torch.zeros(L, B, H); HN = torch.zeros(L, B, H)
  for t in range(L):
    gh = hs[t] @ Whh.t() + bhh
    r = torch.sigmoid(gi[t, :, :H] + gh[:, :H]); z = torch.sigmoid(gi[t, :, H:2*H] + gh[:, H:2*H])
    hn = gh[:, 2*H:]; n = torch.tanh(gi[t, :, 2*H:] + r * hn)
    hnew = (hs[t] - n) * z + n
    v = valid[t].view(-1, 1)
    hs[t + 1] = v * hnew + (1 - v) * hs[t]
    R[t], Z[t], Nn[t], HN[t] = r, z, n, hn
  out = hs[1:] * valid.unsqueeze(-1)                                           # padded outputs are zero
  z1 = out @ W1.t() + b1; a1 = torch.relu(z1); mu = a1 @ W2.t() + b2           # [L,B,D]
  inv = 1.0 / torch.arange(1, L + 1).float()
  avg = torch.cumsum(mu, 0) * inv.view(-1, 1, 1)
  pred = mask * avg[:-1]
  diff = pred - truth
  sq = diff ** 2
  w = 1 / (2 * sigma2)
  rows = float((L - 1) * B)
  nz = (sq.view(-1, D)[:, 0] != 0).float().sum()
  l1 = (sq * w).mean() * D * rows / nz
  nd = (sq.view(-1, D) != 0).float().sum(0)
  l2 = ((2 * alpha_s + nd + 2) / (2 * nd) * torch.log(sigma2)).sum() + (beta_s / (sigma2 * nd)).sum()
  norms = [p.norm() for p in rnn.parameters()]
  l3 = reg * sum(norms)
  print('loss1 %.6f/%.6f loss2 %.6f/%.6f loss3 %.8f/%.8f' % (l1, loss1, l2, loss2, l3, loss3))
  # backward
  davg = torch.zeros(L, B, D)
  davg[:-1] = mask * 2 * diff * w / nz
  dmu = torch.flip(torch.cumsum(torch.flip(davg * inv.view(-1, 1, 1), [0]), 0), [0])
  g_sigma2 = -(sq.view(-1, D).sum(0) / nz) / (2 * sigma2 ** 2) + ((2 * alpha_s + nd + 2) / (2 * nd)) / sigma2 - beta_s / (sigma2 ** 2 * nd)
  dmu2 = dmu.view(-1, D); a1f = a1.view(-1, H); outf = out.reshape(-1, H)
  gW2 = dmu2.t() @ a1f; gb2 = dmu2.sum(0)
  da1 = dmu2 @ W2; dz1 = da1 * (z1.view(-1, H) > 0).float()
  gW1 = dz1.t() @ outf; gb1 = dz1.sum(0)
  dout = (dz1 @ W1).view(L, B, H) * valid.unsqueeze(-1)
  dGi = torch.zeros(L, B, 3 * H); dGh = torch.zeros(L, B, 3 * H)
  carry = torch.zeros(B, H)
  for t in reversed(range(L)):
    v = valid[t].view(-1, 1)
    dh = (dout[t] + carry) * v
    r, z, n, hn, hp = R[t], Z[t], Nn[t], HN[t], hs[t]
    dn = dh * (1 - z); dz = dh * (hp - n)
    dan = dn * (1 - n * n); dr = dan * hn; dhn = dan * r
    daz = dz * z * (1 - z); dar = dr * r * (1 - r)
    dGi[t] = torch.cat([dar, daz, dan], 1); dGh[t] = torch.cat([dar, daz, dhn], 1)
    carry = dh * z + dGh[t] @ Whh + (1 - v) * carry
  gWih = dGi.view(-1, 3 * H).t() @ x.view(-1, D); gWhh = dGh.view(-1, 3 * H).t() @ hs[:-1].reshape(-1, H)
  gbih = dGi.view(-1, 3 * H).sum(0); gbhh = dGh.view(-1, 3 * H).sum(0)
  gh0 = carry.sum(0).view(1, 1, H)
  man = {'gru.weight_ih_l0': gWih, 'gru.weight_hh_l0': gWhh, 'gru.bias_ih_l0': gbih, 'gru.bias_hh_l0': gbhh,
         'linear_mean1.weight': gW1, 'linear_mean1.bias': gb1, 'linear_mean2.weight': gW2, 'linear_mean2.bias': gb2}
  for (name, p), nm in zip(rnn.named_parameters(), norms):
    man[name] = man[name] + reg * p / nm
  man['h0'] = gh0; man['sigma2'] = g_sigma2
  for k in ref:
    err = (man[k] - ref[k]).abs().max().item() / (ref[k].abs().max().item() + 1e-12)
    print('%-22s rel err %.2e' % (k, err))
