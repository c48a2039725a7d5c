"""Loss terms of UIS-RNN training; `weighted_mse_loss` is also the per-candidate Gaussian score
of the beam search (`/root/reference/uisrnn/loss_func.py`).  Same values as the reference, written
without the dense `diag(weight)` matrix product (loss_func.py:37-39 builds a D x D matrix per call)."""
import torch


def weighted_mse_loss(input_tensor, target_tensor, weight=1):
  """sum_rows sum_d weight_d * (input - target)_d^2 / #{rows whose FIRST squared difference != 0}.

  The odd normaliser (rows are counted by their first column only, loss_func.py:36) is kept: it
  is what the training loss and the beam-search score are defined by.
  """
  dim = input_tensor.size()[-1]
  squared = ((input_tensor - target_tensor) ** 2).view(-1, dim)
  rows = float(squared.size()[0])
  non_zero_rows = torch.sum(squared[:, 0] != 0).float()
  weight = weight.float().view(-1) if torch.is_tensor(weight) else torch.full((dim,), float(weight))
  weighted = squared * weight
  return torch.mean(weighted) * weight.nelement() * rows / non_zero_rows


def sigma2_prior_loss(num_non_zero, sigma_alpha, sigma_beta, sigma2):
  """Inverse-gamma prior on sigma^2 (loss_func.py:44-60)."""
  shape_term = (2 * sigma_alpha + num_non_zero + 2) / (2 * num_non_zero) * torch.log(sigma2)
  scale_term = sigma_beta / (sigma2 * num_non_zero)
  return shape_term.sum() + scale_term.sum()


def regularization_loss(params, weight):
  """weight * sum of the (non-squared) L2 norms of the parameter tensors (loss_func.py:63-76)."""
  total = 0
  for param in params:
    total = total + torch.norm(param)
  return weight * total
