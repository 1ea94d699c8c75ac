"""DDIM sampler (gaussian_diffusion.py:134-284): known-answer pins of the CPU restatement.
The reference ships no vectors for it; these follow from its formulas."""
import numpy as np
import pytest
import torch

from oracle import umd_oracle as O


def _gd(name="cosine"):
  return O.gaussian_diffusion_tables(name, 1000)


def test_timesteps_follow_python_floor_division():
  ts = O.ddim_timesteps(1000, 250)          # gaussian_diffusion.py:236-237: arange(999, 0, -4) + [0]
  assert len(ts) == 251 and ts[0] == 999 and ts[1] == 995 and ts[-2] == 3 and ts[-1] == 0
  ts = O.ddim_timesteps(1000, 125)          # configs/ae_i1k.py sampling_timesteps=125 -> stride 8
  assert len(ts) == 126 and ts[1] == 991 and ts[-2] == 7
  assert O.ddim_timesteps(1000, 300)[1] == 995   # -1000 // 300 == -4 (floor), not -3


@pytest.mark.parametrize("sched", ["cosine", "linear"])
def test_eta0_step_moves_along_the_same_x0_eps_line(sched):
  """With the true eps as model output, pred_xstart = x0 and the eta = 0 update lands exactly on
  q_sample(x0, t_next, eps) (DDIM consistency)."""
  gd = _gd(sched)
  g = torch.Generator().manual_seed(0)
  x0 = torch.rand(4, 8, 8, 3, generator=g, dtype=torch.float64) * 2 - 1
  eps = torch.randn(4, 8, 8, 3, generator=g, dtype=torch.float64)
  t = torch.tensor([[999], [500], [37], [4]])
  t_next = torch.tensor([[995], [496], [33], [0]])
  gd64 = {k: np.asarray(v, dtype=np.float64) for k, v in gd.items()}
  x_t = O.q_sample(gd64, x0, t, eps)
  out = O.ddim_sample(gd64, lambda x_t, t: eps, x_t, t, t_next, torch.randn(4, 8, 8, 3, generator=g, dtype=torch.float64), eta=0.0)
  assert torch.allclose(out["pred_xstart"], x0, atol=1e-6)   # 1/sqrt(abar_999) ~ 1e2 amplifies rounding
  assert torch.allclose(out["sample"], O.q_sample(gd64, x0, t_next, eps), atol=1e-6)


def test_final_call_returns_pred_xstart_and_ignores_noise():
  gd = _gd()
  g = torch.Generator().manual_seed(1)
  x = torch.randn(3, 4, 4, 3, generator=g)
  e = torch.randn(3, 4, 4, 3, generator=g)
  t0 = torch.zeros(3, 1, dtype=torch.int64)
  a = O.ddim_sample(gd, lambda x_t, t: e, x, t0, None, torch.randn(3, 4, 4, 3, generator=g), eta=1.0)
  b = O.ddim_sample(gd, lambda x_t, t: e, x, t0, None, torch.zeros(3, 4, 4, 3), eta=1.0)
  assert torch.equal(a["sample"], b["sample"])                      # (t > 0) gates the noise, :209-210
  assert torch.allclose(a["sample"], a["pred_xstart"], atol=1e-6)   # alphas_cumprod_prev[0] = 1


def test_sigma_is_ancestral_at_eta1_and_unit_stride():
  """eta = 1 with t_next = t - 1 is ancestral sampling: sigma^2 = beta_t (1 - abar_{t-1}) / (1 - abar_t).  (The
  reference's own `posterior_variance` table, gaussian_diffusion.py:43, is a different expression — it divides by
  1 - abar_T — and is not used by the sampler.)"""
  gd = _gd()
  t = torch.tensor([700, 20])
  ab = torch.as_tensor(gd["alphas_cumprod"])[t]
  abp = torch.as_tensor(gd["alphas_cumprod"])[t - 1]
  beta = torch.as_tensor(gd["betas"])[t]
  sigma2 = (1 - abp) / (1 - ab) * (1 - ab / abp)
  assert torch.allclose(sigma2, beta * (1 - abp) / (1 - ab), rtol=1e-9)
  # and the sampler uses exactly that sigma: with zero eps-model and x = 0 the sample is sigma * noise
  z = torch.zeros(2, 2, 2, 3, dtype=torch.float64)
  nz = torch.ones(2, 2, 2, 3, dtype=torch.float64)
  gd64 = {k: np.asarray(v, dtype=np.float64) for k, v in gd.items()}
  out = O.ddim_sample(gd64, lambda x_t, t: z, z, t.reshape(-1, 1), (t - 1).reshape(-1, 1), nz, eta=1.0)
  assert torch.allclose(out["sample"][:, 0, 0, 0] ** 2, sigma2, rtol=1e-9)


def test_clip_denoised_bounds_pred_xstart():
  gd = _gd()
  g = torch.Generator().manual_seed(2)
  x = torch.randn(2, 4, 4, 3, generator=g) * 3
  e = torch.randn(2, 4, 4, 3, generator=g)
  t = torch.tensor([[900], [100]])
  out = O.ddim_sample(gd, lambda x_t, t: e, x, t, t - 4, torch.zeros_like(x), clip_denoised=True, eta=0.0)
  assert float(out["pred_xstart"].abs().max()) <= 1.0


def test_loop_consumes_noises_in_reference_order():
  gd = _gd()
  g = torch.Generator().manual_seed(3)
  steps = 5
  noises = [torch.randn(2, 4, 4, 3, generator=g) for _ in range(steps + 2)]
  calls = []

  def apply_fn(*, x_t, t, y=None, cfg_scale=None):
    calls.append(int(t[0, 0]))
    return 0.1 * x_t

  out = O.ddim_sample_loop(gd, apply_fn, noises, sampling_steps=steps, eta=0.5)
  assert calls == O.ddim_timesteps(1000, steps)[:steps] + [0]
  assert out.shape == (2, 4, 4, 3) and bool(torch.isfinite(out).all())
