"""GPU parity of the DDIM sampler (gaussian_diffusion.py:134-284) against the CPU oracle:
the fused update kernel (umd_ddim_step) on random inputs, the sampling loop through the engine's forward with
classifier-free guidance, and the x0/eps-line consistency property at the full bench batch."""
import numpy as np
import pytest
import torch

from tests import util as U
from oracle import umd_oracle as O

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _tables(sched="cosine"):
  from small_vision_b200.diffusion import create_gaussian_diffusion, to_device
  gd = create_gaussian_diffusion(sched, 1000)
  return gd, to_device(gd, DEV)


@pytest.mark.parametrize("channels,sched", [(3, "cosine"), (4, "linear")])
@pytest.mark.parametrize("eta,clip,eps_pred,cfg_scale,with_next", [
    (0.0, False, True, None, True), (1.0, False, True, None, True), (0.5, True, True, None, True),
    (1.0, False, False, None, True), (0.3, False, True, 1.5, True), (0.7, True, False, 4.0, True),
    (1.0, False, True, None, False)])
def test_ddim_step_kernel_matches_oracle(channels, sched, eta, clip, eps_pred, cfg_scale, with_next):
  from small_vision_b200 import diffusion as Dm
  gd, gdd = _tables(sched)
  n, H = 5, 8
  g = torch.Generator().manual_seed(int(eta * 10) + channels + (7 if clip else 0))
  x = torch.randn(n, H, H, channels, generator=g)
  noise = torch.randn(n, H, H, channels, generator=g)
  pred = torch.randn(2 * n if cfg_scale is not None else n, H, H, 2 * channels, generator=g)
  t = torch.tensor([[999], [640], [17], [1], [0]], dtype=torch.int32)
  t_next = torch.tensor([[995], [636], [13], [0], [0]], dtype=torch.int32) if with_next else None

  def combined(p):
    if cfg_scale is None:
      return p
    co, un = p[:n], p[n:]
    return un + cfg_scale * (co - un)           # ae.py:192-195

  def o_apply(*, x_t, t):
    p = combined(pred)
    return p[..., channels:] if eps_pred else O.predict_eps_from_xstart(gd, x_t, t, p[..., :channels])  # train_ae.py:479-482

  ref = O.ddim_sample(gd, o_apply, x, t.long(), None if t_next is None else t_next.long(), noise, clip_denoised=clip, eta=eta)
  got = Dm.ddim_sample(gdd, lambda **kw: Dm.RawPred(pred.to(DEV), cfg_scale, eps_pred), x.to(DEV), t.to(DEV),
                       None if t_next is None else t_next.to(DEV), 0, clip_denoised=clip, eta=eta, noise=noise.to(DEV))
  torch.cuda.synchronize()
  for k in ("sample", "pred_xstart"):
    a, b = got[k].cpu(), ref[k]
    assert torch.isfinite(a).all()
    # fp32 arithmetic in both.  With the x0 head (train_ae.py:482) the round trip x0 -> eps -> x0 subtracts two
    # numbers ~ 1e2 |x| at t = 999 (1/sqrt(abar_999) ~ 1e2), so fp32 evaluation order (FMA contraction) shows at 1e-4
    tol = 2e-5 if eps_pred else 2e-3
    assert U.rel_l2(a, b) < tol, (k, U.rel_l2(a, b))
  # the (t > 0) gate: rows with t == 0 must not depend on the noise
  assert torch.allclose(got["sample"][4].cpu(), ref["sample"][4], rtol=1e-4, atol=1e-5)


def test_ddim_step_accepts_eps_only_prediction():
  """A plain apply_fn that returns the eps tensor [n,H,W,C] (the reference's p_apply contract)."""
  from small_vision_b200 import diffusion as Dm
  gd, gdd = _tables()
  g = torch.Generator().manual_seed(5)
  x, e, nz = (torch.randn(3, 4, 4, 3, generator=g) for _ in range(3))
  t = torch.tensor([[800], [400], [8]], dtype=torch.int32)
  ref = O.ddim_sample(gd, lambda x_t, t: e, x, t.long(), (t - 8).long(), nz, eta=1.0)
  got = Dm.ddim_sample(gdd, lambda **kw: e.to(DEV), x.to(DEV), t.to(DEV), (t - 8).to(DEV), 0, eta=1.0, noise=nz.to(DEV))
  assert U.rel_l2(got["sample"].cpu(), ref["sample"]) < 2e-5


def test_ddim_consistency_at_bench_batch():
  """Size-independent property at 512 x 64 x 64 x 3: with the true eps as the model output an eta = 0 update lands on
  q_sample(x0, t_next, eps) and pred_xstart recovers x0."""
  from small_vision_b200 import diffusion as Dm
  gd, gdd = _tables()
  g = torch.Generator(device=DEV).manual_seed(11)
  n = 512
  x0 = torch.rand(n, 64, 64, 3, device=DEV, generator=g) * 2 - 1
  eps = torch.randn(n, 64, 64, 3, device=DEV, generator=g)
  t = torch.randint(8, 900, (n, 1), device=DEV, generator=g, dtype=torch.int32)
  t_next = t - 8
  x_t = Dm.q_sample(gd=gdd, x_start=x0, t=t, noise=eps)
  got = Dm.ddim_sample(gdd, lambda **kw: eps, x_t, t, t_next, 0, eta=0.0, noise=torch.zeros_like(x0))
  want = Dm.q_sample(gd=gdd, x_start=x0, t=t_next, noise=eps)
  assert float((got["pred_xstart"] - x0).abs().max()) < 2e-4
  assert float((got["sample"] - want).abs().max()) < 2e-4


@pytest.mark.parametrize("cfg_scale,eps_pred", [(None, True), (1.5, True), (2.0, False)])
def test_ddim_sample_loop_matches_oracle(cfg_scale, eps_pred):
  """The whole sampler (train_ae.py:472-509 wiring: model at t + 1, eps head, CFG doubling) on a shallow S/4 model."""
  from small_vision_b200 import diffusion as Dm
  nc = 10
  model, ocfg = U.make_models("S/4", adaln=True, num_classes=nc, depth=2, dec_depth=1)
  params = U.perturb_init(model, 3, DEV)
  oparams = U.cpu_tree(params)
  gd, gdd = _tables()
  n, steps = 2, 4
  g = torch.Generator().manual_seed(21)
  noises = [torch.randn(n, 64, 64, 3, generator=g) for _ in range(steps + 2)]
  ys = torch.tensor([3, 7])
  use_y = cfg_scale is not None
  ref = O.ddim_sample_loop(gd, O.make_apply_fn(oparams, ocfg, gd, eps_pred=eps_pred), noises, ys=ys if use_y else None,
                           sampling_steps=steps, cfg_scale=cfg_scale, eta=0.3)
  apply_fn = Dm.create_apply_fn(model, params, eps_pred=eps_pred)
  got, _ = Dm.ddim_sample_loop(gdd, apply_fn, 0, torch.zeros(n, 64, 64, 3), ys=ys.to(DEV) if use_y else None,
                               sampling_steps=steps, cfg_scale=cfg_scale, eta=0.3, noises=noises)
  torch.cuda.synchronize()
  a = got["sample"].cpu()
  assert a.shape == ref.shape and torch.isfinite(a).all()
  r = U.rel_l2(a, ref)
  assert r <= U.TOL_PRED_REL_L2, f"sample rel-L2 {r}"
