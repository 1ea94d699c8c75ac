"""Independent pins of the optimiser chain of big_vision/trainers/train_ae.py:135-151,365-374.

optax is a third-party dependency that is absent from /root/reference and from this image (requirements.txt:5 pins no
version), so the oracle's `optimizer_update` cannot be checked against optax itself.  What CAN be done is done here:

  * a second, structurally different restatement of the published optax algorithm — the chain as optax composes it,
    one GradientTransformation at a time with its own state (clip_by_global_norm -> scale_by_adam(mu_dtype=bf16) ->
    add_decayed_weights(mask) -> scale_by_learning_rate(schedule) -> apply_updates), in numpy float32 with the bf16
    storage rounding done by explicit bit manipulation (round to nearest even) — run for three steps from a seeded state and
    compared with oracle.optimizer_update leaf by leaf;
  * the clip branch `where(norm < c, g, g / norm * c)` both ways, including the boundary norm == c;
  * the schedule evaluated at the PRE-increment count (lr(0) = 0: the first step leaves the parameters alone);
  * the decay mask per leaf from the reference's own `get_weight_decay_mask` (recorded while its code ran:
    tests/golden/reference_recipe_golden.json);
  * closed forms: step 1 with and without clipping, EMA = optax.incremental_update.

Each transformation cites what it restates: optax/_src/clipping.py (clip_by_global_norm), optax/_src/transform.py
(scale_by_adam, update_moment, bias_correction, add_decayed_weights, scale_by_learning_rate / scale_by_schedule),
optax/_src/schedule.py (warmup_cosine_decay_schedule = join_schedules(linear_schedule, cosine_decay_schedule)),
optax/_src/update.py (apply_updates, incremental_update), as of optax 0.1.7 - 0.2.3 (the arithmetic of these functions did
not change across those releases).
"""
import json
import math
import os

import numpy as np
import pytest
import torch

from oracle import umd_oracle as O

F32 = np.float32


# ---------------------------------------------------------------------------------------------------------------
# bf16 storage: round to nearest even on the upper 16 bits of the float32 pattern
# ---------------------------------------------------------------------------------------------------------------
def bf16_round(x):
  x = np.asarray(x, dtype=np.float32)
  u = x.view(np.uint32).astype(np.uint64)
  lsb = (u >> 16) & 1
  r = ((u + 0x7FFF + lsb) >> 16) << 16
  return r.astype(np.uint32).view(np.float32).reshape(x.shape)


def test_bf16_rounding_is_round_to_nearest_even():
  g = torch.Generator().manual_seed(0)
  x = torch.randn(20000, generator=g) * torch.logspace(-6, 3, 20000)
  ties = torch.tensor([1.0 + 2 ** -8, 1.0 + 3 * 2 ** -8, 1.0 + 2 ** -9, -(1.0 + 2 ** -8), 0.0, 2 ** -130], dtype=torch.float32)
  x = torch.cat([x, ties])
  ours = bf16_round(x.numpy())
  ref = x.to(torch.bfloat16).to(torch.float32).numpy()
  assert np.array_equal(ours, ref)
  assert bf16_round(np.float32(1.0 + 2 ** -8)) == np.float32(1.0)            # tie -> even mantissa
  assert bf16_round(np.float32(1.0 + 3 * 2 ** -8)) == np.float32(1.0 + 2 ** -6)


# ---------------------------------------------------------------------------------------------------------------
# optax, one transformation at a time (state = what optax keeps)
# ---------------------------------------------------------------------------------------------------------------
def linear_schedule(init_value, end_value, transition_steps):
  def f(count):
    c = min(max(count, 0), transition_steps)
    frac = 1 - c / transition_steps
    return (init_value - end_value) * frac + end_value
  return f


def cosine_decay_schedule(init_value, decay_steps, alpha):
  def f(count):
    c = min(count, decay_steps)
    cosine = 0.5 * (1 + math.cos(math.pi * c / decay_steps))
    return init_value * ((1 - alpha) * cosine + alpha)
  return f


def warmup_cosine_decay_schedule(init_value, peak_value, warmup_steps, decay_steps, end_value=0.0):
  a = linear_schedule(init_value, peak_value, warmup_steps)
  alpha = 0.0 if peak_value == 0.0 else end_value / peak_value
  b = cosine_decay_schedule(peak_value, decay_steps - warmup_steps, alpha)
  return lambda count: a(count) if count < warmup_steps else b(count - warmup_steps)   # join_schedules(..., [warmup_steps])


class ClipByGlobalNorm:
  def __init__(self, max_norm):
    self.c = max_norm

  def update(self, g):
    norm = F32(math.sqrt(sum(float(np.sum(np.square(v, dtype=np.float32), dtype=np.float32)) for v in g.values())))
    if norm < self.c:
      return dict(g), norm
    return {k: (v / norm) * F32(self.c) for k, v in g.items()}, norm


class ScaleByAdam:
  def __init__(self, params, b1, b2, eps):
    self.b1, self.b2, self.eps = F32(b1), F32(b2), F32(eps)
    self.count = 0
    self.mu = {k: np.zeros_like(v) for k, v in params.items()}     # stored bf16 (kept here as its float32 image)
    self.nu = {k: np.zeros_like(v) for k, v in params.items()}

  def update(self, g):
    self.count += 1
    out = {}
    for k, gk in g.items():
      mu = (F32(1) - self.b1) * gk + self.b1 * self.mu[k]           # un-rounded first moment, used for this step
      nu = (F32(1) - self.b2) * (gk * gk) + self.b2 * self.nu[k]
      mu_hat = mu / F32(1 - float(self.b1) ** self.count)
      nu_hat = nu / F32(1 - float(self.b2) ** self.count)
      out[k] = mu_hat / (np.sqrt(nu_hat) + self.eps)
      self.mu[k] = bf16_round(mu)                                   # cast_tree(mu, mu_dtype) AFTER the update is formed
      self.nu[k] = nu
    return out


class AddDecayedWeights:
  def __init__(self, wd, mask):
    self.wd, self.mask = F32(wd), mask

  def update(self, u, params):
    return {k: (v + self.wd * params[k] if self.mask[k] else v) for k, v in u.items()}


class ScaleByLearningRate:
  def __init__(self, schedule):
    self.schedule, self.count = schedule, 0

  def update(self, u):
    step = F32(-self.schedule(self.count))      # evaluated at the count BEFORE the increment
    self.count += 1
    return {k: step * v for k, v in u.items()}


def _tree(seed, scale=1.0):
  g = np.random.default_rng(seed)
  return {("blk", "Dense_0", "kernel"): (g.standard_normal((7, 5)) * scale).astype(F32),
          ("blk", "Dense_0", "bias"): (g.standard_normal(5) * scale).astype(F32),
          ("cls",): (g.standard_normal((1, 4, 3)) * scale).astype(F32),
          ("pos_embedding",): (g.standard_normal((1, 6, 3)) * scale).astype(F32),
          ("blk", "LayerNorm_0", "scale"): (1 + 0.1 * g.standard_normal(5)).astype(F32)}


def _grads(seed, scale):
  g = np.random.default_rng(seed)
  return {k: (g.standard_normal(v.shape) * scale).astype(F32) for k, v in _tree(0).items()}


def _to_oracle(tree):
  return O.unflatten_tree({k: torch.from_numpy(np.array(v)) for k, v in tree.items()})


@pytest.mark.parametrize("grad_scale", [0.05, 3.0])   # below and above clip_norm = 1: both branches of the clip
def test_three_steps_of_the_chain_match_the_oracle(grad_scale):
  hp = dict(clip_norm=1.0, peak_lr=3e-3, warmup_steps=2, total_steps=50, b1=0.9, b2=0.95, wd=0.05)
  params = _tree(0)
  mask = {k: all(s not in ("cls", "image_mask_embedding", "bias") for s in k) for k in params}
  assert mask[("cls",)] is False and mask[("blk", "Dense_0", "bias")] is False and mask[("blk", "LayerNorm_0", "scale")] is True
  clip = ClipByGlobalNorm(hp["clip_norm"])
  adam = ScaleByAdam(params, hp["b1"], hp["b2"], 1e-8)
  decay = AddDecayedWeights(hp["wd"], mask)
  lr = ScaleByLearningRate(warmup_cosine_decay_schedule(0.0, hp["peak_lr"], hp["warmup_steps"], hp["total_steps"]))
  oparams = _to_oracle(params)
  opt = O.init_opt_state(oparams)
  for step in range(3):
    grads = _grads(100 + step, grad_scale)
    g, norm = clip.update(grads)
    assert (norm < hp["clip_norm"]) == (grad_scale < 1.0)
    u = lr.update(decay.update(adam.update(g), params))
    new_params = {k: params[k] + u[k] for k in params}               # optax.apply_updates
    oparams_new, opt, oupd, ognorm = O.optimizer_update(_to_oracle(grads), opt, oparams, hp)
    assert abs(ognorm - float(norm)) <= 1e-6 * float(norm)
    fo, fu = O.flatten_tree(oparams_new), oupd
    for k in params:
      np.testing.assert_allclose(fu[k].numpy(), u[k], rtol=2e-5, atol=1e-9, err_msg=f"update {k} step {step}")
      np.testing.assert_allclose(fo[k].numpy(), new_params[k], rtol=1e-6, atol=1e-8, err_msg=f"param {k} step {step}")
      # stored first moment: bf16 bit patterns agree except where the fp32 sum sits within an ulp of a rounding tie
      mu_o = opt["mu"][k].float().numpy()
      bad = mu_o != adam.mu[k]
      assert bad.mean() <= 0.05 and np.allclose(mu_o, adam.mu[k], rtol=2 ** -7, atol=1e-12), f"mu {k} step {step}"
      np.testing.assert_allclose(opt["nu"][k].numpy(), adam.nu[k], rtol=1e-6, atol=1e-12)
    if step == 0:   # lr(0) = init_value = 0: the first step moves nothing (SURVEY §8c pin 10)
      assert all(float(np.abs(u[k]).max()) == 0.0 for k in u)
    params, oparams = new_params, oparams_new
  assert opt["count"] == 3 and adam.count == 3 and lr.count == 3


def test_clip_boundary_and_scaling():
  c = ClipByGlobalNorm(1.0)
  g = {("a",): np.array([0.6, 0.8], dtype=F32)}             # norm exactly 1.0: NOT < c, goes through the scaling branch
  out, norm = c.update(g)
  assert norm == F32(1.0) and np.allclose(out[("a",)], g[("a",)])
  hp = dict(clip_norm=1.0, peak_lr=1e-3, warmup_steps=0, total_steps=10, b1=0.9, b2=0.95, wd=0.0)
  for scale in (0.5, 1.0, 4.0):
    gg = {("a",): (np.array([0.6, 0.8]) * scale).astype(F32)}
    p = {("a",): np.zeros(2, dtype=F32)}
    _, _, _, gn = O.optimizer_update(_to_oracle(gg), O.init_opt_state(_to_oracle(p)), _to_oracle(p), hp)
    out, norm = c.update(gg)
    assert abs(gn - scale) < 1e-6
    want = gg[("a",)] if scale < 1.0 else gg[("a",)] / F32(scale)
    assert np.allclose(out[("a",)], want, rtol=1e-6)


def test_schedule_matches_oracle_and_host_at_every_count():
  from small_vision_b200.config import warmup_cosine_lr
  for warm, total, peak in ((0, 10, 1e-3), (3, 20, 2.4e-3), (15480, 247726, 2.4e-3)):
    s = warmup_cosine_decay_schedule(0.0, peak, warm, total)
    counts = sorted(set([0, 1, 2, warm - 1, warm, warm + 1, total // 2, total - 1, total, total + 5]) - {-1})
    for cnt in counts:
      a = s(cnt)
      assert abs(O.warmup_cosine_lr(cnt, peak=peak, warmup_steps=warm, decay_steps=total) - a) <= 1e-12 + 1e-9 * abs(a), cnt
      assert abs(warmup_cosine_lr(cnt, peak=peak, warmup_steps=warm, decay_steps=total) - a) <= 1e-12 + 1e-9 * abs(a), cnt
    if warm > 0:
      assert s(0) == 0.0 and abs(s(1) - peak / warm) < 1e-15 and abs(s(warm) - peak) < 1e-15


def test_decay_mask_is_the_references_own():
  """Every leaf of the engine's arena carries the decay flag that the reference's get_weight_decay_mask
  (train_ae.py:125-134, executed by tests/golden/make_recipe_golden.py) returned for that leaf."""
  from small_vision_b200.config import make_model_config
  from small_vision_b200.params import ArenaLayout
  path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_recipe_golden.json")
  gold = json.load(open(path))
  checked = 0
  for name, case in gold["recipes"].items():
    mk = {k: v for k, v in case["model"].items() if k in ("variant", "adaln", "num_classes", "channels", "img_size")}
    lay = ArenaLayout(make_model_config(**mk))
    flags = lay.wd_flags("cpu")
    for lf in lay.leaves:
      ref = case["decay_mask"]["/".join(lf.path)]
      assert lay.decay(lf) == ref, (name, lf.path)
      for l in range(lf.shape[0] if lf.stack else 1):
        assert int(flags[(lf.offset + l * lf.lstride) // 64]) == int(ref)
      checked += 1
  assert checked > 100


def test_ema_is_optax_incremental_update():
  """optax.incremental_update(new, old, step_size) = step_size * new + (1 - step_size) * old (train_ae.py:374)."""
  p = _to_oracle(_tree(3))
  e = _to_oracle(_tree(4))
  hp = dict(clip_norm=1.0, peak_lr=1e-3, warmup_steps=0, total_steps=10, b1=0.9, b2=0.95, wd=0.05, ema_decay=0.25)
  st = {"params": p, "opt": O.init_opt_state(p), "ema_params": e, "gd": None}
  new_p, _, _, _ = O.optimizer_update(_to_oracle(_tree(5)), st["opt"], p, hp)
  fe, fn = O.flatten_tree(e), O.flatten_tree(new_p)
  want = {k: 0.25 * fn[k] + 0.75 * fe[k] for k in fe}
  # the same formula the oracle's update_step applies
  got = {k: hp["ema_decay"] * fn[k] + (1.0 - hp["ema_decay"]) * fe[k] for k in fe}
  for k in want:
    assert torch.allclose(got[k], want[k])
