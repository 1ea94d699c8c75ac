"""Shared helpers of the parity tests: build the same model in the CUDA engine and in the CPU oracle,
draw one set of inputs / random draws for both, and compare.  Tolerances follow SURVEY.md App. G: the
engine feeds bf16 operands to fp32-accumulating tensor-core GEMMs (fp32 residual stream, fp32 LayerNorm
statistics, fp32 loss/optimiser) while the oracle is fp32 throughout — the reference default (ae.py:51)."""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

from oracle import umd_oracle as O  # noqa: E402

# App. G tolerances
TOL_PRED_REL_L2 = 3e-2
TOL_LOSS_REL = 1e-2
TOL_GRAD_COS = 0.995
TOL_GRAD_REL_L2 = 5e-2
TOL_GNORM_REL = 2e-2


def rel_l2(a, b):
  a, b = a.double().reshape(-1), b.double().reshape(-1)
  return float((a - b).norm() / (b.norm() + 1e-30))


def cosine(a, b):
  a, b = a.double().reshape(-1), b.double().reshape(-1)
  return float((a @ b) / (a.norm() * b.norm() + 1e-30))


def make_models(variant="S/4", *, adaln=True, num_classes=None, img_size=64, channels=3, depth=None, dec_depth=None,
                **kw):
  """Returns (engine model, oracle cfg dict)."""
  from small_vision_b200.model import Model
  extra = dict(kw)
  if depth is not None:
    extra["depth"] = depth
  if dec_depth is not None:
    extra["dec_depth"] = dec_depth
  model = Model(variant=variant, adaln=adaln, num_classes=num_classes, img_size=img_size, channels=channels, **extra)
  ocfg = O.model_config(variant=variant, adaln=adaln, num_classes=num_classes, img_size=img_size, channels=channels,
                        **extra)
  return model, ocfg


def cpu_tree(tree):
  """Deep copy of a (GPU) parameter tree onto the CPU as plain nested dicts of fp32 tensors."""
  return {k: (cpu_tree(v) if isinstance(v, dict) else v.detach().float().cpu().clone()) for k, v in tree.items()}


def perturb_init(model, seed, device):
  """Flax-like init with non-zero adaLN kernels, plus small random biases / LN params / cls so that no
  gradient path is hidden behind an exact zero or one (SURVEY.md §8d 'Synthetic inputs')."""
  from small_vision_b200.params import tree_from_arena, init_arena
  arena = init_arena(model.layout, seed, "cpu", nonzero_adaln=True)
  g = torch.Generator().manual_seed(seed + 1234)
  for lf in model.layout.init_order:
    if lf.init in ("zeros", "ones"):
      lf.view(arena).add_(0.05 * torch.randn(lf.size, generator=g).view(lf.shape))
  arena = arena.to(device)
  return tree_from_arena(model.layout, arena)


def make_batch(model, B, *, n_noise, seed, use_labels=False, device="cuda"):
  cfg = model.cfg
  g = torch.Generator().manual_seed(seed)
  H, C, L = cfg.img_size, cfg.channels, cfg.num_patches
  n_clean = B - n_noise
  image = torch.rand(B, H, H, C, generator=g) * 2 - 1
  batch = {"image": image, "label": torch.randint(0, max(cfg.num_classes or 1, 1), (B,), generator=g)}
  mask_noise_noise = torch.rand(n_noise, L, generator=g)
  mask_noise_clean = torch.rand(n_clean, L, generator=g)
  # deliberate ties: fp32 uniforms collide in ~0.4 % of rows; make sure the stable tie-break is hit
  if n_noise > 0:
    mask_noise_noise[0, 5] = mask_noise_noise[0, 200]
    mask_noise_noise[0, 17] = mask_noise_noise[0, 3]
  if n_clean > 0:
    mask_noise_clean[0, 100] = mask_noise_clean[0, 7]
  rand = {
      "t": torch.randint(0, 1000, (n_noise, 1), generator=g, dtype=torch.int32),
      "noise": torch.randn(n_noise, H, H, C, generator=g),
      "mask_noise_noise": mask_noise_noise,
      "mask_noise_clean": mask_noise_clean,
  }
  if use_labels:
    rand["label_drop_noise"] = torch.rand(n_noise, generator=g) < 0.1
  return batch, rand


def to_dev(d, device):
  return {k: (v.to(device) if isinstance(v, torch.Tensor) else v) for k, v in d.items()}


def tree_compare(ours, ref, *, what, cos_min=TOL_GRAD_COS, rel_max=TOL_GRAD_REL_L2, abs_floor=0.0, skip=()):
  """Per-leaf comparison.  Leaves whose reference norm is below abs_floor * (global norm) are checked by
  absolute error against that floor instead (cosine of a near-zero vector is noise)."""
  fo, fr = O.flatten_tree(ours), O.flatten_tree(ref)
  assert set(fo) == set(fr), (sorted(set(fo) ^ set(fr))[:5])
  gnorm = math.sqrt(sum(float(v.double().pow(2).sum()) for v in fr.values()))
  worst = {"cos": 1.0, "rel": 0.0}
  fails = []
  for k in sorted(fr):
    if any(s in k for s in skip):
      continue
    a, b = fo[k].detach().float().cpu(), fr[k].detach().float().cpu()
    assert a.shape == b.shape, (k, a.shape, b.shape)
    bn = float(b.double().norm())
    if bn <= abs_floor * gnorm:
      err = float((a.double() - b.double()).norm())
      if err > max(abs_floor * gnorm, 1e-12) * rel_max * 4 + 1e-12:
        fails.append(("/".join(k), "tiny-leaf abs err", err, bn))
      continue
    c, r = cosine(a, b), rel_l2(a, b)
    worst["cos"] = min(worst["cos"], c)
    worst["rel"] = max(worst["rel"], r)
    if c < cos_min or r > rel_max:
      fails.append(("/".join(k), f"cos={c:.5f}", f"rel={r:.4f}", bn))
  assert not fails, f"{what}: {len(fails)} leaves out of tolerance, e.g. {fails[:6]}"
  return worst


def oracle_hp(tcfg):
  t = tcfg.resolved()
  return dict(clip_norm=t.clip_norm, peak_lr=t.scaled_peak_lr, warmup_steps=t.warmup_steps, total_steps=t.total_steps,
              b1=t.betas[0], b2=t.betas[1], wd=t.wd, ema_decay=t.ema_decay)


def _compare_updated(ours, ref, init, what):
  """Updated parameters against the oracle's.  General bound: relative L2 <= 2e-3 of the leaf.  Adam normalises the
  update to ~lr per element, so on leaves whose VALUES are themselves ~lr or smaller — the N(0, 1e-6) MLP biases and the
  zero-initialised biases (every path that ends in "bias"), plus zero-initialised `cls` — the leaf-relative error is
  dominated by sign flips of near-zero gradient elements; only those leaves are bounded against the size of the update
  instead (error <= 25 % of the update norm).  The optimiser arithmetic itself is checked to fp32 round-off in
  tests/test_kernels_gpu.py::test_adamw_step_matches_oracle."""
  fo, fr, f0 = O.flatten_tree(ours), O.flatten_tree(ref), O.flatten_tree(init)
  worst = 0.0
  for k in fr:
    a, b, c = fo[k].detach().float().cpu(), fr[k], f0[k]
    r = rel_l2(a, b)
    upd = float((b.double() - c.double()).norm())
    err = float((a.double() - b.double()).norm())
    small_valued = k[-1] == "bias" or k[0] == "cls"
    if small_valued:
      assert r <= 2e-3 or err <= 0.25 * upd, f"{what} {'/'.join(k)}: rel {r:.4g}, err {err:.4g}, update {upd:.4g}"
      worst = max(worst, min(r, err / (upd + 1e-30)))
    else:
      assert r <= 2e-3, f"{what} {'/'.join(k)}: rel {r:.4g}, err {err:.4g}, update {upd:.4g}"
      worst = max(worst, r)
  return worst


def run_step_parity(*, variant="S/4", batch=8, adaln=True, num_classes=None, use_labels=False, mask_ratio=0.375,
                    mask_ratio_no_noise=0.75, no_noise_prob=0.5, steps=2, seed=0, device="cuda", depth=None,
                    dec_depth=None, ema_decay=None, img_size=64, channels=3, beta_schedule="cosine",
                    check_grads=True, residual_dtype="float32", grad_stream_dtype="float32"):
  """Runs `steps` update_fn steps in the engine and in the oracle from identical state and draws and asserts
  App. G tolerances on loss, gradients (first step), grad-norm, l2 measurements and updated parameters."""
  from small_vision_b200.config import TrainConfig
  from small_vision_b200.train import create_train_state, make_update_fn
  from small_vision_b200.diffusion import create_gaussian_diffusion
  model, ocfg = make_models(variant, adaln=adaln, num_classes=num_classes, depth=depth, dec_depth=dec_depth,
                            img_size=img_size, channels=channels, residual_dtype=residual_dtype,
                            grad_stream_dtype=grad_stream_dtype)
  tcfg = TrainConfig(batch_size=batch, no_noise_prob=no_noise_prob, mask_ratio=mask_ratio,
                     mask_ratio_no_noise=mask_ratio_no_noise, use_labels=use_labels, total_steps=1000, warmup_steps=0,
                     peak_lr=2e-3, ema_decay=ema_decay, beta_schedule=beta_schedule,
                     diffusion_space=(img_size, img_size, channels))
  params = perturb_init(model, seed, device)
  state = create_train_state(model, tcfg, seed=seed, device=device, params=params)
  ostate = {"params": cpu_tree(state["params"]), "gd": create_gaussian_diffusion(beta_schedule, 1000)}
  init_params = cpu_tree(state["params"])
  ostate["opt"] = O.init_opt_state(ostate["params"])
  if ema_decay:
    ostate["ema_params"] = cpu_tree(state["params"])
  update_fn = make_update_fn(model, tcfg)
  hp = oracle_hp(tcfg)
  otc = dict(mask_ratio=mask_ratio, mask_ratio_no_noise=mask_ratio_no_noise, no_noise_prob=no_noise_prob,
             use_labels=use_labels)
  n_clean = int(batch * no_noise_prob)
  n_noise = batch - n_clean
  from small_vision_b200.params import tree_from_arena
  rep = {}
  for s in range(steps):
    b, rand = make_batch(model, batch, n_noise=n_noise, seed=seed + 100 + s, use_labels=use_labels)
    gb = to_dev(b, device)
    gb["_rand"] = to_dev(rand, device)
    state, meas = update_fn(state, gb)
    torch.cuda.synchronize()
    ostate, omeas, oextra = O.update_step(ostate, b, ocfg, otc, hp, rand)
    loss, oloss = float(meas["training_loss"]), omeas["training_loss"]
    assert math.isfinite(loss), loss
    rel = abs(loss - oloss) / abs(oloss)
    assert rel <= TOL_LOSS_REL, f"step {s}: loss {loss} vs oracle {oloss} (rel {rel:.3g})"
    rep[f"loss_rel_{s}"] = rel
    gn, ogn = float(meas["grad_norm"]), omeas["grad_norm"]
    assert abs(gn - ogn) / ogn <= TOL_GNORM_REL, f"step {s}: grad norm {gn} vs {ogn}"
    rep[f"gnorm_rel_{s}"] = abs(gn - ogn) / ogn
    if check_grads and s == 0:
      gtree = tree_from_arena(model.layout, update_fn.grads()[:model.layout.total])
      w = tree_compare(gtree, oextra["grads"], what=f"grads step {s}", abs_floor=1e-4)
      rep["grad_cos_min"], rep["grad_rel_max"] = w["cos"], w["rel"]
    for key in ("l2_params", "l2_updates"):
      a, r = float(meas[key]), omeas[key]
      assert abs(a - r) <= 2e-2 * abs(r) + 1e-6, f"step {s}: {key} {a} vs {r}"
  # parameters after the last step.  Adam normalises the update to ~lr per element, so on leaves whose values are
  # themselves ~lr (the N(0,1e-6) MLP biases) the parameter error is dominated by sign flips of near-zero gradient
  # elements; those leaves are bounded against the size of the update instead.  The optimiser arithmetic itself is
  # checked to fp32 round-off in tests/test_kernels_gpu.py::test_adamw_step_matches_oracle.
  rep["param_err_max"] = _compare_updated(state["params"], ostate["params"], init_params, "param")
  if ema_decay:
    rep["ema_err_max"] = _compare_updated(state["ema_params"], ostate["ema_params"], init_params, "ema")
  return rep
