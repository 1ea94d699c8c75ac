"""The oracle against tests/golden/reference_golden.pt — numbers produced by EXECUTING THE REFERENCE'S OWN SOURCE
(`ae.py`, `vit.py`, `embeddings.py`, `gaussian_diffusion.py` and train_ae.py's `loss_fn`) over a numpy-fp64 stand-in for
jax/flax (tests/golden/make_reference_golden.py, tests/golden/refshim/README.md).  This is what pins the oracle to the
reference: the oracle runs in float64 on the same parameters, inputs and supplied draws and must agree to round-off;
its autograd gradient must reproduce the slopes obtained by differencing the reference's loss.  Nothing here reads
/root/reference."""
import os

import numpy as np
import pytest
import torch

from oracle import umd_oracle as O
from tests import util as U
from tests.golden import make_reference_golden as RG

GOLD = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_golden.pt"))
F64 = torch.float64


def rebuild(name):
  mkw, tkw, B, n_noise = RG.CASES[name]
  model, ocfg = U.make_models(**mkw)
  params = U.cpu_tree(U.perturb_init(model, RG.PARAM_SEED, "cpu"))
  batch, rand = U.make_batch(model, B, n_noise=n_noise, seed=RG.BATCH_SEED, use_labels=tkw["use_labels"], device="cpu")
  want = GOLD["cases"][name]
  assert RG.digest(batch["image"]) + RG.digest(rand["noise"]) + RG.digest(rand["mask_noise_clean"]) == want["input_digest"]
  assert RG.digest(torch.cat([v.reshape(-1) for _, v in sorted(RG.flatten(params).items())])) == want["param_digest"]
  return model, ocfg, tkw, params, batch, rand, n_noise, want


def oracle_loss(params64, ocfg, tkw, batch, rand, n_noise, schedule="cosine"):
  gd = O.gaussian_diffusion_tables(schedule, 1000)
  img = batch["image"].to(F64)
  x0n, x0c = img[:n_noise], img[n_noise:]
  x_t = O.q_sample(gd, x0n, rand["t"], rand["noise"].to(F64))
  labels = batch["label"][:n_noise] if tkw["use_labels"] else None
  loss, aux = O.loss_fn(params64, ocfg, tkw, x0n, x_t, x0c, rand["t"], rand["noise"].to(F64), labels, rand, dtype=F64)
  return loss, aux, x_t


def to64(tree, grad=False):
  return {k: (to64(v, grad) if isinstance(v, dict) else v.double().clone().requires_grad_(grad)) for k, v in tree.items()}


def check_branch(aux, br, want, patch=4):
  pred = aux[f"pred_{br}"].detach()
  out = aux[f"out_{br}"]
  assert U.rel_l2(pred[0], want["pred0"]) <= 1e-6            # the fixture stores sample 0 as float32
  assert torch.allclose(pred.mean(dim=(1, 2)), want["pred_sample_means"], rtol=0, atol=1e-11)
  assert abs(float(pred.abs().mean()) - want["pred_abs_mean"]) <= 1e-11
  assert torch.allclose(out["pre_logits"].detach(), want["pre_logits"], rtol=0, atol=1e-10)
  if "patch_mask" in want:
    seq = out["mask"][:, ::patch, ::patch, 0].reshape(pred.shape[0], -1)
    assert torch.equal(seq.to(torch.uint8), want["patch_mask"])           # bit-exact, ties included
    assert torch.equal((out["ids_restore"] >= (seq == 0).sum(1, keepdim=True)).to(torch.uint8), want["patch_mask"])
  else:
    assert out["mask"] is None


@pytest.mark.parametrize("name", sorted(RG.CASES))
def test_forward_and_loss_match_reference_source(name):
  model, ocfg, tkw, params, batch, rand, n_noise, want = rebuild(name)
  loss, aux, x_t = oracle_loss(to64(params), ocfg, tkw, batch, rand, n_noise, RG.SCHEDULE.get(name, "cosine"))
  assert U.rel_l2(x_t, want["x_t"]) <= 1e-6
  assert abs(float(loss) - want["loss"]) <= 1e-11 * abs(want["loss"]) + 1e-12, (float(loss), want["loss"])
  for br in ("noise", "clean"):
    if br in want:
      check_branch(aux, br, want[br], ocfg["patch_size"][0])
    else:
      assert f"pred_{br}" not in aux


@pytest.mark.parametrize("name", sorted(RG.CASES))
def test_autograd_gradient_reproduces_reference_loss_slopes(name):
  """<grad L_oracle, d> == (L_ref(p + h d) - L_ref(p - h d)) / 2h (Richardson-extrapolated, fp64) for seeded unit
  directions d over the whole tree and over each top-level parameter group."""
  model, ocfg, tkw, params, batch, rand, n_noise, want = rebuild(name)
  p64 = to64(params, grad=True)
  loss, _, _ = oracle_loss(p64, ocfg, tkw, batch, rand, n_noise, RG.SCHEDULE.get(name, "cosine"))
  loss.backward()
  grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in RG.flatten(p64).items()}
  gnorm = float(torch.sqrt(sum((g ** 2).sum() for g in grads.values())))
  dirs = RG.directions(params)
  assert [g for g, _ in dirs] == [g for g, _, _ in want["slopes"]]
  for (grp, d), (_, slope, err) in zip(dirs, want["slopes"]):
    mine = float(sum((grads[k] * d[k]).sum() for k in grads))
    tol = 1e-6 * abs(slope) + 10 * err + 1e-9 * gnorm
    assert abs(mine - slope) <= tol, (name, grp, mine, slope, err)


@pytest.mark.parametrize("name", [n for n in sorted(RG.CASES) if "cfg" in GOLD["cases"][n]])
def test_classifier_free_guidance_forward_matches_reference_source(name):
  model, ocfg, tkw, params, batch, rand, n_noise, want = rebuild(name)
  c = want["cfg"]
  pred, out = O.model_apply(to64(params), ocfg, batch["image"][:2].to(F64), t=c["t"], y=batch["label"][:2],
                            cfg_scale=c["cfg_scale"], dtype=F64)
  assert U.rel_l2(pred, c["pred"]) <= 1e-6
  assert torch.allclose(out["pre_logits"], c["pre_logits"], rtol=0, atol=1e-10)


def test_diffusion_tables_match_reference_source():
  for sched in ("cosine", "linear"):
    mine = O.gaussian_diffusion_tables(sched, 1000)
    want = GOLD["diffusion"][f"tables_{sched}"]
    assert set(want) <= set(mine), sorted(set(want) - set(mine))
    for k, v in want.items():
      np.testing.assert_allclose(np.asarray(mine[k], dtype=np.float64), v.numpy(), rtol=1e-13, atol=0, err_msg=k)


def test_q_sample_and_ddim_match_reference_source():
  D = GOLD["diffusion"]
  gd = O.gaussian_diffusion_tables("cosine", 1000)
  i = D["inputs"]
  x, eps, nz, t, tn = i["x"], i["eps"], i["noise"], i["t"], i["t_next"]
  assert torch.allclose(O.q_sample(gd, x, t, nz), D["q_sample"], rtol=1e-12, atol=1e-13)
  assert torch.allclose(O.predict_xstart_from_eps(gd, x, t, eps), D["xstart_from_eps"], rtol=1e-12, atol=1e-12)
  for s in D["ddim_steps"]:
    r = O.ddim_sample(gd, lambda x_t, t, **kw: eps, x, t.long(), tn.long() if s["use_next"] else None, nz,
                      clip_denoised=s["clip"], eta=s["eta"])
    assert torch.allclose(r["pred_xstart"], s["pred_xstart"], rtol=1e-11, atol=1e-11), s["eta"]
    assert torch.allclose(r["sample"], s["sample"], rtol=1e-10, atol=1e-10), s["eta"]
  L = D["ddim_loop"]

  def apply_fn(*, x_t, t, y=None, cfg_scale=None):
    return 0.3 * x_t + 0.01 * (t.double() / 1000.0)[:, :, None, None] + 0.05 * y.double()[:, None, None, None]
  got = O.ddim_sample_loop(gd, apply_fn, list(L["draws"]), ys=L["ys"], sampling_steps=L["sampling_steps"], eta=L["eta"])
  assert torch.allclose(got, L["sample"], rtol=1e-10, atol=1e-10)
  for key, v in D.items():
    if key.startswith("timesteps_"):
      _, n, s = key.split("_")
      assert O.ddim_timesteps(int(n), int(s)) == v.tolist(), key


def test_sampler_matches_reference_source():
  """create_apply_fn (train_ae.py:472-483) under ddim_sample_loop with classifier-free guidance, both heads."""
  S = RG.SAMPLER
  gold = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_sampler_golden.pt"))["sampler"]
  mkw = RG.CASES[S["case"]][0]
  model, ocfg = U.make_models(**mkw)
  params = to64(U.cpu_tree(U.perturb_init(model, RG.PARAM_SEED, "cpu")))
  g = torch.Generator().manual_seed(S["noise_seed"])
  noises = [torch.randn(S["n"], 64, 64, 3, generator=g) for _ in range(S["steps"] + 2)]
  assert RG.digest(torch.stack(noises)) == gold["noise_digest"]
  gd = O.gaussian_diffusion_tables("cosine", 1000)
  ys = torch.tensor(S["ys"])
  for (cfg_scale, eps_pred), want in zip(S["variants"], gold["samples"]):
    apply_fn = O.make_apply_fn(params, ocfg, gd, eps_pred=eps_pred, dtype=F64)
    got = O.ddim_sample_loop(gd, apply_fn, [z.double() for z in noises], ys=ys if cfg_scale is not None else None,
                             sampling_steps=S["steps"], cfg_scale=cfg_scale, eta=S["eta"])
    assert U.rel_l2(got, want) <= 1e-6, (cfg_scale, eps_pred, U.rel_l2(got, want))


def test_evaluator_functions_match_reference_source():
  """predict_fn / create_noised_pred_fn / eval_patch_fn / eval_loss_fn (train_ae.py:384-470, lifted by ast) against the
  oracle compositions that tests/test_model_gpu.py::test_evaluator_predict_functions_match_oracle holds the engine to."""
  E = RG.EVAL
  gold = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_sampler_golden.pt"))["evaluators"]
  model, ocfg = U.make_models(**E["model"])
  params = to64(U.cpu_tree(U.perturb_init(model, E["param_seed"], "cpu")))
  image, noise, t, mn = RG.eval_inputs()
  assert RG.digest(image) + RG.digest(noise) + RG.digest(mn) == gold["input_digest"]
  image, noise = image.double(), noise.double()
  gd = O.gaussian_diffusion_tables("cosine", 1000)
  n, C = E["n"], 3
  z = torch.zeros(n, 1, dtype=torch.int32)
  _, o = O.model_apply(params, ocfg, image, t=z, dtype=F64)
  assert torch.allclose(o["pre_logits"], gold["predict_pre_logits"], rtol=0, atol=1e-10)
  t50 = torch.full((n, 1), E["t_noised"], dtype=torch.int32)
  _, o = O.model_apply(params, ocfg, O.q_sample(gd, image, t50, noise), t=t50 + 1, dtype=F64)
  assert torch.allclose(o["pre_logits"], gold["noised_pre_logits"], rtol=0, atol=1e-10)
  pred, o = O.model_apply(params, ocfg, image, t=z, mask=E["mask_ratio_no_noise"], mask_noise=mn, dtype=F64)
  assert torch.equal(o["mask"][:, ::4, ::4, 0].reshape(n, -1).to(torch.uint8), gold["patch_mask"])
  assert U.rel_l2(pred[:2, ..., :C], gold["patch_pred_x0"]) <= 1e-6
  x_t = O.q_sample(gd, image, t, noise)
  pred, _ = O.model_apply(params, ocfg, x_t, t=t + 1, dtype=F64)
  loss = (torch.mean((pred[..., C:] - noise) ** 2) + torch.mean((pred[..., :C] - image) ** 2)) / 2
  assert abs(float(loss) - gold["loss"]) <= 1e-11
  assert U.rel_l2(x_t[:2], gold["x_t"]) <= 1e-6 and U.rel_l2(pred[:2, ..., :C], gold["pred_x0"]) <= 1e-6
  assert U.rel_l2(O.predict_xstart_from_eps(gd, x_t, t, pred[..., C:])[:2], gold["pred_x0_eps"]) <= 1e-6


def test_whole_update_fn_matches_reference_source():
  """The reference's entire update_fn (train_ae.py:291-382, lifted; library calls stood in) on the umd_lbl_s4 case: its own RNG
  split tree, batch split, label slicing, q_sample and loss must land on the loss the oracle's update_step computes from
  the same draws; l2_params is the norm of the UPDATED parameters, l2_updates of the updates, and the EMA moves from the
  old average towards the new parameters by ema_decay (train_ae.py:366-377)."""
  import math
  S = RG.STEP
  gold = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_sampler_golden.pt"))["update_fn"]
  model, ocfg, tkw, params, batch, rand, n_noise, want = rebuild(S["case"])
  assert gold["training_loss"] == want["loss"]                       # update_fn and the separately lifted loss_fn agree exactly
  p64 = to64(params)
  state = {"params": p64, "gd": O.gaussian_diffusion_tables("cosine", 1000), "opt": O.init_opt_state(p64)}
  hp = dict(clip_norm=1.0, peak_lr=1e-3, warmup_steps=0, total_steps=1000, b1=0.9, b2=0.95, wd=0.05)
  _, meas, _ = O.update_step(state, batch, ocfg, tkw, hp, rand, dtype=F64)
  # update_step forms x_t in float32 (as the reference does on device, train_ae.py:183-185) even when the model runs in float64
  assert abs(meas["training_loss"] - gold["training_loss"]) <= 1e-9
  upd = RG.step_updates(params)
  flat = RG.flatten(p64)
  assert gold["l2_updates"] == pytest.approx(math.sqrt(sum(float((u ** 2).sum()) for u in upd.values())), rel=1e-12)
  assert gold["l2_params"] == pytest.approx(math.sqrt(sum(float(((flat[k] + upd[k]) ** 2).sum()) for k in flat)), rel=1e-12)
  k = ("final_conv", "bias")
  new_p = flat[k] + upd[k]
  assert torch.allclose(gold["new_param_probe"], new_p, rtol=0, atol=1e-15)
  old_ema = 0.9 * flat[k]
  assert torch.allclose(gold["new_ema_probe"], old_ema + S["ema_decay"] * (new_p - old_ema), rtol=0, atol=1e-15)
  assert gold["state_keys"] == ["ema_params", "gd", "opt", "params", "rng"]
