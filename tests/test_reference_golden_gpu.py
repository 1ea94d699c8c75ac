"""CUDA path against tests/golden/reference_golden.pt — numbers produced by executing the reference's own source
(ae.py, vit.py, embeddings.py, gaussian_diffusion.py, train_ae.py's loss_fn) over the numpy-fp64 jax/flax stand-in
(tests/golden/make_reference_golden.py).  Nothing here runs the oracle or reads /root/reference.  Mask outputs must
match bit-exactly; floating point within SURVEY.md App. G (bf16 tensor-core operands vs exact arithmetic)."""
import os

import pytest
import torch

from tests import util as U
from tests.golden import make_reference_golden as RG

pytestmark = pytest.mark.gpu
DEV = "cuda"
GOLD = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_golden.pt"))


def rebuild(name):
  mkw, tkw, B, n_noise = RG.CASES[name]
  model, _ = U.make_models(**mkw)
  params = U.perturb_init(model, RG.PARAM_SEED, DEV)
  batch, rand = U.make_batch(model, B, n_noise=n_noise, seed=RG.BATCH_SEED, use_labels=tkw["use_labels"])
  want = GOLD["cases"][name]
  assert RG.digest(batch["image"]) + RG.digest(rand["noise"]) + RG.digest(rand["mask_noise_clean"]) == want["input_digest"]
  flat = sorted(RG.flatten(U.cpu_tree(params)).items())
  assert RG.digest(torch.cat([v.reshape(-1) for _, v in flat])) == want["param_digest"]
  return model, tkw, params, batch, rand, n_noise, want


def check_branch(pred, out, want, patch=4):
  pred = pred.float().cpu()
  assert torch.isfinite(pred).all()
  r = U.rel_l2(pred[0], want["pred0"])
  assert r <= U.TOL_PRED_REL_L2, f"pred rel-L2 {r}"
  assert U.rel_l2(pred.double().mean(dim=(1, 2)), want["pred_sample_means"]) <= U.TOL_PRED_REL_L2
  assert abs(float(pred.abs().mean()) - want["pred_abs_mean"]) <= 1e-2 * want["pred_abs_mean"]
  assert U.rel_l2(out["pre_logits"].cpu(), want["pre_logits"]) <= U.TOL_PRED_REL_L2
  if "patch_mask" in want:
    seq = out["mask"][:, ::patch, ::patch, 0].reshape(pred.shape[0], -1).cpu()
    assert torch.equal(seq.to(torch.uint8), want["patch_mask"]), "token mask must match the reference bit-exactly"
    # every pixel of a patch carries the patch's value (ae.py:30-36)
    m = out["mask"].cpu()
    assert torch.equal(m, m[:, ::patch, ::patch].repeat_interleave(patch, 1).repeat_interleave(patch, 2))
  else:
    assert out["mask"] is None


@pytest.mark.parametrize("name", sorted(RG.CASES))
def test_training_forward_matches_reference_source(name):
  """Model.apply(train=True, mask=..., rngs=...) exactly as train_ae.py:325-347 calls it, per branch."""
  from small_vision_b200.diffusion import create_gaussian_diffusion, q_sample, to_device
  model, tkw, params, batch, rand, n_noise, want = rebuild(name)
  img = batch["image"].to(DEV)
  gd = to_device(create_gaussian_diffusion(RG.SCHEDULE.get(name, "cosine"), 1000), DEV)
  patch = model.cfg.patch
  x_t = q_sample(gd=gd, x_start=img[:n_noise].contiguous(), t=rand["t"].to(DEV), noise=rand["noise"].to(DEV))
  assert U.rel_l2(x_t.cpu(), want["x_t"]) <= 1e-5
  if "clean" in want:
    nc = img.shape[0] - n_noise
    pred, out = model.apply({"params": params}, img[n_noise:].contiguous(),
                            t=torch.zeros(nc, 1, dtype=torch.int32, device=DEV), train=True,
                            mask=tkw["mask_ratio_no_noise"], rngs={"mae_noise": rand["mask_noise_clean"].to(DEV)})
    check_branch(pred, out, want["clean"], patch)
  if "noise" in want:
    rngs = {"mae_noise": rand["mask_noise_noise"].to(DEV)}
    if "label_drop_noise" in rand:
      rngs["cfg"] = rand["label_drop_noise"].to(DEV)
    y = batch["label"][:n_noise].to(DEV) if tkw["use_labels"] else None
    pred, out = model.apply({"params": params}, x_t, t=rand["t"].to(DEV) + 1, y=y, train=True, mask=tkw["mask_ratio"],
                            rngs=rngs)
    check_branch(pred, out, want["noise"], patch)


@pytest.mark.parametrize("name", sorted(RG.CASES))
def test_update_fn_loss_and_gradient_slopes_match_reference_source(name):
  """update_fn's loss against the reference's loss_fn, and its gradient against the slopes of the reference's loss
  along the fixture's seeded parameter directions (3 over the whole tree, one per top-level group)."""
  from small_vision_b200.config import TrainConfig
  from small_vision_b200.params import tree_from_arena
  from small_vision_b200.train import create_train_state, make_update_fn
  model, tkw, params, batch, rand, n_noise, want = rebuild(name)
  B = batch["image"].shape[0]
  mkw = RG.CASES[name][0]
  tcfg = TrainConfig(batch_size=B, total_steps=1000, warmup_steps=0, peak_lr=2e-3, beta_schedule=RG.SCHEDULE.get(name, "cosine"),
                     diffusion_space=(mkw.get("img_size", 64), mkw.get("img_size", 64), mkw.get("channels", 3)), **tkw)
  state = create_train_state(model, tcfg, seed=0, device=DEV, params=params)
  dirs = RG.directions(U.cpu_tree(params))          # before the step: update_fn rewrites the arena in place
  gb = U.to_dev(batch, DEV)
  gb["_rand"] = U.to_dev(rand, DEV)
  update_fn = make_update_fn(model, tcfg)
  state, meas = update_fn(state, gb)
  torch.cuda.synchronize()
  loss = float(meas["training_loss"])
  assert abs(loss - want["loss"]) <= U.TOL_LOSS_REL * abs(want["loss"]), (loss, want["loss"])
  grads = RG.flatten(U.cpu_tree(tree_from_arena(model.layout, update_fn.grads()[:model.layout.total])))
  assert [g for g, _ in dirs] == [g for g, _, _ in want["slopes"]]
  mine = torch.tensor([float(sum((grads[k].double() * d[k]).sum() for k in grads)) for _, d in dirs], dtype=torch.float64)
  ref = torch.tensor([s for _, s, _ in want["slopes"]], dtype=torch.float64)
  assert U.cosine(mine, ref) >= U.TOL_GRAD_COS, (mine.tolist(), ref.tolist())
  assert U.rel_l2(mine, ref) <= 2 * U.TOL_GRAD_REL_L2, (mine.tolist(), ref.tolist())


@pytest.mark.parametrize("name", [n for n in sorted(RG.CASES) if "cfg" in GOLD["cases"][n]])
def test_classifier_free_guidance_forward_matches_reference_source(name):
  model, tkw, params, batch, rand, n_noise, want = rebuild(name)
  c = want["cfg"]
  pred, out = model.apply({"params": params}, batch["image"][:2].to(DEV), t=c["t"].to(DEV), y=batch["label"][:2].to(DEV),
                          cfg_scale=c["cfg_scale"])
  assert U.rel_l2(pred.float().cpu(), c["pred"]) <= U.TOL_PRED_REL_L2
  assert U.rel_l2(out["pre_logits"].cpu(), c["pre_logits"]) <= U.TOL_PRED_REL_L2


def test_ddim_step_kernel_matches_reference_source():
  """umd_ddim_step through diffusion.ddim_sample against gaussian_diffusion.py:166-211 run as is."""
  from small_vision_b200.diffusion import create_gaussian_diffusion, ddim_sample, to_device
  D = GOLD["diffusion"]
  gd = to_device(create_gaussian_diffusion("cosine", 1000), DEV)
  i = D["inputs"]
  x, eps, nz = (i[k].float().to(DEV).contiguous() for k in ("x", "eps", "noise"))
  t, tn = i["t"].to(DEV), i["t_next"].to(DEV)
  for s in D["ddim_steps"]:
    r = ddim_sample(gd, lambda **kw: eps, x, t, tn if s["use_next"] else None, 0, clip_denoised=s["clip"], eta=s["eta"],
                    noise=nz)
    # row 3 sits at t = 999 where sqrt(1/abar - 1) ~ 2e4 amplifies fp32 round-off of x_t and eps
    assert U.rel_l2(r["pred_xstart"].cpu(), s["pred_xstart"].float()) <= 1e-4, s
    assert U.rel_l2(r["sample"].cpu(), s["sample"].float()) <= 1e-4, s


def test_sampler_matches_reference_source():
  """diffusion.create_apply_fn + ddim_sample_loop (forward kernels on the doubled batch + umd_ddim_step) against the
  reference's create_apply_fn (train_ae.py:472-483) under its ddim_sample_loop, executed over the jax stand-in."""
  from small_vision_b200 import diffusion as Dm
  S = RG.SAMPLER
  gold = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_sampler_golden.pt"))["sampler"]
  mkw = RG.CASES[S["case"]][0]
  model, _ = U.make_models(**mkw)
  params = U.perturb_init(model, RG.PARAM_SEED, DEV)
  g = torch.Generator().manual_seed(S["noise_seed"])
  noises = [torch.randn(S["n"], 64, 64, 3, generator=g) for _ in range(S["steps"] + 2)]
  assert RG.digest(torch.stack(noises)) == gold["noise_digest"]
  gd = Dm.to_device(Dm.create_gaussian_diffusion("cosine", 1000), DEV)
  ys = torch.tensor(S["ys"]).to(DEV)
  for (cfg_scale, eps_pred), want in zip(S["variants"], gold["samples"]):
    apply_fn = Dm.create_apply_fn(model, params, eps_pred=eps_pred)
    got, _ = Dm.ddim_sample_loop(gd, apply_fn, 0, torch.zeros(S["n"], 64, 64, 3), ys=ys if cfg_scale is not None else None,
                                 sampling_steps=S["steps"], cfg_scale=cfg_scale, eta=S["eta"], noises=noises)
    a = got["sample"].float().cpu()
    assert torch.isfinite(a).all()
    r = U.rel_l2(a, want)
    assert r <= U.TOL_PRED_REL_L2, (cfg_scale, eps_pred, r)


def test_evaluator_functions_match_reference_source():
  """small-vision_b200/evaluators.py against the reference's evaluator closures (train_ae.py:384-470) run over the stand-in."""
  from small_vision_b200 import evaluators as E
  from small_vision_b200.diffusion import create_gaussian_diffusion, to_device
  cfgE = RG.EVAL
  gold = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_sampler_golden.pt"))["evaluators"]
  model, _ = U.make_models(**cfgE["model"])
  params = U.perturb_init(model, cfgE["param_seed"], DEV)
  image, noise, t, mn = RG.eval_inputs(model.cfg.num_patches)
  assert RG.digest(image) + RG.digest(noise) + RG.digest(mn) == gold["input_digest"]
  state = {"params": params, "gd": to_device(create_gaussian_diffusion("cosine", 1000), DEV), "rng": 0}
  batch = {"image": image.to(DEV), "_rand": {"noise": noise.to(DEV), "t": t.to(DEV), "mae_noise": mn.to(DEV)}}
  _, out = E.make_predict_fn(model)(state, batch)
  assert U.rel_l2(out["pre_logits"].cpu(), gold["predict_pre_logits"]) <= U.TOL_PRED_REL_L2
  _, out = E.create_noised_pred_fn(model, cfgE["t_noised"])(state, batch)
  assert U.rel_l2(out["pre_logits"].cpu(), gold["noised_pre_logits"]) <= U.TOL_PRED_REL_L2
  px0, mask = E.make_eval_patch_fn(model, cfgE["mask_ratio_no_noise"])(state, batch)
  assert torch.equal(mask[:, ::4, ::4, 0].reshape(cfgE["n"], -1).cpu().to(torch.uint8), gold["patch_mask"])
  assert U.rel_l2(px0[:2].cpu(), gold["patch_pred_x0"]) <= U.TOL_PRED_REL_L2
  loss, x_t, pred_x0, pred_x0_eps = E.make_eval_loss_fn(model)(state, batch)
  assert abs(float(loss) - gold["loss"]) <= U.TOL_LOSS_REL * abs(gold["loss"])
  assert U.rel_l2(x_t[:2].cpu(), gold["x_t"]) <= 1e-5
  assert U.rel_l2(pred_x0[:2].cpu(), gold["pred_x0"]) <= U.TOL_PRED_REL_L2
  assert U.rel_l2(pred_x0_eps[:2].cpu(), gold["pred_x0_eps"]) <= U.TOL_PRED_REL_L2
