"""GPU parity of the whole path against the CPU oracle (oracle/umd_oracle.py): `Model.apply` forward
(ae.py:176-197) for the five BASELINE.json configurations scaled down to sizes the oracle finishes in
seconds, and full `update_fn` steps (train_ae.py:287-382): loss, every gradient leaf, grad-norm, the
l2 measurements and the updated parameters.  Tolerances: tests/util.py."""
import math

import pytest
import torch

from tests import util as U
from oracle import umd_oracle as O

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _forward_case(variant, *, adaln, num_classes=None, mask=0.0, with_y=False, with_t=True, n=6, seed=0, depth=None,
                  dec_depth=None, img_size=64, channels=3, cfg_scale=None):
  model, ocfg = U.make_models(variant, adaln=adaln, num_classes=num_classes, depth=depth, dec_depth=dec_depth,
                              img_size=img_size, channels=channels)
  params = U.perturb_init(model, seed, DEV)
  oparams = U.cpu_tree(params)
  g = torch.Generator().manual_seed(seed + 7)
  image = torch.rand(n, img_size, img_size, channels, generator=g) * 2 - 1
  t = torch.randint(1, 1001, (n, 1), generator=g, dtype=torch.int32) if with_t else None
  y = torch.randint(0, num_classes, (n,), generator=g) if with_y else None
  mn = torch.rand(n, model.cfg.num_patches, generator=g)
  mn[0, 3] = mn[0, 77]  # a tie
  pred, out = model.apply({"params": params}, image.to(DEV), t=None if t is None else t.to(DEV),
                          y=None if y is None else y.to(DEV), mask=mask, train=False, cfg_scale=cfg_scale,
                          rngs={"mae_noise": mn.to(DEV)})
  torch.cuda.synchronize()
  opred, oout = O.model_apply(oparams, ocfg, image, t=t, y=y, mask=mask, train=False, mask_noise=mn,
                              cfg_scale=cfg_scale)
  assert pred.shape == opred.shape
  assert torch.isfinite(pred).all()
  r = U.rel_l2(pred.cpu(), opred)
  assert r <= U.TOL_PRED_REL_L2, f"pred rel-L2 {r}"
  r2 = U.rel_l2(out["pre_logits"].cpu(), oout["pre_logits"])
  assert r2 <= U.TOL_PRED_REL_L2, f"pre_logits rel-L2 {r2}"
  if mask > 0:
    assert torch.equal(out["mask"].cpu(), oout["mask"]), "pixel mask must match bit-exactly"
  else:
    assert out["mask"] is None and oout["mask"] is None
  return r, r2


def test_forward_umd_s4_noise_branch():
  _forward_case("S/4", adaln=True, mask=0.375)


def test_forward_umd_s4_mae_branch_t0():
  _forward_case("S/4", adaln=True, mask=0.75, with_t=False)


def test_forward_umd_b4_unmasked():
  _forward_case("B/4", adaln=True, mask=0.0, n=4)


def test_forward_mae_b4_no_adaln():
  _forward_case("B/4", adaln=False, mask=0.75, n=4, depth=3, dec_depth=2)


def test_forward_dit_labels():
  _forward_case("S/4", adaln=True, num_classes=10, with_y=True, mask=0.0)


def test_forward_dit_null_label():
  _forward_case("S/4", adaln=True, num_classes=10, with_y=False, mask=0.0)


def test_forward_cfg_scale():
  _forward_case("S/4", adaln=True, num_classes=10, with_y=True, mask=0.0, cfg_scale=1.5, n=4)


def test_forward_latent_l2():
  """Latent-UMD-L/2: 32x32x4 latents, patch 2, width 1024 (depth cut to keep the oracle quick)."""
  _forward_case("L/2", adaln=True, mask=0.375, n=3, img_size=32, channels=4, depth=2, dec_depth=2)


def test_identity_pin_zero_init_adaln():
  """SURVEY.md §8c pin (1): with the reference's zero-init adaLN kernels every block is the identity, so
  pred = ConvT(LN_dec(dec_in)[:, 1:]) and pre_logits = mean of LN_enc(cls)."""
  from small_vision_b200.params import init_arena, tree_from_arena
  model, ocfg = U.make_models("S/4", adaln=True)
  arena = init_arena(model.layout, 3, DEV, nonzero_adaln=False)
  params = tree_from_arena(model.layout, arena)
  g = torch.Generator().manual_seed(0)
  image = torch.rand(4, 64, 64, 3, generator=g) * 2 - 1
  pred, out = model.apply({"params": params}, image.to(DEV), mask=0.0)
  op = U.cpu_tree(params)
  x = torch.einsum("nijabc,abcd->nijd", O.patchify(image, 4), op["embedding"]["kernel"]).reshape(4, 256, -1)
  x = x + op["embedding"]["bias"] + op["pos_embedding"]
  x = torch.cat([op["cls"].expand(4, -1, -1), x], 1)
  x = O.layer_norm(x, op["Encoder"]["encoder_norm"]["scale"], op["Encoder"]["encoder_norm"]["bias"])
  rep = x[:, :4].mean(1)
  xd = torch.cat([rep[:, None], x[:, 4:] + op["dec_pos_embedding"]], 1)
  xd = O.layer_norm(xd, op["Decoder"]["encoder_norm"]["scale"], op["Decoder"]["encoder_norm"]["bias"])[:, 1:]
  ref = O.conv_transpose_unpatchify(xd.reshape(4, 16, 16, -1), op["final_conv"]["kernel"], op["final_conv"]["bias"])
  assert U.rel_l2(pred.cpu(), ref) < 1e-2
  assert U.rel_l2(out["pre_logits"].cpu(), rep) < 1e-4   # no GEMM on this path beyond the fp32 embed


# ------------------------------------------------------------------------------------------ steps
def test_step_umd_s4():
  U.run_step_parity(variant="S/4", batch=8, adaln=True, steps=2, depth=4, dec_depth=2)


def test_step_mae_no_adaln():
  U.run_step_parity(variant="S/4", batch=8, adaln=False, steps=2, depth=3, dec_depth=2)


def test_step_dit_labels_ema():
  U.run_step_parity(variant="S/4", batch=6, adaln=True, num_classes=10, use_labels=True, mask_ratio=0.0,
                    no_noise_prob=0.0, steps=2, depth=3, dec_depth=2, ema_decay=1e-2)


def test_step_umd_with_labels_null_class_on_clean_branch():
  U.run_step_parity(variant="S/4", batch=8, adaln=True, num_classes=10, use_labels=True, steps=1, depth=2, dec_depth=2)


def test_step_latent():
  U.run_step_parity(variant="S/2", batch=4, adaln=True, steps=1, depth=2, dec_depth=2, img_size=32, channels=4,
                    beta_schedule="linear")


def test_step_umd_b4_one_block():
  U.run_step_parity(variant="B/4", batch=4, adaln=True, steps=1, depth=1, dec_depth=1)


def test_first_step_with_warmup_leaves_params_unchanged():
  """SURVEY.md §8c pin (10): optax evaluates the schedule at the pre-increment count, so lr(0) = 0 with
  warm-up and the first step must leave every parameter untouched and report l2_updates = 0."""
  from small_vision_b200.config import TrainConfig
  from small_vision_b200.train import create_train_state, make_update_fn
  model, _ = U.make_models("S/4", adaln=True, depth=2, dec_depth=1)
  tcfg = TrainConfig(batch_size=4, total_steps=100, warmup_steps=10)
  state = create_train_state(model, tcfg, seed=0, device=DEV, nonzero_adaln=True)
  before = state["params"].arena.clone()
  b, rand = U.make_batch(model, 4, n_noise=2, seed=5)
  gb = U.to_dev(b, DEV)
  gb["_rand"] = U.to_dev(rand, DEV)
  state, meas = make_update_fn(model, tcfg)(state, gb)
  assert torch.equal(before, state["params"].arena)
  assert float(meas["l2_updates"]) == 0.0
  assert math.isfinite(float(meas["training_loss"]))


def test_mae_branch_gives_zero_gradient_to_eps_half_of_final_conv():
  """SURVEY.md §8c pin (6): train_ae.py:333 uses only the first C channels of pred on the clean branch."""
  from small_vision_b200.config import TrainConfig
  from small_vision_b200.train import create_train_state, make_update_fn
  from small_vision_b200.params import tree_from_arena
  model, _ = U.make_models("S/4", adaln=True, depth=1, dec_depth=1)
  tcfg = TrainConfig(batch_size=4, no_noise_prob=1.0, total_steps=100, warmup_steps=0)
  state = create_train_state(model, tcfg, seed=0, device=DEV, nonzero_adaln=True)
  b, rand = U.make_batch(model, 4, n_noise=0, seed=5)
  gb = U.to_dev(b, DEV)
  gb["_rand"] = U.to_dev(rand, DEV)
  fn = make_update_fn(model, tcfg)
  fn(state, gb)
  g = tree_from_arena(model.layout, fn.grads()[:model.layout.total])
  gk = g["final_conv"]["kernel"]
  assert float(gk[..., 3:].abs().max()) == 0.0 and float(gk[..., :3].abs().max()) > 0.0
  assert float(g["final_conv"]["bias"][3:].abs().max()) == 0.0


def test_evaluator_predict_functions_match_oracle():
  """train_ae.py:384-470: representation at t = 0 and of a noised input, MAE reconstruction
  with its mask, and the denoising evaluation loss, on supplied draws."""
  from small_vision_b200 import evaluators as E
  from small_vision_b200.diffusion import create_gaussian_diffusion, to_device
  model, ocfg = U.make_models("S/4", adaln=True, num_classes=None, depth=2, dec_depth=1)
  params = U.perturb_init(model, 5, DEV)
  oparams = U.cpu_tree(params)
  gd = create_gaussian_diffusion("cosine", 1000)
  state = {"params": params, "gd": to_device(gd, DEV), "rng": 0}
  g = torch.Generator().manual_seed(9)
  n, C = 4, 3
  image = torch.rand(n, 64, 64, C, generator=g) * 2 - 1
  noise = torch.randn(n, 64, 64, C, generator=g)
  t = torch.randint(0, 1000, (n, 1), generator=g, dtype=torch.int32)
  mn = torch.rand(n, model.cfg.num_patches, generator=g)
  batch = {"image": image.to(DEV), "_rand": {"noise": noise.to(DEV), "t": t.to(DEV), "mae_noise": mn.to(DEV)}}
  # predict_fn
  _, out = E.make_predict_fn(model)(state, batch)
  _, oout = O.model_apply(oparams, ocfg, image, t=torch.zeros(n, 1, dtype=torch.int32))
  assert U.rel_l2(out["pre_logits"].cpu(), oout["pre_logits"]) <= U.TOL_PRED_REL_L2
  # noised representation at t = 50
  _, out = E.create_noised_pred_fn(model, 50)(state, batch)
  t50 = torch.full((n, 1), 50, dtype=torch.int32)
  _, oout = O.model_apply(oparams, ocfg, O.q_sample(gd, image, t50, noise), t=t50 + 1)
  assert U.rel_l2(out["pre_logits"].cpu(), oout["pre_logits"]) <= U.TOL_PRED_REL_L2
  # MAE reconstruction
  px0, mask = E.make_eval_patch_fn(model, 0.75)(state, batch)
  opred, oout = O.model_apply(oparams, ocfg, image, t=torch.zeros(n, 1, dtype=torch.int32), mask=0.75, mask_noise=mn)
  assert torch.equal(mask.cpu(), oout["mask"])
  assert U.rel_l2(px0.cpu(), opred[..., :C]) <= U.TOL_PRED_REL_L2
  # evaluation loss
  loss, x_t, pred_x0, pred_x0_eps = E.make_eval_loss_fn(model)(state, batch)
  ox_t = O.q_sample(gd, image, t, noise)
  opred, _ = O.model_apply(oparams, ocfg, ox_t, t=t + 1)
  oloss = (torch.mean((opred[..., C:] - noise) ** 2) + torch.mean((opred[..., :C] - image) ** 2)) / 2
  assert abs(float(loss) - float(oloss)) <= U.TOL_LOSS_REL * abs(float(oloss))
  assert torch.allclose(x_t.cpu(), ox_t, rtol=1e-5, atol=1e-6)
  assert U.rel_l2(pred_x0.cpu(), opred[..., :C]) <= U.TOL_PRED_REL_L2
  assert U.rel_l2(pred_x0_eps.cpu(), O.predict_xstart_from_eps(gd, ox_t, t, opred[..., C:])) <= U.TOL_PRED_REL_L2
