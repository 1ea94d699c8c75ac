"""Regression tests for state that is mutated through the C ABI behind torch's back (round-1 advisor findings):
the bf16 shadow of an arena that umd_adamw_step rewrites through raw pointers, the evaluator closures on a real
`create_train_state` state (rng is a [seed, step] tensor), and per-rank draws under data parallelism."""
import pytest
import torch

from tests import util as U

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _dit_setup(ema_decay=0.5):
  from small_vision_b200.config import TrainConfig
  from small_vision_b200.train import create_train_state, make_update_fn
  model, _ = U.make_models("S/4", adaln=True, num_classes=10, depth=2, dec_depth=1)
  tcfg = TrainConfig(batch_size=8, total_steps=100, warmup_steps=0, peak_lr=1e-3 * 256 / 8, ema_decay=ema_decay,
                     use_labels=True, mask_ratio=0.0, no_noise_prob=0.0)
  params = U.perturb_init(model, 0, DEV)
  state = create_train_state(model, tcfg, seed=0, device=DEV, params=params)
  return model, tcfg, state, make_update_fn(model, tcfg)


def test_ema_shadow_follows_the_optimiser():
  """train -> sample(ema) -> train -> sample(ema): the second EMA forward must see the EMA arena as the optimiser
  kernel left it, not the bf16 copy cached by the first one."""
  from small_vision_b200.params import tree_from_arena
  model, tcfg, state, fn = _dit_setup()
  b, _ = U.make_batch(model, 8, n_noise=8, seed=1, use_labels=True)
  gb = U.to_dev(b, DEV)
  x = gb["image"][:4]
  t = torch.full((4, 1), 300, dtype=torch.int32, device=DEV)
  y = gb["label"][:4]

  def ema_forward():
    pred, _ = model.apply({"params": state["ema_params"]}, x, t=t, y=y)
    # reference: the same values in a brand-new arena object (never seen by the shadow cache)
    fresh = tree_from_arena(model.layout, state["ema_params"].arena.clone())
    ref, _ = model.apply({"params": fresh}, x, t=t, y=y)
    return pred.clone(), ref.clone()

  state, _ = fn(state, gb)
  p1, r1 = ema_forward()
  assert torch.equal(p1, r1)
  for _ in range(3):
    state, _ = fn(state, gb)
  p2, r2 = ema_forward()
  assert torch.isfinite(p2).all()
  assert torch.equal(p2, r2), "EMA forward ran on a stale bf16 shadow"
  assert not torch.equal(p1, p2), "the EMA parameters should have moved between the two samples"
  # the parameter shadow is refreshed by the optimiser kernel itself
  pp, _ = model.apply({"params": state["params"]}, x, t=t, y=y)
  fresh = tree_from_arena(model.layout, state["params"].arena.clone())
  pr, _ = model.apply({"params": fresh}, x, t=t, y=y)
  assert torch.equal(pp, pr)


def test_shadow_tracks_torch_side_writes_and_new_arenas():
  from small_vision_b200.params import tree_from_arena
  model, _ = U.make_models("S/4", adaln=True, depth=1, dec_depth=1)
  params = U.perturb_init(model, 0, DEV)
  x = torch.rand(2, 64, 64, 3, device=DEV) * 2 - 1
  p0, _ = model.apply({"params": params}, x)
  params["embedding"]["kernel"].mul_(1.5)        # write through a view: bumps the arena's version
  p1, _ = model.apply({"params": params}, x)
  assert not torch.equal(p0, p1)
  # a plain nested dict (e.g. an imported checkpoint) packs into a NEW arena each call; same values -> same result
  plain = U.cpu_tree(params)
  q1, _ = model.apply({"params": plain}, x)
  assert torch.equal(q1, p1)
  plain["embedding"]["kernel"] = plain["embedding"]["kernel"] / 1.5
  q0, _ = model.apply({"params": plain}, x)
  assert not torch.equal(q0, q1)


def test_evaluators_run_on_a_real_train_state():
  """make_eval_loss_fn / create_noised_pred_fn draw from train_state["rng"], the int64 [seed, step] pair that
  create_train_state stores (train_ae.py:395-470)."""
  from small_vision_b200 import evaluators as E
  model, tcfg, state, fn = _dit_setup(ema_decay=None)
  b, _ = U.make_batch(model, 8, n_noise=8, seed=2, use_labels=True)
  gb = U.to_dev(b, DEV)
  loss, x_t, px0, px0e = E.make_eval_loss_fn(model, use_labels=True)(state, gb)
  assert torch.isfinite(loss) and x_t.shape == gb["image"].shape and px0.shape == px0e.shape
  loss2, *_ = E.make_eval_loss_fn(model, use_labels=True)(state, gb)
  assert torch.equal(loss, loss2), "same rng -> same draws"
  _, out = E.create_noised_pred_fn(model, 50)(state, gb)
  assert torch.isfinite(out["pre_logits"]).all()
  state, _ = fn(state, gb)                      # the step advances rng[1]
  loss3, *_ = E.make_eval_loss_fn(model, use_labels=True)(state, gb)
  assert not torch.equal(loss, loss3)


def test_ranks_draw_different_randoms():
  from small_vision_b200.config import TrainConfig
  from small_vision_b200.train import create_train_state, make_update_fn
  model, _ = U.make_models("S/4", adaln=True, depth=1, dec_depth=1)
  tcfg = TrainConfig(batch_size=8, total_steps=100, warmup_steps=0)
  state = create_train_state(model, tcfg, seed=0, device=DEV)
  fn = make_update_fn(model, tcfg)
  d = torch.device(DEV)
  r0 = fn.draw_step_randoms(state, 8, d, rank=0)
  r0b = fn.draw_step_randoms(state, 8, d, rank=0)
  r1 = fn.draw_step_randoms(state, 8, d, rank=1)
  for k in r0:
    assert torch.equal(r0[k], r0b[k])
  assert not torch.equal(r0["noise"], r1["noise"]) and not torch.equal(r0["mask_noise_clean"], r1["mask_noise_clean"])


@pytest.mark.parametrize("fmt", ["npz", "ts"])
def test_checkpoint_file_loads_into_the_arena_and_reproduces_the_forward(tmp_path, fmt):
  """A parameter file in either of the reference's on-disk layouts (flat .npz, utils.py:218-236; tensorstore-style
  directory, utils.py:886-1016) -> load_params -> Model.apply: pre_logits and pred equal those of the tree that was saved."""
  from small_vision_b200 import checkpoint as CK
  model, _ = U.make_models("S/4", adaln=True, num_classes=10, depth=2, dec_depth=1)
  params = U.perturb_init(model, 5, DEV)
  x = torch.rand(3, 64, 64, 3, device=DEV) * 2 - 1
  t = torch.tensor([[5], [400], [999]], dtype=torch.int32, device=DEV)
  y = torch.tensor([1, 7, 3], device=DEV)
  pred0, out0 = model.apply({"params": params}, x, t=t, y=y)
  if fmt == "npz":
    path = str(tmp_path / "ck.npz")
    CK.save_checkpoint_np(path, {"params": params, "opt": {"count": 3}})
  else:
    path = str(tmp_path / "ck")
    CK.save_checkpoint_ts({"params": params, "opt": {"count": torch.tensor(3)}}, path, 100, keep=False)
  loaded = CK.load_params(path)                      # plain nested dict of numpy arrays
  pred1, out1 = model.apply({"params": loaded}, x, t=t, y=y)
  assert torch.equal(pred0, pred1) and torch.equal(out0["pre_logits"], out1["pre_logits"])
  sub = CK.load_params(path + ":Encoder/encoder_norm")
  assert sorted(sub) == ["bias", "scale"]
