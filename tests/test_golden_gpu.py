"""GPU path against the committed known-answer vectors (tests/golden/umd_golden.pt; provenance in
tests/golden/make_golden.py).  Nothing here runs the oracle: the fixture travels to the GPU box.
Tolerances are those of SURVEY.md App. G (bf16 tensor-core operands vs an fp32 reference); integer outputs
(mask permutations) must match bit-exactly."""
import os

import pytest
import torch

from tests import util as U
from tests.golden import make_golden as MG

pytestmark = pytest.mark.gpu
DEV = "cuda"
GOLD = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "umd_golden.pt"))


def test_mask_argsort_matches_golden_bit_exactly():
  from small_vision_b200.model import mask_argsort
  m = GOLD["masking"]
  for keep in (64, 160):
    a, b, mask = mask_argsort(m["noise"].to(DEV), keep)
    assert torch.equal(a.cpu().to(torch.int16), m["ids_shuffle"])
    assert torch.equal(b.cpu().to(torch.int16), m["ids_restore"])
    want_mask = (m["ids_restore"].long() >= keep).float()
    assert torch.equal(mask.cpu(), want_mask)


@pytest.mark.parametrize("name", sorted(MG.CASES))
def test_update_fn_matches_golden_step(name):
  from small_vision_b200.config import TrainConfig
  from small_vision_b200.params import tree_from_arena
  from small_vision_b200.train import create_train_state, make_update_fn
  from oracle.umd_oracle import flatten_tree   # tree helper only; no oracle arithmetic runs here
  mkw, tkw, B, n_noise = MG.CASES[name]
  want = GOLD["cases"][name]
  model, _ = U.make_models(**mkw)
  tcfg = TrainConfig(batch_size=B, total_steps=MG.HP["total_steps"], warmup_steps=0, peak_lr=MG.HP["peak_lr"], **tkw)
  # the golden step used peak_lr as the *scaled* rate: undo train_ae.py:136's B/256 scaling here
  tcfg.peak_lr = MG.HP["peak_lr"] * 256 / B
  params = U.perturb_init(model, 0, DEV)
  state = create_train_state(model, tcfg, seed=0, device=DEV, params=params)
  batch, rand = U.make_batch(model, B, n_noise=n_noise, seed=100, use_labels=tkw["use_labels"])
  assert MG.digest(batch["image"]) + MG.digest(rand["noise"]) + MG.digest(rand["mask_noise_clean"]) == want["input_digest"]
  gb = U.to_dev(batch, DEV)
  gb["_rand"] = U.to_dev(rand, DEV)
  update_fn = make_update_fn(model, tcfg)
  state, meas = update_fn(state, gb)
  torch.cuda.synchronize()
  loss = float(meas["training_loss"])
  assert abs(loss - want["training_loss"]) <= U.TOL_LOSS_REL * abs(want["training_loss"]), (loss, want["training_loss"])
  gn = float(meas["grad_norm"])
  assert abs(gn - want["grad_norm"]) <= U.TOL_GNORM_REL * want["grad_norm"], (gn, want["grad_norm"])
  for k in ("l2_params", "l2_updates"):
    assert abs(float(meas[k]) - want[k]) <= 2e-2 * abs(want[k]) + 1e-6, k
  grads = flatten_tree(tree_from_arena(model.layout, update_fn.grads()[:model.layout.total]))
  total = want["grad_norm"]
  for path, g in grads.items():
    ref = want["grad_leaf_norms"]["/".join(path)]
    mine = float(g.double().norm())
    assert abs(mine - ref) <= 5e-2 * ref + 2e-4 * total, (path, mine, ref)
  got_b = grads[("final_conv", "bias")].cpu()
  assert U.rel_l2(got_b, want["grad_final_conv_bias"]) <= U.TOL_GRAD_REL_L2 or \
      float((got_b - want["grad_final_conv_bias"]).norm()) <= 2e-4 * total
