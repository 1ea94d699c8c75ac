"""Size-independent properties of the CUDA path at the full model size (UMD-B/4, all 12+4 blocks), where the CPU
oracle is too slow to serve as the checker:
  * data parallelism: because every sample masks the same number of patches the loss is a per-sample mean
    (SURVEY.md §8c pin 3, train_ae.py:333-360), so the gradient of a batch is the mean of its shards' gradients;
  * repeatability: the same inputs give the same loss and (up to fp32 atomic-add ordering) the same gradients.
"""
import pytest
import torch

from tests import util as U

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _grads(model, tcfg, params, batch, rand):
  from small_vision_b200.train import create_train_state, make_update_fn
  state = create_train_state(model, tcfg, seed=0, device=DEV, params=params)
  fn = make_update_fn(model, tcfg)
  gb = U.to_dev(batch, DEV)
  gb["_rand"] = U.to_dev(rand, DEV)
  _, meas = fn(state, gb)
  torch.cuda.synchronize()
  return fn.grads()[:model.layout.total].clone(), float(meas["training_loss"])


def test_full_size_gradient_is_mean_of_shard_gradients():
  from small_vision_b200.config import TrainConfig
  from small_vision_b200.params import tree_from_arena
  model, _ = U.make_models("B/4", adaln=True)
  B = 32
  # lr(0) = 0 (warm-up): the step computes gradients but leaves the parameters alone, so all runs see the same ones
  mk = lambda bs: TrainConfig(batch_size=bs, total_steps=1000, warmup_steps=10)
  params = U.perturb_init(model, 0, DEV)
  batch, rand = U.make_batch(model, B, n_noise=B // 2, seed=21)
  g_full, loss_full = _grads(model, mk(B), tree_from_arena(model.layout, params.arena.clone()), batch, rand)
  acc, losses = None, []
  h = B // 4
  for r in range(2):
    idx_n = list(range(r * h, (r + 1) * h))
    idx_c = [B // 2 + i for i in idx_n]
    b = {"image": torch.cat([batch["image"][idx_n], batch["image"][idx_c]]), "label": batch["label"][idx_n + idx_c]}
    rd = {"t": rand["t"][idx_n], "noise": rand["noise"][idx_n], "mask_noise_noise": rand["mask_noise_noise"][idx_n],
          "mask_noise_clean": rand["mask_noise_clean"][idx_n]}
    g, l = _grads(model, mk(B // 2), tree_from_arena(model.layout, params.arena.clone()), b, rd)
    losses.append(l)
    acc = g if acc is None else acc + g
  mean = acc / 2
  assert abs(sum(losses) / 2 - loss_full) <= 2e-3 * abs(loss_full)
  rel = float((mean - g_full).double().norm() / g_full.double().norm())
  assert rel <= 2e-2, rel          # bf16 operand rounding differs with the batch composition of each GEMM tile
  cos = float((mean.double() @ g_full.double()) / (mean.double().norm() * g_full.double().norm()))
  assert cos >= 0.9995, cos


def test_full_size_step_is_repeatable():
  """Same inputs, same loss bit for bit; gradients up to the order of fp32 atomic adds.  Two regimes, measured with
  tools/repeat_check.py (per-leaf deltas over repeated steps):
    * leaves of the main stream (blocks, embeddings, final layers): split-K and column-sum atomics only reorder fp32
      additions -> ~1e-7 relative;
    * the conditioning trunks (time / label Dense layers, label table): their input gradient dcond is the sum of the
      adaLN gradients that the LayerNorm-backward CTAs of one sample add atomically, and is then cast to bf16 for the
      trunk's GEMMs -> an order-dependent last bit occasionally flips a bf16 rounding, which shows as discrete
      ~5e-5 changes in those leaves (1e-6 .. 2e-5 of the whole arena)."""
  from small_vision_b200.config import TrainConfig
  from small_vision_b200.params import tree_from_arena
  model, _ = U.make_models("B/4", adaln=True)
  tcfg = TrainConfig(batch_size=16, total_steps=1000, warmup_steps=10)
  params = U.perturb_init(model, 1, DEV)
  batch, rand = U.make_batch(model, 16, n_noise=8, seed=3)
  g1, l1 = _grads(model, tcfg, tree_from_arena(model.layout, params.arena.clone()), batch, rand)
  g2, l2 = _grads(model, tcfg, tree_from_arena(model.layout, params.arena.clone()), batch, rand)
  assert l1 == l2
  assert torch.isfinite(g1).all()
  cond_leaf = lambda path: path[0] in ("time_trunk", "label_trunk", "label_emb")
  main_d2 = main_n2 = 0.0
  for lf in model.layout.leaves:
    a, b = lf.view(g1).double().reshape(-1), lf.view(g2).double().reshape(-1)
    if cond_leaf(lf.path):
      rel = float((a - b).norm() / (a.norm() + 1e-30))
      assert rel <= 5e-4, ("/".join(lf.path), rel)
    else:
      main_d2 += float((a - b).pow(2).sum())
      main_n2 += float(a.pow(2).sum())
  rel_main = (main_d2 / main_n2) ** 0.5
  assert rel_main <= 2e-6, rel_main
