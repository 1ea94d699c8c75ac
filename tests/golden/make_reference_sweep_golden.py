"""Generates tests/golden/reference_sweep_golden.json: the reference's own forward pass and training loss on the EDGE
configurations that the five recipe-shaped cases of make_reference_golden.py do not reach.

Same provenance as reference_golden.pt: every number is produced by the unmodified files under /root/reference
(`big_vision/models/{ae,vit,embeddings}.py`, `big_vision/gaussian_diffusion.py`, the `loss_fn` closure of
`big_vision/trainers/train_ae.py:323-361` lifted by `ast`) executed over tests/golden/refshim (numpy float64).  What
the sweep adds is breadth along the axes on which `update_fn` branches (train_ae.py:304-361):

  * the batch split `n_no_noise = int(B * no_noise_prob)`: all clean (the noise branch contributes the constant 0.0), all
    noised, an odd batch (5 -> 3 + 2), a quarter clean;
  * mask ratios away from the recipe's (len_keep = int(L (1 - r)) with r = 0.5, 0.9; the unmasked noise branch whose
    loss is the plain mean), tiny token counts (L = 16 -> keep 10 / 4 / 1);
  * shapes: 16 x 16 and 32 x 32 images, one channel, patch 2 with four channels, width 768, three encoder blocks;
  * conditioning: the prepended-token model (adaln=False) WITH labels and label drops, a class-conditional model that is
    trained without labels (every sample takes the null class, ae.py:107-110), both beta schedules, t at both ends of the
    schedule (0 and 999);
  * ties in the mask noise of every case (the stable argsort decides, ae.py:15-16).

tests/test_reference_sweep_cpu.py holds the oracle to these numbers in float64 (<= 1e-11, masks bit for bit, two
finite-difference slopes of the reference's loss) and, in this container, re-executes the reference to check that the
committed file is current.

  python tests/golden/make_reference_sweep_golden.py
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from tests import util as U  # noqa: E402
from tests.golden import make_reference_golden as RG  # noqa: E402

TRAIN = dict(mask_ratio=0.375, mask_ratio_no_noise=0.75, no_noise_prob=0.5, use_labels=False)
SMALL = dict(variant="S/4", adaln=True, depth=1, dec_depth=1, img_size=32)

# name: dict(model=Model kwargs, train=overrides of TRAIN, B=batch, schedule=beta schedule, t=explicit timesteps or None)
SWEEP = {
    "all_clean": dict(model=SMALL, train=dict(no_noise_prob=1.0), B=3),
    "all_noise_masked_half": dict(model=SMALL, train=dict(no_noise_prob=0.0, mask_ratio=0.5), B=3),
    "all_noise_unmasked": dict(model=SMALL, train=dict(no_noise_prob=0.0, mask_ratio=0.0), B=2),
    "odd_batch_5": dict(model=SMALL, train=dict(), B=5),
    "quarter_clean_8": dict(model=dict(SMALL, img_size=16), train=dict(no_noise_prob=0.25), B=8),
    "heavy_masks": dict(model=SMALL, train=dict(mask_ratio=0.75, mask_ratio_no_noise=0.9), B=4),
    "sixteen_tokens": dict(model=dict(SMALL, img_size=16), train=dict(mask_ratio_no_noise=0.95), B=4),
    "one_channel": dict(model=dict(SMALL, channels=1), train=dict(), B=4),
    "patch2_four_channels_linear": dict(model=dict(variant="S/2", adaln=True, depth=1, dec_depth=1, img_size=16, channels=4),
                                        train=dict(), B=4, schedule="linear"),
    "width_768_one_block": dict(model=dict(variant="B/4", adaln=True, depth=1, dec_depth=1, img_size=16), train=dict(), B=4),
    "three_encoder_blocks": dict(model=dict(variant="S/4", adaln=True, depth=3, dec_depth=2, img_size=16), train=dict(), B=4),
    "cond_token_with_labels": dict(model=dict(SMALL, adaln=False, num_classes=5), train=dict(use_labels=True), B=6,
                                   label_drop=[True, False, True]),
    "class_model_trained_without_labels": dict(model=dict(SMALL, num_classes=5), train=dict(), B=4),
    "linear_schedule_pixels": dict(model=SMALL, train=dict(), B=4, schedule="linear"),
    "t_at_both_ends": dict(model=SMALL, train=dict(no_noise_prob=0.0), B=4, t=[0, 999, 1, 998]),
}
PARAM_SEED, BATCH_SEED = 3, 300
FD_STEP = 1e-3
N_DIRECTIONS = 2


def case_config(name):
  c = SWEEP[name]
  tkw = dict(TRAIN, **c["train"])
  B = c["B"]
  n_clean = int(B * tkw["no_noise_prob"])          # train_ae.py:304
  return dict(c["model"]), tkw, B, B - n_clean, c.get("schedule", "cosine")


def make_inputs(name):
  """Parameters, batch and supplied draws of one sweep case (torch CPU generators; regenerated identically by the test)."""
  mkw, tkw, B, n_noise, _ = case_config(name)
  model, ocfg = U.make_models(**mkw)
  params = U.cpu_tree(U.perturb_init(model, PARAM_SEED, "cpu"))
  cfg = model.cfg
  H, C, L = cfg.img_size, cfg.channels, cfg.num_patches
  n_clean = B - n_noise
  g = torch.Generator().manual_seed(BATCH_SEED + sorted(SWEEP).index(name))
  batch = {"image": torch.rand(B, H, H, C, generator=g) * 2 - 1,
           "label": torch.randint(0, max(cfg.num_classes or 1, 1), (B,), generator=g)}
  mn, mc = torch.rand(n_noise, L, generator=g), torch.rand(n_clean, L, generator=g)
  if n_noise > 0:                                   # ties, also across the keep boundary
    mn[0, 1] = mn[0, L - 2]
    mn[-1, 0] = mn[-1, L // 2]
  if n_clean > 0:
    mc[0, L - 1] = mc[0, 2]
  t = SWEEP[name].get("t")
  rand = {"t": (torch.tensor(t, dtype=torch.int32).reshape(-1, 1) if t is not None else
                torch.randint(0, 1000, (n_noise, 1), generator=g, dtype=torch.int32)),
          "noise": torch.randn(n_noise, H, H, C, generator=g), "mask_noise_noise": mn, "mask_noise_clean": mc}
  assert rand["t"].shape[0] == n_noise
  if tkw["use_labels"]:
    drop = SWEEP[name].get("label_drop")
    rand["label_drop_noise"] = torch.tensor(drop) if drop is not None else torch.rand(n_noise, generator=g) < 0.1
    assert rand["label_drop_noise"].shape[0] == n_noise
  return model, ocfg, tkw, params, batch, rand, n_noise


def directions(params, n=N_DIRECTIONS, seed=977):
  """n seeded unit-norm directions over the whole parameter tree (the first n of RG.directions(params, seed), without
  drawing the per-group ones that follow them)."""
  flat = RG.flatten(params)
  g = torch.Generator().manual_seed(seed)
  out = []
  for _ in range(n):
    d = {k: torch.randn(tuple(np.shape(v)), generator=g, dtype=torch.float64) for k, v in flat.items()}
    nrm = float(torch.sqrt(sum((x ** 2).sum() for x in d.values())))
    out.append(("*", {k: x / nrm for k, x in d.items()}))
  return out


def run_reference(name, ae, gdm, loss_code, loss_only=False):
  """The reference's forward of both branches and its loss_fn on one sweep case (loss_only: just the loss)."""
  import jax
  import jax.numpy as jnp
  mkw, tkw, B, n_noise, schedule = case_config(name)
  _, _, _, params_t, batch, rand, _ = make_inputs(name)
  n_clean = B - n_noise
  np64, Config = RG.np64, RG.Config
  params = RG.tree64(params_t)
  model = ae.Model(**mkw)                                                       # ae.py:220-222
  gd = gdm.create_gaussian_diffusion(schedule, 1000)                             # gaussian_diffusion.py:32-66
  images = jnp.asarray(np64(batch["image"]))
  x0_noise, x0_clean = images[:n_noise], images[n_noise:]                        # train_ae.py:307-308
  t = jnp.asarray(rand["t"].numpy().astype(np.int32))
  noise = jnp.asarray(np64(rand["noise"]))
  x_t = gdm.q_sample(gd=gd, x_start=x0_noise, t=t, noise=noise) if n_noise > 0 else x0_noise
  labels = jnp.asarray(batch["label"][:n_noise].numpy().astype(np.int32)) if tkw["use_labels"] else None
  img_size, channels = mkw.get("img_size", 64), mkw.get("channels", 3)
  patch = int(mkw["variant"].split("/")[1])

  def keys():
    drop = rand.get("label_drop_noise")
    return dict(
        rng_model=jax.Key(), rng_model_noise=jax.Key(),
        mae_noise_rng=jax.Key({"uniform": np64(rand["mask_noise_clean"])}),
        mae_noise_rng_noise=jax.Key({"uniform": np64(rand["mask_noise_noise"])}),
        cfg_rng=jax.Key({"bernoulli": np.zeros((n_clean,))}),
        cfg_rng_noise=jax.Key({"bernoulli": np64(drop.double()) if drop is not None else np.zeros((n_noise,))}))

  def reference_loss(p):
    env = dict(jnp=jnp, model=model, config=Config(diffusion_space=(img_size, img_size, channels), **tkw), B=B, n_noise=n_noise,
               n_no_noise=n_clean, x_0_noise=x0_noise, x_0_no_noise=x0_clean, x_t_noise=x_t, batched_t=t,
               labels_t=labels, noise=noise, **keys())
    exec(loss_code, env)
    return float(env["loss_fn"](p))

  out = {"loss": reference_loss(params), "n_noise": n_noise, "n_clean": n_clean,
         "input_digest": RG.digest(batch["image"]) + RG.digest(rand["mask_noise_noise"]) + RG.digest(rand["mask_noise_clean"]),
         "param_digest": RG.digest(torch.cat([v.reshape(-1) for _, v in sorted(RG.flatten(params_t).items())]))}
  if loss_only:
    return out
  k = keys()

  def pack(pred, o):
    pred = np.asarray(pred, dtype=np.float64)
    pl = np.asarray(o["pre_logits"], dtype=np.float64)
    d = {"pred_sample_means": pred.mean(axis=(1, 2)).tolist(), "pred_abs_mean": float(np.abs(pred).mean()),
         "pre_logits_sample_means": pl.mean(axis=1).tolist(), "pre_logits_abs_mean": float(np.abs(pl).mean())}
    if o["mask"] is not None:
      m = np.asarray(o["mask"])[:, ::patch, ::patch, 0]
      assert set(np.unique(m)) <= {0.0, 1.0}
      d["patch_mask"] = ["".join(str(int(v)) for v in row.reshape(-1)) for row in m]
    return d
  if n_clean > 0:
    pred, o = model.apply({"params": params}, x0_clean, t=jnp.zeros((n_clean, 1), dtype=jnp.int32), train=True,
                          mask=tkw["mask_ratio_no_noise"],
                          rngs={"dropout": k["rng_model"], "cfg": k["cfg_rng"], "mae_noise": k["mae_noise_rng"]})
    out["clean"] = pack(pred, o)
  if n_noise > 0:
    pred, o = model.apply({"params": params}, x_t, t=t + 1, y=labels, train=True, mask=tkw["mask_ratio"],
                          rngs={"dropout": k["rng_model_noise"], "cfg": k["cfg_rng_noise"], "mae_noise": k["mae_noise_rng_noise"]})
    out["noise"] = pack(pred, o)
  slopes = []
  for _, d in directions(params_t):
    lp, lm = reference_loss(RG.shifted(params, d, FD_STEP)), reference_loss(RG.shifted(params, d, -FD_STEP))
    lp2, lm2 = reference_loss(RG.shifted(params, d, FD_STEP / 2)), reference_loss(RG.shifted(params, d, -FD_STEP / 2))
    d1, d2 = (lp - lm) / (2 * FD_STEP), (lp2 - lm2) / FD_STEP
    slopes.append([(4 * d2 - d1) / 3, abs(d2 - d1)])          # Richardson-extrapolated slope and its error scale
  out["slopes"] = slopes
  return out


def main():
  ae, gdm, loss_code, loss_lines = RG.load_reference()
  out = {"provenance": "models/ae.py, models/vit.py, models/embeddings.py, gaussian_diffusion.py and loss_fn of "
                       "trainers/train_ae.py:%d-%d of the reference, executed over tests/golden/refshim (numpy float64)" % loss_lines,
         "cases": {}}
  for name in sorted(SWEEP):
    out["cases"][name] = run_reference(name, ae, gdm, loss_code)
    print(name, out["cases"][name]["loss"], out["cases"][name]["slopes"])
  path = os.path.join(HERE, "reference_sweep_golden.json")
  with open(path, "w") as f:
    json.dump(out, f, indent=1, sort_keys=True)
  print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
  main()
