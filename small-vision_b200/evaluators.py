"""Evaluator predict functions of the reference trainer (big_vision/trainers/train_ae.py:384-470; SURVEY.md §8f
rank 3): thin callers of the same forward path — representation at t = 0, representation of a noised input,
MAE reconstruction with its pixel mask, and the denoising evaluation loss.  Same names, arguments
(`train_state`, `batch`) and return values as the reference; the random draws the reference takes from
`train_state["rng"]` may be supplied in `batch["_rand"]` (keys "noise", "t", "mae_noise") to reproduce a run.

The few-shot ridge probe that consumes `pre_logits` (evaluators/fewshot_lsr.py) lives in fewshot.py.  The
latent-diffusion VAE (`vae_encode` / `vae_decode`) is not part of this path: with latents, pass the encoded batch.
"""
from __future__ import annotations

import torch

from . import diffusion as _d


def _rand(batch, key):
  r = batch.get("_rand") if isinstance(batch, dict) else None
  return None if r is None else r.get(key)


def _images(model, batch):
  return model._as_input(batch["image"])


def make_predict_fn(model):
  """train_ae.py:384-393: (None, {"pre_logits", "mask"}) at t = 0."""
  def predict_fn(train_state, batch):
    images = _images(model, batch)
    t = torch.zeros((images.shape[0], 1), dtype=torch.int32, device=images.device)
    _, out = model.apply({"params": train_state["params"]}, images, t=t)
    return None, out
  return predict_fn


def create_noised_pred_fn(model, t):
  """train_ae.py:395-413: representation of q_sample(x, t) evaluated at t + 1."""
  def predict_fn(train_state, batch):
    images = _images(model, batch)
    B = images.shape[0]
    batched_t = torch.full((B, 1), int(t), dtype=torch.int32, device=images.device)
    noise = _rand(batch, "noise")
    if noise is None:
      noise = torch.randn(images.shape, device=images.device,
                          generator=_d._as_generator(train_state.get("rng"), images.device, stream=1))
    x_t = _d.q_sample(gd=train_state["gd"], x_start=images, t=batched_t, noise=noise.to(images.device).contiguous())
    _, out = model.apply({"params": train_state["params"]}, x_t, t=batched_t + 1)
    return None, out
  return predict_fn


def make_eval_patch_fn(model, mask_ratio_no_noise, channels=None):
  """train_ae.py:415-434: (pred_x0 [B,H,W,C], mask [B,H,W,1]) of the masked clean input at t = 0."""
  C = channels or model.cfg.channels

  def eval_patch_fn(train_state, batch):
    images = _images(model, batch)
    B = images.shape[0]
    t = torch.zeros((B, 1), dtype=torch.int32, device=images.device)
    mn = _rand(batch, "mae_noise")
    pred, out = model.apply({"params": train_state["params"]}, images, t=t, mask=mask_ratio_no_noise,
                            rngs={"mae_noise": mn if mn is not None else train_state.get("rng")})
    return pred[..., :C], out["mask"]
  return eval_patch_fn


def make_eval_loss_fn(model, use_labels=False, channels=None):
  """train_ae.py:436-470: (loss, x_t, pred_x0, pred_x0_eps) with t ~ U{0..T-1}, the model evaluated at t + 1."""
  C = channels or model.cfg.channels

  def eval_loss_fn(train_state, batch):
    images = _images(model, batch)
    B = images.shape[0]
    gd = train_state["gd"]
    dev = images.device
    labels = batch["label"].to(dev) if use_labels else None
    g = None   # built only when a draw is needed (train_state["rng"] is the [seed, step] pair of create_train_state)
    t = _rand(batch, "t")
    if t is None:
      g = _d._as_generator(train_state.get("rng"), dev, stream=2)
      t = torch.randint(0, int(gd["betas"].numel()), (B, 1), device=dev, generator=g, dtype=torch.int32)
    t = t.to(device=dev, dtype=torch.int32).reshape(B, 1)
    noise = _rand(batch, "noise")
    if noise is None:
      g = g if g is not None else _d._as_generator(train_state.get("rng"), dev, stream=2)
      noise = torch.randn(images.shape, device=dev, generator=g)
    noise = noise.to(dev).contiguous()
    x_t = _d.q_sample(gd=gd, x_start=images, t=t, noise=noise)
    pred, _ = model.apply({"params": train_state["params"]}, x_t, y=labels, t=t + 1)
    pred_eps, pred_x0 = pred[..., C:], pred[..., :C]
    # evaluation-only scalar (two plain means, :462): not on the training path, left to torch
    loss = (torch.mean((pred_eps - noise) ** 2) + torch.mean((pred_x0 - images) ** 2)) / 2
    sra = gd["sqrt_recip_alphas_cumprod"][t.reshape(-1).long()].reshape(B, 1, 1, 1)
    srm1 = gd["sqrt_recipm1_alphas_cumprod"][t.reshape(-1).long()].reshape(B, 1, 1, 1)
    pred_x0_eps = sra * x_t - srm1 * pred_eps    # _predict_xstart_from_eps, gaussian_diffusion.py:122-127
    return loss, x_t, pred_x0, pred_x0_eps
  return eval_loss_fn
