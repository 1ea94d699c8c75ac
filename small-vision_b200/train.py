"""`update_fn(train_state, batch)` — the training step of big_vision/trainers/train_ae.py:287-382 on the
CUDA engine: q_sample noising, the two-branch (noise / clean-MAE) loss, backward, the data-parallel
gradient all-reduce, global-norm clip + AdamW (bf16 mu) + EMA, and the step measurements.

train_state keeps the reference's keys: params, opt, rng, gd, [ema_params].  It is *donated*
(train_ae.py:289): the arenas are updated in place and the same objects are returned.
"""
from __future__ import annotations

import ctypes as C
import dataclasses

import torch

from . import lib
from .config import TrainConfig, warmup_cosine_lr
from .diffusion import create_gaussian_diffusion, q_sample, seed_from_rng, to_device
from .model import ViTAE, mask_argsort
from .params import ParamTree, arena_from_tree, tree_from_arena
from .sharding import GradientReducer

N_EXTRA = 64  # scalar slots appended to the gradient arena (slot 0 = loss) so they ride on the last all-reduce bucket


def create_train_state(model: ViTAE, tcfg: TrainConfig, *, seed: int = 0, device="cuda", nonzero_adaln: bool = False,
                       params=None):
  """train_ae.py:172-201,263-274: params (+ bf16 shadow), optimiser state, rng, schedule tables, optional EMA."""
  layout = model.layout
  if params is None:
    arena = model.init({"params": seed}, device=device, nonzero_adaln=nonzero_adaln)["params"].arena
  else:
    arena = arena_from_tree(layout, params, device)
  state = {
      "params": tree_from_arena(layout, arena),
      "opt": {
          "count": 0,  # ScaleByAdamState.count == ScaleByScheduleState.count (optax.py:30-41 reads it back)
          "mu": tree_from_arena(layout, torch.zeros(layout.total, dtype=torch.bfloat16, device=device)),
          "nu": tree_from_arena(layout, torch.zeros(layout.total, dtype=torch.float32, device=device)),
      },
      "rng": torch.tensor([seed + 1, 0], dtype=torch.int64),
      "gd": to_device(create_gaussian_diffusion(tcfg.beta_schedule, tcfg.timesteps), device),
  }
  if tcfg.ema_decay:
    state["ema_params"] = tree_from_arena(layout, arena.clone())  # train_ae.py:276-282
  return state


class _Scratch:
  pass


def make_update_fn(model: ViTAE, tcfg: TrainConfig, *, process_group=None):
  """Builds update_fn(train_state, batch) -> (train_state, measurements).

  batch: {"image": f32[B,H,W,C] in [-1,1] (this rank's shard), "label": int[B]} on the GPU.
  For parity runs batch may carry "_rand": {"t": int[n_noise], "noise": f32[n_noise,H,W,C],
  "mask_noise_noise": f32[n_noise,L], "mask_noise_clean": f32[n_clean,L], "label_drop_noise": bool[n_noise]},
  the draws the reference makes at train_ae.py:302-317 / ae.py:14 / embeddings.py:44.
  """
  tcfg = tcfg.resolved()
  cfg = model.cfg
  layout = model.layout
  reducer = GradientReducer(layout, process_group)
  sc = _Scratch()
  sc.grads = None
  wd_flags = {}

  def draw_step_randoms(train_state, B, dev, *, supplied=None, rank=0):
    """The draws the reference makes inside one step (train_ae.py:302-317, ae.py:14, embeddings.py:44), in a fixed
    order from one generator: t ~ U{0..T-1} [n_noise], noise ~ N(0,1), the two branches' mask uniforms [n, L] and the
    label-drop mask.  Entries of `supplied` (batch["_rand"]) replace the corresponding draws.  train_state["rng"] stays
    replicated; every data-parallel rank draws for its own shard (i.i.d. over the global batch like the reference), so
    the rank is mixed into the seed."""
    supplied = supplied or {}
    n_clean = int(B * tcfg.no_noise_prob)          # train_ae.py:304
    n_noise = B - n_clean
    L = cfg.num_patches
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed_from_rng(train_state["rng"], rank=rank))
    T = train_state["gd"]["betas"].numel()
    out = {}
    out["t"] = supplied["t"].reshape(-1).to(device=dev, dtype=torch.int32) if "t" in supplied else \
        torch.randint(0, T, (n_noise,), device=dev, generator=gen, dtype=torch.int32)
    shape = (n_noise, cfg.img_size, cfg.img_size, cfg.channels)
    out["noise"] = supplied["noise"].to(dev).contiguous() if "noise" in supplied else \
        torch.randn(shape, device=dev, generator=gen)
    if tcfg.mask_ratio > 0.0 and n_noise > 0:
      out["mask_noise_noise"] = supplied["mask_noise_noise"].to(dev).float().contiguous() \
          if "mask_noise_noise" in supplied else torch.rand(n_noise, L, device=dev, generator=gen)
    if n_clean > 0:
      out["mask_noise_clean"] = supplied["mask_noise_clean"].to(dev).float().contiguous() \
          if "mask_noise_clean" in supplied else torch.rand(n_clean, L, device=dev, generator=gen)
    if cfg.num_classes is not None and tcfg.use_labels and n_noise > 0:
      out["label_drop_noise"] = supplied["label_drop_noise"].to(dev) if "label_drop_noise" in supplied else \
          (torch.rand(n_noise, device=dev, generator=gen) < cfg.cfg_dropout_rate)
    return out

  def forward_backward(train_state, batch, *, rand_rank=None, reduce=True):
    """Draws, q_sample, forward of both branches, loss and backward; with reduce the gradient arena is mean-all-reduced
    over the process group.  Returns (arena, shadow, grads, loss_slot).  rand_rank overrides the rank that seeds the
    draws (bench.py's data-parallel check replays other ranks' shards on rank 0)."""
    images = batch["image"]
    assert images.is_cuda and images.dtype == torch.float32, "batch['image'] must be a float32 CUDA tensor"
    images = images.contiguous()
    dev = images.device
    B = images.shape[0]
    n_clean = int(B * tcfg.no_noise_prob)          # train_ae.py:304
    n_noise = B - n_clean
    L = cfg.num_patches
    masked0 = tcfg.mask_ratio > 0.0 and n_noise > 0
    masked1 = n_clean > 0
    keep0 = cfg.len_keep(tcfg.mask_ratio) if masked0 else L
    keep1 = cfg.len_keep(tcfg.mask_ratio_no_noise) if masked1 else L
    rand = draw_step_randoms(train_state, B, dev, supplied=batch.get("_rand"),
                             rank=reducer.rank if rand_rank is None else rand_rank)
    t, noise = rand["t"], rand["noise"]
    ids_shuffle = torch.empty(B, L, dtype=torch.int32, device=dev)
    ids_restore = torch.empty(B, L, dtype=torch.int32, device=dev)
    if masked0:
      a, b, _ = mask_argsort(rand["mask_noise_noise"], keep0)
      ids_shuffle[:n_noise], ids_restore[:n_noise] = a, b
    if masked1:
      a, b, _ = mask_argsort(rand["mask_noise_clean"], keep1)
      ids_shuffle[n_noise:], ids_restore[n_noise:] = a, b

    # ---- model inputs: x_t for the noise branch (q_sample, :318-321), x_0 for the clean branch
    model_in = images.clone()
    if n_noise > 0:
      q_sample(gd=train_state["gd"], x_start=images[:n_noise], t=t, noise=noise, out=model_in[:n_noise])
    tm = torch.zeros(B, dtype=torch.int32, device=dev)
    tm[:n_noise] = t + 1                              # :341 (clean branch sees t = 0, :327)
    labels = None
    if cfg.num_classes is not None:
      labels = torch.full((B,), cfg.num_classes, dtype=torch.int32, device=dev)   # y=None -> null class (ae.py:107-110)
      if tcfg.use_labels and n_noise > 0:
        y = batch["label"][:n_noise].to(device=dev, dtype=torch.int32)
        labels[:n_noise] = torch.where(rand["label_drop_noise"], torch.full_like(y, cfg.num_classes), y)

    # ---- forward + loss + backward
    params = train_state["params"]
    arena = arena_from_tree(layout, params, dev)
    if not (isinstance(params, ParamTree) and params.arena is arena):
      train_state["params"] = params = tree_from_arena(layout, arena)
    shadow = model.shadow_of(arena)
    if sc.grads is None or sc.grads.device != dev:
      sc.grads = torch.empty(layout.total + N_EXTRA, dtype=torch.float32, device=dev)
      sc.opt_scratch = torch.empty(4096, dtype=torch.float32, device=dev)
      sc.meas = torch.zeros(4, dtype=torch.float32, device=dev)
    grads = sc.grads
    grads.zero_()
    loss_slot = grads[layout.total:layout.total + 1]
    model.forward_arena(arena, shadow, image=model_in, t=tm, labels=labels, n0=n_noise, n1=n_clean, keep0=keep0,
                        keep1=keep1, masked0=masked0, masked1=masked1, ids_shuffle=ids_shuffle, ids_restore=ids_restore,
                        want_pred=False, train=True, x0=images, noise=noise, loss_out=loss_slot)
    model.backward_arena(arena, shadow, grads, bucket_cb=(lambda e: reducer.on_event(grads, e)) if reduce else None)
    if reduce:
      reducer.finish()                                # implicit GSPMD all-reduce of train_ae.py:364
    return arena, shadow, grads, loss_slot

  def update_fn(train_state, batch):
    arena, shadow, grads, loss_slot = forward_backward(train_state, batch)
    dev = arena.device
    rng = train_state["rng"]

    # ---- optimiser (train_ae.py:365-366 with the chain of :135-151)
    opt = train_state["opt"]
    count = int(opt["count"])
    lr = warmup_cosine_lr(count, peak=tcfg.scaled_peak_lr, warmup_steps=tcfg.warmup_steps, decay_steps=tcfg.total_steps)
    b1, b2 = tcfg.betas
    if dev not in wd_flags:
      wd_flags[dev] = layout.wd_flags(dev)
    a = lib.AdamwArgs()
    a.params, a.grads = lib.ptr(arena), lib.ptr(grads)
    a.mu, a.nu = lib.ptr(opt["mu"].arena), lib.ptr(opt["nu"].arena)
    a.params_bf16 = lib.ptr(shadow)
    a.ema = lib.ptr(train_state["ema_params"].arena) if "ema_params" in train_state else None
    a.wd_flags = lib.ptr(wd_flags[dev])
    a.n = layout.total
    a.clip_norm, a.lr, a.b1, a.b2, a.eps, a.wd = tcfg.clip_norm, lr, b1, b2, 1e-8, tcfg.wd
    a.bias_corr1, a.bias_corr2 = 1.0 - b1 ** (count + 1), 1.0 - b2 ** (count + 1)
    a.ema_decay = tcfg.ema_decay or 0.0
    a.scratch, a.scratch_floats = lib.ptr(sc.opt_scratch), sc.opt_scratch.numel()
    a.measurements = lib.ptr(sc.meas)
    lib.check(lib.load().umd_adamw_step(C.byref(a), lib.current_stream()), "umd_adamw_step")
    model.set_shadow(arena, shadow)                     # the kernel refreshed the parameter shadow ...
    if "ema_params" in train_state:                     # ... but moved the EMA arena behind torch's back
      model.invalidate_shadow(train_state["ema_params"].arena)
    opt["count"] = count + 1
    rng = rng.clone()
    rng[1] += 1
    train_state["rng"] = rng
    # ---- measurements (train_ae.py:367-371); device scalars, read them only at log steps (:643-652)
    m = torch.cat([loss_slot, sc.meas[:3]])
    measurements = {"training_loss": m[0], "l2_params": m[1], "l2_updates": m[2], "grad_norm": m[3]}
    return train_state, measurements

  update_fn.grads = lambda: sc.grads
  update_fn.forward_backward = forward_backward
  update_fn.draw_step_randoms = draw_step_randoms
  return update_fn
