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
  sc.comm = None
  sc.ws = {}

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

  def run_step(train_state, batch, *, rand_rank=None, reduce=True, optimise=True):
    """One umd_train_step call (include/umd_b200.h): draws -> q_sample -> masking -> forward of both branches -> loss
    -> backward -> gradient mean over the ranks -> clip + AdamW + EMA.  Returns (arena, shadow, grads, meas) with meas =
    device float[4]: loss, l2_params, l2_updates, grad_norm (the last three only with optimise)."""
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
    comm = communicator(dev) if reduce else None
    rand = draw_step_randoms(train_state, B, dev, supplied=batch.get("_rand"),
                             rank=reducer.rank if rand_rank is None else rand_rank)
    params = train_state["params"]
    arena = arena_from_tree(layout, params, dev)
    if not (isinstance(params, ParamTree) and params.arena is arena):
      train_state["params"] = params = tree_from_arena(layout, arena)
    shadow = model.shadow_of(arena)
    if sc.grads is None or sc.grads.device != dev:
      sc.grads = torch.empty(layout.total + N_EXTRA, dtype=torch.float32, device=dev)
      sc.opt_scratch = torch.empty(4096, dtype=torch.float32, device=dev)
      sc.wd_flags = layout.wd_flags(dev)
      sc.bounds = (C.c_longlong * (2 * len(layout.bucket_bounds)))(*[x for lo_hi in layout.bucket_bounds for x in lo_hi])
      sc.events = (C.c_int * len(layout.bucket_events))(*layout.bucket_events)
    grads = sc.grads
    shape = model.step_shape(n_noise, n_clean, keep0, keep1, masked0, masked1)
    wkey = (n_noise, n_clean, keep0, keep1, str(dev))
    ws = sc.ws.get(wkey)
    if ws is None:
      ws = sc.ws[wkey] = torch.empty(lib.train_workspace_bytes(model._mcfg, shape), dtype=torch.uint8, device=dev)
    meas = torch.empty(4, dtype=torch.float32, device=dev)   # fresh per step: the caller may read it at a later log step
    gd = train_state["gd"]
    a = lib.TrainStepArgs()
    a.cfg = C.pointer(model._mcfg)
    a.shape = shape
    a.offsets = C.cast(model._offsets, C.c_void_p)
    o = a.opt
    o.params = lib.ptr(arena)
    o.params_bf16 = lib.ptr(shadow)
    o.n = layout.total
    o.measurements = lib.ptr(meas)
    a.grads = lib.ptr(grads)
    if optimise:
      opt = train_state["opt"]
      count = int(opt["count"])
      b1, b2 = tcfg.betas
      o.mu, o.nu = lib.ptr(opt["mu"].arena), lib.ptr(opt["nu"].arena)
      o.ema = lib.ptr(train_state["ema_params"].arena) if "ema_params" in train_state else None
      o.wd_flags = lib.ptr(sc.wd_flags)
      o.clip_norm, o.b1, o.b2, o.eps, o.wd = tcfg.clip_norm, b1, b2, 1e-8, tcfg.wd
      # optax evaluates the schedule at the pre-increment count (train_ae.py:135-151)
      o.lr = warmup_cosine_lr(count, peak=tcfg.scaled_peak_lr, warmup_steps=tcfg.warmup_steps, decay_steps=tcfg.total_steps)
      o.bias_corr1, o.bias_corr2 = 1.0 - b1 ** (count + 1), 1.0 - b2 ** (count + 1)
      o.ema_decay = tcfg.ema_decay or 0.0
      o.scratch, o.scratch_floats = lib.ptr(sc.opt_scratch), sc.opt_scratch.numel()
    else:
      a.flags = lib.UMD_STEP_NO_OPTIMIZER
    a.image = lib.ptr(images)
    label = None
    if cfg.num_classes is not None and tcfg.use_labels and n_noise > 0:
      label = batch["label"].to(device=dev, dtype=torch.int64).contiguous()
      a.use_labels = 1
    a.label = lib.ptr(label)
    drop = rand["label_drop_noise"].to(torch.uint8) if "label_drop_noise" in rand else None
    a.label_drop = lib.ptr(drop)
    a.t, a.noise = lib.ptr(rand["t"]), lib.ptr(rand["noise"])
    a.mask_noise0 = lib.ptr(rand.get("mask_noise_noise"))
    a.mask_noise1 = lib.ptr(rand.get("mask_noise_clean"))
    a.sqrt_alphas_cumprod = lib.ptr(gd["sqrt_alphas_cumprod"])
    a.sqrt_one_minus_alphas_cumprod = lib.ptr(gd["sqrt_one_minus_alphas_cumprod"])
    a.workspace, a.workspace_bytes = lib.ptr(ws), ws.numel()
    if comm is not None:
      a.comm = comm.handle
      a.bucket_bounds = C.cast(sc.bounds, C.c_void_p)
      a.bucket_events = C.cast(sc.events, C.c_void_p)
      a.num_buckets = len(layout.bucket_bounds)
    lib.check(lib.load().umd_train_step(C.byref(a), lib.current_stream()), "umd_train_step")
    return arena, shadow, grads, meas

  def communicator(dev):
    """The C-level NCCL communicator of this rank, created on first use (None on a single GPU)."""
    if process_group is None or reducer.world == 1:
      return None
    if sc.comm is None:
      from .sharding import Communicator
      sc.comm = Communicator(process_group)
    return sc.comm

  def forward_backward(train_state, batch, *, rand_rank=None, reduce=True):
    """Draws, q_sample, forward of both branches, loss and backward, no optimiser; with reduce the gradient arena is
    mean-all-reduced over the process group.  Returns (arena, shadow, grads, loss_slot).  rand_rank overrides the rank
    that seeds the draws (bench.py's data-parallel check replays other ranks' shards on rank 0)."""
    arena, shadow, grads, _ = run_step(train_state, batch, rand_rank=rand_rank, reduce=reduce, optimise=False)
    return arena, shadow, grads, grads[layout.total:layout.total + 1]

  def update_fn(train_state, batch):
    arena, shadow, grads, meas = run_step(train_state, batch)
    model.set_shadow(arena, shadow)                     # the kernel refreshed the parameter shadow ...
    if "ema_params" in train_state:                     # ... but moved the EMA arena behind torch's back
      model.invalidate_shadow(train_state["ema_params"].arena)
    train_state["opt"]["count"] = int(train_state["opt"]["count"]) + 1
    rng = train_state["rng"].clone()
    rng[1] += 1
    train_state["rng"] = rng
    # measurements (train_ae.py:367-371): device scalars, read them only at log steps (:643-652)
    measurements = {"training_loss": meas[0], "l2_params": meas[1], "l2_updates": meas[2], "grad_norm": meas[3]}
    return train_state, measurements

  update_fn.grads = lambda: sc.grads
  update_fn.forward_backward = forward_backward
  update_fn.draw_step_randoms = draw_step_randoms
  update_fn.communicator = communicator
  return update_fn
