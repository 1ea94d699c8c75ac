"""`Model(...)` façade with the reference's `init` / `apply` surface (big_vision/models/ae.py:176-222;
call sites train_ae.py:106-112,325-346,384-483 — SURVEY.md App. F), backed by the CUDA engine.

Differences that a caller can see:
  * arrays are torch CUDA tensors instead of jax Arrays;
  * `rngs` values are integer seeds (or torch Generators); for exact reproduction of a reference run
    the draws themselves may be passed: rngs={"mae_noise": f32[n, L] uniforms, "cfg": bool[n] drop mask};
  * there is no CPU path: without the built extension or without a GPU this raises.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import lib
from .config import ModelConfig, make_model_config
from .params import ArenaLayout, ParamTree, arena_from_tree, init_arena, tree_from_arena


class ForwardResult:
  __slots__ = ("pred", "pre_logits", "mask_seq", "ids_restore", "ids_shuffle")


class ViTAE:
  """The auto-encoder module object (`_ViTAE`, ae.py:38-197)."""

  def __init__(self, cfg: ModelConfig):
    self.cfg = cfg
    self.layout = ArenaLayout(cfg)
    self._mcfg = lib.model_cfg_struct(cfg)
    self._offsets = self.layout.offsets_tensor()
    self._ws = {}       # (shape tuple, train) -> workspace tensor
    self.no_decay_list = cfg.no_decay_list

  # ---- reference surface -----------------------------------------------------------------
  def init(self, rngs, image=None, *, t=None, train=False, mask=0.0, device="cuda", nonzero_adaln=False):
    """module.init(rngs, image, t=, train=, mask=) (train_ae.py:106-112) -> {"params": tree}.
    Parameter shapes do not depend on the example inputs, so they are optional here."""
    seed = rngs["params"] if isinstance(rngs, dict) else rngs
    if isinstance(seed, torch.Tensor):
      seed = int(seed.reshape(-1)[0].item())
    arena = init_arena(self.layout, int(seed), device, nonzero_adaln=nonzero_adaln)
    return {"params": tree_from_arena(self.layout, arena)}

  def apply_raw(self, variables, image, *, t=None, y=None, cfg_scale=None):
    """Inference forward that leaves the classifier-free-guidance combine (ae.py:192-195) to the consumer: with
    cfg_scale the result is the doubled batch [2n,H,W,2C], conditional rows first (the DDIM kernel combines them)."""
    pred, _ = self.apply(variables, image, t=t, y=y, cfg_scale=cfg_scale, _combine=False)
    return pred

  def apply(self, variables, image, *, t=None, y=None, cfg_scale=None, mask=0.0, train=False, rngs=None, _combine=True):
    """ae.py:176-197.  Returns (pred [n,H,W,2C], {"mask": [n,H,W,1] | None, "pre_logits": [n,D]})."""
    cfg = self.cfg
    rngs = rngs or {}
    image = self._as_input(image)
    if cfg_scale is not None:  # ae.py:177-187
      assert y is not None, "y must be provided if cfg_scale is not None"
      assert cfg.num_classes is not None, "num_classes must be provided if cfg_scale is not None"
      assert not train, "cfg_scale is only used during inference"
      nh = image.shape[0]
      image = torch.cat([image, image], 0)
      t = torch.cat([t, t], 0)
      y = torch.cat([y.to(image.device), torch.full((nh,), cfg.num_classes, dtype=y.dtype, device=image.device)], 0)
    n = image.shape[0]
    dev = image.device
    tt = torch.zeros(n, dtype=torch.int32, device=dev) if t is None else t.reshape(-1).to(device=dev, dtype=torch.int32)
    labels = self._labels(y, n, train, rngs.get("cfg"), dev)
    masked = mask > 0.0
    keep = cfg.len_keep(mask) if masked else cfg.num_patches
    ids_shuffle = ids_restore = mask_seq = None
    if masked:
      noise = self._mask_noise(rngs.get("mae_noise"), n, dev)
      ids_shuffle, ids_restore, mask_seq = mask_argsort(noise, keep)
    arena = arena_from_tree(self.layout, variables["params"], dev)
    res = self.forward_arena(arena, self.shadow_of(arena), image=image, t=tt, labels=labels, n0=n, n1=0, keep0=keep,
                             keep1=cfg.num_patches, masked0=masked, masked1=False, ids_shuffle=ids_shuffle,
                             ids_restore=ids_restore, want_pred=True, train=False)
    pred = res.pred
    if cfg_scale is not None and _combine:  # ae.py:192-195
      un, co = pred[n // 2:], pred[:n // 2]
      pred = un + cfg_scale * (co - un)
    out = {"mask": None, "pre_logits": res.pre_logits}
    if masked:
      g, p = cfg.grid, cfg.patch
      m = mask_seq.reshape(n, g, g)
      out["mask"] = m.repeat_interleave(p, 1).repeat_interleave(p, 2)[..., None]  # ae.py:30-36
    return pred, out

  # ---- engine plumbing -------------------------------------------------------------------
  def _as_input(self, image):
    if not torch.cuda.is_available():
      raise lib.UmdError("the UMD hot path needs a CUDA device (no CPU fallback)")
    image = torch.as_tensor(image)
    if not image.is_cuda:
      image = image.cuda()
    cfg = self.cfg
    assert tuple(image.shape[1:]) == (cfg.img_size, cfg.img_size, cfg.channels), image.shape
    return image.to(torch.float32).contiguous()

  def _labels(self, y, n, train, cfg_rng, dev):
    cfg = self.cfg
    if cfg.num_classes is None:
      assert y is None, "num_classes must be provided if y is not None"  # ae.py:112
      return None
    if y is None:  # ae.py:107-110: the null class
      return torch.full((n,), cfg.num_classes, dtype=torch.int32, device=dev)
    y = y.reshape(-1).to(device=dev, dtype=torch.int32)
    if train:  # embeddings.py:43-45
      drop = self._label_drop(cfg_rng, n, dev)
      y = torch.where(drop, torch.full_like(y, cfg.num_classes), y)
    return y.contiguous()

  def _label_drop(self, cfg_rng, n, dev):
    if isinstance(cfg_rng, torch.Tensor) and cfg_rng.dtype == torch.bool:
      return cfg_rng.to(dev)
    g = _generator(cfg_rng, dev)
    return torch.rand(n, device=dev, generator=g) < self.cfg.cfg_dropout_rate

  def _mask_noise(self, rng, n, dev):
    L = self.cfg.num_patches
    if isinstance(rng, torch.Tensor) and rng.is_floating_point():
      assert tuple(rng.shape) == (n, L), rng.shape
      return rng.to(device=dev, dtype=torch.float32).contiguous()
    return torch.rand(n, L, device=dev, generator=_generator(rng, dev))

  def shadow_of(self, arena):
    """bf16 copy of the arena consumed by the GEMMs.  The shadow lives ON the arena tensor object (so it dies with it
    and a new arena that reuses the address can never hit it) together with the torch version it was cast at and a
    validity flag: writes through torch bump the version, writes through the C ABI (umd_adamw_step updates the
    parameter and EMA arenas through raw pointers) must be announced with set_shadow / invalidate_shadow."""
    ent = getattr(arena, "_umd_shadow", None)
    if ent is not None and ent[0] and ent[1] == arena._version and ent[2].numel() == arena.numel():
      return ent[2]
    sh = ent[2] if ent is not None and ent[2].numel() == arena.numel() and ent[2].device == arena.device else \
        torch.empty_like(arena, dtype=torch.bfloat16)
    lib.cast_bf16(arena, sh)
    arena._umd_shadow = (True, arena._version, sh)
    return sh

  def set_shadow(self, arena, shadow):
    """The caller's kernel has just rewritten `shadow` from `arena` (umd_adamw_step refreshes params_bf16 itself)."""
    arena._umd_shadow = (True, arena._version, shadow)

  def invalidate_shadow(self, arena):
    """`arena` was modified through the C ABI without refreshing its shadow (the EMA arena after every step)."""
    ent = getattr(arena, "_umd_shadow", None)
    if ent is not None:
      arena._umd_shadow = (False, ent[1], ent[2])

  def workspace(self, shape, train, device):
    key = (shape.n0, shape.n1, shape.keep0, shape.keep1, shape.masked0, shape.masked1, bool(train), str(device))
    ws = self._ws.get(key)
    if ws is None:
      nbytes = lib.workspace_bytes(self._mcfg, shape, train)
      ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
      self._ws[key] = ws
    return ws

  def step_shape(self, n0, n1, keep0, keep1, masked0, masked1):
    s = lib.StepShape()
    s.n0, s.n1, s.keep0, s.keep1, s.masked0, s.masked1 = n0, n1, keep0, keep1, int(masked0), int(masked1)
    return s

  def forward_arena(self, arena, shadow, *, image, t, labels, n0, n1, keep0, keep1, masked0, masked1, ids_shuffle,
                    ids_restore, want_pred, train, x0=None, noise=None, loss_out=None):
    """One umd_forward call.  Returns ForwardResult; with train=True the workspace keeps the activations
    for backward_arena."""
    cfg = self.cfg
    n = n0 + n1
    dev = image.device
    shape = self.step_shape(n0, n1, keep0, keep1, masked0, masked1)
    ws = self.workspace(shape, train, dev)
    res = ForwardResult()
    res.pred = torch.empty(n, cfg.img_size, cfg.img_size, 2 * cfg.channels, device=dev) if want_pred else None
    res.pre_logits = torch.empty(n, cfg.width, device=dev)
    res.ids_shuffle, res.ids_restore = ids_shuffle, ids_restore
    io = lib.IO()
    io.image, io.t, io.labels = lib.ptr(image), lib.ptr(t), lib.ptr(labels)
    io.ids_shuffle, io.ids_restore = lib.ptr(ids_shuffle), lib.ptr(ids_restore)
    io.x0, io.noise = lib.ptr(x0), lib.ptr(noise)
    io.pred, io.pre_logits, io.loss = lib.ptr(res.pred), lib.ptr(res.pre_logits), lib.ptr(loss_out)
    L = lib.load()
    lib.check(L.umd_forward(C.byref(self._mcfg), C.byref(shape), self._offsets, lib.ptr(arena), lib.ptr(shadow),
                            C.byref(io), lib.ptr(ws), C.c_size_t(ws.numel()), C.c_int(int(train)), lib.current_stream()),
              "umd_forward")
    self._last = (shape, io, ws, (image, t, labels, ids_shuffle, ids_restore, x0, noise, res, loss_out))
    return res

  def backward_arena(self, arena, shadow, grads, bucket_cb=None):
    """umd_backward for the preceding forward_arena(train=True)."""
    shape, io, ws, _keep = self._last
    L = lib.load()
    err = []

    def _cb(_user, k):
      try:
        if bucket_cb is not None:
          bucket_cb(int(k))
      except BaseException as e:  # never let an exception cross the C frame
        err.append(e)

    cb = lib.BUCKET_CB(_cb)
    rc = L.umd_backward(C.byref(self._mcfg), C.byref(shape), self._offsets, lib.ptr(arena), lib.ptr(shadow),
                        lib.ptr(grads), C.byref(io), lib.ptr(ws), C.c_size_t(ws.numel()), cb, None, lib.current_stream())
    if err:
      raise err[0]
    lib.check(rc, "umd_backward")


def _generator(seed, dev):
  if isinstance(seed, torch.Generator):
    return seed
  if isinstance(seed, torch.Tensor):
    seed = int(seed.reshape(-1)[0].item())
  g = torch.Generator(device=dev)
  g.manual_seed(int(seed) if seed is not None else 0)
  return g


def mask_argsort(noise, keep):
  """random_masking's index work (ae.py:14-16,25-27) on the GPU: stable argsort, inverse permutation and
  the 0/1 sequence mask (0 = keep, 1 = remove).  Bit-exact with jnp.argsort (stable) by construction."""
  n, Lp = noise.shape
  dev = noise.device
  ids_shuffle = torch.empty(n, Lp, dtype=torch.int32, device=dev)
  ids_restore = torch.empty(n, Lp, dtype=torch.int32, device=dev)
  mask = torch.empty(n, Lp, dtype=torch.float32, device=dev)
  L = lib.load()
  lib.check(L.umd_mask_argsort(lib.ptr(noise), C.c_int(n), C.c_int(Lp), C.c_int(keep), lib.ptr(ids_shuffle),
                               lib.ptr(ids_restore), lib.ptr(mask), lib.current_stream()), "umd_mask_argsort")
  return ids_shuffle, ids_restore, mask


def Model(*, variant=None, **kw):  # pylint: disable=invalid-name
  """Factory with the reference's signature (ae.py:220-222)."""
  return ViTAE(make_model_config(variant=variant, **kw))
