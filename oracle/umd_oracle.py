"""CPU oracle: a plain PyTorch (fp32, fp64 switchable) restatement of the reference's UMD
auto-encoder training step.  TEST INFRASTRUCTURE ONLY — imported by tests/, by
__graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs; the
product path (small-vision_b200/) never imports it.

PARITY STATUS: pinned to the reference's own source for the model, diffusion and loss;
the optimiser restates optax's published algorithm (library release unpinned, see below).  The reference (philippe-eecs/small-vision) is JAX/Flax/Optax
code; none of the three can be installed in this image (no wheels, no network) and the
repository ships no golden vectors or tests for this path.
  * Model forward, masking, conditioning, classifier-free guidance, q_sample, the DDIM
    step and loop, and the training loss: tests/golden/reference_golden.pt holds the
    outputs of the reference's UNMODIFIED files (models/ae.py, models/vit.py,
    models/embeddings.py, gaussian_diffusion.py, and the loss_fn closure of
    trainers/train_ae.py:323-361) executed over a numpy-float64 stand-in for the jax /
    flax.linen names they use (tests/golden/refshim/, generator
    tests/golden/make_reference_golden.py).  This oracle run in float64 agrees with those
    numbers to 1e-11, and its autograd gradient reproduces the slopes obtained by
    differencing the reference's loss along seeded parameter directions
    (tests/test_reference_golden_cpu.py).  What that does NOT pin is the arithmetic of the
    library layers the stand-in itself restates (Dense, LayerNorm eps 1e-6, dot-product
    attention, Conv / ConvTranspose orientation, gelu-tanh, silu); those are checked
    against torch's independent implementations in tests/test_oracle_invariants.py.
  * Few-shot probe, sharding declarations, checkpoint naming, value_range and the recipe
    wiring (step counts, schedule / AdamW arguments, decay mask) are pinned the same way
    (tests/golden/refshim/README.md lists every fixture and the reference code behind it).
  * Optimiser (optax: clip_by_global_norm, adamw with bf16 mu, masked decay,
    warmup_cosine_decay_schedule; train_ae.py:135-151): optax is absent from /root/reference
    and from this image, and requirements.txt:5 pins no version.  optimizer_update restates
    optax/_src/clipping.py::clip_by_global_norm (select(norm < c, g, g / norm * c)),
    optax/_src/transform.py::{scale_by_adam (update_moment, bias_correction, the mu_dtype cast
    AFTER the update is formed), add_decayed_weights, scale_by_schedule (step size from the
    pre-increment count)}, optax/_src/schedule.py::warmup_cosine_decay_schedule
    (join_schedules(linear_schedule, cosine_decay_schedule)) and optax/_src/update.py::
    {apply_updates, incremental_update} as published in optax 0.1.7 - 0.2.3 (their arithmetic
    is the same across those releases).  tests/test_optimizer_pins_cpu.py checks it against a
    second, transformation-by-transformation restatement in numpy (bf16 rounding by bit
    manipulation), closed forms, torch.optim.AdamW and the reference's own decay mask —
    PARITY UNPINNED only with respect to the library release itself.
  * The same code runs on the GPU (fp32, TF32 off) as the secondary oracle for full-size
    shapes (tests/test_fullsize_gpu.py): every tensor it creates follows its inputs' device.

Each function cites the reference lines it follows (paths relative to /root/reference).
Every random quantity the reference draws inside the step (mask noise, t, noise, label
dropout) is an explicit argument here.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

# ------------------------------------------------------------------------------------------
# gaussian_diffusion.py
# ------------------------------------------------------------------------------------------


def beta_schedule(name: str, T: int) -> np.ndarray:
  """big_vision/gaussian_diffusion.py:10-30 — float64 betas, 'linear' or 'cosine'."""
  if name == "linear":
    s = 1000.0 / T
    return np.linspace(s * 1e-4, s * 2e-2, T, dtype=np.float64)
  if name == "cosine":
    def abar(u):
      return math.cos((u + 0.008) / 1.008 * math.pi / 2) ** 2
    out = np.empty(T, dtype=np.float64)
    for i in range(T):
      out[i] = min(1.0 - abar((i + 1) / T) / abar(i / T), 0.999)
    return out
  raise NotImplementedError(name)


def gaussian_diffusion_tables(beta_type: str = "cosine", training_steps: int = 1000) -> dict:
  """big_vision/gaussian_diffusion.py:32-67 — the 13 float64 schedule tables."""
  b = beta_schedule(beta_type, training_steps).astype(np.float64)
  a = 1.0 - b
  ac = np.cumprod(a)
  ac_prev = np.concatenate([[1.0], ac[:-1]])
  ac_next = np.concatenate([ac[1:], [0.0]])
  post_var = b * (1.0 - ac) / (1.0 - ac[-1])
  return {
      "betas": b,
      "alphas": a,
      "alphas_cumprod": ac,
      "alphas_cumprod_prev": ac_prev,
      "alphas_cumprod_next": ac_next,
      "sqrt_alphas_cumprod": np.sqrt(ac),
      "sqrt_one_minus_alphas_cumprod": np.sqrt(1.0 - ac),
      "sqrt_recip_alphas_cumprod": np.sqrt(1.0 / ac),
      "sqrt_recipm1_alphas_cumprod": np.sqrt(1.0 / ac - 1.0),
      "posterior_variance": post_var,
      "posterior_log_variance_clipped": np.log(np.concatenate([post_var[1:2], post_var[1:]])),
      "posterior_mean_coef1": b * np.sqrt(ac_prev) / (1.0 - ac),
      "posterior_mean_coef2": (1.0 - ac_prev) * np.sqrt(a) / (1.0 - ac),
  }


def q_sample(gd: dict, x_start: torch.Tensor, t: torch.Tensor, noise: torch.Tensor) -> torch.Tensor:
  """gaussian_diffusion.py:85-98,286-289.  Tables are float32 on device in the reference
  (train_ae.py:183-185 with x64 off), so they are cast to x_start's dtype before the gather."""
  dt = x_start.dtype
  dv = x_start.device   # the oracle also runs in fp32 on the GPU as the full-size secondary oracle
  ca = torch.as_tensor(np.asarray(gd["sqrt_alphas_cumprod"])).to(device=dv, dtype=dt)[t.reshape(-1).long().to(dv)]
  cb = torch.as_tensor(np.asarray(gd["sqrt_one_minus_alphas_cumprod"])).to(device=dv, dtype=dt)[t.reshape(-1).long().to(dv)]
  shp = (-1,) + (1,) * (x_start.dim() - 1)
  return ca.reshape(shp) * x_start + cb.reshape(shp) * noise


def _extract(gd, key, t, x):
  """gaussian_diffusion.py:286-289 _extract_into_tensor (tables float32 on device, train_ae.py:183-185)."""
  arr = torch.as_tensor(np.asarray(gd[key])).to(device=x.device, dtype=x.dtype)
  return arr[t.reshape(-1).long()].reshape((-1,) + (1,) * (x.dim() - 1))


def predict_xstart_from_eps(gd, x_t, t, eps):
  """gaussian_diffusion.py:122-127."""
  return _extract(gd, "sqrt_recip_alphas_cumprod", t, x_t) * x_t - _extract(gd, "sqrt_recipm1_alphas_cumprod", t, x_t) * eps


def predict_eps_from_xstart(gd, x_t, t, pred_xstart):
  """gaussian_diffusion.py:129-132."""
  return (_extract(gd, "sqrt_recip_alphas_cumprod", t, x_t) * x_t - pred_xstart) / _extract(gd, "sqrt_recipm1_alphas_cumprod", t, x_t)


def q_posterior_mean(gd, x_start, x_t, t):
  """gaussian_diffusion.py:100-120 (mean only)."""
  return _extract(gd, "posterior_mean_coef1", t, x_t) * x_start + _extract(gd, "posterior_mean_coef2", t, x_t) * x_t


def ddim_sample(gd, p_apply, x, t, t_next, noise, clip_denoised=False, model_kwargs=None, eta=1.0):
  """gaussian_diffusion.py:166-211 on top of p_mean_variance (:134-164).  `noise` replaces the N(0,1) draw of
  :196-197; p_apply(x_t=, t=, **model_kwargs) returns the model's eps."""
  model_output = p_apply(x_t=x, t=t, **(model_kwargs or {}))
  pred_xstart = predict_xstart_from_eps(gd, x, t, model_output)
  if clip_denoised:
    pred_xstart = pred_xstart.clip(-1, 1)
  eps = predict_eps_from_xstart(gd, x, t, pred_xstart)
  alpha_bar = _extract(gd, "alphas_cumprod", t, x)
  alpha_bar_prev = _extract(gd, "alphas_cumprod", t_next, x) if t_next is not None else _extract(gd, "alphas_cumprod_prev", t, x)
  sigma = eta * torch.sqrt((1 - alpha_bar_prev) / (1 - alpha_bar)) * torch.sqrt(1 - alpha_bar / alpha_bar_prev)
  mean_pred = pred_xstart * torch.sqrt(alpha_bar_prev) + torch.sqrt(1 - alpha_bar_prev - sigma ** 2) * eps
  nonzero = (t.reshape((-1,) + (1,) * (x.dim() - 1)) > 0).to(x.dtype)
  return {"sample": mean_pred + nonzero * sigma * noise, "pred_xstart": pred_xstart}


def ddim_timesteps(num_train_steps, sampling_steps):
  """gaussian_diffusion.py:236-237."""
  ts = list(range(num_train_steps - 1, 0, -num_train_steps // sampling_steps))
  return ts + [0]


def ddim_sample_loop(gd, apply_fn, noises, ys=None, clip_denoised=False, sampling_steps=250, cfg_scale=None, eta=1.0):
  """gaussian_diffusion.py:213-280.  noises: the draws in reference order (initial image, one per scan step, one for
  the final t = 0 call).  Returns the final call's pred_xstart (:271-275)."""
  model_kwargs = dict(y=ys, cfg_scale=cfg_scale)
  img = noises[0]
  n = img.shape[0]
  ts = ddim_timesteps(len(np.asarray(gd["betas"])), sampling_steps)
  for k in range(sampling_steps):
    t_curr = torch.full((n, 1), ts[k], dtype=torch.int64)
    t_next = torch.full((n, 1), ts[k + 1], dtype=torch.int64)
    img = ddim_sample(gd, apply_fn, img, t_curr, t_next, noises[1 + k], clip_denoised, model_kwargs, eta)["sample"]
  final = ddim_sample(gd, apply_fn, img, torch.zeros((n, 1), dtype=torch.int64), None, noises[1 + sampling_steps],
                      clip_denoised, model_kwargs, eta)
  return final["pred_xstart"]


def make_apply_fn(params, cfg, gd=None, *, eps_pred=True, dtype=torch.float32):
  """train_ae.py:472-483 create_apply_fn over model_apply: the model sees t + 1; the eps head is the last C channels
  (eps_pred=False converts the x0 head instead and needs the tables)."""
  C = cfg["channels"]

  def apply_fn(*, x_t, t, y=None, cfg_scale=None):
    pred, _ = model_apply(params, cfg, x_t, t=t.reshape(-1) + 1, y=y, cfg_scale=cfg_scale, dtype=dtype)
    if eps_pred:
      return pred[..., C:]
    return predict_eps_from_xstart(gd, x_t, t, pred[..., :C])
  return apply_fn


# ------------------------------------------------------------------------------------------
# models/ae.py helpers
# ------------------------------------------------------------------------------------------

VARIANTS = {  # ae.py:205-215
    "S": dict(width=384, depth=12, dec_depth=4, num_heads=6),
    "B": dict(width=768, depth=12, dec_depth=4, num_heads=12),
    "L": dict(width=1024, depth=24, dec_depth=8, num_heads=16),
}


def model_config(variant=None, **kw) -> dict:
  """ae.py:38-55,200-222 — defaults of _ViTAE merged with the decoded variant and kwargs."""
  cfg = dict(num_classes=None, channels=3, img_size=64, patch_size=(4, 4), width=768, depth=12, dec_depth=4,
             mlp_dim=None, num_heads=12, adaln=False, cfg_dropout_rate=0.1, num_cls=4,
             no_decay_list=("cls", "image_mask_embedding", "bias"))
  if variant is not None:
    v = variant
    if "/" in variant:
      v, p = variant.split("/")
      cfg["patch_size"] = (int(p), int(p))
    cfg.update(VARIANTS[v])
  for k, val in kw.items():
    if k in ("scan", "remat_policy", "dtype_mm", "dropout"):
      continue  # performance / unused knobs
    cfg[k] = val
  return cfg


def len_keep_of(L: int, mask_ratio: float) -> int:
  """ae.py:11 — Python double arithmetic, truncation."""
  return int(L * (1 - mask_ratio))


def random_masking(x, mask_ratio, noise):
  """ae.py:9-28 with the uniform draw supplied.  argsort is stable (jnp.argsort default)."""
  N, L, _ = x.shape
  keep = len_keep_of(L, mask_ratio)
  ids_shuffle = torch.argsort(noise, dim=1, stable=True)
  ids_restore = torch.argsort(ids_shuffle, dim=1, stable=True)
  ids_keep = ids_shuffle[:, :keep]
  x_masked = torch.gather(x, 1, ids_keep[:, :, None].expand(-1, -1, x.shape[2]))
  mask = torch.ones(N, L, dtype=x.dtype, device=x.device)
  mask[:, :keep] = 0
  mask = torch.gather(mask, 1, ids_restore)
  return x_masked, mask, ids_restore


def sequence_mask_to_image_mask(seq_mask, patch, img_size):
  """ae.py:30-36."""
  g = img_size // patch
  m = seq_mask.reshape(-1, g, g)
  m = m.repeat_interleave(patch, dim=1).repeat_interleave(patch, dim=2)
  return m[..., None]


# ------------------------------------------------------------------------------------------
# flax.linen layer semantics
# ------------------------------------------------------------------------------------------


def layer_norm(x, scale, bias, eps=1e-6):
  """nn.LayerNorm(): fast-variance form, eps 1e-6 (vit.py:78,96,163)."""
  mean = x.mean(-1, keepdim=True)
  mean2 = (x * x).mean(-1, keepdim=True)
  var = torch.clamp(mean2 - mean * mean, min=0.0)
  return (x - mean) * torch.rsqrt(var + eps) * scale + bias


def gelu_tanh(x):
  """nn.gelu default approximate=True (vit.py:55)."""
  return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * x ** 3)))


def dense(x, p):
  return x @ p["kernel"] + p["bias"]


def time_embedding(t, width, dtype):
  """embeddings.py:13-31."""
  half = width // 2
  step = math.log(10000) / (half - 1)
  freq = torch.exp(torch.arange(half, dtype=dtype, device=t.device) * -step)
  e = t.to(dtype) * freq
  return torch.cat([torch.sin(e), torch.cos(e)], dim=-1)


def embedding_trunk(x, p):
  """embeddings.py:50-59."""
  return dense(F.silu(dense(x, p["Dense_0"])), p["Dense_1"])


def attention(y, p, num_heads):
  """nn.MultiHeadDotProductAttention(y, y) (vit.py:82-87): q scaled by 1/sqrt(Dh) before QK^T."""
  q = torch.einsum("bsd,dhk->bshk", y, p["query"]["kernel"]) + p["query"]["bias"]
  k = torch.einsum("bsd,dhk->bshk", y, p["key"]["kernel"]) + p["key"]["bias"]
  v = torch.einsum("bsd,dhk->bshk", y, p["value"]["kernel"]) + p["value"]["bias"]
  q = q / math.sqrt(q.shape[-1])
  logits = torch.einsum("bqhd,bkhd->bhqk", q, k)
  w = torch.softmax(logits, dim=-1)
  o = torch.einsum("bhqk,bkhd->bqhd", w, v)
  return torch.einsum("bqhd,hdo->bqo", o, p["out"]["kernel"]) + p["out"]["bias"]


def block(x, cond, p, l, *, adaln, num_heads):
  """Encoder1DBlock (vit.py:60-113); `p` holds scanned params with a leading [depth] axis."""
  def at(tree):
    return {k: (at(v) if isinstance(v, dict) else v[l]) for k, v in tree.items()}
  p = at(p)
  if adaln:
    ada = dense(cond, p["Dense_0"])
    sh0, sc0, g0, sh1, sc1, g1 = torch.chunk(ada, 6, dim=-1)
  else:
    x = torch.cat([cond[:, None, :], x], dim=1)
  y = layer_norm(x, p["LayerNorm_0"]["scale"], p["LayerNorm_0"]["bias"])
  if adaln:
    y = y * (1 + sc0[:, None, :]) + sh0[:, None, :]
  y = attention(y, p["MultiHeadDotProductAttention_0"], num_heads)
  if adaln:
    y = g0[:, None, :] * y
  x = x + y
  y = layer_norm(x, p["LayerNorm_1"]["scale"], p["LayerNorm_1"]["bias"])
  if adaln:
    y = y * (1 + sc1[:, None, :]) + sh1[:, None, :]
  y = dense(gelu_tanh(dense(y, p["MlpBlock_0"]["Dense_0"])), p["MlpBlock_0"]["Dense_1"])
  if adaln:
    y = g1[:, None, :] * y
  x = x + y
  if not adaln:
    x = x[:, 1:]
  return x


def scanned_key(tree):
  ks = [k for k in tree if k != "encoder_norm"]
  assert len(ks) == 1, ks  # the auto-generated scan/remat name is flax-version dependent
  return ks[0]


def encoder(x, cond, p, depth, *, adaln, num_heads):
  """Encoder (vit.py:115-163): scanned blocks then LayerNorm 'encoder_norm'."""
  bp = p[scanned_key(p)]
  for l in range(depth):
    x = block(x, cond, bp, l, adaln=adaln, num_heads=num_heads)
  return layer_norm(x, p["encoder_norm"]["scale"], p["encoder_norm"]["bias"])


def patchify(image, patch):
  n, H, W, C = image.shape
  h, w = H // patch, W // patch
  return image.reshape(n, h, patch, w, patch, C).permute(0, 1, 3, 2, 4, 5)  # [n,h,w,a,b,c]


def conv_transpose_unpatchify(x, kernel, bias, *, flip=True):
  """nn.ConvTranspose(2C, (p,p), strides=(p,p), 'VALID') with transpose_kernel=False
  (ae.py:95-97): pred[n,ip+a,jp+b,o] = b[o] + sum_c x[n,i,j,c] K[p-1-a,p-1-b,c,o]."""
  n, h, w, _ = x.shape
  p = kernel.shape[0]
  K = kernel.flip(0, 1) if flip else kernel
  y = torch.einsum("nijc,abco->niajbo", x, K)
  return y.reshape(n, h * p, w * p, kernel.shape[-1]) + bias


def model_apply(params, cfg, image, *, t=None, y=None, cfg_scale=None, mask=0.0, train=False, mask_noise=None,
                label_drop=None, dtype=torch.float32, flip_final_conv=True):
  """_ViTAE.__call__ (ae.py:176-197) = embed (:99-125) -> encode (:127-145) -> decode (:147-174).

  mask_noise: f32[n, L] uniforms replacing jax.random.uniform (ae.py:14) — required when mask > 0.
  label_drop: bool[n] replacing the bernoulli draw of LabelEmbedder (embeddings.py:43-45); only
              consulted when train=True.
  Returns (pred [n,H,W,2C], {"mask", "pre_logits", "ids_restore"}).
  """
  D, ps, C = cfg["width"], cfg["patch_size"][0], cfg["channels"]
  adaln, nc = cfg["adaln"], cfg["num_classes"]
  cast = lambda tree: {k: (cast(v) if isinstance(v, dict) else v.to(dtype)) for k, v in tree.items()}
  params = cast(params)
  image = image.to(dtype)
  if cfg_scale is not None:
    assert y is not None and nc is not None and not train
    nh = image.shape[0]
    image = torch.cat([image, image], 0)
    t = torch.cat([t, t], 0)
    y = torch.cat([y, torch.full((nh,), nc, dtype=y.dtype, device=y.device)], 0)
  n = image.shape[0]
  # ---- embed
  pt = patchify(image, ps)
  x = torch.einsum("nijabc,abcd->nijd", pt, params["embedding"]["kernel"]) + params["embedding"]["bias"]
  h, w = x.shape[1], x.shape[2]
  L = h * w
  x = x.reshape(n, L, D)
  if t is None:
    t = torch.zeros(n, 1, dtype=torch.int32, device=image.device)
  if y is None and nc is not None:
    y = torch.full((n,), nc, dtype=torch.long, device=image.device)
  if y is not None:
    assert nc is not None, "num_classes must be provided if y is not None"
    yy = y.long()
    if train and label_drop is not None:
      yy = torch.where(label_drop.bool(), torch.full_like(yy, nc), yy)
    y_cond = embedding_trunk(params["label_emb"]["embedding"]["embedding"][yy], params["label_trunk"])
  else:
    y_cond = torch.zeros(n, D, dtype=dtype, device=image.device)
  time_cond = embedding_trunk(time_embedding(t.reshape(n, 1), D, dtype), params["time_trunk"])
  cond = F.silu(time_cond + y_cond) if adaln else time_cond + y_cond
  # ---- encode
  out = {}
  x = x + params["pos_embedding"]
  ids_restore = None
  if mask > 0.0:
    assert mask_noise is not None
    x, seq_mask, ids_restore = random_masking(x, mask, mask_noise)
    out["mask"] = sequence_mask_to_image_mask(seq_mask, ps, cfg["img_size"])
  else:
    out["mask"] = None
  x = torch.cat([params["cls"].expand(n, -1, -1), x], dim=1)
  x = encoder(x, cond, params["Encoder"], cfg["depth"], adaln=adaln, num_heads=cfg["num_heads"])
  rep = x[:, :cfg["num_cls"]].mean(dim=1)
  enc = x[:, cfg["num_cls"]:]
  out["pre_logits"] = rep
  out["ids_restore"] = ids_restore
  # ---- decode
  if ids_restore is not None:
    n_masked = L - int(L * (1.0 - mask))
    enc = torch.cat([enc, params["image_mask_embedding"].expand(n, n_masked, -1)], dim=1)
    enc = torch.gather(enc, 1, ids_restore[:, :, None].expand(-1, -1, D))
  xd = enc + params["dec_pos_embedding"]
  xd = torch.cat([rep[:, None, :], xd], dim=1)
  xd = encoder(xd, cond, params["Decoder"], cfg["dec_depth"], adaln=adaln, num_heads=cfg["num_heads"])
  xd = xd[:, 1:, :]
  if adaln:
    fm = dense(cond, params["final_modulation"])
    shift, scale = torch.chunk(fm[:, None, :], 2, dim=-1)
    xd = xd * (1 + scale) + shift
  xd = xd.reshape(n, h, w, D)
  pred = conv_transpose_unpatchify(xd, params["final_conv"]["kernel"], params["final_conv"]["bias"],
                                   flip=flip_final_conv)
  if cfg_scale is not None:
    un, co = pred[n // 2:], pred[:n // 2]
    pred = un + cfg_scale * (co - un)
  return pred, out


# ------------------------------------------------------------------------------------------
# trainers/train_ae.py: loss, optimiser, update_fn
# ------------------------------------------------------------------------------------------


def loss_fn(params, cfg, tc, x0_noise, x_t_noise, x0_clean, t, noise, labels, rand, dtype=torch.float32):
  """train_ae.py:323-361.  `tc` carries mask_ratio, mask_ratio_no_noise; `rand` the supplied draws:
  mask_noise_noise [n_noise, L], mask_noise_clean [n_clean, L], label_drop_noise [n_noise] (optional)."""
  C = cfg["channels"]
  n_noise, n_clean = x0_noise.shape[0], x0_clean.shape[0]
  B = n_noise + n_clean
  aux = {}
  if n_clean > 0:
    pred, out = model_apply(params, cfg, x0_clean, t=torch.zeros(n_clean, 1, dtype=torch.int32, device=x0_clean.device), train=True,
                            mask=tc["mask_ratio_no_noise"], mask_noise=rand.get("mask_noise_clean"), dtype=dtype)
    m = out["mask"]
    se = (pred[..., :C] - x0_clean.to(dtype)) ** 2
    mae_loss = (se * m).mean() / m.mean()
    aux["pred_clean"], aux["out_clean"] = pred, out
  else:
    mae_loss = 0.0
  if n_noise > 0:
    pred, out = model_apply(params, cfg, x_t_noise, t=t + 1, y=labels, train=True, mask=tc["mask_ratio"],
                            mask_noise=rand.get("mask_noise_noise"), label_drop=rand.get("label_drop_noise"),
                            dtype=dtype)
    x0_se = (pred[..., :C] - x0_noise.to(dtype)) ** 2
    eps_se = (pred[..., C:] - noise.to(dtype)) ** 2
    m = out["mask"]
    if m is not None:
      eps_loss = (eps_se * (1 - m)).mean() / (1 - m).mean()
      x0_loss = (x0_se * m).mean() / m.mean()
      dit_loss = (eps_loss + x0_loss) / 2
    else:
      dit_loss = (eps_se.mean() + x0_se.mean()) / 2
    aux["pred_noise"], aux["out_noise"] = pred, out
  else:
    dit_loss = 0.0
  loss = dit_loss * (1 - n_clean / B) + mae_loss * (n_clean / B)
  return loss, aux


def flatten_tree(tree, prefix=()):
  out = {}
  for k, v in tree.items():
    if isinstance(v, dict):
      out.update(flatten_tree(v, prefix + (k,)))
    else:
      out[prefix + (k,)] = v
  return out


def unflatten_tree(flat):
  out = {}
  for path, v in flat.items():
    d = out
    for k in path[:-1]:
      d = d.setdefault(k, {})
    d[path[-1]] = v
  return out


def weight_decay_mask(params, no_decay_list=("cls", "image_mask_embedding", "bias")):
  """train_ae.py:125-134: decay a leaf iff none of its path components is in no_decay_list."""
  return {path: all(k not in no_decay_list for k in path) for path in flatten_tree(params)}


def warmup_cosine_lr(count, *, peak, warmup_steps, decay_steps, init_value=0.0, end_value=0.0):
  """optax.warmup_cosine_decay_schedule (train_ae.py:135-138): linear warm-up joined at
  `warmup_steps` with a cosine decay over decay_steps - warmup_steps."""
  if count < warmup_steps:
    c = min(max(count, 0), warmup_steps)
    return init_value + (peak - init_value) * (c / warmup_steps)
  dsteps = decay_steps - warmup_steps
  c = min(count - warmup_steps, dsteps)
  cos = 0.5 * (1 + math.cos(math.pi * c / dsteps))
  alpha = end_value / peak if peak != 0 else 0.0
  return peak * ((1 - alpha) * cos + alpha)


def init_opt_state(params):
  """optax.adamw init with mu_dtype=bfloat16 (train_ae.py:140-146)."""
  flat = flatten_tree(params)
  return {"count": 0,
          "mu": {k: torch.zeros_like(v, dtype=torch.bfloat16) for k, v in flat.items()},
          "nu": {k: torch.zeros_like(v, dtype=torch.float32) for k, v in flat.items()}}


def optimizer_update(grads, opt, params, hp):
  """optax.chain(clip_by_global_norm(c), adamw(lr, b1, b2, eps=1e-8, wd, mask, mu_dtype=bf16))
  followed by apply_updates (train_ae.py:148-151,365-366).
  mu is accumulated in fp32 from the bf16-stored state, used un-rounded for this step's update
  and stored back as bf16 (optax casts after computing the update)."""
  fg, fp = flatten_tree(grads), flatten_tree(params)
  gnorm = math.sqrt(sum(float((g.double() ** 2).sum()) for g in fg.values()))
  c = hp["clip_norm"]
  count = opt["count"]
  count_inc = count + 1
  lr = warmup_cosine_lr(count, peak=hp["peak_lr"], warmup_steps=hp["warmup_steps"], decay_steps=hp["total_steps"])
  b1, b2, eps, wd = hp["b1"], hp["b2"], 1e-8, hp["wd"]
  mask = weight_decay_mask(params, hp.get("no_decay_list", ("cls", "image_mask_embedding", "bias")))
  new_p, new_mu, new_nu, upd = {}, {}, {}, {}
  for k, g in fg.items():
    g = g.float()
    if not gnorm < c:
      g = (g / gnorm) * c
    mu = (1 - b1) * g + b1 * opt["mu"][k].float()
    nu = (1 - b2) * (g * g) + b2 * opt["nu"][k]
    mu_hat = mu / (1 - b1 ** count_inc)
    nu_hat = nu / (1 - b2 ** count_inc)
    u = mu_hat / (torch.sqrt(nu_hat) + eps)
    if mask[k]:
      u = u + wd * fp[k]
    u = -lr * u
    upd[k] = u
    new_p[k] = fp[k] + u
    new_mu[k] = mu.to(torch.bfloat16)
    new_nu[k] = nu
  return unflatten_tree(new_p), {"count": count_inc, "mu": new_mu, "nu": new_nu}, upd, gnorm


def update_step(state, batch, cfg, tc, hp, rand, dtype=torch.float32):
  """update_fn (train_ae.py:287-382) with every RNG draw supplied in `rand`:
  t int32[n_noise,1], noise f32[n_noise,H,W,C], mask_noise_noise, mask_noise_clean, label_drop_noise."""
  images = batch["image"]
  B = images.shape[0]
  n_clean = int(B * tc["no_noise_prob"])
  n_noise = B - n_clean
  x0_noise, x0_clean = images[:n_noise], images[n_noise:]
  labels = batch["label"][:n_noise] if tc.get("use_labels", False) else None
  t, noise = rand["t"], rand["noise"]
  x_t = q_sample(state["gd"], x0_noise.float(), t, noise.float())
  flat = flatten_tree(state["params"])
  leaves = {k: v.detach().clone().requires_grad_(True) for k, v in flat.items()}
  loss, aux = loss_fn(unflatten_tree(leaves), cfg, tc, x0_noise, x_t, x0_clean, t, noise, labels, rand, dtype=dtype)
  loss.backward()
  grads = unflatten_tree({k: v.grad.float() if v.grad is not None else torch.zeros_like(v) for k, v in leaves.items()})
  new_params, new_opt, upd, gnorm = optimizer_update(grads, state["opt"], state["params"], hp)
  meas = {"training_loss": float(loss.detach()),
          "l2_params": math.sqrt(sum(float((p.double() ** 2).sum()) for p in flatten_tree(new_params).values())),
          "l2_updates": math.sqrt(sum(float((u.double() ** 2).sum()) for u in upd.values())),
          "grad_norm": gnorm}
  new_state = {"params": new_params, "opt": new_opt, "gd": state["gd"]}
  if "ema_params" in state:
    d = hp["ema_decay"]
    fe = flatten_tree(state["ema_params"])
    fn = flatten_tree(new_params)
    new_state["ema_params"] = unflatten_tree({k: d * fn[k] + (1.0 - d) * fe[k] for k in fe})
  return new_state, meas, {"grads": grads, "aux": aux, "x_t": x_t}


# ------------------------------------------------------------------------------------------
# evaluators/fewshot_lsr.py: few-shot ridge probe on pre_logits
# ------------------------------------------------------------------------------------------

FEWSHOT_BIAS_CONSTANT = 100.0  # fewshot_lsr.py:31


def fewshot_precompute_cache(x, y, num_classes):
  """fewshot_lsr.py:43-97 (_precompute_cache): whiten, append the bias feature, one-hot targets in {-1, 1}, and the
  eigendecomposition of x^T x (N >= D) or x x^T (D > N)."""
  mean = x.mean(dim=0, keepdim=True)
  std = x.std(dim=0, keepdim=True, unbiased=False) + 1e-5
  x = (x - mean) / std
  x = torch.cat([x, torch.full((x.shape[0], 1), FEWSHOT_BIAS_CONSTANT, dtype=x.dtype)], dim=1)
  yy = 2.0 * F.one_hot(y.long(), num_classes).to(x.dtype) - 1.0
  n, dim = x.shape
  if n >= dim:
    eigs, q = torch.linalg.eigh(x.T @ x)
    rhs, lhs = q.T @ (x.T @ yy), q
  else:
    eigs, q = torch.linalg.eigh(x @ x.T)
    rhs, lhs = q.T @ yy, x.T @ q
  return {"eigs": eigs, "rhs": rhs, "lhs": lhs, "mean": mean, "std": std}


def fewshot_weights(cache, l2_reg):
  """fewshot_lsr.py:103-108."""
  scaling = (1.0 / (cache["eigs"] + l2_reg)).reshape(1, -1)
  return (cache["lhs"] * scaling) @ cache["rhs"]


def fewshot_acc(cache, x_test, y_test, l2_reg):
  """fewshot_lsr.py:94-112 (_eig_fewshot_acc_fn).  Returns (accuracy, preds, scores)."""
  x_test = (x_test - cache["mean"]) / cache["std"]
  x_test = torch.cat([x_test, torch.full((x_test.shape[0], 1), FEWSHOT_BIAS_CONSTANT, dtype=x_test.dtype)], dim=1)
  scores = x_test @ fewshot_weights(cache, l2_reg)
  preds = scores.argmax(dim=1)
  return float((preds == y_test.long()).double().mean()), preds, scores


# ------------------------------------------------------------------------------------------
# Input stage after JPEG decoding.  TensorFlow owns this arithmetic in the reference
# (tf.image.resize / random_flip_left_right / tf.cast; tensorflow is absent from /root/reference and from this image):
# PARITY UNPINNED — restated from TensorFlow's published half-pixel-centre bilinear kernel (resize_bilinear_op.cc:
# in = (out + 0.5) * (in_size / out_size) - 0.5, lower = max(floor(in), 0), upper = min(ceil(in), in_size - 1),
# lerp = in - floor(in); top/bottom row interpolated along x first, then along y), all in float32.
# ------------------------------------------------------------------------------------------


def _bilinear_axis(in_size, out_size):
  f32 = np.float32
  scale = f32(in_size) / f32(out_size)
  src = (np.arange(out_size, dtype=f32) + f32(0.5)) * scale - f32(0.5)
  fl = np.floor(src)
  lower = np.maximum(fl.astype(np.int64), 0)
  upper = np.minimum(np.ceil(src).astype(np.int64), in_size - 1)
  return lower, upper, (src - fl).astype(f32)


def preprocess_train(images_u8, boxes=None, flips=None, size=None, vmin=-1.0, vmax=1.0, in_min=0.0, in_max=255.0,
                     clip_values=False):
  """configs/ae_i1k.py:64-69 after decoding: crop (pp/ops_image.py:234-238) -> get_resize (:75-85: bilinear, clip,
  cast back to uint8) -> flip_lr (:306-314) -> value_range (pp/ops_general.py:51-60).  numpy float32, no fused
  multiply-adds.  size: None (source size), int or (h, w).  Returns (float32 [n, Sh, Sw, C], uint8 intermediate)."""
  f32 = np.float32
  x = np.asarray(images_u8)
  n, H, W, C = x.shape
  Sh, Sw = (H, W) if size is None else ((int(size), int(size)) if np.isscalar(size) else (int(size[0]), int(size[1])))
  out = np.empty((n, Sh, Sw, C), dtype=f32)
  mid = np.empty((n, Sh, Sw, C), dtype=np.uint8)
  for i in range(n):
    y0, x0, bh, bw = (0, 0, H, W) if boxes is None else [int(v) for v in boxes[i]]
    crop = x[i, y0:y0 + bh, x0:x0 + bw].astype(f32)
    yl, yu, ly = _bilinear_axis(bh, Sh)
    xl, xu, lx = _bilinear_axis(bw, Sw)
    lx_, ly_ = lx[None, :, None], ly[:, None, None]
    tl, tr = crop[yl][:, xl], crop[yl][:, xu]
    bl, br = crop[yu][:, xl], crop[yu][:, xu]
    top = tl + (tr - tl) * lx_
    bot = bl + (br - bl) * lx_
    v = top + (bot - top) * ly_
    q = np.clip(v, f32(0), f32(255)).astype(np.uint8)            # tf.cast truncates
    if flips is not None and bool(flips[i]):
      q = q[:, ::-1]
    mid[i] = q
    f = (q.astype(f32) - f32(in_min)) / (f32(in_max) - f32(in_min))
    f = f32(vmin) + f * f32(float(vmax) - float(vmin))      # Python-scalar difference, rounded once (ops_general.py:57)
    if clip_values:
      f = np.clip(f, f32(vmin), f32(vmax))
    out[i] = f
  return out, mid
