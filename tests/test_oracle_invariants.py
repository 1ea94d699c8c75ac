"""Pins of the CPU oracle (oracle/umd_oracle.py) that can be derived from the reference's own source
(numbered as there) plus piecewise cross-checks against independent implementations that
exist in this image (torch.nn.functional, torch.optim.AdamW, transformers' ViT-MAE masking).

The reference (JAX/Flax/Optax) cannot be imported here and ships no golden vectors for this path, so these
are the strongest checks available: "parity unpinned" in the sense of the task statement.  CPU only.
"""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import umd_oracle as O
from tests import util as U


def _tiny(adaln=True, num_classes=None, depth=2, dec_depth=1, zero_adaln=False, seed=0):
  model, ocfg = U.make_models("S/4", adaln=adaln, num_classes=num_classes, depth=depth, dec_depth=dec_depth)
  if zero_adaln:
    from small_vision_b200.params import init_arena, tree_from_arena
    params = U.cpu_tree(tree_from_arena(model.layout, init_arena(model.layout, seed, "cpu", nonzero_adaln=False)))
  else:
    params = U.cpu_tree(U.perturb_init(model, seed, "cpu"))
  return model, ocfg, params


# ------------------------------------------------------------------------------ (1) adaLN-Zero identity
def test_pin1_adaln_zero_blocks_are_identity():
  """vit.py:71 zero-init Dense(6D) + ae.py:94 zero-init final_modulation: every block is the identity, so
  pred = ConvT(LN_dec(dec_in)[:, 1:]) with dec_in built from LN_enc(embed)."""
  model, ocfg, p = _tiny(zero_adaln=True)
  g = torch.Generator().manual_seed(3)
  img = torch.rand(3, 64, 64, 3, generator=g) * 2 - 1
  t = torch.randint(1, 1000, (3, 1), generator=g, dtype=torch.int32)
  pred, out = O.model_apply(p, ocfg, img, t=t, mask=0.0)
  D = ocfg["width"]
  x = torch.einsum("nijabc,abcd->nijd", O.patchify(img, 4), p["embedding"]["kernel"]) + p["embedding"]["bias"]
  x = x.reshape(3, 256, D) + p["pos_embedding"]
  x = torch.cat([p["cls"].expand(3, -1, -1), x], 1)
  enc = O.layer_norm(x, p["Encoder"]["encoder_norm"]["scale"], p["Encoder"]["encoder_norm"]["bias"])
  rep = enc[:, :4].mean(1)
  xd = torch.cat([rep[:, None], enc[:, 4:] + p["dec_pos_embedding"]], 1)
  xd = O.layer_norm(xd, p["Decoder"]["encoder_norm"]["scale"], p["Decoder"]["encoder_norm"]["bias"])[:, 1:]
  want = O.conv_transpose_unpatchify(xd.reshape(3, 16, 16, D), p["final_conv"]["kernel"], p["final_conv"]["bias"])
  assert torch.allclose(pred, want, atol=1e-5, rtol=1e-5)
  assert torch.allclose(out["pre_logits"], rep, atol=1e-6)


# ------------------------------------------------------------------------------ (2) masking algebra
@pytest.mark.parametrize("ratio,keep", [(0.375, 160), (0.75, 64), (0.0, 256)])
def test_pin2_len_keep_values(ratio, keep):
  assert O.len_keep_of(256, ratio) == keep


def test_pin2_random_masking_algebra_and_stable_ties():
  g = torch.Generator().manual_seed(0)
  n, L, D = 5, 256, 8
  noise = torch.rand(n, L, generator=g)
  noise[0, 5] = noise[0, 200]
  noise[1, :] = 0.25            # a fully tied row: stable argsort must return arange
  x = torch.randn(n, L, D, generator=g)
  xm, mask, ids_restore = O.random_masking(x, 0.375, noise)
  ids_shuffle = torch.argsort(noise, dim=1, stable=True)
  ar = torch.arange(L).expand(n, L)
  assert torch.equal(torch.gather(ids_restore, 1, ids_shuffle), ar)
  assert torch.equal(ids_shuffle[1], torch.arange(L))
  assert torch.equal(mask.sum(1), torch.full((n,), float(L - 160)))
  keep = ids_shuffle[:, :160]
  assert torch.equal(xm, torch.gather(x, 1, keep[:, :, None].expand(-1, -1, D)))
  assert torch.equal(torch.gather(mask, 1, keep), torch.zeros(n, 160))
  # tie-break by index: position 5 sorts before 200
  pos = {int(v): i for i, v in enumerate(ids_shuffle[0].tolist())}
  assert pos[5] + 1 == pos[200]


def test_pin2_masking_matches_transformers_vit_mae():
  """Independent implementation of the same MAE masking (transformers ViTMAEEmbeddings.random_masking)."""
  tr = pytest.importorskip("transformers")
  from transformers.models.vit_mae.modeling_vit_mae import ViTMAEEmbeddings
  cfg = tr.ViTMAEConfig(hidden_size=8, image_size=64, patch_size=4, num_channels=3, mask_ratio=0.75)
  emb = ViTMAEEmbeddings(cfg)
  g = torch.Generator().manual_seed(1)
  noise = torch.rand(4, 256, generator=g)
  x = torch.randn(4, 256, 8, generator=g)
  xm, mask, ids_restore = O.random_masking(x, 0.75, noise)
  hx, hmask, hrestore = emb.random_masking(x, noise=noise)
  assert torch.equal(xm, hx) and torch.equal(mask, hmask) and torch.equal(ids_restore, hrestore)


def test_pixel_mask_layout():
  m = torch.zeros(1, 256)
  m[0, 17] = 1            # patch (1, 1)
  pm = O.sequence_mask_to_image_mask(m, 4, 64)
  assert pm.shape == (1, 64, 64, 1)
  assert float(pm.sum()) == 16 and float(pm[0, 4:8, 4:8, 0].sum()) == 16


# ------------------------------------------------------------------------------ (3) DP = mean of shard grads
def test_pin3_loss_is_per_sample_mean_so_dp_gradient_is_mean_of_shards():
  model, ocfg, p = _tiny()
  tc = dict(mask_ratio=0.375, mask_ratio_no_noise=0.75, no_noise_prob=0.5, use_labels=False)
  gd = O.gaussian_diffusion_tables("cosine", 1000)
  hp = dict(clip_norm=1.0, peak_lr=1e-3, warmup_steps=0, total_steps=100, b1=0.9, b2=0.95, wd=0.05)
  B = 8
  batch, rand = U.make_batch(model, B, n_noise=4, seed=5, device="cpu")
  st = {"params": p, "gd": gd, "opt": O.init_opt_state(p)}
  _, meas, ex = O.update_step(st, batch, ocfg, tc, hp, rand)
  full = O.flatten_tree(ex["grads"])
  # two shards, each with 2 noised + 2 clean samples
  acc, losses = None, []
  for r in range(2):
    idx_n, idx_c = [2 * r, 2 * r + 1], [4 + 2 * r, 5 + 2 * r]
    b = {"image": torch.cat([batch["image"][idx_n], batch["image"][idx_c]]), "label": batch["label"][idx_n + idx_c]}
    rd = {"t": rand["t"][idx_n], "noise": rand["noise"][idx_n], "mask_noise_noise": rand["mask_noise_noise"][idx_n],
          "mask_noise_clean": rand["mask_noise_clean"][[i - 4 for i in idx_c]]}
    _, m, e = O.update_step(st, b, ocfg, tc, hp, rd)
    losses.append(m["training_loss"])
    fg = O.flatten_tree(e["grads"])
    acc = fg if acc is None else {k: acc[k] + fg[k] for k in fg}
  assert abs(sum(losses) / 2 - meas["training_loss"]) < 1e-5
  for k in full:
    assert torch.allclose(acc[k] / 2, full[k], atol=2e-6, rtol=2e-4), k


# ------------------------------------------------------------------------------ (4) time embedding at t = 0
def test_pin4_time_embedding():
  e = O.time_embedding(torch.zeros(2, 1, dtype=torch.int32), 384, torch.float32)
  assert torch.equal(e[:, :192], torch.zeros(2, 192)) and torch.equal(e[:, 192:], torch.ones(2, 192))
  e = O.time_embedding(torch.tensor([[7]], dtype=torch.int32), 384, torch.float64)
  k = torch.arange(192, dtype=torch.float64)
  want = 7 * torch.exp(-k * math.log(10000) / 191)
  assert torch.allclose(e[0, :192], torch.sin(want)) and torch.allclose(e[0, 192:], torch.cos(want))


# ------------------------------------------------------------------------------ (5) schedules
@pytest.mark.parametrize("name", ["cosine", "linear"])
def test_pin5_schedule_tables(name):
  gd = O.gaussian_diffusion_tables(name, 1000)
  assert len(gd) == 13 and all(v.dtype == np.float64 and v.shape == (1000,) for v in gd.values())
  assert gd["betas"].max() <= 0.999 and gd["betas"].min() > 0
  assert np.all(np.diff(gd["alphas_cumprod"]) < 0)
  assert np.allclose(gd["sqrt_alphas_cumprod"] ** 2 + gd["sqrt_one_minus_alphas_cumprod"] ** 2, 1.0, atol=1e-12)
  if name == "linear":
    assert gd["betas"][0] == pytest.approx(1e-4) and gd["betas"][-1] == pytest.approx(2e-2)
  from small_vision_b200.diffusion import create_gaussian_diffusion
  ours = create_gaussian_diffusion(name, 1000)
  for k, v in gd.items():
    assert np.array_equal(np.asarray(ours[k]), v), k


def test_q_sample_closed_form():
  gd = O.gaussian_diffusion_tables("cosine", 1000)
  g = torch.Generator().manual_seed(0)
  x0, nz = torch.randn(4, 8, 8, 3, generator=g), torch.randn(4, 8, 8, 3, generator=g)
  t = torch.tensor([[0], [10], [500], [999]], dtype=torch.int32)
  xt = O.q_sample(gd, x0, t, nz)
  for i, ti in enumerate([0, 10, 500, 999]):
    a = np.float32(gd["sqrt_alphas_cumprod"][ti]); b = np.float32(gd["sqrt_one_minus_alphas_cumprod"][ti])
    assert torch.equal(xt[i], float(a) * x0[i] + float(b) * nz[i])


# ------------------------------------------------------------------------------ (6) MAE branch: eps half untouched
def test_pin6_mae_branch_gives_zero_gradient_to_eps_half_of_final_conv():
  model, ocfg, p = _tiny()
  tc = dict(mask_ratio=0.375, mask_ratio_no_noise=0.75, no_noise_prob=1.0, use_labels=False)
  batch, rand = U.make_batch(model, 4, n_noise=0, seed=2, device="cpu")
  rand["t"] = torch.zeros(0, 1, dtype=torch.int32)
  st = {"params": p, "gd": O.gaussian_diffusion_tables(), "opt": O.init_opt_state(p)}
  hp = dict(clip_norm=1.0, peak_lr=1e-3, warmup_steps=0, total_steps=100, b1=0.9, b2=0.95, wd=0.05)
  _, _, ex = O.update_step(st, batch, ocfg, tc, hp, rand)
  gk = ex["grads"]["final_conv"]["kernel"]
  assert float(gk[..., 3:].abs().max()) == 0.0 and float(gk[..., :3].abs().max()) > 0
  assert float(ex["grads"]["final_conv"]["bias"][3:].abs().max()) == 0.0


# ------------------------------------------------------------------------------ (7) adaln=False cond token
def test_pin7_cond_token_row_is_stripped():
  """vit.py:73-74,111-112: the block prepends cond and strips row 0; the output length is unchanged and the
  result depends on cond (through attention) but not on anything written to the stripped row."""
  model, ocfg, p = _tiny(adaln=False)
  g = torch.Generator().manual_seed(0)
  x = torch.randn(2, 20, 384, generator=g)
  c1, c2 = torch.randn(2, 384, generator=g), torch.randn(2, 384, generator=g)
  bp = p["Encoder"][O.scanned_key(p["Encoder"])]
  y1 = O.block(x, c1, bp, 0, adaln=False, num_heads=6)
  y2 = O.block(x, c2, bp, 0, adaln=False, num_heads=6)
  assert y1.shape == x.shape and not torch.allclose(y1, y2)
  assert "Dense_0" not in bp and "final_modulation" not in p


# ------------------------------------------------------------------------------ (8) fp64 finite differences
def test_pin8_finite_difference_gradients_fp64():
  model, ocfg, p = _tiny(depth=1, dec_depth=1)
  tc = dict(mask_ratio=0.375, mask_ratio_no_noise=0.75, no_noise_prob=0.5, use_labels=False)
  batch, rand = U.make_batch(model, 2, n_noise=1, seed=9, device="cpu")
  gd = O.gaussian_diffusion_tables()
  x_t = O.q_sample(gd, batch["image"][:1].double(), rand["t"], rand["noise"].double())

  def loss_of(params):
    l, _ = O.loss_fn(params, ocfg, tc, batch["image"][:1], x_t, batch["image"][1:], rand["t"], rand["noise"], None, rand,
                     dtype=torch.float64)
    return l

  flat = {k: v.double().clone().requires_grad_(True) for k, v in O.flatten_tree(p).items()}
  loss_of(O.unflatten_tree(flat)).backward()
  gen = torch.Generator().manual_seed(0)
  probes = [("final_conv", "kernel"), ("pos_embedding",), ("cls",), ("image_mask_embedding",),
            ("Encoder", O.scanned_key(p["Encoder"]), "Dense_0", "kernel"),
            ("Decoder", O.scanned_key(p["Decoder"]), "MlpBlock_0", "Dense_0", "bias"),
            ("time_trunk", "Dense_0", "kernel"), ("final_modulation", "kernel"), ("embedding", "kernel")]
  for path in probes:
    v = flat[path]
    for _ in range(2):
      i = int(torch.randint(0, v.numel(), (1,), generator=gen))
      eps = 1e-5
      with torch.no_grad():
        base = {k: w.detach().clone() for k, w in flat.items()}
        base[path].view(-1)[i] += eps
        lp = float(loss_of(O.unflatten_tree(base)))
        base[path].view(-1)[i] -= 2 * eps
        lm = float(loss_of(O.unflatten_tree(base)))
      fd = (lp - lm) / (2 * eps)
      an = float(v.grad.view(-1)[i])
      assert abs(fd - an) <= 1e-6 + 1e-4 * abs(an), (path, i, fd, an)


# ------------------------------------------------------------------------------ (9) patchify / un-patchify
def test_pin9_unpatchify_orientation():
  """App. A.7: flax ConvTranspose(transpose_kernel=False) == lhs-dilated cross-correlation with the un-flipped
  kernel == torch conv_transpose2d with the spatially flipped kernel."""
  g = torch.Generator().manual_seed(0)
  x = torch.randn(2, 4, 4, 5, generator=g)
  K = torch.randn(4, 4, 5, 6, generator=g)
  b = torch.randn(6, generator=g)
  got = O.conv_transpose_unpatchify(x, K, b, flip=True)
  # emulate jax.lax.conv_transpose: dilate the input by the stride, pad p-1, cross-correlate with K as stored
  xin = x.permute(0, 3, 1, 2)
  dil = torch.zeros(2, 5, 13, 13)
  dil[:, :, ::4, ::4] = xin
  w = K.permute(3, 2, 0, 1)                        # [O, I, kh, kw], no flip
  want = F.conv2d(F.pad(dil, (3, 3, 3, 3)), w).permute(0, 2, 3, 1) + b
  assert torch.allclose(got, want, atol=1e-5)
  tw = F.conv_transpose2d(xin, K.flip(0, 1).permute(2, 3, 0, 1), stride=4).permute(0, 2, 3, 1) + b
  assert torch.allclose(got, tw, atol=1e-5)
  assert not torch.allclose(got, O.conv_transpose_unpatchify(x, K, b, flip=False), atol=1e-3)
  # index identity: one-hot kernel entry (a', b') lands at pixel offset (p-1-a', p-1-b')
  K1 = torch.zeros(4, 4, 1, 1); K1[1, 2, 0, 0] = 1
  y = O.conv_transpose_unpatchify(torch.ones(1, 1, 1, 1), K1, torch.zeros(1))
  assert float(y[0, 2, 1, 0]) == 1 and float(y.sum()) == 1


def test_patch_embed_is_strided_cross_correlation():
  g = torch.Generator().manual_seed(0)
  img = torch.randn(2, 16, 16, 3, generator=g)
  K = torch.randn(4, 4, 3, 7, generator=g)
  ours = torch.einsum("nijabc,abcd->nijd", O.patchify(img, 4), K)
  ref = F.conv2d(img.permute(0, 3, 1, 2), K.permute(3, 2, 0, 1), stride=4).permute(0, 2, 3, 1)
  assert torch.allclose(ours, ref, atol=1e-5)


# ------------------------------------------------------------------------------ (10) optimiser closed forms
def test_pin10_first_step_with_warmup_leaves_params_unchanged():
  p = {"w": {"kernel": torch.randn(4, 4), "bias": torch.randn(4)}}
  g = {"w": {"kernel": torch.randn(4, 4) * 10, "bias": torch.randn(4)}}
  hp = dict(clip_norm=1.0, peak_lr=1e-3, warmup_steps=10, total_steps=100, b1=0.9, b2=0.95, wd=0.05)
  np_, opt, upd, gn = O.optimizer_update(g, O.init_opt_state(p), p, hp)
  assert all(float(u.abs().max()) == 0 for u in upd.values())
  assert torch.equal(np_["w"]["kernel"], p["w"]["kernel"]) and opt["count"] == 1
  # second step: lr(1) = peak / warmup_steps
  assert O.warmup_cosine_lr(1, peak=1e-3, warmup_steps=10, decay_steps=100) == pytest.approx(1e-4)


def test_pin10_adamw_step1_closed_form_and_wd_mask():
  g0 = torch.Generator().manual_seed(0)
  p = {"a": {"kernel": torch.randn(6, 6, generator=g0), "bias": torch.randn(6, generator=g0)},
       "cls": torch.randn(1, 4, 6, generator=g0), "pos_embedding": torch.randn(1, 5, 6, generator=g0),
       "ln": {"scale": torch.ones(6)}}
  g = {k: ({kk: torch.randn(vv.shape, generator=g0) for kk, vv in v.items()} if isinstance(v, dict)
           else torch.randn(v.shape, generator=g0)) for k, v in p.items()}
  hp = dict(clip_norm=1.0, peak_lr=2e-3, warmup_steps=0, total_steps=100, b1=0.9, b2=0.95, wd=0.05)
  newp, opt, upd, gn = O.optimizer_update(g, O.init_opt_state(p), p, hp)
  fg, fp = O.flatten_tree(g), O.flatten_tree(p)
  assert gn == pytest.approx(math.sqrt(sum(float((v.double() ** 2).sum()) for v in fg.values())))
  mask = O.weight_decay_mask(p)
  assert mask == {("a", "kernel"): True, ("a", "bias"): False, ("cls",): False, ("pos_embedding",): True,
                  ("ln", "scale"): True}
  lr = 2e-3  # count 0, no warm-up
  for k in fg:
    gc = fg[k] / gn  # gn > 1: clipped to unit norm
    want = -lr * (gc / (gc.abs() + 1e-8) + (0.05 * fp[k] if mask[k] else 0))
    assert torch.allclose(upd[k], want, atol=1e-7, rtol=1e-5), k
    assert opt["mu"][k].dtype == torch.bfloat16 and torch.equal(opt["mu"][k], (0.1 * gc).to(torch.bfloat16))


def test_adamw_matches_torch_optim_when_mu_is_exact():
  """torch.optim.AdamW applies the same update when the bf16 storage of mu is lossless (power-of-two grads)."""
  p0 = torch.tensor([0.5, -1.0, 2.0, 4.0])
  grads = [torch.tensor([0.25, -0.5, 0.125, 0.0625]), torch.tensor([0.5, 0.25, -0.25, 0.125])]
  hp = dict(clip_norm=1e9, peak_lr=1e-2, warmup_steps=0, total_steps=10 ** 9, b1=0.5, b2=0.95, wd=0.05)
  p = {"k": p0.clone()}
  opt = O.init_opt_state(p)
  tp = torch.nn.Parameter(p0.clone())
  topt = torch.optim.AdamW([tp], lr=1e-2, betas=(0.5, 0.95), eps=1e-8, weight_decay=0.05)
  for g in grads:
    p, opt, _, _ = O.optimizer_update({"k": g}, opt, p, hp)
    tp.grad = g.clone()
    topt.step()
  assert torch.allclose(p["k"], tp.detach(), atol=1e-6, rtol=1e-5)


def test_warmup_cosine_schedule_shape():
  kw = dict(peak=1.0, warmup_steps=10, decay_steps=110)
  assert O.warmup_cosine_lr(0, **kw) == 0 and O.warmup_cosine_lr(10, **kw) == pytest.approx(1.0)
  assert O.warmup_cosine_lr(60, **kw) == pytest.approx(0.5) and O.warmup_cosine_lr(110, **kw) == pytest.approx(0.0, abs=1e-12)
  from small_vision_b200.config import warmup_cosine_lr
  for c in (0, 1, 5, 10, 11, 60, 109, 110, 500):
    assert warmup_cosine_lr(c, **kw) == pytest.approx(O.warmup_cosine_lr(c, **kw), abs=1e-12)


# ------------------------------------------------------------------------------ layer semantics vs torch.nn.functional
def test_layers_match_torch_functional():
  g = torch.Generator().manual_seed(0)
  x = torch.randn(3, 10, 384, generator=g) * 2 + 0.5
  sc, bi = torch.randn(384, generator=g), torch.randn(384, generator=g)
  assert torch.allclose(O.layer_norm(x, sc, bi), F.layer_norm(x, (384,), sc, bi, eps=1e-6), atol=2e-5)
  assert torch.allclose(O.gelu_tanh(x), F.gelu(x, approximate="tanh"), atol=1e-6)
  H, Dh = 6, 64
  ap = {n: {"kernel": torch.randn(384, H, Dh, generator=g) * 0.05, "bias": torch.randn(H, Dh, generator=g) * 0.1}
        for n in ("query", "key", "value")}
  ap["out"] = {"kernel": torch.randn(H, Dh, 384, generator=g) * 0.05, "bias": torch.randn(384, generator=g) * 0.1}
  got = O.attention(x, ap, H)
  q, k, v = (torch.einsum("bsd,dhk->bhsk", x, ap[n]["kernel"]) + ap[n]["bias"][None, :, None, :] for n in ("query", "key", "value"))
  o = F.scaled_dot_product_attention(q, k, v)      # scale 1/sqrt(Dh), no mask
  want = torch.einsum("bhsk,hko->bso", o, ap["out"]["kernel"]) + ap["out"]["bias"]
  assert torch.allclose(got, want, atol=2e-5)


def test_cfg_forward_doubles_the_batch():
  """ae.py:177-195: pred = uncond + s * (cond - uncond) with the null class for the unconditional half."""
  model, ocfg, p = _tiny(num_classes=10)
  g = torch.Generator().manual_seed(0)
  img = torch.rand(2, 64, 64, 3, generator=g) * 2 - 1
  t = torch.tensor([[5], [900]], dtype=torch.int32)
  y = torch.tensor([3, 7])
  co, _ = O.model_apply(p, ocfg, img, t=t, y=y)
  un, _ = O.model_apply(p, ocfg, img, t=t, y=torch.tensor([10, 10]))
  mix, _ = O.model_apply(p, ocfg, img, t=t, y=y, cfg_scale=1.5)
  assert torch.allclose(mix, un + 1.5 * (co - un), atol=1e-5)
  none_y, _ = O.model_apply(p, ocfg, img, t=t, y=None)
  assert torch.allclose(none_y, un, atol=1e-6)


def test_loss_reduces_to_masked_pixel_means():
  model, ocfg, p = _tiny()
  tc = dict(mask_ratio=0.375, mask_ratio_no_noise=0.75, no_noise_prob=0.5, use_labels=False)
  batch, rand = U.make_batch(model, 4, n_noise=2, seed=1, device="cpu")
  gd = O.gaussian_diffusion_tables()
  x_t = O.q_sample(gd, batch["image"][:2], rand["t"], rand["noise"])
  loss, aux = O.loss_fn(p, ocfg, tc, batch["image"][:2], x_t, batch["image"][2:], rand["t"], rand["noise"], None, rand)
  pc, mc = aux["pred_clean"], aux["out_clean"]["mask"]
  mae = ((pc[..., :3] - batch["image"][2:]) ** 2 * mc).sum() / (3 * mc.sum())
  pn, mn = aux["pred_noise"], aux["out_noise"]["mask"]
  x0l = ((pn[..., :3] - batch["image"][:2]) ** 2 * mn).sum() / (3 * mn.sum())
  epl = ((pn[..., 3:] - rand["noise"]) ** 2 * (1 - mn)).sum() / (3 * (1 - mn).sum())
  assert float(loss) == pytest.approx(float(0.5 * (x0l + epl) / 2 + 0.5 * mae), rel=1e-5)
  assert float(mc.mean()) == pytest.approx(192 / 256) and float(mn.mean()) == pytest.approx(96 / 256)
