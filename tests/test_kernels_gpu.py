"""GPU parity of the stand-alone C-ABI entry points (include/umd_b200.h) against the oracle's functions or a
plain fp32 torch restatement of the same op on identical inputs.  Integer/index results are bit-exact; fp32
kernels match to fp32 round-off; kernels with bf16 inputs/outputs to bf16 round-off (2^-8 relative)."""
import ctypes as C
import math

import pytest
import torch

from tests import util as U
from oracle import umd_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _lib():
  from small_vision_b200 import lib
  return lib


# ------------------------------------------------------------------------------------------ q_sample
@pytest.mark.parametrize("n,shape", [(1, (64, 64, 3)), (7, (64, 64, 3)), (5, (32, 32, 4)), (256, (64, 64, 3))])
@pytest.mark.parametrize("sched", ["cosine", "linear"])
def test_qsample_matches_oracle(n, shape, sched):
  from small_vision_b200.diffusion import create_gaussian_diffusion, q_sample, to_device
  gd = create_gaussian_diffusion(sched, 1000)
  g = torch.Generator().manual_seed(n)
  x0 = torch.rand((n,) + shape, generator=g) * 2 - 1
  noise = torch.randn((n,) + shape, generator=g)
  t = torch.randint(0, 1000, (n, 1), generator=g, dtype=torch.int32)
  t[0, 0] = 999
  ref = O.q_sample(O.gaussian_diffusion_tables(sched, 1000), x0, t, noise)
  out = q_sample(gd=to_device(gd, DEV), x_start=x0.to(DEV), t=t.to(DEV), noise=noise.to(DEV))
  # fp32 a*x + b*y: the kernel may contract to an fma => <= 1 ulp of the larger term
  assert torch.allclose(out.cpu(), ref, rtol=0, atol=4e-7 * float(ref.abs().max() + 1))


def test_qsample_linearity_full_size():
  """size-independent property at the BASELINE size (256 noised images per GPU): q(x0, 0-noise) + q(0, noise) = q(x0, noise)"""
  from small_vision_b200.diffusion import create_gaussian_diffusion, q_sample, to_device
  gd = to_device(create_gaussian_diffusion("cosine", 1000), DEV)
  g = torch.Generator(device=DEV).manual_seed(0)
  x0 = torch.rand(256, 64, 64, 3, device=DEV, generator=g)
  nz = torch.randn(256, 64, 64, 3, device=DEV, generator=g)
  t = torch.randint(0, 1000, (256,), device=DEV, generator=g, dtype=torch.int32)
  z = torch.zeros_like(x0)
  a = q_sample(gd=gd, x_start=x0, t=t, noise=z) + q_sample(gd=gd, x_start=z, t=t, noise=nz)
  b = q_sample(gd=gd, x_start=x0, t=t, noise=nz)
  assert torch.allclose(a, b, rtol=0, atol=1e-6)


# ------------------------------------------------------------------------------------------ mask argsort
@pytest.mark.parametrize("n,L,keep", [(1, 256, 160), (33, 256, 64), (4, 256, 0), (3, 256, 256), (5, 64, 16), (2, 1024, 640)])
def test_mask_argsort_bit_exact(n, L, keep):
  from small_vision_b200.model import mask_argsort
  g = torch.Generator().manual_seed(L + n)
  noise = torch.rand(n, L, generator=g)
  # heavy ties (quantised values), signed zeros and an all-equal row
  noise[0] = (noise[0] * 8).floor() / 8
  if n > 1:
    noise[1] = 0.25
    noise[1, 3] = -0.0
    noise[1, 9] = 0.0
  ids_shuffle, ids_restore, mask = mask_argsort(noise.to(DEV), keep)
  ref_shuffle = torch.argsort(noise, dim=1, stable=True)
  ref_restore = torch.argsort(ref_shuffle, dim=1, stable=True)
  assert torch.equal(ids_shuffle.cpu().long(), ref_shuffle)
  assert torch.equal(ids_restore.cpu().long(), ref_restore)
  x = torch.zeros(n, L, 1)
  _, ref_mask, _ = O.random_masking(x, 1 - keep / L, noise) if keep not in (0, L) else (None, (ref_restore >= keep).float(), None)
  assert torch.equal(mask.cpu(), (ref_restore >= keep).float())
  if keep not in (0, L):
    assert torch.equal(mask.cpu(), ref_mask)
  # pins (2) of SURVEY.md §8c
  ar = torch.arange(L).expand(n, L)
  assert torch.equal(torch.gather(ids_restore.cpu().long(), 1, ids_shuffle.cpu().long()), ar)
  assert torch.equal(mask.cpu().sum(1), torch.full((n,), float(L - keep)))


def test_mask_argsort_full_size_is_a_permutation():
  from small_vision_b200.model import mask_argsort
  noise = torch.rand(2048, 256, device=DEV)
  ids_shuffle, ids_restore, mask = mask_argsort(noise, 160)
  s = torch.sort(ids_shuffle.long(), dim=1).values
  assert torch.equal(s, torch.arange(256, device=DEV).expand(2048, 256))
  sorted_noise = torch.gather(noise, 1, ids_shuffle.long())
  assert bool((sorted_noise[:, 1:] >= sorted_noise[:, :-1]).all())
  assert torch.equal(mask.sum(1), torch.full((2048,), 96.0, device=DEV))


# ------------------------------------------------------------------------------------------ LayerNorm + modulate
def _ln_ref(x, gamma, beta, shift, scale, n0, s0, n1, s1):
  y = O.layer_norm(x, gamma, beta)
  if scale is not None:
    rows = torch.cat([torch.arange(n0).repeat_interleave(s0), n0 + torch.arange(n1).repeat_interleave(s1)])
    y = y * (1 + scale[rows]) + shift[rows]
  return y


@pytest.mark.parametrize("D", [384, 768, 1024])
@pytest.mark.parametrize("n0,s0,n1,s1", [(3, 164, 2, 68), (2, 257, 0, 0), (0, 0, 3, 69)])
@pytest.mark.parametrize("mod", [True, False])
def test_ln_modulate_fwd_bwd(D, n0, s0, n1, s1, mod):
  lib = _lib()
  L = lib.load()
  g = torch.Generator().manual_seed(D + n0)
  rows, B = n0 * s0 + n1 * s1, n0 + n1
  x = (torch.randn(rows, D, generator=g) * 1.5 + 0.3).requires_grad_(True)
  gamma = (1 + 0.1 * torch.randn(D, generator=g)).requires_grad_(True)
  beta = (0.1 * torch.randn(D, generator=g)).requires_grad_(True)
  modt = (0.2 * torch.randn(B, 6 * D, generator=g)).requires_grad_(True)
  shift = modt[:, 0:D] if mod else None
  scale = modt[:, D:2 * D] if mod else None
  y = _ln_ref(x, gamma, beta, shift, scale, n0, s0, n1, s1)
  dy = torch.randn(rows, D, generator=g)
  y.backward(dy)
  xg, gg, bg, mg = x.detach().to(DEV), gamma.detach().to(DEV), beta.detach().to(DEV), modt.detach().to(DEV)
  out = torch.empty(rows, D, device=DEV)
  mean = torch.empty(rows, device=DEV)
  rstd = torch.empty(rows, device=DEV)
  null = C.c_void_p(0)
  sh_p = C.c_void_p(mg.data_ptr()) if mod else null
  sc_p = C.c_void_p(mg.data_ptr() + 4 * D) if mod else null
  lib.check(L.umd_ln_modulate_fwd(lib.ptr(xg), lib.ptr(gg), lib.ptr(bg), sh_p, sc_p, C.c_longlong(6 * D), n0, s0, n1, s1, D,
                                  lib.ptr(out), 0, lib.ptr(mean), lib.ptr(rstd), lib.current_stream()), "ln fwd")
  assert torch.allclose(out.cpu(), y.detach(), rtol=1e-4, atol=1e-4)
  outb = torch.empty(rows, D, device=DEV, dtype=torch.bfloat16)
  lib.check(L.umd_ln_modulate_fwd(lib.ptr(xg), lib.ptr(gg), lib.ptr(bg), sh_p, sc_p, C.c_longlong(6 * D), n0, s0, n1, s1, D,
                                  lib.ptr(outb), 1, null, null, lib.current_stream()), "ln fwd bf16")
  assert torch.allclose(outb.float().cpu(), y.detach(), rtol=2 ** -7, atol=1e-2)
  # backward (fp32 dy), accumulate into a non-zero dx
  dx0 = torch.randn(rows, D, generator=g)
  dx = dx0.to(DEV).clone()
  dmod = torch.zeros(B, 6 * D, device=DEV)
  dgamma = torch.zeros(D, device=DEV)
  dbeta = torch.zeros(D, device=DEV)
  dsh_p = C.c_void_p(dmod.data_ptr()) if mod else null
  dsc_p = C.c_void_p(dmod.data_ptr() + 4 * D) if mod else null
  lib.check(L.umd_ln_modulate_bwd(lib.ptr(dy.to(DEV)), 0, lib.ptr(xg), lib.ptr(mean), lib.ptr(rstd), lib.ptr(gg), lib.ptr(bg),
                                  sc_p, C.c_longlong(6 * D), n0, s0, n1, s1, D, lib.ptr(dx), 1, dsh_p, dsc_p,
                                  C.c_longlong(6 * D), lib.ptr(dgamma), lib.ptr(dbeta), lib.current_stream()), "ln bwd")
  assert U.rel_l2(dx.cpu() - dx0, x.grad) < 1e-4
  assert U.rel_l2(dgamma.cpu(), gamma.grad) < 1e-4
  assert U.rel_l2(dbeta.cpu(), beta.grad) < 1e-4
  if mod:
    assert U.rel_l2(dmod.cpu()[:, :2 * D], modt.grad[:, :2 * D]) < 1e-4
    assert float(dmod[:, 2 * D:].abs().max()) == 0.0


@pytest.mark.parametrize("D", [384, 768])
@pytest.mark.parametrize("n0,s0,n1,s1", [(3, 164, 2, 68), (2, 257, 0, 0)])
@pytest.mark.parametrize("with_gate", [True, False])
def test_ln_modulate_bwd_with_fused_gate_stage(D, n0, s0, n1, s1, with_gate):
  """x = x_prev + gate[sample] * (h + bias) feeds the LayerNorm (vit.py:89-98): one launch returns the stream gradient,
  dz = gate * dx (bf16), dgate and the bias gradient next to the LayerNorm's own parameter gradients (App. E 1, 4-6)."""
  lib = _lib()
  L = lib.load()
  g = torch.Generator().manual_seed(D + s0)
  rows, B = n0 * s0 + n1 * s1, n0 + n1
  sample = torch.cat([torch.arange(n0).repeat_interleave(s0), n0 + torch.arange(n1).repeat_interleave(s1)])
  x_prev = torch.randn(rows, D, generator=g)
  h = torch.randn(rows, D, generator=g).to(torch.bfloat16).float()
  bias = torch.zeros(D, requires_grad=True)
  gate = (0.5 * torch.randn(B, D, generator=g)).requires_grad_(True) if with_gate else None
  gamma = (1 + 0.1 * torch.randn(D, generator=g)).requires_grad_(True)
  beta = (0.1 * torch.randn(D, generator=g)).requires_grad_(True)
  modt = (0.2 * torch.randn(B, 2 * D, generator=g)).requires_grad_(True)
  z = (h + bias).requires_grad_(True)
  z.retain_grad()
  x = x_prev + (gate[sample] * z if with_gate else z)
  x.retain_grad()
  y = _ln_ref(x, gamma, beta, modt[:, :D], modt[:, D:], n0, s0, n1, s1)
  dy = torch.randn(rows, D, generator=g).to(torch.bfloat16)
  dx0 = torch.randn(rows, D, generator=g)
  (y * dy.float()).sum().backward(retain_graph=True)
  ln_dx = x.grad.clone()
  for leaf in (x, z, bias, gamma, beta, modt) + ((gate,) if with_gate else ()):
    leaf.grad = None
  ((y * dy.float()).sum() + (x * dx0).sum()).backward()
  xd = x.detach()
  mean = xd.mean(-1)
  rstd = 1.0 / torch.sqrt(xd.var(-1, unbiased=False) + 1e-6)
  dx = dx0.to(DEV).clone()
  dmod = torch.zeros(B, 2 * D, device=DEV)
  dgamma, dbeta, dbias = (torch.zeros(D, device=DEV) for _ in range(3))
  dgate = torch.zeros(B, D, device=DEV)
  dz = torch.empty(rows, D, device=DEV, dtype=torch.bfloat16)
  mg, gg = modt.detach().to(DEV), (gate.detach().to(DEV) if with_gate else None)
  null = C.c_void_p(0)
  # device copies are kept alive in named tensors: a temporary would be freed (and its memory reused) before the launch
  dyg, xg, meang, rstdg = dy.to(DEV), xd.to(DEV), mean.to(DEV), rstd.to(DEV)
  gammag, betag, zg = gamma.detach().to(DEV), beta.detach().to(DEV), z.detach().to(torch.bfloat16).to(DEV)
  lib.check(L.umd_ln_modulate_bwd_gated(
      lib.ptr(dyg), 1, lib.ptr(xg), lib.ptr(meang), lib.ptr(rstdg), lib.ptr(gammag), lib.ptr(betag),
      C.c_void_p(mg.data_ptr() + 4 * D), C.c_longlong(2 * D), n0, s0, n1, s1, D, lib.ptr(dx), 1,
      C.c_void_p(dmod.data_ptr()), C.c_void_p(dmod.data_ptr() + 4 * D), C.c_longlong(2 * D), lib.ptr(dgamma), lib.ptr(dbeta),
      lib.ptr(dz), lib.ptr(zg), lib.ptr(gg) if with_gate else null, C.c_longlong(D),
      lib.ptr(dgate) if with_gate else null, C.c_longlong(D), lib.ptr(dbias), lib.current_stream()), "ln bwd gated")
  torch.cuda.synchronize()
  assert U.rel_l2(dx.cpu(), x.grad) < 1e-4                     # dx_new = dx0 + LayerNorm backward
  assert U.rel_l2(dx.cpu() - dx0, ln_dx) < 1e-3
  assert U.rel_l2(dz.float().cpu(), z.grad) < 5e-3             # bf16 output
  assert U.rel_l2(dbias.cpu(), bias.grad) < 1e-4
  assert U.rel_l2(dgamma.cpu(), gamma.grad) < 1e-4 and U.rel_l2(dbeta.cpu(), beta.grad) < 1e-4
  assert U.rel_l2(dmod.cpu(), modt.grad) < 1e-4
  if with_gate:
    assert U.rel_l2(dgate.cpu(), gate.grad) < 1e-4


# ------------------------------------------------------------------------------------------ attention
def _attn_ref(qkv, n0, s0, n1, s1, H, Dh):
  D = H * Dh
  outs = []
  r = 0
  for n, s in ((n0, s0), (n1, s1)):
    if n == 0:
      continue
    blk = qkv[r:r + n * s].reshape(n, s, 3, H, Dh)
    q, k, v = blk[:, :, 0], blk[:, :, 1], blk[:, :, 2]
    logits = torch.einsum("bqhd,bkhd->bhqk", q / math.sqrt(Dh), k)
    o = torch.einsum("bhqk,bkhd->bqhd", torch.softmax(logits, -1), v)
    outs.append(o.reshape(n * s, D))
    r += n * s
  return torch.cat(outs)


@pytest.mark.parametrize("impl", ["default", "simt"])
@pytest.mark.parametrize("n0,s0,n1,s1,H", [(2, 164, 2, 68, 6), (1, 257, 0, 0, 12), (0, 0, 2, 69, 6), (2, 260, 0, 0, 6),
                                           (3, 165, 1, 258, 16), (1, 16, 1, 1, 2)])
def test_attention_fwd_bwd(n0, s0, n1, s1, H, impl, monkeypatch):
  lib = _lib()
  L = lib.load()
  Dh, D = 64, H * 64
  rows = n0 * s0 + n1 * s1
  g = torch.Generator().manual_seed(rows + H)
  qkv = (torch.randn(rows, 3 * D, generator=g) * 1.2).to(torch.bfloat16)
  dout = torch.randn(rows, D, generator=g).to(torch.bfloat16)
  ref_in = qkv.float().requires_grad_(True)
  ref = _attn_ref(ref_in, n0, s0, n1, s1, H, Dh)
  ref.backward(dout.float())
  fwd = L.umd_attention_fwd_simt if impl == "simt" else L.umd_attention_fwd
  bwd = L.umd_attention_bwd_simt if impl == "simt" else L.umd_attention_bwd
  qg, dg = qkv.to(DEV), dout.to(DEV)
  out = torch.empty(rows, D, device=DEV, dtype=torch.bfloat16)
  lse = torch.empty(rows, H, device=DEV)
  lib.check(fwd(lib.ptr(qg), lib.ptr(out), lib.ptr(lse), n0, s0, n1, s1, H, Dh, lib.current_stream()), "attn fwd")
  assert U.rel_l2(out.float().cpu(), ref.detach()) < 1e-2
  assert torch.allclose(out.float().cpu(), ref.detach(), rtol=2 ** -6, atol=2e-2)
  dqkv = torch.zeros(rows, 3 * D, device=DEV, dtype=torch.bfloat16)
  lib.check(bwd(lib.ptr(qg), lib.ptr(out), lib.ptr(dg), lib.ptr(lse), lib.ptr(dqkv), n0, s0, n1, s1, H, Dh,
                lib.current_stream()), "attn bwd")
  gr = ref_in.grad
  for j, name in enumerate("qkv"):
    r = U.rel_l2(dqkv.float().cpu()[:, j * D:(j + 1) * D], gr[:, j * D:(j + 1) * D])
    assert r < 2e-2, f"d{name} rel-L2 {r}"


@pytest.mark.parametrize("n0,s0,n1,s1,H", [(2, 164, 2, 68, 6), (3, 257, 0, 0, 12), (0, 0, 2, 69, 6), (2, 260, 0, 0, 6),
                                           (3, 165, 1, 258, 16), (1, 16, 1, 1, 2), (40, 257, 64, 68, 12)])
def test_attention_bwd_with_supplied_delta(n0, s0, n1, s1, H):
  """umd_attention_bwd_delta (delta = rowsum(dO o O) per head from the out-projection dgrad epilogue, no O operand)
  gives the gradients of umd_attention_bwd, which forms delta itself from O."""
  lib = _lib()
  L = lib.load()
  Dh, D = 64, H * 64
  rows = n0 * s0 + n1 * s1
  g = torch.Generator().manual_seed(rows + 3 * H)
  qg = (torch.randn(rows, 3 * D, generator=g) * 1.2).to(torch.bfloat16).to(DEV)
  dg = torch.randn(rows, D, generator=g).to(torch.bfloat16).to(DEV)
  out = torch.empty(rows, D, device=DEV, dtype=torch.bfloat16)
  lse = torch.empty(rows, H, device=DEV)
  lib.check(L.umd_attention_fwd(lib.ptr(qg), lib.ptr(out), lib.ptr(lse), n0, s0, n1, s1, H, Dh, lib.current_stream()), "attn fwd")
  d_o = torch.full((rows, 3 * D), float("nan"), device=DEV, dtype=torch.bfloat16)
  lib.check(L.umd_attention_bwd(lib.ptr(qg), lib.ptr(out), lib.ptr(dg), lib.ptr(lse), lib.ptr(d_o), n0, s0, n1, s1, H, Dh,
                                lib.current_stream()), "attn bwd")
  delta = (dg.float() * out.float()).reshape(rows, H, Dh).sum(-1).contiguous()
  d_d = torch.full((rows, 3 * D), float("nan"), device=DEV, dtype=torch.bfloat16)
  lib.check(L.umd_attention_bwd_delta(lib.ptr(qg), lib.ptr(dg), lib.ptr(lse), lib.ptr(delta), lib.ptr(d_d), n0, s0, n1, s1, H,
                                      Dh, lib.current_stream()), "attn bwd delta")
  torch.cuda.synchronize()
  assert torch.isfinite(d_d.float()).all()
  assert U.rel_l2(d_d.float().cpu(), d_o.float().cpu()) < 2e-3
  # and against autograd
  ref_in = qg.float().cpu().requires_grad_(True)
  _attn_ref(ref_in, n0, s0, n1, s1, H, Dh).backward(dg.float().cpu())
  assert U.rel_l2(d_d.float().cpu(), ref_in.grad) < 2e-2


def test_attention_rejects_shapes_outside_the_tensor_core_path():
  """No silent CUDA-core fallback: the product entry points return UMD_ERR_UNSUPPORTED (3)."""
  lib = _lib()
  L = lib.load()
  for (n, s, H, Dh) in ((1, 300, 2, 64), (1, 64, 2, 32)):
    rows, D = n * s, H * Dh
    qkv = torch.zeros(rows, 3 * D, device=DEV, dtype=torch.bfloat16)
    out = torch.zeros(rows, D, device=DEV, dtype=torch.bfloat16)
    lse = torch.zeros(rows, H, device=DEV)
    rc = L.umd_attention_fwd(lib.ptr(qkv), lib.ptr(out), lib.ptr(lse), n, s, 0, 0, H, Dh, lib.current_stream())
    assert rc == 3 and b"unsupported shape" in L.umd_last_error()
    rc = L.umd_attention_bwd(lib.ptr(qkv), lib.ptr(out), lib.ptr(out), lib.ptr(lse), lib.ptr(qkv), n, s, 0, 0, H, Dh,
                             lib.current_stream())
    assert rc == 3
  # the checker kernels stay reachable by name
  n, s, H, Dh = 1, 64, 2, 64
  qkv = torch.randn(n * s, 3 * H * Dh, device=DEV).to(torch.bfloat16)
  out = torch.zeros(n * s, H * Dh, device=DEV, dtype=torch.bfloat16)
  lse = torch.zeros(n * s, H, device=DEV)
  lib.check(L.umd_attention_fwd_simt(lib.ptr(qkv), lib.ptr(out), lib.ptr(lse), n, s, 0, 0, H, Dh, lib.current_stream()), "simt")
  torch.cuda.synchronize()
  assert torch.isfinite(out.float()).all()


@pytest.mark.parametrize("n0,s0,n1,s1,H", [(2, 258, 0, 0, 6), (2, 260, 1, 129, 4), (1, 131, 2, 257, 12), (3, 132, 0, 0, 2)])
def test_attention_tail_rows_on_control_warps(n0, s0, n1, s1, H):
  """Tails of 1..4 rows behind a full 128-row tile computed on the idle control warps, forward and backward (the
  default only routes a 1-row tail there; the hook raises the limit to the kernels' maximum)."""
  lib = _lib()
  L = lib.load()
  L.umd_debug_attn_tail_limits(4, 4)
  try:
    Dh, D = 64, H * 64
    rows = n0 * s0 + n1 * s1
    g = torch.Generator().manual_seed(rows)
    qkv = (torch.randn(rows, 3 * D, generator=g) * 1.2).to(torch.bfloat16)
    dout = torch.randn(rows, D, generator=g).to(torch.bfloat16)
    ref_in = qkv.float().requires_grad_(True)
    ref = _attn_ref(ref_in, n0, s0, n1, s1, H, Dh)
    ref.backward(dout.float())
    qg, dg = qkv.to(DEV), dout.to(DEV)
    out = torch.empty(rows, D, device=DEV, dtype=torch.bfloat16)
    lse = torch.empty(rows, H, device=DEV)
    lib.check(L.umd_attention_fwd(lib.ptr(qg), lib.ptr(out), lib.ptr(lse), n0, s0, n1, s1, H, Dh, lib.current_stream()), "attn fwd")
    assert U.rel_l2(out.float().cpu(), ref.detach()) < 1e-2
    assert torch.allclose(out.float().cpu(), ref.detach(), rtol=2 ** -6, atol=2e-2)
    dqkv = torch.zeros(rows, 3 * D, device=DEV, dtype=torch.bfloat16)
    lib.check(L.umd_attention_bwd(lib.ptr(qg), lib.ptr(out), lib.ptr(dg), lib.ptr(lse), lib.ptr(dqkv), n0, s0, n1, s1, H, Dh,
                                  lib.current_stream()), "attn bwd")
    torch.cuda.synchronize()
    gr = ref_in.grad
    for j, name in enumerate("qkv"):
      r = U.rel_l2(dqkv.float().cpu()[:, j * D:(j + 1) * D], gr[:, j * D:(j + 1) * D])
      assert r < 2e-2, f"d{name} rel-L2 {r}"
    # the tail rows themselves (written by the control warps, not by the TMA-store epilogues)
    for n, s, base in ((n0, s0, 0), (n1, s1, n0 * s0)):
      if n == 0 or s % 128 == 0 or s % 128 > 4:
        continue
      idx = torch.cat([torch.arange(base + k * s + (s // 128) * 128, base + (k + 1) * s) for k in range(n)])
      assert U.rel_l2(out.float().cpu()[idx], ref.detach()[idx]) < 1e-2
      assert U.rel_l2(dqkv.float().cpu()[idx], gr[idx]) < 2e-2
  finally:
    L.umd_debug_attn_tail_limits(-1, -1)


@pytest.mark.parametrize("n0,s0,n1,s1,H", [(40, 257, 64, 68, 12), (30, 260, 0, 0, 12), (28, 164, 0, 0, 12), (0, 0, 52, 129, 6)])
def test_attention_persistent_ctas_walk_over_many_items(n0, s0, n1, s1, H):
  """More (sample, head) items than resident CTAs (148, or 296 for the two-CTA-per-SM variant): every CTA of the
  forward kernel processes several items back to back (barrier phases, TMEM accumulators and shared-memory tiles are
  reused across items); one- and three-tile shapes, with and without a tail row.  Backward on the same inputs."""
  lib = _lib()
  L = lib.load()
  Dh, D = 64, H * 64
  rows = n0 * s0 + n1 * s1
  g = torch.Generator().manual_seed(rows)
  qkv = (torch.randn(rows, 3 * D, generator=g) * 1.2).to(torch.bfloat16)
  dout = torch.randn(rows, D, generator=g).to(torch.bfloat16)
  ref_in = qkv.float().requires_grad_(True)
  ref = _attn_ref(ref_in, n0, s0, n1, s1, H, Dh)
  ref.backward(dout.float())
  qg, dg = qkv.to(DEV), dout.to(DEV)
  out = torch.empty(rows, D, device=DEV, dtype=torch.bfloat16)
  lse = torch.empty(rows, H, device=DEV)
  for _ in range(2):   # twice: the second call starts from warm caches and a different CTA / item timing
    out.zero_()
    lib.check(L.umd_attention_fwd(lib.ptr(qg), lib.ptr(out), lib.ptr(lse), n0, s0, n1, s1, H, Dh, lib.current_stream()), "attn fwd")
    torch.cuda.synchronize()
    assert U.rel_l2(out.float().cpu(), ref.detach()) < 1e-2
    assert torch.allclose(out.float().cpu(), ref.detach(), rtol=2 ** -6, atol=2e-2)
  # lse against the reference's log-sum-exp of the scaled logits
  r = 0
  for n, s in ((n0, s0), (n1, s1)):
    if n == 0:
      continue
    blk = qkv.float()[r:r + n * s].reshape(n, s, 3, H, Dh)
    logits = torch.einsum("bqhd,bkhd->bqhk", blk[:, :, 0] / math.sqrt(Dh), blk[:, :, 1])
    want = torch.logsumexp(logits, -1).reshape(n * s, H)
    assert torch.allclose(lse.cpu()[r:r + n * s], want, rtol=1e-3, atol=2e-2)
    r += n * s
  dqkv = torch.zeros(rows, 3 * D, device=DEV, dtype=torch.bfloat16)
  lib.check(L.umd_attention_bwd(lib.ptr(qg), lib.ptr(out), lib.ptr(dg), lib.ptr(lse), lib.ptr(dqkv), n0, s0, n1, s1, H, Dh,
                                lib.current_stream()), "attn bwd")
  torch.cuda.synchronize()
  for j, name in enumerate("qkv"):
    rr = U.rel_l2(dqkv.float().cpu()[:, j * D:(j + 1) * D], ref_in.grad[:, j * D:(j + 1) * D])
    assert rr < 2e-2, f"d{name} rel-L2 {rr}"


# ------------------------------------------------------------------------------------------ optimiser
@pytest.mark.parametrize("count,clip_active,ema", [(0, True, False), (5, False, True), (200, True, True)])
def test_adamw_step_matches_oracle(count, clip_active, ema):
  """train_ae.py:148-151,365-374 (optax chain, App. A.13) on a synthetic arena with three leaves."""
  lib = _lib()
  g = torch.Generator().manual_seed(count)
  shapes = {("a", "kernel"): (300, 64), ("a", "bias"): (192,), ("cls",): (1, 4, 64)}
  params = {k: torch.randn(s, generator=g) * 0.1 for k, s in shapes.items()}
  gscale = 3.0 if clip_active else 1e-3
  grads = {k: torch.randn(s, generator=g) * gscale for k, s in shapes.items()}
  opt = {"count": count, "mu": {k: (torch.randn(s, generator=g) * 1e-2).to(torch.bfloat16) for k, s in shapes.items()},
         "nu": {k: torch.rand(s, generator=g) * 1e-4 for k, s in shapes.items()}}
  hp = dict(clip_norm=1.0, peak_lr=1e-3, warmup_steps=0, total_steps=10_000, b1=0.9, b2=0.95, wd=0.05)
  ptree, gtree = O.unflatten_tree(params), O.unflatten_tree(grads)
  new_p, new_opt, upd, gnorm = O.optimizer_update(gtree, opt, ptree, hp)
  # arena: each leaf padded to a multiple of 64
  offs, total = {}, 0
  for k, s in shapes.items():
    offs[k] = total
    total += (math.prod(s) + 63) // 64 * 64
  def pack(d, dtype=torch.float32):
    a = torch.zeros(total, dtype=dtype)
    for k, v in d.items():
      a[offs[k]:offs[k] + v.numel()] = v.reshape(-1).to(dtype)
    return a.to(DEV)
  P, G, MU, NU = pack(params), pack(grads), pack(opt["mu"], torch.bfloat16), pack(opt["nu"])
  flags = torch.zeros(total // 64, dtype=torch.uint8)
  flags[offs[("a", "kernel")] // 64:(offs[("a", "kernel")] + 300 * 64) // 64] = 1
  flags = flags.to(DEV)
  ema_t = pack(params) * 0.5 if ema else None
  ema0 = ema_t.clone() if ema else None
  shadow = torch.zeros(total, dtype=torch.bfloat16, device=DEV)
  scratch = torch.empty(4096, device=DEV)
  meas = torch.zeros(4, device=DEV)
  a = lib.AdamwArgs()
  a.params, a.grads, a.mu, a.nu = lib.ptr(P), lib.ptr(G), lib.ptr(MU), lib.ptr(NU)
  a.params_bf16, a.ema, a.wd_flags, a.n = lib.ptr(shadow), lib.ptr(ema_t), lib.ptr(flags), total
  lr = O.warmup_cosine_lr(count, peak=hp["peak_lr"], warmup_steps=0, decay_steps=hp["total_steps"])
  a.clip_norm, a.lr, a.b1, a.b2, a.eps, a.wd = 1.0, lr, 0.9, 0.95, 1e-8, 0.05
  a.bias_corr1, a.bias_corr2 = 1 - 0.9 ** (count + 1), 1 - 0.95 ** (count + 1)
  a.ema_decay = 0.01
  a.scratch, a.scratch_floats, a.measurements = lib.ptr(scratch), 4096, lib.ptr(meas)
  lib.check(lib.load().umd_adamw_step(C.byref(a), lib.current_stream()), "adamw")
  torch.cuda.synchronize()
  fp, fmu, fnu = O.flatten_tree(new_p), new_opt["mu"], new_opt["nu"]
  for k, s in shapes.items():
    n = math.prod(s)
    sl = slice(offs[k], offs[k] + n)
    # the update itself (p_new - p_old) to fp32 round-off of the chain
    du = (P[sl].cpu() - params[k].reshape(-1))
    assert U.rel_l2(du, upd[k].reshape(-1)) < 2e-4, k
    assert torch.allclose(P[sl].cpu(), fp[k].reshape(-1), rtol=1e-6, atol=1e-7), k
    assert torch.allclose(NU[sl].cpu(), fnu[k].reshape(-1), rtol=1e-5, atol=1e-12), k
    # bf16 mu: identical up to one bf16 ulp where the fp32 value sits on a rounding boundary
    assert torch.allclose(MU[sl].float().cpu(), fmu[k].float().reshape(-1), rtol=2 ** -7, atol=1e-8), k
    assert torch.equal(shadow[sl].cpu(), P[sl].cpu().to(torch.bfloat16)), k
    if ema:
      ref_e = 0.01 * fp[k].reshape(-1) + 0.99 * ema0[sl].cpu()
      assert torch.allclose(ema_t[sl].cpu(), ref_e, rtol=1e-6, atol=1e-8), k
  l2p = math.sqrt(sum(float(v.double().pow(2).sum()) for v in fp.values()))
  l2u = math.sqrt(sum(float(v.double().pow(2).sum()) for v in upd.values()))
  assert abs(float(meas[0]) - l2p) <= 1e-5 * l2p
  assert abs(float(meas[1]) - l2u) <= 1e-4 * l2u
  assert abs(float(meas[2]) - gnorm) <= 1e-5 * gnorm
