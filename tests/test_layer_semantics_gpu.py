"""GPU-side pins of the Flax layer semantics that the oracle and the refshim both restate (a shared misreading
of ConvTranspose orientation, Conv orientation or the LayerNorm epsilon would be invisible to an oracle-vs-engine test).
The checker here is torch.nn.functional alone — conv2d, conv_transpose2d, layer_norm — on asymmetric kernels and
small-variance rows; oracle/umd_oracle.py is not involved.

With the reference's own initialisation (adaLN Dense_0 kernel and bias zero, vit.py:71; final_modulation zero, ae.py:94)
every block is the identity, so the whole model reduces to  patch-embed conv -> +pos -> [cls | tokens] -> LayerNorm ->
rep = mean(cls outputs) -> [rep | tokens + dec_pos] -> LayerNorm -> ConvTranspose  (SURVEY.md §8c pin 1): exactly the
layers whose orientation / epsilon conventions are at stake.
"""
import pytest
import torch
import torch.nn.functional as F

from tests import util as U

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("variant,img,ch", [("S/4", 64, 3), ("S/2", 32, 4), ("B/4", 64, 3)])
def test_identity_blocks_reduce_to_torch_conv_layernorm_convtranspose(variant, img, ch):
  from small_vision_b200.params import init_arena, tree_from_arena
  model, _ = U.make_models(variant, adaln=True, img_size=img, channels=ch, depth=2, dec_depth=1)
  cfg = model.cfg
  arena = init_arena(model.layout, 0, "cpu", nonzero_adaln=False)     # the reference's zero-init: blocks are identities
  p = tree_from_arena(model.layout, arena)
  g = torch.Generator().manual_seed(1)
  ps, D, C = cfg.patch, cfg.width, cfg.channels
  # asymmetric kernels (no spatial or channel symmetry), non-trivial LayerNorm affine, non-zero cls / biases
  p["embedding"]["kernel"].copy_(torch.randn(ps, ps, C, D, generator=g) * 0.2)
  p["embedding"]["bias"].copy_(torch.randn(D, generator=g) * 0.1)
  p["final_conv"]["kernel"].copy_(torch.randn(ps, ps, D, 2 * C, generator=g) * 0.05)
  p["final_conv"]["bias"].copy_(torch.randn(2 * C, generator=g) * 0.1)
  p["cls"].copy_(torch.randn(1, cfg.num_cls, D, generator=g) * 0.3)
  for nm in ("Encoder", "Decoder"):
    p[nm]["encoder_norm"]["scale"].copy_(1 + 0.2 * torch.randn(D, generator=g))
    p[nm]["encoder_norm"]["bias"].copy_(0.1 * torch.randn(D, generator=g))
  n = 3
  image = torch.rand(n, img, img, C, generator=g) * 2 - 1
  t = torch.randint(1, 1000, (n, 1), generator=g, dtype=torch.int32)
  pred, out = model.apply({"params": tree_from_arena(model.layout, arena.to(DEV))}, image.to(DEV), t=t.to(DEV))
  torch.cuda.synchronize()
  # ---- torch.nn.functional restatement (fp64)
  f = lambda a: a.double()
  x = F.conv2d(f(image).permute(0, 3, 1, 2), f(p["embedding"]["kernel"]).permute(3, 2, 0, 1), f(p["embedding"]["bias"]), stride=ps)
  h = img // ps
  tok = x.permute(0, 2, 3, 1).reshape(n, h * h, D) + f(p["pos_embedding"])
  seq = torch.cat([f(p["cls"]).expand(n, -1, -1), tok], 1)
  enc = F.layer_norm(seq, (D,), f(p["Encoder"]["encoder_norm"]["scale"]), f(p["Encoder"]["encoder_norm"]["bias"]), eps=1e-6)
  rep = enc[:, :cfg.num_cls].mean(1)
  xd = torch.cat([rep[:, None], enc[:, cfg.num_cls:] + f(p["dec_pos_embedding"])], 1)
  xd = F.layer_norm(xd, (D,), f(p["Decoder"]["encoder_norm"]["scale"]), f(p["Decoder"]["encoder_norm"]["bias"]), eps=1e-6)[:, 1:]
  # flax ConvTranspose(transpose_kernel=False) = torch conv_transpose2d with the spatially flipped kernel
  w = f(p["final_conv"]["kernel"]).flip(0, 1).permute(2, 3, 0, 1)
  want = F.conv_transpose2d(xd.reshape(n, h, h, D).permute(0, 3, 1, 2), w, f(p["final_conv"]["bias"]), stride=ps).permute(0, 2, 3, 1)
  r = U.rel_l2(pred.cpu(), want)
  assert r <= 1e-2, f"pred rel-L2 {r}"
  assert U.rel_l2(out["pre_logits"].cpu(), rep) <= 1e-2
  # the other orientation is far away: an un-flipped kernel must NOT explain the output
  w_bad = f(p["final_conv"]["kernel"]).permute(2, 3, 0, 1)
  bad = F.conv_transpose2d(xd.reshape(n, h, h, D).permute(0, 3, 1, 2), w_bad, f(p["final_conv"]["bias"]), stride=ps).permute(0, 2, 3, 1)
  assert U.rel_l2(pred.cpu(), bad) > 0.3


@pytest.mark.parametrize("D", [384, 768, 1024])
def test_layernorm_epsilon_is_1e_6(D):
  """Rows whose variance is of the order of the epsilon: eps = 1e-6 (flax default) and eps = 1e-5 (torch default) differ by
  tens of percent there.  The mean is kept small so that the fast-variance form E[x^2] - E[x]^2 stays well conditioned."""
  import ctypes as C
  from small_vision_b200 import lib
  L = lib.load()
  g = torch.Generator().manual_seed(D)
  rows = 64
  x = (0.05 + 2e-3 * torch.randn(rows, D, generator=g)).float()          # variance 4e-6
  gamma, beta = 1 + 0.1 * torch.randn(D, generator=g), 0.1 * torch.randn(D, generator=g)
  out = torch.empty(rows, D, device=DEV)
  mean, rstd = torch.empty(rows, device=DEV), torch.empty(rows, device=DEV)
  xg, gg, bg = x.to(DEV), gamma.to(DEV), beta.to(DEV)      # named: the buffers must outlive the asynchronous launch
  lib.check(L.umd_ln_modulate_fwd(lib.ptr(xg), lib.ptr(gg), lib.ptr(bg), None, None, C.c_longlong(0),
                                  C.c_int(1), C.c_int(rows), C.c_int(0), C.c_int(0), C.c_int(D), lib.ptr(out), C.c_int(0),
                                  lib.ptr(mean), lib.ptr(rstd), lib.current_stream()), "ln fwd")
  torch.cuda.synchronize()
  want = F.layer_norm(x.double(), (D,), gamma.double(), beta.double(), eps=1e-6)
  wrong = F.layer_norm(x.double(), (D,), gamma.double(), beta.double(), eps=1e-5)
  assert U.rel_l2(out.cpu(), want) <= 2e-3
  assert U.rel_l2(out.cpu(), wrong) >= 0.1
