"""ctypes binding of libumd_b200.so (the C ABI declared in include/umd_b200.h).

The product path never falls back to PyTorch or to the oracle: if the shared library is
missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libumd_b200.so")


class UmdError(RuntimeError):
  pass


class GemmArgs(C.Structure):
  _fields_ = [
      ("A", C.c_void_p), ("B", C.c_void_p),
      ("M", C.c_int), ("N", C.c_int), ("K", C.c_int), ("batch", C.c_int),
      ("a_mn", C.c_int), ("b_mn", C.c_int),
      ("lda", C.c_longlong), ("ldb", C.c_longlong),
      ("a_bs", C.c_longlong), ("b_bs", C.c_longlong),
      ("epi", C.c_int), ("split_k", C.c_int),
      ("out0", C.c_void_p), ("ld0", C.c_longlong), ("bs0", C.c_longlong),
      ("out1", C.c_void_p), ("ld1", C.c_longlong),
      ("bias", C.c_void_p), ("bias_bs", C.c_longlong),
      ("aux", C.c_void_p), ("ldaux", C.c_longlong),
      ("gate", C.c_void_p), ("ldgate", C.c_longlong),
      ("split_row", C.c_int), ("s0", C.c_int), ("s1", C.c_int), ("n0", C.c_int),
      ("b_kchunk", C.c_int),
  ]


EPI_BF16, EPI_F32, EPI_GELU, EPI_GATE_RES, EPI_DGELU, EPI_ATOMIC, EPI_BF16_DELTA = range(7)

_lib = None


def load():
  """Loads the shared library once; raises UmdError when it has not been built."""
  global _lib
  if _lib is not None:
    return _lib
  if not os.path.exists(LIB_PATH):
    raise UmdError(
        f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(there is no CPU or PyTorch fallback for the UMD hot path)")
  lib = C.CDLL(LIB_PATH)
  lib.umd_last_error.restype = C.c_char_p
  lib.umd_launch_count.restype = C.c_longlong
  _lib = lib
  return lib


def check(rc, what=""):
  if rc != 0:
    msg = load().umd_last_error().decode("utf-8", "replace")
    raise UmdError(f"{what} failed (code {rc}): {msg}")


def ptr(t):
  return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def current_stream():
  import torch
  return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def launch_count():
  return int(load().umd_launch_count())


def gemm(A, B, *, M, N, K, a_mn=False, b_mn=False, lda=None, ldb=None, batch=1, a_bs=0, b_bs=0,
         epi=EPI_BF16, split_k=1, out0=None, ld0=0, bs0=0, out1=None, ld1=0, bias=None, bias_bs=0,
         aux=None, ldaux=0, gate=None, ldgate=0, rowmap=None, b_kchunk=0):
  """Thin wrapper over umd_gemm_bf16 for tests and host-side orchestration."""
  lib = load()
  a = GemmArgs()
  a.A, a.B = ptr(A), ptr(B)
  a.M, a.N, a.K, a.batch = M, N, K, batch
  a.a_mn, a.b_mn = int(a_mn), int(b_mn)
  a.lda = lda if lda is not None else (M if a_mn else K)
  a.ldb = ldb if ldb is not None else (N if b_mn else K)
  a.a_bs, a.b_bs = a_bs, b_bs
  a.epi, a.split_k = epi, split_k
  a.out0, a.ld0, a.bs0 = ptr(out0), ld0 or N, bs0
  a.out1, a.ld1 = ptr(out1), ld1 or N
  a.bias, a.bias_bs = ptr(bias), bias_bs
  a.aux, a.ldaux = ptr(aux), ldaux or N
  a.gate, a.ldgate = ptr(gate), ldgate
  if rowmap is None:
    rowmap = (M, M, 1, 1)
  a.split_row, a.s0, a.s1, a.n0 = rowmap
  a.b_kchunk = b_kchunk
  check(lib.umd_gemm_bf16(C.byref(a), current_stream()), "umd_gemm_bf16")


class ModelCfg(C.Structure):
  _fields_ = [(k, C.c_int) for k in ("img_size", "patch", "channels", "width", "depth", "dec_depth", "heads", "mlp_dim",
                                      "num_cls", "num_classes", "adaln", "flip_final_conv", "residual_bf16")]


class StepShape(C.Structure):
  _fields_ = [(k, C.c_int) for k in ("n0", "n1", "keep0", "keep1", "masked0", "masked1")]


class IO(C.Structure):
  _fields_ = [(k, C.c_void_p) for k in ("image", "t", "labels", "ids_shuffle", "ids_restore", "x0", "noise", "pred",
                                        "pre_logits", "loss")]


class AdamwArgs(C.Structure):
  _fields_ = [
      ("params", C.c_void_p), ("grads", C.c_void_p), ("mu", C.c_void_p), ("nu", C.c_void_p),
      ("params_bf16", C.c_void_p), ("ema", C.c_void_p), ("wd_flags", C.c_void_p), ("n", C.c_longlong),
      ("clip_norm", C.c_float), ("lr", C.c_float), ("b1", C.c_float), ("b2", C.c_float), ("eps", C.c_float),
      ("wd", C.c_float), ("bias_corr1", C.c_float), ("bias_corr2", C.c_float), ("ema_decay", C.c_float),
      ("scratch", C.c_void_p), ("scratch_floats", C.c_int), ("measurements", C.c_void_p),
  ]


BUCKET_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_int)

UMD_STEP_NO_OPTIMIZER = 1


class TrainStepArgs(C.Structure):
  """umd_train_step_args (include/umd_b200.h)."""
  _fields_ = [
      ("cfg", C.POINTER(ModelCfg)), ("shape", StepShape), ("offsets", C.c_void_p), ("opt", AdamwArgs), ("grads", C.c_void_p),
      ("image", C.c_void_p), ("label", C.c_void_p), ("use_labels", C.c_int), ("t", C.c_void_p), ("noise", C.c_void_p),
      ("mask_noise0", C.c_void_p), ("mask_noise1", C.c_void_p), ("label_drop", C.c_void_p),
      ("sqrt_alphas_cumprod", C.c_void_p), ("sqrt_one_minus_alphas_cumprod", C.c_void_p),
      ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t), ("comm", C.c_void_p),
      ("bucket_bounds", C.c_void_p), ("bucket_events", C.c_void_p), ("num_buckets", C.c_int), ("flags", C.c_int),
  ]


def train_workspace_bytes(mcfg, shape):
  lib = load()
  lib.umd_train_workspace_bytes.restype = C.c_size_t
  n = lib.umd_train_workspace_bytes(C.byref(mcfg), C.byref(shape))
  if n == 0:
    raise UmdError("umd_train_workspace_bytes: " + lib.umd_last_error().decode())
  return int(n)


def model_cfg_struct(cfg):
  m = ModelCfg()
  m.img_size, m.patch, m.channels = cfg.img_size, cfg.patch, cfg.channels
  m.width, m.depth, m.dec_depth, m.heads, m.mlp_dim = cfg.width, cfg.depth, cfg.dec_depth, cfg.num_heads, cfg.mlp
  m.num_cls = cfg.num_cls
  m.num_classes = cfg.num_classes or 0
  m.adaln = int(cfg.adaln)
  m.flip_final_conv = int(cfg.flip_final_conv)
  m.residual_bf16 = (2 if cfg.grad_stream_dtype == "bfloat16" else 1) if cfg.residual_dtype == "bfloat16" else 0
  return m


def workspace_bytes(mcfg, shape, train):
  lib = load()
  lib.umd_workspace_bytes.restype = C.c_size_t
  n = lib.umd_workspace_bytes(C.byref(mcfg), C.byref(shape), C.c_int(int(train)))
  if n == 0:
    raise UmdError("umd_workspace_bytes: " + lib.umd_last_error().decode())
  return int(n)


def cast_bf16(src, dst):
  lib = load()
  check(lib.umd_cast_f32_to_bf16(ptr(src), C.c_longlong(src.numel()), ptr(dst), current_stream()), "umd_cast_f32_to_bf16")
