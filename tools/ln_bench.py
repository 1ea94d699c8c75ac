"""Micro-benchmark of the LayerNorm + modulate kernels through the C ABI (no fused residual / gate stage) at several
row counts: GB/s over the algorithmic bytes (forward: fp32 row in, bf16 row out; backward: dy bf16 + x fp32 in,
dx fp32 read-modify-write)."""
import os
import sys

import ctypes as C

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from small_vision_b200 import lib  # noqa: E402


def timeit(fn, iters=10):
  for _ in range(3):
    fn()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(iters):
    fn()
  e1.record()
  torch.cuda.synchronize()
  return e0.elapsed_time(e1) / iters


def main():
  L = lib.load()
  D = 768
  st = lib.current_stream()
  for n, S in ((256, 232), (512, 257), (2048, 257)):
    rows = n * S
    x = torch.randn(rows, D, device="cuda")
    gamma, beta = torch.randn(D, device="cuda"), torch.randn(D, device="cuda")
    mod = torch.randn(n, 2 * D, device="cuda") * 0.1
    out = torch.empty(rows, D, device="cuda", dtype=torch.bfloat16)
    mean, rstd = torch.empty(rows, device="cuda"), torch.empty(rows, device="cuda")
    f = lambda: lib.check(L.umd_ln_modulate_fwd(lib.ptr(x), lib.ptr(gamma), lib.ptr(beta), lib.ptr(mod), lib.ptr(mod[:, D:]),
                                                 C.c_longlong(2 * D), n, S, 0, 0, D, lib.ptr(out), 1, lib.ptr(mean), lib.ptr(rstd), st))
    tf = timeit(f)
    dy = torch.randn(rows, D, device="cuda").to(torch.bfloat16)
    dx = torch.zeros(rows, D, device="cuda")
    dmod = torch.zeros(n, 2 * D, device="cuda")
    dg, db = torch.zeros(D, device="cuda"), torch.zeros(D, device="cuda")
    b = lambda: lib.check(L.umd_ln_modulate_bwd(lib.ptr(dy), 1, lib.ptr(x), lib.ptr(mean), lib.ptr(rstd), lib.ptr(gamma), lib.ptr(beta),
                                                 lib.ptr(mod[:, D:]), C.c_longlong(2 * D), n, S, 0, 0, D, lib.ptr(dx), 1, lib.ptr(dmod),
                                                 lib.ptr(dmod[:, D:]), C.c_longlong(2 * D), lib.ptr(dg), lib.ptr(db), st))
    tb = timeit(b)
    z = torch.randn(rows, D, device="cuda").to(torch.bfloat16)
    dz = torch.empty_like(z)
    gate = torch.randn(n, D, device="cuda")
    dgate, dbias = torch.zeros(n, D, device="cuda"), torch.zeros(D, device="cuda")
    bg = lambda: lib.check(L.umd_ln_modulate_bwd_gated(
        lib.ptr(dy), 1, lib.ptr(x), lib.ptr(mean), lib.ptr(rstd), lib.ptr(gamma), lib.ptr(beta), lib.ptr(mod[:, D:]), C.c_longlong(2 * D),
        n, S, 0, 0, D, lib.ptr(dx), 1, lib.ptr(dmod), lib.ptr(dmod[:, D:]), C.c_longlong(2 * D), lib.ptr(dg), lib.ptr(db), lib.ptr(dz),
        lib.ptr(z), lib.ptr(gate), C.c_longlong(D), lib.ptr(dgate), C.c_longlong(D), lib.ptr(dbias), st))
    tg = timeit(bg)
    print(f"rows {rows:7d}: fwd {tf * 1e3:7.1f} us {rows * D * 6 / tf / 1e6:7.1f} GB/s | bwd {tb * 1e3:7.1f} us {rows * D * 14 / tb / 1e6:7.1f} GB/s"
          f" | bwd + gate stage {tg * 1e3:7.1f} us {rows * D * 18 / tg / 1e6:7.1f} GB/s")


if __name__ == "__main__":
  main()
