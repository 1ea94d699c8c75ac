"""Micro-benchmark of the tcgen05 attention kernels at the UMD-B/4 step shapes (decoder 512 x 257, encoder
256 x 164 + 256 x 68; 12 heads of 64).  Prints ms, algorithmic TFLOP/s and GB/s per call."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from small_vision_b200 import lib  # noqa: E402


def bench(fn, iters, warm=2):
  for _ in range(warm):
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
  iters = int(sys.argv[1]) if len(sys.argv) > 1 else 10
  L = lib.load()
  H, Dh = 12, 64
  D = H * Dh
  for name, n0, s0, n1, s1 in (("decoder", 512, 257, 0, 0), ("encoder", 256, 164, 256, 68), ("enc_noise", 256, 164, 0, 0),
                               ("enc_mae", 0, 0, 256, 68), ("dit_enc", 256, 260, 0, 0), ("mae_dec", 512, 258, 0, 0)):
    rows = n0 * s0 + n1 * s1
    qkv = torch.randn(rows, 3 * D, device="cuda").to(torch.bfloat16)
    dout = torch.randn(rows, D, device="cuda").to(torch.bfloat16)
    out = torch.empty(rows, D, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(rows, H, device="cuda")
    dqkv = torch.empty(rows, 3 * D, device="cuda", dtype=torch.bfloat16)
    st = lib.current_stream()
    f = lambda: lib.check(L.umd_attention_fwd(lib.ptr(qkv), lib.ptr(out), lib.ptr(lse), n0, s0, n1, s1, H, Dh, st))
    b = lambda: lib.check(L.umd_attention_bwd(lib.ptr(qkv), lib.ptr(out), lib.ptr(dout), lib.ptr(lse), lib.ptr(dqkv), n0, s0,
                                              n1, s1, H, Dh, st))
    delta = torch.randn(rows, H, device="cuda")
    bd = lambda: lib.check(L.umd_attention_bwd_delta(lib.ptr(qkv), lib.ptr(dout), lib.ptr(lse), lib.ptr(delta), lib.ptr(dqkv),
                                                     n0, s0, n1, s1, H, Dh, st))
    flops = 4.0 * Dh * H * (n0 * s0 * s0 + n1 * s1 * s1)
    tf, tb, td = bench(f, iters), bench(b, iters), bench(bd, iters)
    by_f = rows * D * 2 * 4 + rows * H * 4            # q, k, v in; o, lse out
    by_b = rows * D * 2 * 8 + rows * H * 4            # q, k, v, o, dO, lse in; dq, dk, dv out
    by_d = rows * D * 2 * 7 + rows * H * 8            # ... with delta [rows, H] instead of o
    print(f"{name:10s} fwd {tf:7.3f} ms {flops / tf / 1e9:7.1f} TF/s {by_f / tf / 1e6:7.1f} GB/s | "
          f"bwd {tb:7.3f} ms {2.5 * flops / tb / 1e9:7.1f} TF/s {by_b / tb / 1e6:7.1f} GB/s | "
          f"bwd(delta) {td:7.3f} ms {by_d / td / 1e6:7.1f} GB/s")


if __name__ == "__main__":
  main()
