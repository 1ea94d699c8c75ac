"""Micro-benchmark of the tcgen05 GEMM at the UMD-B/4 step shapes.
Prints achieved TFLOP/s per shape against MEASURED_PEAKS.json."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from small_vision_b200 import lib  # noqa: E402


def bench(fn, iters=20, warm=3):
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
  peak = 1685.0
  p = os.path.join(ROOT, "MEASURED_PEAKS.json")
  if os.path.exists(p):
    peak = json.load(open(p))["bf16_tflops"]
  rows = []
  Me, Md = 59392, 131584
  shapes = [("qkv_enc", Me, 2304, 768), ("out_enc", Me, 768, 768), ("fc1_enc", Me, 3072, 768), ("fc2_enc", Me, 768, 3072),
            ("qkv_dec", Md, 2304, 768), ("fc1_dec", Md, 3072, 768), ("fc2_dec", Md, 768, 3072)]
  for name, M, N, K in shapes:
    X = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    W = (torch.randn(K, N, device="cuda") * K ** -0.5).to(torch.bfloat16)
    Wt = W.t().contiguous()
    Y = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    b = torch.zeros(N, device="cuda")
    ms = bench(lambda: lib.gemm(X, W, M=M, N=N, K=K, b_mn=True, epi=lib.EPI_BF16, out0=Y, bias=b))
    tf = 2.0 * M * N * K / ms / 1e9
    rows.append((name + "_fwd(K,MN)", M, N, K, ms, tf))
    ms = bench(lambda: lib.gemm(X, Wt, M=M, N=N, K=K, b_mn=False, epi=lib.EPI_BF16, out0=Y))
    tf = 2.0 * M * N * K / ms / 1e9
    rows.append((name + "_dgrad(K,K)", M, N, K, ms, tf))
    dW = torch.zeros(K, N, device="cuda")
    ms = bench(lambda: lib.gemm(X, Y, M=K, N=N, K=M, a_mn=True, b_mn=True, epi=lib.EPI_ATOMIC, split_k=8, out0=dW))
    tf = 2.0 * M * N * K / ms / 1e9
    rows.append((name + "_wgrad(MN,MN)", K, N, M, ms, tf))
    # cuBLAS for context (not on the product path)
    ms = bench(lambda: torch.matmul(X, W, out=Y))
    rows.append((name + "_cublas", M, N, K, ms, 2.0 * M * N * K / ms / 1e9))
  print(f"{'shape':28s} {'M':>7s} {'N':>6s} {'K':>7s} {'ms':>8s} {'TF/s':>8s} {'frac':>6s}")
  for name, M, N, K, ms, tf in rows:
    print(f"{name:28s} {M:7d} {N:6d} {K:7d} {ms:8.3f} {tf:8.1f} {tf / peak:6.3f}")


if __name__ == "__main__":
  main()
