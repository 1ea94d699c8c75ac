"""Debug aid: per-row comparison of the tcgen05 attention kernels against the CUDA-core checker kernels."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from small_vision_b200 import lib

def run(n, S, H=2, seed=0):
  L = lib.load(); Dh = 64; D = H * Dh; rows = n * S
  g = torch.Generator().manual_seed(seed)
  qkv = (torch.randn(rows, 3 * D, generator=g) * 1.2).to(torch.bfloat16).cuda()
  dout = torch.randn(rows, D, generator=g).to(torch.bfloat16).cuda()
  res = {}
  for name, f, b in (("tc", L.umd_attention_fwd, L.umd_attention_bwd), ("simt", L.umd_attention_fwd_simt, L.umd_attention_bwd_simt)):
    out = torch.zeros(rows, D, device="cuda", dtype=torch.bfloat16); lse = torch.zeros(rows, H, device="cuda")
    dqkv = torch.zeros(rows, 3 * D, device="cuda", dtype=torch.bfloat16)
    st = lib.current_stream()
    lib.check(f(lib.ptr(qkv), lib.ptr(out), lib.ptr(lse), n, S, 0, 0, H, Dh, st))
    lib.check(b(lib.ptr(qkv), lib.ptr(out), lib.ptr(dout), lib.ptr(lse), lib.ptr(dqkv), n, S, 0, 0, H, Dh, st))
    torch.cuda.synchronize()
    res[name] = (out.float().cpu(), lse.cpu(), dqkv.float().cpu())
  for k, nm in enumerate(("out", "lse", "dqkv")):
    a, b = res["tc"][k], res["simt"][k]
    err = (a - b).norm(dim=1) / (b.norm(dim=1) + 1e-9)
    bad = (err > 0.03).nonzero().flatten().tolist()
    print(f"n={n} S={S} {nm}: max row rel err {float(err.max()):.4f}; bad rows {bad[:12]} ({len(bad)})")
    if nm == "dqkv" and bad:
      for j, part in enumerate("qkv"):
        e2 = (a[:, j*D:(j+1)*D] - b[:, j*D:(j+1)*D]).norm(dim=1) / (b[:, j*D:(j+1)*D].norm(dim=1) + 1e-9)
        bb = (e2 > 0.03).nonzero().flatten().tolist()
        print(f"    d{part}: bad rows {bb[:12]} ({len(bb)})")

for n, S in ((1, 257), (2, 257), (1, 260), (1, 256), (1, 272), (1, 164), (2, 69), (1, 129)):
  run(n, S)
