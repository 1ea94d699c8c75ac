"""Bitwise repeatability of the attention kernels: same inputs, repeated launches; reports how many launches differ
from the first one in out / lse (forward) and dqkv (backward)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from small_vision_b200 import lib

def run(n0, s0, n1, s1, H=12, reps=20):
  L = lib.load(); Dh = 64; D = H * Dh; rows = n0 * s0 + n1 * s1
  g = torch.Generator(device="cuda").manual_seed(rows)
  qkv = (torch.randn(rows, 3 * D, device="cuda", generator=g) * 1.2).to(torch.bfloat16)
  dout = torch.randn(rows, D, device="cuda", generator=g).to(torch.bfloat16)
  st = lib.current_stream()
  outs, lses, dqs = [], [], []
  for i in range(reps):
    out = torch.zeros(rows, D, device="cuda", dtype=torch.bfloat16); lse = torch.zeros(rows, H, device="cuda")
    lib.check(L.umd_attention_fwd(lib.ptr(qkv), lib.ptr(out), lib.ptr(lse), n0, s0, n1, s1, H, Dh, st))
    dq = torch.zeros(rows, 3 * D, device="cuda", dtype=torch.bfloat16)
    lib.check(L.umd_attention_bwd(lib.ptr(qkv), lib.ptr(outs[0] if outs else out), lib.ptr(dout), lib.ptr(lses[0] if lses else lse), lib.ptr(dq), n0, s0, n1, s1, H, Dh, st))
    torch.cuda.synchronize()
    outs.append(out); lses.append(lse); dqs.append(dq)
  do = sum(int(not torch.equal(outs[0], o)) for o in outs[1:])
  dl = sum(int(not torch.equal(lses[0], o)) for o in lses[1:])
  dd = sum(int(not torch.equal(dqs[0], o)) for o in dqs[1:])
  bad = ""
  for o in lses[1:]:
    if not torch.equal(lses[0], o):
      idx = (lses[0] != o).nonzero()
      bad = f" first lse mismatch rows {idx[:4].tolist()} vals {lses[0][idx[0,0], idx[0,1]].item()} vs {o[idx[0,0], idx[0,1]].item()}"
      break
  for o in dqs[1:]:
    if not torch.equal(dqs[0], o) and not bad:
      idx = (dqs[0] != o).nonzero()
      bad = f" dqkv mismatches {idx.shape[0]} first {idx[:3].tolist()}"
      break
  print(f"n0={n0} s0={s0} n1={n1} s1={s1}: differing launches out {do} lse {dl} dqkv {dd} of {reps - 1}{bad}")

for shape in ((16, 257, 0, 0), (8, 164, 8, 68), (512, 257, 0, 0), (256, 164, 256, 68)):
  run(*shape)
