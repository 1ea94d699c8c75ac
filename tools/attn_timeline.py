"""Debug aid: clock64() timeline of CTA 0 of the attention backward kernel (cycles relative to the softmax warps' start)."""
import ctypes as C, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from small_vision_b200 import lib

def run(n, S, H=12):
  L = lib.load(); Dh = 64; D = H * Dh; rows = n * S
  qkv = torch.randn(rows, 3 * D, device="cuda").to(torch.bfloat16)
  dout = torch.randn(rows, D, device="cuda").to(torch.bfloat16)
  out = torch.empty(rows, D, device="cuda", dtype=torch.bfloat16); lse = torch.empty(rows, H, device="cuda")
  dqkv = torch.empty(rows, 3 * D, device="cuda", dtype=torch.bfloat16)
  st = lib.current_stream()
  lib.check(L.umd_attention_fwd(lib.ptr(qkv), lib.ptr(out), lib.ptr(lse), n, S, 0, 0, H, Dh, st))
  tl = torch.zeros(11 * 64, dtype=torch.int64, device="cuda")
  for it in range(2):
    L.umd_debug_attn_timeline(C.c_void_p(tl.data_ptr()))
    delta = (dout.float() * out.float()).reshape(rows, H, Dh).sum(-1).contiguous()
    lib.check(L.umd_attention_bwd_delta(lib.ptr(qkv), lib.ptr(dout), lib.ptr(lse), lib.ptr(delta), lib.ptr(dqkv), n, S, 0, 0, H, Dh, st))
    torch.cuda.synchronize()
  L.umd_debug_attn_timeline(None)
  t = tl.cpu().reshape(11, 64)
  t0 = int(t[4, 0])
  names = ["mma:loaded", "mma:scores_issued[step]", "mma:p_ready[blk]", "mma:acc_issued[blk]", "sm:start/prologue_done", "sm:s_ready[step]",
           "sm:ld_done[step]", "sm:tile_free[step]", "sm:stored[step]", "sm:acc_ready[j]", "sm:dq_start/end, then before_fence[step]"]
  print(f"--- n={n} S={S}")
  for k, nm in enumerate(names):
    vals = [int(v) - t0 for v in t[k] if int(v) != 0]
    print(f"{nm:28s}", vals)

def run_fwd(n, S, H=12):
  L = lib.load(); Dh = 64; D = H * Dh; rows = n * S
  qkv = torch.randn(rows, 3 * D, device="cuda").to(torch.bfloat16)
  out = torch.empty(rows, D, device="cuda", dtype=torch.bfloat16); lse = torch.empty(rows, H, device="cuda")
  st = lib.current_stream()
  tl = torch.zeros(11 * 64, dtype=torch.int64, device="cuda")
  for it in range(2):
    L.umd_debug_attn_timeline(C.c_void_p(tl.data_ptr()))
    lib.check(L.umd_attention_fwd(lib.ptr(qkv), lib.ptr(out), lib.ptr(lse), n, S, 0, 0, H, Dh, st))
    torch.cuda.synchronize()
  L.umd_debug_attn_timeline(None)
  t = tl.cpu().reshape(11, 64)
  t0 = int(t[4, 0])
  names = ["mma:loaded", "mma:tile_start[i]", "mma:qk_issued[i]", "mma:p_ready[i]", "sm:start", "sm:s_ready[i]", "sm:max_done[i]",
           "sm:p_stored[i]", "sm:o_ready[i]", "sm:tile_done[i]"]
  print(f"--- fwd n={n} S={S}")
  for k, nm in enumerate(names):
    print(f"{nm:28s}", [int(v) - t0 for v in t[k] if int(v) != 0])


if len(sys.argv) > 1 and sys.argv[1] == "fwd":
  for n, S in ((512, 257), (256, 164), (256, 68)):
    run_fwd(n, S)
else:
  for n, S in ((512, 257), (256, 164), (256, 68)):
    run(n, S)
