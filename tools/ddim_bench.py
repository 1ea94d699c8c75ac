"""Sampling throughput of the DDIM loop: DiT-B/4 64x64, class-conditional with
classifier-free guidance (every step runs the forward on the doubled batch), eta = 1, clip_denoised, cosine schedule —
the reference's sampling recipe (configs/ae_i1k.py:44-53: 125 steps, 1024 samples per call).  Prints one JSON line:
sampled images/sec with the model forward and the fused DDIM update timed by CUDA events, the achieved forward
TFLOP/s against the measured bf16 peak, and the achieved GB/s of the DDIM update kernel alone.

  python tools/ddim_bench.py [--batch 256] [--steps 125] [--variant B/4] [--cfg-scale 1.5]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from small_vision_b200 import diffusion as Dm, lib  # noqa: E402
from small_vision_b200.model import Model  # noqa: E402
from small_vision_b200.params import tree_from_arena, init_arena  # noqa: E402
import bench  # noqa: E402


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--batch", type=int, default=256)
  ap.add_argument("--steps", type=int, default=125)
  ap.add_argument("--variant", default="B/4")
  ap.add_argument("--cfg-scale", type=float, default=1.5)
  a = ap.parse_args()
  dev = "cuda"
  model = Model(variant=a.variant, adaln=True, num_classes=1000, channels=3, img_size=64)
  params = tree_from_arena(model.layout, init_arena(model.layout, 0, dev, nonzero_adaln=True))
  gd = Dm.to_device(Dm.create_gaussian_diffusion("cosine", 1000), dev)
  apply_fn = Dm.create_apply_fn(model, params)
  ys = torch.randint(0, 1000, (a.batch,), device=dev)
  shape = torch.zeros(a.batch, 64, 64, 3)

  def run(steps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    l0 = lib.launch_count()
    e0.record()
    out, _ = Dm.ddim_sample_loop(gd, apply_fn, 1, shape, ys=ys, clip_denoised=True, sampling_steps=steps,
                                 cfg_scale=a.cfg_scale, eta=1.0)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1), out["sample"], lib.launch_count() - l0

  run(3)  # warm-up (workspace allocation, tensor maps)
  ms, sample, launches = run(a.steps)
  assert torch.isfinite(sample).all()
  cfg = model.cfg
  tkw = dict(no_noise_prob=0.0, mask_ratio=0.0, mask_ratio_no_noise=0.75, use_labels=True)
  fwd_flops_img = bench.step_flops_per_image(cfg, tkw) / 3.0
  nfwd = a.steps + 1
  tflops = fwd_flops_img * 2 * a.batch * nfwd / (ms * 1e-3) / 1e12
  # the update kernel alone: x, noise in; 2 x eps head of the doubled batch in; sample, pred_xstart out (fp32)
  x = torch.randn(a.batch, 64, 64, 3, device=dev)
  pred = torch.randn(2 * a.batch, 64, 64, 6, device=dev)
  t = torch.full((a.batch, 1), 500, dtype=torch.int32, device=dev)
  fn = lambda: Dm.ddim_sample(gd, lambda **kw: Dm.RawPred(pred, a.cfg_scale, True), x, t, t - 8, None, eta=1.0, noise=x)
  for _ in range(3):
    fn()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(20):
    fn()
  e1.record()
  torch.cuda.synchronize()
  k_ms = e0.elapsed_time(e1) / 20
  k_bytes = x.numel() * 4 * (2 + 2 + 2)   # x, noise, 2 eps reads (sector-granular reads of the 2C rows fetch more), 2 outputs
  peaks = bench.load_peaks() if hasattr(bench, "load_peaks") else {}
  print(json.dumps({
      "metric": "DDIM sampled images/sec (DiT-%s 64x64, CFG)" % a.variant, "value": a.batch / (ms * 1e-3), "unit": "images/sec",
      "config": {"batch": a.batch, "sampling_steps": a.steps, "forwards": nfwd, "cfg_scale": a.cfg_scale, "eta": 1.0,
                 "forward_batch": 2 * a.batch},
      "ms_total": ms, "ms_per_sampler_step": ms / nfwd, "forward_tflops": tflops, "gpu_launches": launches,
      "ddim_step_kernel": {"ms_with_host_glue": k_ms, "algorithmic_bytes": k_bytes, "gbs": k_bytes / (k_ms * 1e-3) / 1e9},
      "peaks": peaks}))


if __name__ == "__main__":
  main()
