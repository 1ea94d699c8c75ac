"""Timings of the §8f rank 3 / rank 4 components on one B200 (CUDA events on the launching stream, warm-up first):
the few-shot ridge probe at the reference's ImageNet shapes (configs/eval_ae_i1k.py:122: 10 and 100 shots x 1000
classes, width 768) and the fused input stage at the bench batch.  Prints one JSON line.
  python tools/widening_bench.py > gpurun_out/widening_bench.json
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from small_vision_b200 import fewshot as FS, pp  # noqa: E402


def timed(fn, reps=5, warm=2):
  for _ in range(warm):
    fn()
  torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(reps):
    fn()
  b.record()
  torch.cuda.synchronize()
  return a.elapsed_time(b) / reps


def main():
  dev = "cuda"
  out = {"device": torch.cuda.get_device_name(0)}
  g = torch.Generator(device=dev).manual_seed(0)
  c, d = 1000, 768
  for shots in (10, 100):
    n, nt = c * shots, 50_000
    y = torch.arange(c, device=dev, dtype=torch.int32).repeat_interleave(shots)
    centres = torch.randn(c, d, device=dev, generator=g) * 0.25
    x = centres[y.long()] + torch.randn(n, d, device=dev, generator=g)
    yt = torch.randint(0, c, (nt,), device=dev, generator=g, dtype=torch.int32)
    xt = centres[yt.long()] + torch.randn(nt, d, device=dev, generator=g)
    cache = FS._precompute_cache(x, y, c)
    t_cache = timed(lambda: FS._precompute_cache(x, y, c), reps=3, warm=1)
    t_solve = timed(lambda: FS.ridge_weights(cache, 1024.0), reps=3, warm=1)
    t_acc = timed(lambda: FS._eig_fewshot_acc_fn(cache, xt, yt, 1024.0), reps=3, warm=1)
    gram_flops = 2.0 * n * (d + 1) ** 2
    score_flops = 2.0 * nt * (d + 1) * c
    out[f"fewshot_{shots}shot"] = {
        "support": n, "query": nt, "classes": c, "width": d,
        "precompute_cache_ms": t_cache, "gram_fp32_tflops": gram_flops / (t_cache * 1e-3) / 1e12,
        "ridge_solve_ms": t_solve, "acc_fn_ms": t_acc,
        "score_matmul_fp32_tflops_lower_bound": score_flops / (t_acc * 1e-3) / 1e12,
        "accuracy": float(FS._eig_fewshot_acc_fn(cache, xt, yt, 1024.0))}
  # input stage: 512 images per step; cached 64x64 sources (downsampled ImageNet) and 256x256 sources
  for H in (64, 256):
    n, S = 512, 64
    img = torch.randint(0, 256, (n, H, H, 3), device=dev, dtype=torch.uint8, generator=g)
    boxes = torch.as_tensor(pp.sample_inception_boxes(n, H, H, seed=1)).to(dev)
    flips = torch.rand(n, device=dev, generator=g) < 0.5
    b_cpu = boxes.cpu()
    ms = timed(lambda: pp.augment(img, boxes=b_cpu, flips=flips, size=S), reps=20, warm=3)
    wr = n * S * S * 3 * 4
    rd = int((b_cpu[:, 2] * b_cpu[:, 3]).sum()) * 3
    out[f"augment_src{H}"] = {"images": n, "ms_incl_host_box_check_and_upload": ms, "bytes_written": wr,
                               "bytes_read_upper": rd, "images_per_s": n / (ms * 1e-3)}
  print(json.dumps(out))


if __name__ == "__main__":
  main()
