"""Writes tests/golden/bench_loss_golden.json: for every bench.py workload, the loss the fp32 oracle (oracle/umd_oracle.py
run on the GPU with TF32 off, chunked — tests/test_fullsize_gpu.py) gives for bench.py's first step: seed-0 state,
rank-0 batch, rank-0 draws.  bench.py compares the loss of its first step with this value (N = 1).

Run on a B200 box:  python tests/golden/make_bench_loss_golden.py   (needs the built library; ~1 min)
The draws come from torch's CUDA Philox generator, so the file is tied to this image's torch build.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from tests.test_fullsize_gpu import bench_shape_parity  # noqa: E402

CHUNKS = {"umd_b4": 8, "mae_b4": 8, "dit_b4": 8, "latent_umd_l2": 4, "umd_s4": 4}


def main():
  out = {"_meta": {"torch": torch.__version__, "gpu": torch.cuda.get_device_name(0),
                   "how": "oracle.loss_fn on the GPU, fp32, TF32 off, mean over equal chunks of the bench batch"}}
  for wl, ch in CHUNKS.items():
    rep = bench_shape_parity(wl, n_chunks=ch)
    out[wl] = {"per_gpu_batch": rep["batch"], "oracle_loss": rep["oracle_loss"], "engine_loss": rep["loss"],
               "loss_rel": rep["loss_rel"], "gnorm_rel": rep["gnorm_rel"], "grad_cos_min": rep["grad_cos_min"],
               "grad_rel_max": rep["grad_rel_max"]}
    torch.cuda.empty_cache()
  path = os.path.join(ROOT, "tests", "golden", "bench_loss_golden.json")
  if os.environ.get("GOLDEN_OUT"):
    path = os.environ["GOLDEN_OUT"]
  with open(path, "w") as f:
    json.dump(out, f, indent=1)
  print(json.dumps(out, indent=1))


if __name__ == "__main__":
  main()
