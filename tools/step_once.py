"""Runs exactly N training steps of a bench.py workload (default 1, no warm-up) — the process that
tools/profile_round.sh puts under ncu to capture every launch of one step.  With UMD_GEMM_LOG=<file> the library logs the
shape of every GEMM launch in order (joined with the ncu launch list by tools/traffic_table.py)."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--workload", default="umd_b4")
  ap.add_argument("--steps", type=int, default=1)
  ap.add_argument("--per-gpu-batch", type=int, default=None)
  a = ap.parse_args()
  from small_vision_b200.config import TrainConfig
  from small_vision_b200.model import Model
  from small_vision_b200.train import create_train_state, make_update_fn
  mkw, tkw, n = bench.WORKLOADS[a.workload]
  n = a.per_gpu_batch or n
  model = Model(**mkw)
  tcfg = TrainConfig(batch_size=n, **tkw)
  state = create_train_state(model, tcfg, seed=0, device="cuda", nonzero_adaln=True)
  state["opt"]["count"] = 10
  fn = make_update_fn(model, tcfg)
  batch = bench.make_device_batches(model.cfg, n, torch.device("cuda"), 0, 1)[0]
  for _ in range(a.steps):
    state, meas = fn(state, batch)
  torch.cuda.synchronize()
  print("loss", float(meas["training_loss"]))


if __name__ == "__main__":
  main()
