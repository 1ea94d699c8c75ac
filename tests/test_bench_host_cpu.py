"""Host-side pieces of bench.py that need no GPU: the FLOP model behind `step_frac_of_bf16_peak`, the
algorithmic attention bytes behind the HBM fraction, the first-step loss gate against the committed fp32-oracle value, and
the no-progress watchdog."""
import json
import os
import time

import pytest

import bench
from small_vision_b200.config import make_model_config


def test_flop_model_matches_the_survey():
  mkw, tkw, per_gpu = bench.WORKLOADS["umd_b4"]
  cfg = make_model_config(**mkw)
  assert per_gpu == 512
  assert bench.step_flops_per_image(cfg, tkw) == pytest.approx(107.48e9, rel=1e-3)
  for name, want in (("mae_b4", 107.86e9), ("dit_b4", 186.60e9), ("umd_s4", 27.96e9), ("latent_umd_l2", 377.87e9)):
    mkw, tkw, _ = bench.WORKLOADS[name]
    assert bench.step_flops_per_image(make_model_config(**mkw), tkw) == pytest.approx(want, rel=2e-3), name


def test_attention_bytes_model():
  mkw, tkw, per_gpu = bench.WORKLOADS["umd_b4"]
  cfg = make_model_config(**mkw)
  b = bench.attention_bytes_per_step(cfg, tkw, per_gpu)
  rows = 12 * (256 * 164 + 256 * 68) + 4 * 512 * 257          # encoder + decoder token rows over all layers
  assert b["attention_fwd"] == rows * (4 * 768 * 2 + 12 * 4)  # q, k, v in, o out (bf16) + lse
  assert b["attention_bwd"] == rows * (7 * 768 * 2 + 2 * 12 * 4)   # q, k, v, dO in, dq, dk, dv out + lse, delta


def test_first_loss_gate(tmp_path, monkeypatch):
  gold = json.load(open(os.path.join(bench.ROOT, "tests", "golden", "bench_loss_golden.json")))
  g = gold["umd_b4"]
  ok = bench.first_loss_check("umd_b4", g["per_gpu_batch"], 1, g["oracle_loss"] * (1 + 2e-4))
  assert ok["ok"] and ok["rel_err"] == pytest.approx(2e-4, rel=1e-2)
  with pytest.raises(SystemExit):
    bench.first_loss_check("umd_b4", g["per_gpu_batch"], 1, g["oracle_loss"] * 1.05)
  assert bench.first_loss_check("umd_b4", g["per_gpu_batch"], 2, g["oracle_loss"]) is None       # N > 1: the loss slot is a mean over ranks
  assert bench.first_loss_check("umd_b4", 64, 1, g["oracle_loss"]) is None                        # another batch: no committed value
  for wl in ("umd_b4", "mae_b4", "dit_b4", "latent_umd_l2", "umd_s4"):
    assert gold[wl]["per_gpu_batch"] == bench.WORKLOADS[wl][2] and gold[wl]["loss_rel"] <= 1e-2


def test_watchdog_stays_quiet_while_beaten():
  dog = bench.Watchdog(limit=30.0)
  dog.beat("phase a")
  time.sleep(0.05)
  dog.beat()
  assert dog.phase == "phase a" and time.time() - dog.t < 5
  dog.done = True


def test_both_arms_print_the_same_config_object():
  """bench.py --impl reference must carry the b200 arm's `config`, `metric`, `unit` (the driver compares them): both come
  from workload_config / metric_name, and nothing arm-specific (precision, sample size) lives inside `config`."""
  for wl, (mkw, tkw, per_gpu) in bench.WORKLOADS.items():
    for world in (1, 8):
      c = bench.workload_config(wl, per_gpu, world, 123)
      assert set(c) == {"workload", "global_batch", "per_gpu_batch", "parallelism", "params", "l2"}
      assert c["global_batch"] == per_gpu * world and c["parallelism"] == f"dp{world}"
      assert c["workload"].startswith(wl + ": " + mkw["variant"])
  assert bench.metric_name("umd_b4") == bench.METRIC
  c = bench.workload_config("umd_b4", 512, 1, 1)
  assert "256 noised + 256 clean" in c["workload"]
  src = open(os.path.join(bench.ROOT, "bench.py")).read()
  assert src.count('"config": workload_config(') == 2          # one per arm, no hand-built config dict left
