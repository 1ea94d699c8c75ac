#!/usr/bin/env python
"""Headline benchmark: training images/sec of the UMD auto-encoder step (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload umd_b4|...]

One "step" = one call of update_fn (train_ae.py:287-382) on one synthetic batch: draws, q_sample, forward of
both branches, loss, backward, gradient all-reduce (N > 1), global-norm clip + AdamW.  Weak scaling: 512 images
per GPU (256 noised + 256 clean-MAE), the per-GPU share of the reference recipe's global batch 4096 on 8 GPUs.
Prints ONE JSON line on rank 0 (see README / DESIGN.md for the keys).

--impl reference times the CPU restatement of the reference step (oracle/umd_oracle.py) on the host cores:
the reference itself is JAX/Flax/Optax code and none of those can be installed in this image.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

METRIC = "train images/sec (UMD-B/4 64x64)"

WORKLOADS = {
    # name: (model kwargs, train kwargs, per-GPU batch)
    "umd_b4": (dict(variant="B/4", adaln=True, num_classes=None, channels=3, img_size=64),
               dict(no_noise_prob=0.5, mask_ratio=0.375, mask_ratio_no_noise=0.75, use_labels=False), 512),
    "umd_s4": (dict(variant="S/4", adaln=True, num_classes=None, channels=3, img_size=64),
               dict(no_noise_prob=0.5, mask_ratio=0.375, mask_ratio_no_noise=0.75, use_labels=False), 512),
    "mae_b4": (dict(variant="B/4", adaln=False, num_classes=None, channels=3, img_size=64),
               dict(no_noise_prob=0.5, mask_ratio=0.375, mask_ratio_no_noise=0.75, use_labels=False), 512),
    "dit_b4": (dict(variant="B/4", adaln=True, num_classes=1000, channels=3, img_size=64),
               dict(no_noise_prob=0.0, mask_ratio=0.0, mask_ratio_no_noise=0.75, use_labels=True), 256),
    "latent_umd_l2": (dict(variant="L/2", adaln=True, num_classes=None, channels=4, img_size=32),
                      dict(no_noise_prob=0.5, mask_ratio=0.375, mask_ratio_no_noise=0.75, use_labels=False,
                           beta_schedule="linear", diffusion_space=(32, 32, 4)), 128),
}


def workload_config(workload, per_gpu, world, num_params):
  """The `config` object of the JSON line.  Both arms print the SAME object for the same command line (the reference arm
  times a bounded sample of this workload and says which in `cpu_baseline.sample`); what differs between the arms —
  arithmetic type, where the step runs — is in `dtype`, `precision` and `impl`, outside `config`."""
  mkw, tkw, _ = WORKLOADS[workload]
  H, C = mkw["img_size"], mkw["channels"]
  n_clean = int(per_gpu * tkw["no_noise_prob"])
  return {"workload": f"{workload}: {mkw['variant']} {H}x{H}x{C} update_fn step, {per_gpu} img/GPU "
                      f"({per_gpu - n_clean} noised + {n_clean} clean)",
          "global_batch": per_gpu * world, "per_gpu_batch": per_gpu, "parallelism": f"dp{world}", "params": num_params,
          "l2": "inputs rotate over 4 batches; each step writes > 10 GB of activations (>> 126 MB L2)"}


def metric_name(workload):
  return METRIC if workload == "umd_b4" else f"train images/sec ({workload})"


def step_flops_per_image(cfg, tkw):
  """Algorithmic FLOPs of one training step per image: 3 x forward, 2MNK per GEMM,
  4 S^2 D per attention layer; no rematerialisation."""
  D, M = cfg.width, cfg.mlp
  L, p, C = cfg.num_patches, cfg.patch, cfg.channels
  ad = 1 if cfg.adaln else 0
  tok0 = 0 if cfg.adaln else 1
  Sd = L + 1 + tok0

  def fwd(keep):
    Se = keep + cfg.num_cls + tok0
    enc = cfg.depth * (Se * (8 * D * D + 4 * D * M) + 4 * Se * Se * D + 12 * D * D * ad)
    dec = cfg.dec_depth * (Sd * (8 * D * D + 4 * D * M) + 4 * Sd * Sd * D + 12 * D * D * ad)
    return enc + dec + L * 2 * (p * p * C) * D + L * 2 * D * (2 * p * p * C) + 8 * D * D + 4 * D * D * ad

  pn = tkw["no_noise_prob"]
  k0 = cfg.len_keep(tkw["mask_ratio"]) if tkw["mask_ratio"] > 0 else L
  k1 = cfg.len_keep(tkw["mask_ratio_no_noise"])
  return 3.0 * ((1 - pn) * fwd(k0) + pn * fwd(k1))


def make_device_batches(cfg, per_gpu, dev, rank, count=4):
  """Synthetic batches of SURVEY.md §8d: image ~ U(-1, 1) f32[B,H,W,C], label ~ U{0..num_classes-1}, generated on the
  device from seed 1 + rank (every rank gets its own shard of the global batch)."""
  import torch
  H, C = cfg.img_size, cfg.channels
  g = torch.Generator(device=dev).manual_seed(1 + rank)
  return [{"image": torch.rand(per_gpu, H, H, C, device=dev, generator=g) * 2 - 1,
           "label": torch.randint(0, max(cfg.num_classes or 1, 1), (per_gpu,), device=dev, generator=g)}
          for _ in range(count)]


def first_loss_check(workload, per_gpu, world, first_loss):
  """The loss of the first step (initial parameters, rank-0 batch and draws) against the value the fp32 oracle gives
  for the same state, batch and draws (tests/golden/bench_loss_golden.json, written on a B200 by
  tests/golden/make_bench_loss_golden.py and re-checked by tests/test_fullsize_gpu.py).  One GPU only: with N > 1 the
  loss slot is the mean over ranks."""
  path = os.path.join(ROOT, "tests", "golden", "bench_loss_golden.json")
  if world != 1 or not os.path.exists(path):
    return None
  with open(path) as f:
    g = json.load(f).get(workload)
  if not g or g.get("per_gpu_batch") != per_gpu:
    return None
  rel = abs(first_loss - g["oracle_loss"]) / abs(g["oracle_loss"])
  out = {"first_loss": first_loss, "oracle_fp32_loss": g["oracle_loss"], "rel_err": rel, "tolerance": 1e-2, "ok": rel <= 1e-2}
  if not out["ok"]:
    raise SystemExit(f"bench: first-step loss {first_loss} differs from the fp32 oracle's {g['oracle_loss']} by {rel:.3g} (> 1e-2)")
  return out


def dp_check(update_fn, state, model, dev_batches_of, world, rank, pg, dev):
  """Data-parallel correctness on NCCL (train_ae.py:159-170,287-290,364), run once after the timed region at N > 1:
    * the all-reduced gradient arena of one extra step against rank 0's own replay of EVERY rank's shard (same batches,
      same per-rank draws) averaged on one GPU — i.e. the N-rank step against the 1-rank computation of the global batch;
    * bit-equality across ranks of the parameters (and of the reduced gradients) after the timed steps."""
  import torch
  import torch.distributed as dist
  n = model.layout.total
  fb = update_fn.forward_backward
  batch = dev_batches_of(rank)[0]
  _, _, grads, _ = fb(state, batch, reduce=True)
  torch.cuda.synchronize()
  g_dp = grads[:n + 1].clone()           # slot n carries the loss
  out = {}
  # cross-rank equality: compare every rank's checksum pair with rank 0's
  arena = state["params"].arena
  sums = torch.stack([arena.double().sum(), arena.double().abs().sum(), g_dp[:n].double().sum(), g_dp[:n].double().abs().sum()])
  gathered = [torch.empty_like(sums) for _ in range(world)]
  dist.all_gather(gathered, sums, group=pg)
  out["param_checksums_equal"] = all(torch.equal(gathered[0][:2], x[:2]) for x in gathered)
  out["reduced_grad_checksums_equal"] = all(torch.equal(gathered[0][2:], x[2:]) for x in gathered)
  out["param_checksum"] = float(gathered[0][0])
  if rank == 0:
    acc = torch.zeros(n + 1, dtype=torch.float64, device=dev)
    for r in range(world):
      b = dev_batches_of(r)[0]
      _, _, g, _ = fb(state, b, rand_rank=r, reduce=False)
      acc += g[:n + 1].double()
    acc /= world
    torch.cuda.synchronize()
    d, ref = (g_dp[:n].double() - acc[:n]), acc[:n]
    out["grad_rel_l2_vs_one_rank"] = float(d.norm() / ref.norm())
    out["grad_cosine_vs_one_rank"] = float((g_dp[:n].double() @ ref) / (g_dp[:n].double().norm() * ref.norm()))
    out["loss_dp"] = float(g_dp[n])
    out["loss_one_rank"] = float(acc[n])
    out["loss_rel_err"] = abs(out["loss_dp"] - out["loss_one_rank"]) / abs(out["loss_one_rank"])
    # fp32 atomics reorder additions between runs (~1e-7 .. 5e-5 on a few leaves, tests/test_properties_gpu.py)
    out["ok"] = bool(out["param_checksums_equal"] and out["reduced_grad_checksums_equal"] and
                     out["grad_rel_l2_vs_one_rank"] <= 1e-3 and out["loss_rel_err"] <= 1e-5)
  dist.barrier()
  return out


def attention_bytes_per_step(cfg, tkw, per_gpu):
  """Algorithmic HBM bytes of the fused attention kernels per step: forward reads q, k, v and writes o (bf16) + lse
  (fp32); backward reads q, k, v, dO, lse, delta (the out-projection dgrad epilogue supplies delta = rowsum(dO o O), so
  O is not read) and writes dq, dk, dv.  rows x D x 2 B per bf16 tensor."""
  L, D, H = cfg.num_patches, cfg.width, cfg.num_heads
  tok0 = 0 if cfg.adaln else 1
  n1 = int(per_gpu * tkw["no_noise_prob"])
  n0 = per_gpu - n1
  k0 = cfg.len_keep(tkw["mask_ratio"]) if tkw["mask_ratio"] > 0 else L
  k1 = cfg.len_keep(tkw["mask_ratio_no_noise"])
  rows_enc = n0 * (k0 + cfg.num_cls + tok0) + n1 * (k1 + cfg.num_cls + tok0)
  rows_dec = per_gpu * (L + 1 + tok0)
  rows = cfg.depth * rows_enc + cfg.dec_depth * rows_dec
  return {"attention_fwd": rows * (4 * D * 2 + H * 4), "attention_bwd": rows * (7 * D * 2 + 2 * H * 4)}


class Watchdog:
  """Fail fast instead of hanging: if the benchmark makes no progress for `limit` seconds (a collective that never
  completes, a rendezvous that never forms) every thread's stack goes to stderr and the process exits with code 3."""

  def __init__(self, limit=600.0):
    import threading
    self.limit = limit
    self.phase = "start"
    self.t = time.time()
    self.done = False
    threading.Thread(target=self._run, daemon=True).start()

  def beat(self, phase=None):
    self.t = time.time()
    if phase:
      self.phase = phase

  def _run(self):
    import faulthandler
    while not self.done:
      time.sleep(5.0)
      if time.time() - self.t > self.limit:
        sys.stderr.write(f"bench.py watchdog: no progress for {self.limit:.0f} s in phase '{self.phase}' "
                         f"(rank {os.environ.get('RANK', '0')}); stacks follow\n")
        faulthandler.dump_traceback(file=sys.stderr, all_threads=True)
        sys.stderr.flush()
        os._exit(3)


def load_peaks():
  p = os.path.join(ROOT, "MEASURED_PEAKS.json")
  if os.path.exists(p):
    with open(p) as f:
      d = json.load(f)
    return dict(hbm=d.get("hbm_gbs", 6650.0), tf_burst=d.get("bf16_tflops", 1590.0),
                tf_sustained=d.get("bf16_tflops_sustained", 1400.0), source="measured")
  return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


class ClockSampler:
  """nvidia-smi clocks/throttle reasons sampled every 200 ms during the timed region (B200_PROFILING.md)."""
  Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
       "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
       "clocks_event_reasons.sw_power_cap")

  def __init__(self, gpu_index):
    self.idx = gpu_index
    self.proc = None
    self.path = None

  def start(self):
    try:
      fd, self.path = tempfile.mkstemp(prefix="clocks_", suffix=".csv")
      os.close(fd)
      self.f = open(self.path, "w")
      self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                    "-i", str(self.idx)], stdout=self.f, stderr=subprocess.DEVNULL)
    except Exception:  # nvidia-smi missing
      self.proc = None

  def stop(self):
    out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
    if self.proc is None:
      return out
    self.proc.terminate()
    try:
      self.proc.wait(timeout=5)
    except Exception:
      self.proc.kill()
    self.f.close()
    sm, mx, pw, reasons = [], [], [], set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    with open(self.path) as f:
      for line in f:
        parts = [x.strip() for x in line.split(",")]
        if len(parts) < 8:
          continue
        try:
          sm.append(float(parts[1])); mx.append(float(parts[2])); pw.append(float(parts[3]))
        except ValueError:
          continue
        for nm, v in zip(names, parts[4:8]):
          if v.lower().startswith("active"):
            reasons.add(nm)
    os.unlink(self.path)
    if sm:
      s = sorted(sm)
      out.update(sm_mhz=s[len(s) // 2], sm_max_mhz=max(mx), power_w_max=max(pw), reasons=sorted(reasons), samples=len(sm))
    return out


def run_reference(args):
  """CPU restatement of the reference step on the host cores (rank 0 only)."""
  rank = int(os.environ.get("RANK", "0"))
  if rank != 0:
    return
  import torch
  from oracle import umd_oracle as O
  from small_vision_b200.config import TrainConfig, make_model_config
  from small_vision_b200.diffusion import create_gaussian_diffusion
  from tests import util as U
  mkw, tkw, per_gpu = WORKLOADS[args.workload]
  cores = os.cpu_count() or 1
  torch.set_num_threads(cores)
  B = args.cpu_batch
  model, ocfg = U.make_models(mkw["variant"], adaln=mkw["adaln"], num_classes=mkw["num_classes"],
                              img_size=mkw["img_size"], channels=mkw["channels"])
  params = U.cpu_tree(U.perturb_init(model, 0, "cpu"))
  tcfg = TrainConfig(batch_size=B, total_steps=1000, warmup_steps=0, **{k: v for k, v in tkw.items()})
  state = {"params": params, "gd": create_gaussian_diffusion(tkw.get("beta_schedule", "cosine"), 1000)}
  state["opt"] = O.init_opt_state(params)
  hp = U.oracle_hp(tcfg)
  n_clean = int(B * tkw["no_noise_prob"])
  times = []
  for s in range(args.warmup + args.steps):
    b, rand = U.make_batch(model, B, n_noise=B - n_clean, seed=s, use_labels=tkw["use_labels"])
    t0 = time.perf_counter()
    state, meas, _ = O.update_step(state, b, ocfg, tkw, hp, rand)
    dt = time.perf_counter() - t0
    if s >= args.warmup:
      times.append(dt)
  ms = 1e3 * sum(times) / len(times)
  v = B / (ms / 1e3)
  sample = f"{args.steps} steps of batch {B} ({B - n_clean} noised + {n_clean} clean) after {args.warmup} warm-up, fp32, torch-CPU"
  if args.per_gpu_batch:
    per_gpu = args.per_gpu_batch
  line = {"metric": metric_name(args.workload), "value": v, "unit": "images/sec", "n_gpus": args.gpus, "steps": args.steps,
          "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
          "data": "synthetic", "impl": "reference",
          "config": workload_config(args.workload, per_gpu, args.gpus, model.layout.num_params),
          "precision": "fp32 throughout (the reference's default dtype_mm), torch-CPU",
          "note": f"CPU restatement of the reference step (oracle port; JAX is not installable here), timed on rank 0's host cores on a "
                  f"bounded sample of the workload: batch {B} per step instead of {per_gpu} per GPU; images/sec does not depend on N",
          "cpu_baseline": {"value": v, "unit": "images/sec", "cores": cores, "kind": "port", "sample": sample},
          "e2e": {"value": v, "unit": "images/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
          "gpu_launches": 0}
  print(json.dumps(line), flush=True)


def cpu_baseline(args, mkw, tkw):
  """Bounded oracle sample on the host cores (rank 0, N = 1): one warm-up + four timed steps at batch 8 (~10 s of CPU
  work on the GPU box's 16 cores)."""
  import torch
  from oracle import umd_oracle as O
  from small_vision_b200.config import TrainConfig
  from small_vision_b200.diffusion import create_gaussian_diffusion
  from tests import util as U
  cores = os.cpu_count() or 1
  torch.set_num_threads(cores)
  B = args.cpu_batch
  model, ocfg = U.make_models(mkw["variant"], adaln=mkw["adaln"], num_classes=mkw["num_classes"],
                              img_size=mkw["img_size"], channels=mkw["channels"])
  params = U.cpu_tree(U.perturb_init(model, 0, "cpu"))
  tcfg = TrainConfig(batch_size=B, total_steps=1000, warmup_steps=0, **tkw)
  state = {"params": params, "gd": create_gaussian_diffusion(tkw.get("beta_schedule", "cosine"), 1000)}
  state["opt"] = O.init_opt_state(params)
  hp = U.oracle_hp(tcfg)
  n_clean = int(B * tkw["no_noise_prob"])
  times = []
  for s in range(5):
    b, rand = U.make_batch(model, B, n_noise=B - n_clean, seed=s, use_labels=tkw["use_labels"])
    t0 = time.perf_counter()
    state, _, _ = O.update_step(state, b, ocfg, tkw, hp, rand)
    if s >= 1:
      times.append(time.perf_counter() - t0)
  dt = sum(times) / len(times)
  return {"value": B / dt, "unit": "images/sec", "cores": cores, "kind": "port",
          "sample": f"{len(times)} steps of batch {B} after 1 warm-up, fp32 torch-CPU restatement of update_fn (oracle/umd_oracle.py)"}


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=10)
  ap.add_argument("--warmup", type=int, default=3)
  ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
  ap.add_argument("--workload", default="umd_b4", choices=sorted(WORKLOADS))
  ap.add_argument("--per-gpu-batch", type=int, default=None)
  ap.add_argument("--cpu-batch", type=int, default=8)
  ap.add_argument("--no-cpu-baseline", action="store_true")
  ap.add_argument("--no-e2e", action="store_true")
  ap.add_argument("--no-dp-check", action="store_true")
  ap.add_argument("--residual", default="bfloat16", choices=["float32", "bfloat16", "bfloat16+grad"],
                  help="dtype of the residual stream between the blocks.  bfloat16 (default) is the stream of the reference's "
                       "dtype_mm='bfloat16' flow, i.e. of the 'bf16 pre-training' configuration BASELINE.json names; float32 is "
                       "the stream of its default dtype_mm.  Both pass the same parity tests (tests/test_fullsize_gpu.py)")
  args = ap.parse_args()
  args.warmup = max(args.warmup, 1)
  if args.impl == "reference":
    run_reference(args)
    return

  import torch
  import torch.distributed as dist
  from small_vision_b200 import lib
  from small_vision_b200.config import TrainConfig
  from small_vision_b200.model import Model
  from small_vision_b200.train import create_train_state, make_update_fn

  world = int(os.environ.get("WORLD_SIZE", "1"))
  rank = int(os.environ.get("RANK", "0"))
  local = int(os.environ.get("LOCAL_RANK", "0"))
  dog = Watchdog(float(os.environ.get("UMD_BENCH_WATCHDOG_S", "600")))
  if not torch.cuda.is_available():
    raise SystemExit("bench.py needs a CUDA device: the UMD hot path has no CPU fallback")
  torch.cuda.set_device(local)
  dev = torch.device("cuda", local)
  pg = None
  if world > 1:
    # NCCL prints a version banner on stdout when a process creates its first communicator; stdout carries the one JSON
    # line of this benchmark, so the banner goes to stderr
    import ctypes
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
      dog.beat("init_process_group")
      dist.init_process_group("nccl", device_id=dev)
      dist.barrier()
      torch.cuda.synchronize()
      dog.beat("process group up")
    finally:
      ctypes.CDLL(None).fflush(None)
      os.dup2(saved, 1)
      os.close(saved)
    pg = dist.group.WORLD
  assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"

  mkw, tkw, per_gpu = WORKLOADS[args.workload]
  if args.per_gpu_batch:
    per_gpu = args.per_gpu_batch
  model = Model(**mkw, residual_dtype=args.residual.split("+")[0],
                grad_stream_dtype="bfloat16" if args.residual.endswith("+grad") else "float32")
  cfg = model.cfg
  B_global = per_gpu * world
  tcfg = TrainConfig(batch_size=B_global, **tkw)
  state = create_train_state(model, tcfg, seed=0, device=dev, nonzero_adaln=True)
  # start the optimiser past the lr(0) = 0 warm-up step so that every timed step really moves the parameters
  state["opt"]["count"] = 10
  update_fn = make_update_fn(model, tcfg, process_group=pg)
  if world > 1:
    # create the C-level NCCL communicator now, while the GPUs are idle, instead of inside the first step
    dog.beat("umd_comm init")
    update_fn.communicator(dev)
    dist.barrier()
    torch.cuda.synchronize()
    dog.beat("umd_comm up")

  H, C = cfg.img_size, cfg.channels
  n_dev_batches = 4   # 4 x 25 MB of inputs; activations written per step (tens of GB) flush the 126 MB L2 anyway
  dev_batches = make_device_batches(cfg, per_gpu, dev, rank, n_dev_batches)
  host_batches = [{k: v.cpu().pin_memory() for k, v in b.items()} for b in dev_batches]

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  def timed(fn, steps):
    dog.beat()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(steps):
      fn(s)
      dog.beat()
    e1.record()
    barrier()
    dog.beat()
    ms = e0.elapsed_time(e1)
    if world > 1:
      t = torch.tensor([ms], device=dev)
      dist.all_reduce(t, op=dist.ReduceOp.MAX)
      ms = float(t.item())
    return ms

  losses = []

  def step_resident(s):
    nonlocal state
    state, meas = update_fn(state, dev_batches[s % n_dev_batches])
    losses.append(meas["training_loss"])

  # End-to-end step: every step copies its inputs host -> device from pinned memory and reads its loss back
  # device -> host.  Both are asynchronous the way a training loop would do them: the next step's inputs are staged
  # on a copy stream while this step's kernels run, and the loss lands in a pinned ring slot that is checked one step
  # later, so the launch queue never drains (a blocking read per step costs ~1 ms of idle GPU at ~400 launches per step).
  copy_stream = torch.cuda.Stream()
  e2e_loss = [torch.zeros(1, dtype=torch.float32).pin_memory() for _ in range(2)]
  e2e_ev = [torch.cuda.Event() for _ in range(2)]
  staged = {}
  e2e_seen = []

  def stage_inputs(s):
    hb = host_batches[s % n_dev_batches]
    with torch.cuda.stream(copy_stream):
      b = {k: v.to(dev, non_blocking=True) for k, v in hb.items()}
      ev = torch.cuda.Event()
      ev.record(copy_stream)
    return b, ev

  def step_e2e(s):
    nonlocal state
    b, ev = staged.pop(s) if s in staged else stage_inputs(s)
    cur = torch.cuda.current_stream()
    cur.wait_event(ev)
    for v in b.values():
      v.record_stream(cur)
    staged[s + 1] = stage_inputs(s + 1)          # host -> device copy of the next step's inputs, overlapped
    state, meas = update_fn(state, b)
    slot = s % 2
    if s >= 2:                                   # the loss of step s - 2 has long arrived: consume it before reuse
      e2e_ev[slot].synchronize()
      e2e_seen.append(float(e2e_loss[slot][0]))
    e2e_loss[slot].copy_(meas["training_loss"].reshape(1), non_blocking=True)   # device -> host read of the loss
    e2e_ev[slot].record(cur)

  dog.beat("warm-up (first step creates the umd_comm communicator)")
  for s in range(args.warmup):
    step_resident(s)
    dog.beat()
  dog.beat("timed steps")
  L = lib.load()
  import ctypes as Ct
  ncat = L.umd_profile_num_categories()
  L.umd_profile_category_name.restype = Ct.c_char_p
  launches0 = lib.launch_count()
  sampler = ClockSampler(local)
  if rank == 0:
    sampler.start()
  # headline: K un-instrumented steps.  Then the same K steps again with the library's per-kernel CUDA-event scopes
  # switched on (two event records per launch cost ~1 % of the step) for the roofline / breakdown.
  ms_total = timed(step_resident, args.steps)
  launches = lib.launch_count() - launches0
  # (the side stream is switched off for this pass so that the per-kernel event times do not overlap one another)
  L.umd_debug_side_stream(0)
  L.umd_profile_enable(1)
  ms_profiled = timed(step_resident, args.steps)
  L.umd_profile_enable(0)
  L.umd_debug_side_stream(1)
  prof_ms = (Ct.c_float * ncat)()
  prof_work = (Ct.c_double * ncat)()
  prof_n = (Ct.c_longlong * ncat)()
  dropped = L.umd_profile_read(prof_ms, prof_work, prof_n, ncat)
  e2e = None
  if not args.no_e2e:
    for s in range(2):
      step_e2e(s)
    ms_e2e = timed(step_e2e, args.steps)
    e2e = {"value": B_global * args.steps / (ms_e2e / 1e3), "unit": "images/sec",
           "h2d_bytes_per_step": sum(v.numel() * v.element_size() for v in host_batches[0].values()) * world,
           "d2h_bytes_per_step": 4 * world, "ms_per_step": ms_e2e / args.steps}
    assert e2e_seen and all(math.isfinite(x) for x in e2e_seen), "non-finite loss read back on the end-to-end path"
  clocks = sampler.stop() if rank == 0 else None
  final_loss = float(losses[-1])
  if not math.isfinite(final_loss):
    raise SystemExit(f"non-finite training loss {final_loss}")
  loss_check = first_loss_check(args.workload, per_gpu, world, float(losses[0])) if rank == 0 else None
  dog.beat("checks")
  dp = None
  if world > 1 and not args.no_dp_check:
    dp = dp_check(update_fn, state, model, lambda r: make_device_batches(cfg, per_gpu, dev, r, 1), world, rank, pg, dev)
    if rank == 0 and not dp["ok"]:
      raise SystemExit(f"bench: data-parallel check failed: {dp}")

  if rank == 0:
    peaks = load_peaks()
    ms_step = ms_total / args.steps
    ms_step_prof = ms_profiled / args.steps
    value = B_global * args.steps / (ms_total / 1e3)
    fl_img = step_flops_per_image(cfg, tkw)
    cats = {}
    for i in range(ncat):
      nm = L.umd_profile_category_name(i).decode()
      cats[nm] = {"ms_per_step": prof_ms[i] / args.steps, "work_per_step": prof_work[i] / args.steps,
                  "launches_per_step": prof_n[i] / args.steps}
    gm = cats["gemm"]
    gemm_tf = gm["work_per_step"] / (gm["ms_per_step"] * 1e-3) / 1e12 if gm["ms_per_step"] > 0 else 0.0
    # dram__bytes_read.sum + dram__bytes_write.sum per launch, launch-weighted over the forward / dgrad GEMM shapes of
    # one step, from the committed per-shape table (tools/traffic_table.py; regenerated by tools/profile_round.sh)
    traffic, traffic_src, traffic_alg = None, None, None
    tpath = os.path.join(ROOT, "profiles", "r02_traffic_table.json")
    if os.path.exists(tpath) and args.workload == "umd_b4" and per_gpu == WORKLOADS["umd_b4"][2]:
      with open(tpath) as f:
        tj = json.load(f)
      traffic = tj.get("fwd_dgrad_traffic_bytes_per_launch")
      traffic_alg = tj.get("fwd_dgrad_algorithmic_bytes_per_launch")
      traffic_src = "profiles/r02_traffic_table.json (commit %s: %d launches of one step under ncu, %d shapes)" % (
          tj.get("commit"), tj.get("fwd_dgrad_launches_per_step", 0), len(tj.get("shapes", {})))
    roofline = {"bound": "tensor", "kernel": "gemm_kernel (tcgen05 forward/dgrad GEMMs)", "achieved": gemm_tf,
                "peak": peaks["tf_sustained"], "unit": "TFLOP/s", "frac": gemm_tf / peaks["tf_sustained"],
                "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes_per_launch": traffic_alg,
                "algorithmic_flops_per_launch": gm["work_per_step"] / max(gm["launches_per_step"], 1), "peak_source": peaks["source"] + " bf16_tflops_sustained (kernel timed inside a long step)",
                "launches_per_step": gm["launches_per_step"], "share_of_step": gm["ms_per_step"] / ms_step_prof}
    breakdown = {}
    attn_bytes = attention_bytes_per_step(cfg, tkw, per_gpu)
    for nm, c in cats.items():
      if c["ms_per_step"] <= 0:
        continue
      tensor = nm.startswith("gemm") or nm.startswith("attention")
      rate = c["work_per_step"] / (c["ms_per_step"] * 1e-3)
      breakdown[nm] = {"ms": round(c["ms_per_step"], 3), "share": round(c["ms_per_step"] / ms_step_prof, 4),
                       "launches": c["launches_per_step"],
                       ("tflops" if tensor else "gbs"): round(rate / (1e12 if tensor else 1e9), 1),
                       "frac_of_peak": round(rate / ((peaks["tf_sustained"] * 1e12) if tensor else (peaks["hbm"] * 1e9)), 3)}
      if nm in attn_bytes:
        # arithmetic intensity ~ S / 2 flop/B (128 for the 257-token decoder) is below the ridge (1414.5 TF / 6.55 TB/s =
        # 216 flop/B): at these sequence lengths fused attention is HBM-bound, so the HBM fraction is the one that counts
        gbs = attn_bytes[nm] / (c["ms_per_step"] * 1e-3) / 1e9
        breakdown[nm].update(bound="hbm", gbs=round(gbs, 1), frac_of_hbm_peak=round(gbs / peaks["hbm"], 3),
                             frac_of_tensor_peak=breakdown[nm].pop("frac_of_peak"))
    line = {
        "metric": metric_name(args.workload), "value": value, "unit": "images/sec", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": workload_config(args.workload, per_gpu, world, model.layout.num_params),
        "precision": "bf16 GEMM/attention operands, fp32 accumulate, %s residual stream, fp32 LayerNorm statistics / gradient stream / loss / AdamW (bf16 mu)" % {"float32": "fp32", "bfloat16": "bf16", "bfloat16+grad": "bf16 (forward and gradient)"}[args.residual],
        "step_tflops_per_gpu": fl_img * value / world / 1e12,
        "step_frac_of_bf16_peak": fl_img * value / world / 1e12 / peaks["tf_sustained"],
        "flops_per_image": fl_img, "final_loss": final_loss, "loss_check": loss_check, "dp_check": dp,
        "roofline": roofline, "breakdown": breakdown, "profile_scopes_dropped": dropped,
        "breakdown_note": "per-kernel CUDA-event times of a second pass of K steps with the side stream off (no overlap between kernels)",
        "ms_per_step_profiled": ms_step_prof,
        "e2e": e2e, "gpu_launches": int(launches), "gpu_launches_per_step": launches / args.steps,
        "clocks": clocks,
    }
    if world == 1 and not args.no_cpu_baseline:
      line["cpu_baseline"] = cpu_baseline(args, mkw, tkw)
    print(json.dumps(line), flush=True)
  dog.beat("shutdown")
  if world > 1:
    dist.barrier()
    dist.destroy_process_group()
  dog.done = True


if __name__ == "__main__":
  main()
