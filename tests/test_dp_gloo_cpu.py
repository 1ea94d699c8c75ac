"""World-size-2 test of the data-parallel plumbing on CPU (gloo): the bucketed gradient all-reduce of
small-vision_b200/sharding.py (the explicit form of the all-reduce GSPMD inserts at train_ae.py:364 because the
batch is sharded on "data" and the parameters are replicated, train_ae.py:159-170).

Each rank runs the CPU oracle on its shard of one global batch, packs the gradient tree into the flat arena,
reduces it bucket by bucket through GradientReducer, and must end with the gradient of the whole batch."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
  s = socket.socket()
  s.bind(("127.0.0.1", 0))
  p = s.getsockname()[1]
  s.close()
  return p


def _worker(rank, world, port, out_dir):
  if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
  torch.set_num_threads(2)
  dist.init_process_group("gloo", rank=rank, world_size=world)
  try:
    from oracle import umd_oracle as O
    from small_vision_b200.params import arena_from_tree, tree_from_arena
    from small_vision_b200.sharding import GradientReducer, local_batch_slice
    from tests import util as U
    model, ocfg = U.make_models("S/4", adaln=True, depth=1, dec_depth=1)
    params = U.cpu_tree(U.perturb_init(model, 0, "cpu"))
    tc = dict(mask_ratio=0.375, mask_ratio_no_noise=0.75, no_noise_prob=0.5, use_labels=False)
    hp = dict(clip_norm=1.0, peak_lr=1e-3, warmup_steps=0, total_steps=100, b1=0.9, b2=0.95, wd=0.05)
    B, per = 4 * world, 4
    batch, rand = U.make_batch(model, B, n_noise=B // 2, seed=11, device="cpu")
    st = {"params": params, "gd": O.gaussian_diffusion_tables(), "opt": O.init_opt_state(params)}
    # this rank's shard: `per/2` noised + `per/2` clean samples (first half of each local shard is the noise branch)
    hn = per // 2
    sl = local_batch_slice(B // 2, rank, world)
    idx_n = list(range(sl.start, sl.stop))
    idx_c = [B // 2 + i for i in idx_n]
    b = {"image": torch.cat([batch["image"][idx_n], batch["image"][idx_c]]), "label": batch["label"][idx_n + idx_c]}
    rd = {"t": rand["t"][idx_n], "noise": rand["noise"][idx_n], "mask_noise_noise": rand["mask_noise_noise"][idx_n],
          "mask_noise_clean": rand["mask_noise_clean"][idx_n]}
    assert b["image"].shape[0] == per and len(idx_n) == hn
    _, meas, ex = O.update_step(st, b, ocfg, tc, hp, rd)
    lay = model.layout
    n_extra = 64
    grads = torch.zeros(lay.total + n_extra)
    grads[:lay.total] = arena_from_tree(lay, ex["grads"], "cpu")
    grads[lay.total] = meas["training_loss"]           # the loss scalar rides on the last bucket (train.py N_EXTRA)
    red = GradientReducer(lay, dist.group.WORLD)
    assert red.world == world
    for e in range(lay.num_events):                    # the engine's backward events, in order
      red.on_event(grads, e)
    red.finish()
    if rank == 0:
      _, gmeas, gex = O.update_step(st, batch, ocfg, tc, hp, rand)
      want = arena_from_tree(lay, gex["grads"], "cpu")
      err = float((grads[:lay.total] - want).abs().max())
      scale = float(want.abs().max())
      torch.save({"err": err, "scale": scale, "loss": float(grads[lay.total]), "want_loss": gmeas["training_loss"],
                  "buckets": lay.bucket_bounds}, os.path.join(out_dir, "result.pt"))
    # every rank must hold bit-identical reduced gradients (clip + AdamW then stay replicated without a broadcast)
    mine = grads.clone()
    dist.broadcast(grads, src=0)
    assert torch.equal(mine, grads)
  finally:
    dist.destroy_process_group()


def test_bucketed_gradient_allreduce_world2(tmp_path):
  world = 2
  mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
  r = torch.load(os.path.join(tmp_path, "result.pt"))
  assert r["err"] <= 2e-4 * r["scale"] + 1e-7, r
  assert abs(r["loss"] - r["want_loss"]) <= 1e-5 * abs(r["want_loss"]), r
  assert len(r["buckets"]) == 2     # depth 1: decoder side, then embeddings + the encoder layer


def test_single_process_reducer_is_a_noop():
  from small_vision_b200.config import make_model_config
  from small_vision_b200.params import ArenaLayout
  from small_vision_b200.sharding import GradientReducer
  lay = ArenaLayout(make_model_config(variant="S/4", adaln=True, depth=1, dec_depth=1))
  g = torch.arange(lay.total + 64, dtype=torch.float32)
  red = GradientReducer(lay, None)
  before = g.clone()
  for e in range(lay.num_events):
    red.on_event(g, e)
  red.finish()
  assert torch.equal(g, before)
  views = [red.bucket_view(g, k) for k in range(len(lay.bucket_bounds))]
  assert sum(v.numel() for v in views) == g.numel()    # the buckets (+ trailing scalars) tile the arena
