"""Sharding declarations (big_vision/sharding.py:33-78) and the data-parallel plumbing they imply.

The reference declares shardings and lets GSPMD insert the gradient all-reduce
(train_ae.py:159-170,287-290,364).  Here one process drives one GPU: parameters, optimiser state,
schedule tables and RNG are replicated, the batch is split on axis 0 in rank order, and the flat
gradient arena is all-reduced (mean) over NCCL in buckets (decoder side, groups of encoder layers, embeddings)
launched from the backward's event callback so that they overlap the rest of the backward pass.
"""
from __future__ import annotations

import numpy as np


class Mesh:
  """Minimal stand-in for jax.sharding.Mesh: a 1-D list of ranks and its axis names."""

  def __init__(self, devices, axis_names=("data",)):
    self.devices = np.asarray(devices)
    self.axis_names = tuple(axis_names)


class PartitionSpec(tuple):
  def __new__(cls, *parts):
    return super().__new__(cls, parts)

  def __repr__(self):
    return "P(" + ", ".join(repr(p) for p in self) + ")"


P = PartitionSpec


def _tree_map(fn, tree):
  if isinstance(tree, dict):
    return {k: _tree_map(fn, v) for k, v in tree.items()}
  return fn(tree)


def replicated(params, mesh, axis_name):
  """sharding.py:53-55."""
  del axis_name, mesh
  return _tree_map(lambda _: P(), params)


def fully_sharded(params, mesh, axis_name, too_small_to_shard_thr=2 ** 18):
  """sharding.py:58-78: shard the largest evenly divisible dim of every array above the threshold."""
  idx = mesh.axis_names.index(axis_name)
  axis_size = np.shape(mesh.devices)[idx]

  def spec(x):
    shape = tuple(x.shape)
    if np.prod(shape) <= too_small_to_shard_thr:
      return P()
    for i in np.argsort(shape)[::-1]:
      if shape[i] % axis_size == 0:
        return P(*((None,) * int(i) + (axis_name,)))
    return P()

  return _tree_map(spec, params)


def infer_sharding(params, mesh, axis_name, strategy, extra_strategy_args):
  """sharding.py:33-50 — same signature, returns a pytree of PartitionSpecs."""
  fn = {"replicated": replicated, "fully_sharded": fully_sharded}[strategy]
  return fn(params, mesh, axis_name, **extra_strategy_args)


def check_batch_divisible(batch_size, world):
  """train_ae.py:65-68."""
  if batch_size % world != 0:
    raise ValueError(f"Batch size ({batch_size}) must be divisible by device number ({world})")


def local_batch_slice(global_batch, rank, world):
  """input_pipeline.py:205-218: axis-0 split in mesh order."""
  check_batch_divisible(global_batch, world)
  per = global_batch // world
  return slice(rank * per, (rank + 1) * per)


class GradientReducer:
  """Mean all-reduce of the gradient arena in the buckets of params.ArenaLayout (the decoder side, groups of encoder
  layer blocks from the top down, the embeddings), each issued asynchronously from the backward's event callback as
  soon as the engine has enqueued the last kernel that writes it (C1 in SURVEY.md §2.3), so that the reduction of
  every bucket but the last overlaps the rest of the backward pass.  With no process group (single GPU) every method
  is a no-op."""

  def __init__(self, layout, process_group=None):
    self.layout = layout
    self.pg = process_group
    self.works = []
    self.world = 1
    self.rank = 0
    if process_group is not None:
      import torch.distributed as dist
      self.world = dist.get_world_size(process_group)
      self.rank = dist.get_rank(process_group)

  def bucket_view(self, grads, k):
    lo, hi = self.layout.bucket_bounds[k]
    if hi == self.layout.total:
      hi = grads.numel()  # trailing scalar slots (loss) ride on the bucket that ends the arena
    return grads[lo:hi]

  def on_event(self, grads, event):
    """Backward event `event` (include/umd_b200.h, umd_backward) has been enqueued: launch the buckets it completes."""
    for k, e in enumerate(self.layout.bucket_events):
      if e == event:
        self.launch(grads, k)

  def launch(self, grads, k):
    if self.pg is None or self.world == 1:
      return
    import torch.distributed as dist
    view = self.bucket_view(grads, k)
    if grads.is_cuda:
      self.works.append(dist.all_reduce(view, op=dist.ReduceOp.AVG, group=self.pg, async_op=True))
    else:  # gloo (CPU tests) has no AVG
      w = dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.pg, async_op=True)
      self.works.append((w, view))

  def finish(self):
    for w in self.works:
      if isinstance(w, tuple):
        w[0].wait()
        w[1].div_(self.world)
      else:
        w.wait()
    self.works = []


class Communicator:
  """The C-level data-parallel communicator (umd_comm_*, NCCL below the C ABI) of one rank.  torch.distributed is only
  the side channel that carries rank 0's ncclUniqueId to the other ranks; the gradient all-reduce itself is issued by
  umd_train_step on the communicator's own stream behind the backward events."""

  def __init__(self, process_group):
    import ctypes as C
    import torch.distributed as dist
    from . import lib
    self._lib = lib
    L = lib.load()
    if not L.umd_comm_available():
      raise lib.UmdError("libnccl.so.2 could not be loaded: no data-parallel path without NCCL")
    self.world = dist.get_world_size(process_group)
    self.rank = dist.get_rank(process_group)
    buf = (C.c_ubyte * 128)()
    if self.rank == 0:
      lib.check(L.umd_comm_unique_id(buf), "umd_comm_unique_id")
    box = [bytes(buf)]
    dist.broadcast_object_list(box, src=dist.get_global_rank(process_group, 0), group=process_group)
    idb = (C.c_ubyte * 128).from_buffer_copy(box[0])
    h = C.c_void_p()
    # NCCL announces its version on stdout when the first communicator of a process is created (NCCL_DEBUG=VERSION/WARN);
    # stdout belongs to the caller (bench.py prints exactly one JSON line there), so the banner is sent to stderr
    import os
    import sys
    sys.stdout.flush()
    libc = C.CDLL(None)
    libc.fflush(None)
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
      rc = L.umd_comm_init(C.c_int(self.rank), C.c_int(self.world), idb, C.byref(h))
    finally:
      libc.fflush(None)
      os.dup2(saved, 1)
      os.close(saved)
    lib.check(rc, "umd_comm_init")
    self.handle = h

  def allreduce_mean(self, t):
    import ctypes as C
    lib = self._lib
    assert t.is_cuda and t.dtype.is_floating_point and t.element_size() == 4 and t.is_contiguous()
    lib.check(lib.load().umd_comm_allreduce_mean(self.handle, lib.ptr(t), C.c_longlong(t.numel()), lib.current_stream()),
              "umd_comm_allreduce_mean")
    return t

  def close(self):
    if getattr(self, "handle", None):
      self._lib.load().umd_comm_destroy(self.handle)
      self.handle = None

  def __del__(self):
    try:
      self.close()
    except Exception:
      pass
