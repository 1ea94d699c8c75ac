"""Parameter tree <-> flat arena.

The reference keeps parameters as a Flax nested dict (shapes from
big_vision/models/ae.py:57-97, vit.py:60-163, embeddings.py).  Here the same nested dict is a set
of *views* into one flat fp32 arena, so that the CUDA engine, the optimiser and the gradient
all-reduce each see one contiguous buffer.  The arena order groups leaves by the moment their
gradients become final during backward (decoder side, encoder, embeddings/conditioning), which
is what the bucketed all-reduce overlaps with (sharding.py here, train_ae.py:287-290,364 there).
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import torch

from .config import ModelConfig

# --- leaf ids: must match the enums in include/umd_b200.h -------------------------------------
(P_CLS, P_POS, P_DEC_POS, P_MASK_TOKEN, P_EMBED_W, P_EMBED_B, P_TT_W0, P_TT_B0, P_TT_W1, P_TT_B1, P_LABEL_TABLE,
 P_LT_W0, P_LT_B0, P_LT_W1, P_LT_B1, P_FMOD_W, P_FMOD_B, P_FCONV_W, P_FCONV_B) = range(19)
P_ENC_BASE = 19
P_DEC_BASE = P_ENC_BASE + 20
P_COUNT = P_DEC_BASE + 20
(S_ADA_W, S_ADA_B, S_LN0_S, S_LN0_B, S_LN1_S, S_LN1_B, S_Q_W, S_K_W, S_V_W, S_Q_B, S_K_B, S_V_B, S_O_W, S_O_B,
 S_FC1_W, S_FC1_B, S_FC2_W, S_FC2_B, S_NORM_S, S_NORM_B) = range(20)

SCAN_NAME = "ScanCheckpointEncoder1DBlock_0"  # flax auto-name of nn.scan(nn.remat(Encoder1DBlock)) (vit.py:131-148)
ALIGN = 64  # elements; keeps every leaf 256-byte aligned (TMA needs 16) and wd flags per 64-element block exact


class Leaf:
  __slots__ = ("leaf_id", "path", "shape", "init", "bucket", "offset", "size")

  def __init__(self, leaf_id, path, shape, init, bucket):
    self.leaf_id, self.path, self.shape, self.init, self.bucket = leaf_id, tuple(path), tuple(shape), init, bucket
    self.offset = -1
    self.size = int(math.prod(shape))


def _stack_leaves(cfg: ModelConfig, name: str, base: int, depth: int, bucket: int) -> List[Leaf]:
  D, H, M = cfg.width, cfg.num_heads, cfg.mlp
  Dh = D // H
  blk = (name, SCAN_NAME)
  att = blk + ("MultiHeadDotProductAttention_0",)
  out = []
  if cfg.adaln:
    out += [Leaf(base + S_ADA_W, blk + ("Dense_0", "kernel"), (depth, D, 6 * D), "adaln", bucket),
            Leaf(base + S_ADA_B, blk + ("Dense_0", "bias"), (depth, 6 * D), "zeros", bucket)]
  out += [
      Leaf(base + S_LN0_S, blk + ("LayerNorm_0", "scale"), (depth, D), "ones", bucket),
      Leaf(base + S_LN0_B, blk + ("LayerNorm_0", "bias"), (depth, D), "zeros", bucket),
      Leaf(base + S_LN1_S, blk + ("LayerNorm_1", "scale"), (depth, D), "ones", bucket),
      Leaf(base + S_LN1_B, blk + ("LayerNorm_1", "bias"), (depth, D), "zeros", bucket),
      # query / key / value kept adjacent and equally sized: the engine runs them as one batch-3 GEMM
      Leaf(base + S_Q_W, att + ("query", "kernel"), (depth, D, H, Dh), ("xavier", D, D), bucket),
      Leaf(base + S_K_W, att + ("key", "kernel"), (depth, D, H, Dh), ("xavier", D, D), bucket),
      Leaf(base + S_V_W, att + ("value", "kernel"), (depth, D, H, Dh), ("xavier", D, D), bucket),
      Leaf(base + S_Q_B, att + ("query", "bias"), (depth, H, Dh), "zeros", bucket),
      Leaf(base + S_K_B, att + ("key", "bias"), (depth, H, Dh), "zeros", bucket),
      Leaf(base + S_V_B, att + ("value", "bias"), (depth, H, Dh), "zeros", bucket),
      Leaf(base + S_O_W, att + ("out", "kernel"), (depth, H, Dh, D), ("xavier", D, D), bucket),
      Leaf(base + S_O_B, att + ("out", "bias"), (depth, D), "zeros", bucket),
      Leaf(base + S_FC1_W, blk + ("MlpBlock_0", "Dense_0", "kernel"), (depth, D, M), ("xavier", D, M), bucket),
      Leaf(base + S_FC1_B, blk + ("MlpBlock_0", "Dense_0", "bias"), (depth, M), ("normal", 1e-6), bucket),
      Leaf(base + S_FC2_W, blk + ("MlpBlock_0", "Dense_1", "kernel"), (depth, M, D), ("xavier", M, D), bucket),
      Leaf(base + S_FC2_B, blk + ("MlpBlock_0", "Dense_1", "bias"), (depth, D), ("normal", 1e-6), bucket),
      Leaf(base + S_NORM_S, (name, "encoder_norm", "scale"), (D,), "ones", bucket),
      Leaf(base + S_NORM_B, (name, "encoder_norm", "bias"), (D,), "zeros", bucket),
  ]
  return out


def leaf_specs(cfg: ModelConfig) -> List[Leaf]:
  """All leaves in arena order: bucket 0 (decoder side), bucket 1 (encoder), bucket 2 (the rest)."""
  D, L, p, C = cfg.width, cfg.num_patches, cfg.patch, cfg.channels
  leaves: List[Leaf] = []
  leaves += [Leaf(P_FCONV_W, ("final_conv", "kernel"), (p, p, D, 2 * C), ("normal", 0.02), 0),
             Leaf(P_FCONV_B, ("final_conv", "bias"), (2 * C,), "zeros", 0)]
  if cfg.adaln:
    leaves += [Leaf(P_FMOD_W, ("final_modulation", "kernel"), (D, 2 * D), "adaln", 0),
               Leaf(P_FMOD_B, ("final_modulation", "bias"), (2 * D,), "zeros", 0)]
  leaves += _stack_leaves(cfg, "Decoder", P_DEC_BASE, cfg.dec_depth, 0)
  leaves += [Leaf(P_DEC_POS, ("dec_pos_embedding",), (1, L, D), ("normal", 1 / math.sqrt(L)), 0),
             Leaf(P_MASK_TOKEN, ("image_mask_embedding",), (1, 1, D), ("normal", 0.02), 0)]
  leaves += _stack_leaves(cfg, "Encoder", P_ENC_BASE, cfg.depth, 1)
  leaves += [Leaf(P_CLS, ("cls",), (1, cfg.num_cls, D), "zeros", 2),
             Leaf(P_POS, ("pos_embedding",), (1, L, D), ("normal", 1 / math.sqrt(L)), 2),
             Leaf(P_EMBED_W, ("embedding", "kernel"), (p, p, C, D), ("lecun", p * p * C), 2),
             Leaf(P_EMBED_B, ("embedding", "bias"), (D,), "zeros", 2),
             Leaf(P_TT_W0, ("time_trunk", "Dense_0", "kernel"), (D, 2 * D), ("lecun", D), 2),
             Leaf(P_TT_B0, ("time_trunk", "Dense_0", "bias"), (2 * D,), "zeros", 2),
             Leaf(P_TT_W1, ("time_trunk", "Dense_1", "kernel"), (2 * D, D), ("lecun", 2 * D), 2),
             Leaf(P_TT_B1, ("time_trunk", "Dense_1", "bias"), (D,), "zeros", 2)]
  if cfg.num_classes is not None:
    nc = cfg.num_classes
    leaves += [Leaf(P_LABEL_TABLE, ("label_emb", "embedding", "embedding"), (nc + 1, D), ("normal", 1 / math.sqrt(D)), 2),
               Leaf(P_LT_W0, ("label_trunk", "Dense_0", "kernel"), (D, 2 * D), ("lecun", D), 2),
               Leaf(P_LT_B0, ("label_trunk", "Dense_0", "bias"), (2 * D,), "zeros", 2),
               Leaf(P_LT_W1, ("label_trunk", "Dense_1", "kernel"), (2 * D, D), ("lecun", 2 * D), 2),
               Leaf(P_LT_B1, ("label_trunk", "Dense_1", "bias"), (D,), "zeros", 2)]
  return leaves


class ArenaLayout:
  """Offsets of every leaf in the flat arena, bucket boundaries and the weight-decay flags."""

  def __init__(self, cfg: ModelConfig):
    self.cfg = cfg
    self.leaves = leaf_specs(cfg)
    off = 0
    self.bucket_bounds: List[Tuple[int, int]] = []
    cur_bucket, start = 0, 0
    for lf in self.leaves:
      if lf.bucket != cur_bucket:
        self.bucket_bounds.append((start, off))
        cur_bucket, start = lf.bucket, off
      lf.offset = off
      off += (lf.size + ALIGN - 1) // ALIGN * ALIGN
    self.bucket_bounds.append((start, off))
    self.total = off
    self.num_params = sum(lf.size for lf in self.leaves)
    self.by_path: Dict[Tuple[str, ...], Leaf] = {lf.path: lf for lf in self.leaves}
    self.offsets = [-1] * P_COUNT
    for lf in self.leaves:
      self.offsets[lf.leaf_id] = lf.offset

  def decay(self, lf: Leaf) -> bool:
    """train_ae.py:125-134: decayed iff no path component is in model.no_decay_list."""
    return all(k not in self.cfg.no_decay_list for k in lf.path)

  def wd_flags(self, device) -> torch.Tensor:
    flags = torch.zeros(self.total // ALIGN, dtype=torch.uint8)
    for lf in self.leaves:
      if self.decay(lf):
        flags[lf.offset // ALIGN:(lf.offset + lf.size + ALIGN - 1) // ALIGN] = 1
    return flags.to(device)

  def offsets_tensor(self):
    import ctypes as C
    return (C.c_longlong * P_COUNT)(*self.offsets)


class ParamTree(dict):
  """A nested dict whose leaves are views into `arena` (dict subclass so reference-style code that
  walks params as a plain pytree keeps working)."""
  arena: torch.Tensor = None
  layout: ArenaLayout = None


def tree_from_arena(layout: ArenaLayout, arena: torch.Tensor) -> ParamTree:
  root = ParamTree()
  root.arena, root.layout = arena, layout
  for lf in layout.leaves:
    d = root
    for k in lf.path[:-1]:
      d = d.setdefault(k, {})
    d[lf.path[-1]] = arena[lf.offset:lf.offset + lf.size].view(lf.shape)
  return root


def flatten(tree, prefix=()):
  out = {}
  for k, v in tree.items():
    if isinstance(v, dict):
      out.update(flatten(v, prefix + (k,)))
    else:
      out[prefix + (k,)] = v
  return out


def _canonical_path(layout: ArenaLayout, path):
  """Maps a reference path to ours, tolerating a different auto-generated scan-block name."""
  if path in layout.by_path:
    return path
  if len(path) >= 3 and path[0] in ("Encoder", "Decoder") and path[1] != "encoder_norm":
    alt = (path[0], SCAN_NAME) + tuple(path[2:])
    if alt in layout.by_path:
      return alt
  raise KeyError(f"unexpected parameter leaf {'/'.join(path)}")


def arena_from_tree(layout: ArenaLayout, tree, device, dtype=torch.float32) -> torch.Tensor:
  """Packs any nested dict with the reference's leaf paths (e.g. an imported checkpoint) into a new arena."""
  if isinstance(tree, ParamTree) and tree.layout is not None and tree.arena is not None \
      and tree.arena.device == torch.device(device) and tree.arena.dtype == dtype and tree.layout.total == layout.total:
    return tree.arena
  arena = torch.zeros(layout.total, dtype=dtype, device=device)
  flat = flatten(tree)
  seen = set()
  for path, v in flat.items():
    cp = _canonical_path(layout, path)
    lf = layout.by_path[cp]
    t = torch.as_tensor(v)
    if tuple(t.shape) != lf.shape:
      raise ValueError(f"leaf {'/'.join(path)} has shape {tuple(t.shape)}, expected {lf.shape}")
    arena[lf.offset:lf.offset + lf.size] = t.reshape(-1).to(device=device, dtype=dtype)
    seen.add(cp)
  missing = set(layout.by_path) - seen
  if missing:
    raise KeyError(f"missing parameter leaves: {sorted('/'.join(p) for p in missing)[:5]} ...")
  return arena


def init_arena(layout: ArenaLayout, seed: int, device, *, nonzero_adaln: bool = False) -> torch.Tensor:
  """Initialisers of SURVEY.md App. C (flax defaults: lecun_normal for Conv/Dense, xavier_uniform where
  the reference asks for it, zeros for the adaLN projections).  nonzero_adaln=True draws the adaLN /
  final-modulation kernels from N(0, 0.02) instead (parity and throughput runs: zero-init turns every
  block into the identity)."""
  g = torch.Generator(device="cpu").manual_seed(int(seed))
  arena = torch.zeros(layout.total, dtype=torch.float32)
  for lf in layout.leaves:
    kind = lf.init
    n = lf.size
    if kind == "zeros":
      continue
    if kind == "ones":
      v = torch.ones(n)
    elif kind == "adaln":
      v = torch.randn(n, generator=g) * 0.02 if nonzero_adaln else torch.zeros(n)
    elif kind[0] == "normal":
      v = torch.randn(n, generator=g) * kind[1]
    elif kind[0] == "lecun":
      std = math.sqrt(1.0 / kind[1]) / 0.87962566103423978
      v = torch.fmod(torch.randn(n, generator=g), 2.0) * std
    elif kind[0] == "xavier":
      lim = math.sqrt(6.0 / (kind[1] + kind[2]))
      v = (torch.rand(n, generator=g) * 2 - 1) * lim
    else:  # pragma: no cover
      raise ValueError(kind)
    arena[lf.offset:lf.offset + n] = v
  return arena.to(device)
