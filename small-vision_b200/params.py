"""Parameter tree <-> flat arena.

The reference keeps parameters as a Flax nested dict (shapes from
big_vision/models/ae.py:57-97, vit.py:60-163, embeddings.py).  Here the same nested dict is a set
of *views* into one flat fp32 arena, so that the CUDA engine, the optimiser and the gradient
all-reduce each see one contiguous buffer.

Arena order (= the order in which gradients become final during backward, which is what the bucketed
all-reduce overlaps with; sharding.py here, train_ae.py:287-290,364 there):
    decoder side | embeddings / conditioning | Encoder layer 0 .. depth-1 | Encoder/encoder_norm
The scanned blocks are stored LAYER-MAJOR: all leaves of one layer are contiguous (a "layer block"), the
blocks of a stack follow each other at a constant stride.  Flax stacks every scanned leaf over depth
(`[depth, ...]`, vit.py:131-148), which would make no encoder gradient final before layer 0's backward has
run; here the tree leaf of a scanned parameter is a STRIDED view `[depth, ...]` with stride(0) = the layer
stride, so the tree still has the reference's shapes while the gradients of layer l are one contiguous range
that can be all-reduced as soon as layer l's backward has been enqueued.
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
P_ENC_LAYER_STRIDE, P_DEC_LAYER_STRIDE = P_COUNT, P_COUNT + 1   # trailing entries of the offsets array
OFFSETS_LEN = P_COUNT + 2
(S_ADA_W, S_ADA_B, S_LN0_S, S_LN0_B, S_LN1_S, S_LN1_B, S_Q_W, S_K_W, S_V_W, S_Q_B, S_K_B, S_V_B, S_O_W, S_O_B,
 S_FC1_W, S_FC1_B, S_FC2_W, S_FC2_B, S_NORM_S, S_NORM_B) = range(20)

SCAN_NAME = "ScanCheckpointEncoder1DBlock_0"  # flax auto-name of nn.scan(nn.remat(Encoder1DBlock)) (vit.py:131-148)
ALIGN = 64  # elements; keeps every leaf 256-byte aligned (TMA needs 16) and wd flags per 64-element block exact


class Leaf:
  """One parameter leaf.  `stack` is None for a plain leaf (contiguous at `offset`), else "Encoder" / "Decoder": then
  shape[0] is the depth, layer l's slice of `psize` elements starts at offset + l * lstride."""
  __slots__ = ("leaf_id", "path", "shape", "init", "bucket", "offset", "size", "stack", "psize", "lstride")

  def __init__(self, leaf_id, path, shape, init, bucket, stack=None):
    self.leaf_id, self.path, self.shape, self.init, self.bucket = leaf_id, tuple(path), tuple(shape), init, bucket
    self.offset = -1
    self.size = int(math.prod(shape))
    self.stack = stack
    self.psize = self.size // shape[0] if stack else self.size
    self.lstride = 0

  def view(self, arena: torch.Tensor) -> torch.Tensor:
    """This leaf of `arena` (any flat tensor in the arena layout: parameters, gradients, mu, nu) in its tree shape."""
    if not self.stack:
      return arena[self.offset:self.offset + self.size].view(self.shape)
    inner = self.shape[1:]
    strides, acc = [], 1
    for d in reversed(inner):
      strides.append(acc)
      acc *= d
    return arena.as_strided(self.shape, (self.lstride, *reversed(strides)), arena.storage_offset() + self.offset)


def _stack_leaves(cfg: ModelConfig, name: str, base: int, depth: int, bucket: int) -> List[Leaf]:
  D, H, M = cfg.width, cfg.num_heads, cfg.mlp
  Dh = D // H
  blk = (name, SCAN_NAME)
  att = blk + ("MultiHeadDotProductAttention_0",)
  out = []
  if cfg.adaln:
    out += [Leaf(base + S_ADA_W, blk + ("Dense_0", "kernel"), (depth, D, 6 * D), "adaln", bucket, name),
            Leaf(base + S_ADA_B, blk + ("Dense_0", "bias"), (depth, 6 * D), "zeros", bucket, name)]
  out += [
      Leaf(base + S_LN0_S, blk + ("LayerNorm_0", "scale"), (depth, D), "ones", bucket, name),
      Leaf(base + S_LN0_B, blk + ("LayerNorm_0", "bias"), (depth, D), "zeros", bucket, name),
      Leaf(base + S_LN1_S, blk + ("LayerNorm_1", "scale"), (depth, D), "ones", bucket, name),
      Leaf(base + S_LN1_B, blk + ("LayerNorm_1", "bias"), (depth, D), "zeros", bucket, name),
      # query / key / value kept adjacent and equally sized: the engine runs them as one batch-3 GEMM
      Leaf(base + S_Q_W, att + ("query", "kernel"), (depth, D, H, Dh), ("xavier", D, D), bucket, name),
      Leaf(base + S_K_W, att + ("key", "kernel"), (depth, D, H, Dh), ("xavier", D, D), bucket, name),
      Leaf(base + S_V_W, att + ("value", "kernel"), (depth, D, H, Dh), ("xavier", D, D), bucket, name),
      Leaf(base + S_Q_B, att + ("query", "bias"), (depth, H, Dh), "zeros", bucket, name),
      Leaf(base + S_K_B, att + ("key", "bias"), (depth, H, Dh), "zeros", bucket, name),
      Leaf(base + S_V_B, att + ("value", "bias"), (depth, H, Dh), "zeros", bucket, name),
      Leaf(base + S_O_W, att + ("out", "kernel"), (depth, H, Dh, D), ("xavier", D, D), bucket, name),
      Leaf(base + S_O_B, att + ("out", "bias"), (depth, D), "zeros", bucket, name),
      Leaf(base + S_FC1_W, blk + ("MlpBlock_0", "Dense_0", "kernel"), (depth, D, M), ("xavier", D, M), bucket, name),
      Leaf(base + S_FC1_B, blk + ("MlpBlock_0", "Dense_0", "bias"), (depth, M), ("normal", 1e-6), bucket, name),
      Leaf(base + S_FC2_W, blk + ("MlpBlock_0", "Dense_1", "kernel"), (depth, M, D), ("xavier", M, D), bucket, name),
      Leaf(base + S_FC2_B, blk + ("MlpBlock_0", "Dense_1", "bias"), (depth, D), ("normal", 1e-6), bucket, name),
      Leaf(base + S_NORM_S, (name, "encoder_norm", "scale"), (D,), "ones", bucket),
      Leaf(base + S_NORM_B, (name, "encoder_norm", "bias"), (D,), "zeros", bucket),
  ]
  return out


def leaf_specs(cfg: ModelConfig) -> List[Leaf]:
  """All leaves in arena order: decoder side, embeddings / conditioning ("rest"), encoder."""
  D, L, p, C = cfg.width, cfg.num_patches, cfg.patch, cfg.channels
  leaves: List[Leaf] = []
  leaves += [Leaf(P_FCONV_W, ("final_conv", "kernel"), (p, p, D, 2 * C), ("normal", 0.02), "dec"),
             Leaf(P_FCONV_B, ("final_conv", "bias"), (2 * C,), "zeros", "dec")]
  if cfg.adaln:
    leaves += [Leaf(P_FMOD_W, ("final_modulation", "kernel"), (D, 2 * D), "adaln", "dec"),
               Leaf(P_FMOD_B, ("final_modulation", "bias"), (2 * D,), "zeros", "dec")]
  leaves += _stack_leaves(cfg, "Decoder", P_DEC_BASE, cfg.dec_depth, "dec")
  leaves += [Leaf(P_DEC_POS, ("dec_pos_embedding",), (1, L, D), ("normal", 1 / math.sqrt(L)), "dec"),
             Leaf(P_MASK_TOKEN, ("image_mask_embedding",), (1, 1, D), ("normal", 0.02), "dec")]
  leaves += [Leaf(P_CLS, ("cls",), (1, cfg.num_cls, D), "zeros", "rest"),
             Leaf(P_POS, ("pos_embedding",), (1, L, D), ("normal", 1 / math.sqrt(L)), "rest"),
             Leaf(P_EMBED_W, ("embedding", "kernel"), (p, p, C, D), ("lecun", p * p * C), "rest"),
             Leaf(P_EMBED_B, ("embedding", "bias"), (D,), "zeros", "rest"),
             Leaf(P_TT_W0, ("time_trunk", "Dense_0", "kernel"), (D, 2 * D), ("lecun", D), "rest"),
             Leaf(P_TT_B0, ("time_trunk", "Dense_0", "bias"), (2 * D,), "zeros", "rest"),
             Leaf(P_TT_W1, ("time_trunk", "Dense_1", "kernel"), (2 * D, D), ("lecun", 2 * D), "rest"),
             Leaf(P_TT_B1, ("time_trunk", "Dense_1", "bias"), (D,), "zeros", "rest")]
  if cfg.num_classes is not None:
    nc = cfg.num_classes
    leaves += [Leaf(P_LABEL_TABLE, ("label_emb", "embedding", "embedding"), (nc + 1, D), ("normal", 1 / math.sqrt(D)), "rest"),
               Leaf(P_LT_W0, ("label_trunk", "Dense_0", "kernel"), (D, 2 * D), ("lecun", D), "rest"),
               Leaf(P_LT_B0, ("label_trunk", "Dense_0", "bias"), (2 * D,), "zeros", "rest"),
               Leaf(P_LT_W1, ("label_trunk", "Dense_1", "kernel"), (2 * D, D), ("lecun", 2 * D), "rest"),
               Leaf(P_LT_B1, ("label_trunk", "Dense_1", "bias"), (D,), "zeros", "rest")]
  leaves += _stack_leaves(cfg, "Encoder", P_ENC_BASE, cfg.depth, "enc")
  return leaves


def _align(n: int) -> int:
  return (n + ALIGN - 1) // ALIGN * ALIGN


ENC_LAYERS_PER_BUCKET = 3   # encoder gradient buckets: ~32 M parameters (127 MB fp32) each for B/4


class ArenaLayout:
  """Offsets of every leaf in the flat arena, the layer strides of the two stacks, the all-reduce buckets and the
  weight-decay flags.

  Backward events (the engine's bucket callback, include/umd_b200.h): 0 = the decoder side is final; 1 + j = the j-th
  encoder layer in backward order (layer depth-1-j) is final; 1 + depth = everything is final.  `bucket_bounds[i]` is a
  contiguous arena range that may be all-reduced once event `bucket_events[i]` has been signalled, listed in that order:
  the decoder side, then groups of ENC_LAYERS_PER_BUCKET encoder layers from the top down (the first one also carries
  Encoder/encoder_norm and whatever trails the arena), and last — the only bucket whose reduction nothing overlaps, so
  it is kept small — layer 0 together with the embeddings / conditioning leaves that precede it in the arena (their
  gradients are the last to become final)."""

  def __init__(self, cfg: ModelConfig):
    self.cfg = cfg
    self.leaves = leaf_specs(cfg)
    off = 0
    self.layer_stride = {"Encoder": 0, "Decoder": 0}
    self.stack_start = {}
    i = 0
    group_start = {}
    while i < len(self.leaves):
      lf = self.leaves[i]
      group_start.setdefault(lf.bucket, off)
      if lf.stack:
        j = i
        inner = 0
        while j < len(self.leaves) and self.leaves[j].stack == lf.stack:
          self.leaves[j].offset = off + inner
          inner += _align(self.leaves[j].psize)
          j += 1
        depth = lf.shape[0]
        for k in range(i, j):
          self.leaves[k].lstride = inner
        self.layer_stride[lf.stack] = inner
        self.stack_start[lf.stack] = off
        off += depth * inner
        i = j
      else:
        lf.offset = off
        off += _align(lf.size)
        i += 1
    self.total = off
    # order in which the seeded initialisers consume their generator: independent of the arena order, so that a seed
    # keeps producing the same tree whatever the layout (the committed golden vectors were made with this order)
    self.init_order = [lf for grp in ("dec", "enc", "rest") for lf in self.leaves if lf.bucket == grp]
    self.num_params = sum(lf.size for lf in self.leaves)
    self.by_path: Dict[Tuple[str, ...], Leaf] = {lf.path: lf for lf in self.leaves}
    self.offsets = [-1] * OFFSETS_LEN
    for lf in self.leaves:
      self.offsets[lf.leaf_id] = lf.offset
    self.offsets[P_ENC_LAYER_STRIDE] = self.layer_stride["Encoder"]
    self.offsets[P_DEC_LAYER_STRIDE] = self.layer_stride["Decoder"]
    # ---- all-reduce buckets in launch order
    depth, k = cfg.depth, ENC_LAYERS_PER_BUCKET
    enc0, es = self.stack_start["Encoder"], self.layer_stride["Encoder"]
    rest0 = group_start["rest"]
    self.num_events = depth + 2
    self.bucket_bounds: List[Tuple[int, int]] = [(0, rest0)]
    self.bucket_events: List[int] = [0]
    hi_layer, hi = depth, self.total
    while hi_layer > 0:
      lo_layer = max(hi_layer - k, 0)
      if lo_layer == 0 and hi_layer > 1:
        lo_layer = 1      # keep the bucket that cannot overlap anything small: layer 0 alone travels with the embeddings
      if lo_layer == 0:   # the last bucket: layer 0 merged with the embeddings / conditioning leaves in front of it
        self.bucket_bounds.append((rest0, hi))
        self.bucket_events.append(depth + 1)
      else:
        self.bucket_bounds.append((enc0 + lo_layer * es, hi))
        self.bucket_events.append(1 + (depth - 1 - lo_layer))
      hi_layer, hi = lo_layer, enc0 + lo_layer * es

  def decay(self, lf: Leaf) -> bool:
    """train_ae.py:125-134: decayed iff no path component is in model.no_decay_list."""
    return all(k not in self.cfg.no_decay_list for k in lf.path)

  def wd_flags(self, device) -> torch.Tensor:
    flags = torch.zeros(self.total // ALIGN, dtype=torch.uint8)
    for lf in self.leaves:
      if self.decay(lf):
        for l in range(lf.shape[0] if lf.stack else 1):
          o = lf.offset + l * lf.lstride
          flags[o // ALIGN:(o + lf.psize + ALIGN - 1) // ALIGN] = 1
    return flags.to(device)

  def offsets_tensor(self):
    import ctypes as C
    return (C.c_longlong * OFFSETS_LEN)(*self.offsets)


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
    d[lf.path[-1]] = lf.view(arena)
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
    lf.view(arena).copy_(t.to(device=device, dtype=dtype))
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
  for lf in layout.init_order:
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
    lf.view(arena).copy_(v.view(lf.shape))
  return arena.to(device)
