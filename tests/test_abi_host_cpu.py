"""CPU-only checks of the drop-in boundary: the C-ABI library loads and exports every symbol declared in
include/umd_b200.h, argument validation fails loudly (no compute without a GPU), and the host-side mirror of
the reference interface (config, parameter tree, sharding declarations) matches SURVEY.md App. B/C."""
import ctypes as C
import os
import re
import subprocess

import pytest
import torch

from small_vision_b200 import lib
from small_vision_b200.config import TrainConfig, decode_variant, make_model_config
from small_vision_b200.params import ArenaLayout, arena_from_tree, init_arena, tree_from_arena
from small_vision_b200.sharding import Mesh, PartitionSpec, infer_sharding, local_batch_slice, check_batch_divisible

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "umd_b200.h")


@pytest.fixture(scope="module")
def L():
  if not os.path.exists(lib.LIB_PATH):
    subprocess.run(["make", "-C", os.path.join(ROOT, "small-vision_b200", "csrc"), "-j", "8"], check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
  return lib.load()


def _declared_functions():
  src = open(HEADER).read()
  src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
  return sorted(set(re.findall(r"\b(umd_[a-z0-9_]+)\s*\(", src)) - {"umd_bucket_cb"})


def test_library_exports_every_declared_symbol(L):
  names = _declared_functions()
  assert len(names) >= 20, names
  missing = [n for n in names if not hasattr(L, n)]
  assert not missing, f"declared in include/umd_b200.h but not exported: {missing}"


def test_header_is_plain_c(tmp_path):
  """The drop-in boundary is a C ABI: include/umd_b200.h must compile as strict C99 on its own (no C++ or torch types)."""
  import shutil
  gcc = shutil.which("gcc")
  if gcc is None:
    pytest.skip("gcc not present")
  src = tmp_path / "abi.c"
  src.write_text('#include "umd_b200.h"\nint main(void) { return (int)sizeof(umd_model_cfg) == 0; }\n')
  r = subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I", os.path.dirname(HEADER),
                      str(src)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
  assert r.returncode == 0, r.stdout


def test_library_is_sm100a_native():
  """cuobjdump shows tcgen05 / TMA SASS mnemonics (B200_PROFILING.md 'What proves a Blackwell-native kernel')."""
  exe = "/usr/local/cuda/bin/cuobjdump"
  if not (os.path.exists(exe) and os.path.exists(lib.LIB_PATH)):
    pytest.skip("cuobjdump or library not present")
  sass = subprocess.run([exe, "-sass", lib.LIB_PATH], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
  assert "sm_100a" in sass
  for mnem in ("UTCHMMA", "UTMALDG", "LDTM"):
    assert mnem in sass, mnem


def test_error_paths_without_a_gpu(L):
  assert L.umd_version() >= 100
  assert L.umd_gemm_bf16(None, None) == 1                      # UMD_ERR_INVALID
  assert b"null args" in L.umd_last_error()
  a = lib.GemmArgs()
  a.M, a.N, a.K, a.batch = 128, 12, 64, 1                      # N not a multiple of 8
  assert L.umd_gemm_bf16(C.byref(a), None) == 1 and b"multiple of 8" in L.umd_last_error()
  assert lib.launch_count() >= 0


def test_workspace_bytes_is_host_arithmetic(L):
  cfg = make_model_config(variant="B/4", adaln=True)
  m = lib.model_cfg_struct(cfg)
  sh = lib.StepShape(256, 256, 160, 64, 1, 1)
  train = lib.workspace_bytes(m, sh, True)
  infer = lib.workspace_bytes(m, sh, False)
  assert 20e9 < train < 80e9 and infer < train / 4            # no remat: every activation of 16 blocks is kept (F9)
  bad = lib.StepShape(256, 256, 300, 64, 1, 1)
  with pytest.raises(lib.UmdError):
    lib.workspace_bytes(m, bad, True)
  m.width = 100
  with pytest.raises(lib.UmdError, match="width"):
    lib.workspace_bytes(m, sh, True)


def test_missing_library_fails_loudly(monkeypatch):
  monkeypatch.setattr(lib, "_lib", None)
  monkeypatch.setattr(lib, "LIB_PATH", "/nonexistent/libumd_b200.so")
  with pytest.raises(lib.UmdError, match="no CPU or PyTorch fallback"):
    lib.load()


def test_product_path_never_imports_the_oracle():
  pkg = os.path.join(ROOT, "small-vision_b200")
  for dp, _, fs in os.walk(pkg):
    for f in fs:
      if f.endswith((".py", ".cu", ".cuh")):
        txt = open(os.path.join(dp, f)).read()
        assert "umd_oracle" not in txt and "import oracle" not in txt and "from oracle" not in txt, f


# ------------------------------------------------------------------------------------------ config / variants
def test_decode_variant_and_len_keep():
  assert decode_variant("B/4") == dict(width=768, depth=12, dec_depth=4, num_heads=12, patch_size=(4, 4))
  assert decode_variant("L") == dict(width=1024, depth=24, dec_depth=8, num_heads=16)
  with pytest.raises(KeyError):
    decode_variant(":B/4")          # README.md:37 writes "variant=:B/4"; ae.py:205-215 rejects it the same way
  cfg = make_model_config(variant="B/4", adaln=True)
  assert (cfg.num_patches, cfg.len_keep(0.375), cfg.len_keep(0.75), cfg.len_keep(0.0)) == (256, 160, 64, 256)
  assert make_model_config(variant="L/2", img_size=32, channels=4).num_patches == 256
  assert make_model_config(variant="B/16", img_size=256).num_patches == 256


def test_train_config_schedule_constants():
  t = TrainConfig(batch_size=4096).resolved()
  assert t.scaled_peak_lr == pytest.approx(15e-5 * 16)
  # utils.steps rounds to the nearest step (utils.py:1059-1061; 247 725.6 -> 247 726, pinned by tests/test_reference_recipe_cpu.py)
  assert t.total_steps == round(800 * 1_268_355 / 4096) == 247_726 and t.warmup_steps == 40 * 1_268_355 // 4096


@pytest.mark.parametrize("kw,count", [
    (dict(variant="S/4", adaln=True), 43.73e6), (dict(variant="B/4", adaln=True), 174.16e6),
    (dict(variant="B/4", adaln=False), 116.28e6), (dict(variant="B/4", adaln=True, num_classes=1000), 177.29e6),
    (dict(variant="L/2", adaln=True, img_size=32, channels=4), 611.48e6)])
def test_parameter_counts_match_survey_appendix_b(kw, count):
  lay = ArenaLayout(make_model_config(**kw))
  assert lay.num_params == pytest.approx(count, rel=2e-4)


def test_parameter_tree_is_the_flax_layout():
  cfg = make_model_config(variant="S/4", adaln=True, num_classes=10)
  lay = ArenaLayout(cfg)
  tree = tree_from_arena(lay, init_arena(lay, 0, "cpu"))
  D, H = 384, 6
  blk = tree["Encoder"]["ScanCheckpointEncoder1DBlock_0"]
  assert tuple(tree["cls"].shape) == (1, 4, D) and tuple(tree["pos_embedding"].shape) == (1, 256, D)
  assert tuple(tree["embedding"]["kernel"].shape) == (4, 4, 3, D)
  assert tuple(blk["Dense_0"]["kernel"].shape) == (12, D, 6 * D)
  assert tuple(blk["MultiHeadDotProductAttention_0"]["query"]["kernel"].shape) == (12, D, H, 64)
  assert tuple(blk["MultiHeadDotProductAttention_0"]["out"]["kernel"].shape) == (12, H, 64, D)
  assert tuple(blk["MlpBlock_0"]["Dense_1"]["kernel"].shape) == (12, 4 * D, D)
  assert tuple(tree["Decoder"]["encoder_norm"]["scale"].shape) == (D,)
  assert tuple(tree["final_modulation"]["kernel"].shape) == (D, 2 * D)
  assert tuple(tree["final_conv"]["kernel"].shape) == (4, 4, D, 6)
  assert tuple(tree["label_emb"]["embedding"]["embedding"].shape) == (11, D)
  # zero-init leaves (vit.py:71, ae.py:94) and identity LayerNorm
  assert float(blk["Dense_0"]["kernel"].abs().max()) == 0 and float(tree["final_modulation"]["kernel"].abs().max()) == 0
  assert float(blk["LayerNorm_0"]["scale"].min()) == 1
  # leaves are views: writing the tree writes the arena; a re-named scan block still packs (flax-version dependent name)
  tree["cls"].fill_(2.0)
  lf = lay.by_path[("cls",)]
  assert float(tree.arena[lf.offset]) == 2.0
  plain = {k: v for k, v in tree.items()}
  plain["Encoder"] = {"SomeOtherScanName_0": tree["Encoder"]["ScanCheckpointEncoder1DBlock_0"],
                      "encoder_norm": tree["Encoder"]["encoder_norm"]}
  again = arena_from_tree(lay, plain, "cpu")
  assert torch.equal(again, tree.arena)
  with pytest.raises(ValueError):
    bad = dict(plain); bad["cls"] = torch.zeros(1, 3, D)
    arena_from_tree(lay, bad, "cpu")


def test_weight_decay_flags_follow_no_decay_list():
  cfg = make_model_config(variant="S/4", adaln=True)
  lay = ArenaLayout(cfg)
  dec = {"/".join(lf.path): lay.decay(lf) for lf in lay.leaves}
  assert dec["cls"] is False and dec["image_mask_embedding"] is False and dec["pos_embedding"] is True
  assert dec["Encoder/ScanCheckpointEncoder1DBlock_0/LayerNorm_0/scale"] is True      # LN scale IS decayed (a17)
  assert dec["Encoder/ScanCheckpointEncoder1DBlock_0/LayerNorm_0/bias"] is False
  assert dec["final_conv/kernel"] is True and dec["final_conv/bias"] is False
  flags = lay.wd_flags("cpu")
  assert flags.numel() * 64 == lay.total
  for lf in lay.leaves:
    for l in range(lf.shape[0] if lf.stack else 1):
      assert int(flags[(lf.offset + l * lf.lstride) // 64]) == int(lay.decay(lf))
  # buckets tile the arena; launch order = backward order: decoder side, encoder layer groups from the top down (the
  # first one ends the arena: it carries Encoder/encoder_norm and the trailing loss slot), the last group merged with
  # the embeddings / conditioning leaves in front of layer 0
  b, ev = lay.bucket_bounds, lay.bucket_events
  assert len(b) == 6 and ev == [0, 3, 6, 9, 11, 13] and lay.num_events == 14
  assert b[0][0] == 0 and b[1][1] == lay.total and b[-1][0] == b[0][1]
  cov = sorted(b)
  assert all(cov[i][1] == cov[i + 1][0] for i in range(len(cov) - 1)) and cov[-1][1] == lay.total
  enc0, es = lay.stack_start["Encoder"], lay.layer_stride["Encoder"]
  assert b[1][0] == enc0 + 9 * es and b[2] == (enc0 + 6 * es, enc0 + 9 * es) and b[4] == (enc0 + es, enc0 + 3 * es)
  assert b[5][1] == enc0 + es
  # a scanned leaf is a strided [depth, ...] view: layer l's slice is contiguous inside layer block l
  q = lay.by_path[("Encoder", "ScanCheckpointEncoder1DBlock_0", "MultiHeadDotProductAttention_0", "query", "kernel")]
  v = q.view(torch.arange(lay.total, dtype=torch.float32))
  assert tuple(v.shape) == (12, 384, 6, 64) and v.stride(0) == es and v[0].is_contiguous()
  assert float(v[5, 0, 0, 0]) == q.offset + 5 * es


# ------------------------------------------------------------------------------------------ sharding.py
def test_infer_sharding_strategies():
  mesh = Mesh(list(range(8)), ("data",))
  params = {"a": {"kernel": torch.zeros(768, 3072), "bias": torch.zeros(3072)}, "pos": torch.zeros(1, 256, 768),
            "odd": torch.zeros(1001, 769)}
  rep = infer_sharding(params, mesh, "data", "replicated", {})
  assert rep == {"a": {"kernel": PartitionSpec(), "bias": PartitionSpec()}, "pos": PartitionSpec(), "odd": PartitionSpec()}
  fs = infer_sharding(params, mesh, "data", "fully_sharded", {})
  assert fs["a"]["kernel"] == PartitionSpec(None, "data")     # largest divisible dim (sharding.py:70-76)
  assert fs["a"]["bias"] == PartitionSpec()                   # below the 2**18 threshold
  assert fs["pos"] == PartitionSpec()
  assert fs["odd"] == PartitionSpec()                         # no dim divisible by 8
  with pytest.raises(KeyError):
    infer_sharding(params, mesh, "data", "nope", {})


def test_batch_partition_rules():
  assert local_batch_slice(4096, 3, 8) == slice(1536, 2048)
  with pytest.raises(ValueError, match="divisible"):
    check_batch_divisible(100, 8)


def test_seed_from_rng_mixes_rank_and_stream():
  """train_state["rng"] is replicated; data-parallel ranks and the evaluator closures must not share draws."""
  import torch
  from small_vision_b200.diffusion import seed_from_rng
  rng = torch.tensor([1, 7], dtype=torch.int64)
  base = seed_from_rng(rng)
  assert base == seed_from_rng(rng.clone()) == 1 * 1_000_003 + 7
  seeds = {seed_from_rng(rng, rank=r, stream=s) for r in range(8) for s in range(3)}
  assert len(seeds) == 24
  assert seed_from_rng(torch.tensor([1, 8])) != base and seed_from_rng(5) == 5 and seed_from_rng(None) == 0
  assert all(0 <= x < 2 ** 63 for x in seeds)


def test_ctypes_structures_have_the_headers_layout(tmp_path):
  """Every ctypes.Structure of lib.py against the struct of the same role in include/umd_b200.h: gcc compiles a program
  that prints sizeof and the offset of every field BY NAME (a renamed or missing field is a compile error), and the numbers
  must equal ctypes' own.  A drifted field would otherwise shift every later pointer of umd_train_step's argument block."""
  import ctypes as C
  import subprocess
  from small_vision_b200 import lib
  pairs = [("umd_gemm_args", lib.GemmArgs), ("umd_adamw_args", lib.AdamwArgs), ("umd_model_cfg", lib.ModelCfg),
           ("umd_step_shape", lib.StepShape), ("umd_io", lib.IO), ("umd_train_step_args", lib.TrainStepArgs)]
  lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "umd_b200.h"', 'int main(void) {']
  want = []
  for cname, cls in pairs:
    lines.append(f'  printf("%zu\\n", sizeof({cname}));')
    want.append(C.sizeof(cls))
    for fname, _ in cls._fields_:
      lines.append(f'  printf("%zu\\n", offsetof({cname}, {fname}));')
      want.append(getattr(cls, fname).offset)
  from small_vision_b200 import params as P
  # enumerators, again by name: leaf ids (params.py), epilogue ids and flags (lib.py)
  enums = [("UMD_" + k, getattr(P, k)) for k in dir(P) if (k.startswith("P_") or k.startswith("S_")) and isinstance(getattr(P, k), int)]
  enums += [("UMD_OFFSETS_LEN", P.OFFSETS_LEN), ("UMD_STEP_NO_OPTIMIZER", lib.UMD_STEP_NO_OPTIMIZER)]
  enums += [("UMD_" + k, getattr(lib, k)) for k in dir(lib) if k.startswith("EPI_")]
  assert len(enums) >= 19 + 5 + 20 + 2 + 7
  for name, _ in enums:
    lines.append(f'  printf("%d\\n", (int){name});')
  lines += ['  return 0;', '}']
  src = tmp_path / "layout.c"
  src.write_text("\n".join(lines) + "\n")
  exe = tmp_path / "layout"
  inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")
  r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", inc, str(src), "-o", str(exe)], capture_output=True, text=True)
  assert r.returncode == 0, r.stderr
  out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split("\n")
  got = [int(x) for x in out[:len(want)]]
  assert got == want
  got_enums = [int(x) for x in out[len(want):len(want) + len(enums)]]
  assert got_enums == [v for _, v in enums], [(n, v, g) for (n, v), g in zip(enums, got_enums) if v != g]
