"""The oracle against the reference's own source on the EDGE configurations of update_fn (train_ae.py:304-361): all-clean
and all-noised batches, an odd batch, mask ratios away from the recipe's, the unmasked noise branch, 16-token images, one
channel, patch 2, width 768, the prepended-token model with labels and label drops, a class-conditional model trained
without labels, both schedules, t = 0 and t = 999.  The fixture (tests/golden/reference_sweep_golden.json) holds what the
reference's unmodified files computed over tests/golden/refshim (generator: tests/golden/make_reference_sweep_golden.py);
the oracle runs in float64 on the same parameters, inputs and supplied draws.

Where /root/reference exists (the build container) the reference is executed again and must reproduce the committed
losses, so the fixture cannot drift from the generator or from the reference."""
import json
import os

import pytest
import torch

from tests.golden import make_reference_golden as RG
from tests.golden import make_reference_sweep_golden as RS
from tests.test_reference_golden_cpu import oracle_loss, to64

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "reference_sweep_golden.json")) as f:
  GOLD = json.load(f)["cases"]
F64 = torch.float64


def test_fixture_covers_the_sweep():
  assert sorted(GOLD) == sorted(RS.SWEEP)
  splits = {(c["n_noise"], c["n_clean"]) for c in GOLD.values()}
  assert (0, 3) in splits and (3, 0) in splits and (3, 2) in splits and (6, 2) in splits      # all clean, all noised, odd, quarter


def rebuild(name):
  model, ocfg, tkw, params, batch, rand, n_noise = RS.make_inputs(name)
  want = GOLD[name]
  assert RG.digest(batch["image"]) + RG.digest(rand["mask_noise_noise"]) + RG.digest(rand["mask_noise_clean"]) == want["input_digest"]
  assert RG.digest(torch.cat([v.reshape(-1) for _, v in sorted(RG.flatten(params).items())])) == want["param_digest"]
  assert n_noise == want["n_noise"]
  return model, ocfg, tkw, params, batch, rand, n_noise, want


def check_branch(aux, br, want, patch):
  pred = aux[f"pred_{br}"].detach()
  out = aux[f"out_{br}"]
  T = lambda v: torch.tensor(v, dtype=F64)
  assert torch.allclose(pred.mean(dim=(1, 2)), T(want["pred_sample_means"]), rtol=0, atol=1e-11)
  assert abs(float(pred.abs().mean()) - want["pred_abs_mean"]) <= 1e-11
  pl = out["pre_logits"].detach()
  assert torch.allclose(pl.mean(dim=1), T(want["pre_logits_sample_means"]), rtol=0, atol=1e-11)
  assert abs(float(pl.abs().mean()) - want["pre_logits_abs_mean"]) <= 1e-11
  if "patch_mask" in want:
    seq = out["mask"][:, ::patch, ::patch, 0].reshape(pred.shape[0], -1)
    got = ["".join(str(int(v)) for v in row) for row in seq.to(torch.uint8).tolist()]
    assert got == want["patch_mask"]                      # bit for bit, ties included
  else:
    assert out["mask"] is None


@pytest.mark.parametrize("name", sorted(RS.SWEEP))
def test_oracle_matches_reference_source_on_edge_configurations(name):
  model, ocfg, tkw, params, batch, rand, n_noise, want = rebuild(name)
  schedule = RS.case_config(name)[4]
  loss, aux, _ = oracle_loss(to64(params), ocfg, tkw, batch, rand, n_noise, schedule)
  assert abs(float(loss) - want["loss"]) <= 1e-11 * abs(want["loss"]) + 1e-12, (float(loss), want["loss"])
  for br in ("noise", "clean"):
    if br in want:
      check_branch(aux, br, want[br], ocfg["patch_size"][0])
    else:
      assert f"pred_{br}" not in aux


@pytest.mark.parametrize("name", sorted(RS.SWEEP))
def test_oracle_gradient_reproduces_reference_loss_slopes_on_edge_configurations(name):
  model, ocfg, tkw, params, batch, rand, n_noise, want = rebuild(name)
  p64 = to64(params, grad=True)
  loss, _, _ = oracle_loss(p64, ocfg, tkw, batch, rand, n_noise, RS.case_config(name)[4])
  loss.backward()
  grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in RG.flatten(p64).items()}
  gnorm = float(torch.sqrt(sum((g ** 2).sum() for g in grads.values())))
  for (_, d), (slope, err) in zip(RS.directions(params), want["slopes"]):
    mine = float(sum((grads[k] * d[k]).sum() for k in grads))
    assert abs(mine - slope) <= 1e-6 * abs(slope) + 10 * err + 1e-9 * gnorm, (name, mine, slope, err)


LIVE = ["all_clean", "all_noise_unmasked", "cond_token_with_labels", "odd_batch_5", "patch2_four_channels_linear", "sixteen_tokens"]


@pytest.mark.skipif(not os.path.isdir(os.path.join(RG.REF, "big_vision")), reason="needs the reference tree (build container only)")
def test_committed_losses_are_what_the_reference_computes_now():
  """Re-executes the reference (forward of both branches inside its own loss_fn) on the sweep cases named below."""
  import subprocess
  import sys
  # in a child process: the shim shadows the names `jax` / `flax` on sys.path, which must not leak into this test session
  code = ("import json, sys; sys.path.insert(0, %r)\n"
          "from tests.golden import make_reference_golden as RG, make_reference_sweep_golden as RS\n"
          "ae, gdm, code, _ = RG.load_reference()\n"
          "print(json.dumps({n: RS.run_reference(n, ae, gdm, code, loss_only=True)['loss'] for n in %r}))\n") % (os.path.dirname(HERE), LIVE)
  r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900)
  assert r.returncode == 0, r.stderr[-2000:]
  live = json.loads(r.stdout.strip().splitlines()[-1])
  for name in LIVE:
    want = GOLD[name]
    assert abs(live[name] - want["loss"]) <= 1e-13 * abs(want["loss"]), (name, live[name], want["loss"])
