"""Generates tests/golden/umd_golden.pt: small known-answer vectors for the hot path.

PROVENANCE: these vectors come from the CPU oracle (oracle/umd_oracle.py, fp32), NOT from the reference
itself — philippe-eecs/small-vision is JAX/Flax/Optax code that cannot be imported in this image and ships no
golden vectors for this path.  They pin the oracle against silent drift (CPU test) and give
the GPU tests a fixture that needs no oracle run.  Integer outputs (mask permutations) are exact by definition
of a stable argsort; floating-point outputs carry the tolerances written in tests/test_golden*.py.

  python tests/golden/make_golden.py
"""
import hashlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import umd_oracle as O  # noqa: E402
from tests import util as U  # noqa: E402

CASES = {
    # name: (model kwargs, train kwargs, batch, n_noise)
    "umd_s4": (dict(variant="S/4", adaln=True, depth=2, dec_depth=1),
               dict(mask_ratio=0.375, mask_ratio_no_noise=0.75, no_noise_prob=0.5, use_labels=False), 4, 2),
    "mae_s4": (dict(variant="S/4", adaln=False, depth=2, dec_depth=1),
               dict(mask_ratio=0.375, mask_ratio_no_noise=0.75, no_noise_prob=0.5, use_labels=False), 4, 2),
    "dit_s4": (dict(variant="S/4", adaln=True, num_classes=10, depth=2, dec_depth=1),
               dict(mask_ratio=0.0, mask_ratio_no_noise=0.75, no_noise_prob=0.0, use_labels=True), 4, 4),
}
HP = dict(clip_norm=1.0, peak_lr=2e-3, warmup_steps=0, total_steps=1000, b1=0.9, b2=0.95, wd=0.05)


def digest(t):
  return hashlib.sha256(t.contiguous().numpy().tobytes()).hexdigest()[:16]


def build_case(name):
  mkw, tkw, B, n_noise = CASES[name]
  model, ocfg = U.make_models(**mkw)
  params = U.cpu_tree(U.perturb_init(model, 0, "cpu"))
  batch, rand = U.make_batch(model, B, n_noise=n_noise, seed=100, use_labels=tkw["use_labels"], device="cpu")
  st = {"params": params, "gd": O.gaussian_diffusion_tables(), "opt": O.init_opt_state(params)}
  st2, meas, ex = O.update_step(st, batch, ocfg, tkw, HP, rand)
  fg = O.flatten_tree(ex["grads"])
  out = {
      "input_digest": digest(batch["image"]) + digest(rand["noise"]) + digest(rand["mask_noise_clean"]),
      "param_digest": digest(torch.cat([v.reshape(-1) for _, v in sorted(O.flatten_tree(params).items())])),
      "training_loss": meas["training_loss"], "grad_norm": meas["grad_norm"],
      "l2_params": meas["l2_params"], "l2_updates": meas["l2_updates"],
      "grad_leaf_norms": {"/".join(k): float(v.double().norm()) for k, v in fg.items()},
      "grad_final_conv_bias": fg[("final_conv", "bias")].clone(),
      "x_t_head": ex["x_t"].reshape(-1)[:64].clone(),
  }
  for br in ("noise", "clean"):
    if f"pred_{br}" in ex["aux"]:
      pred = ex["aux"][f"pred_{br}"].detach()
      out[f"pred_{br}_sample0_row0"] = pred[0, 0].clone()           # [W, 2C]
      out[f"pred_{br}_abs_mean"] = float(pred.abs().mean())
      ir = ex["aux"][f"out_{br}"]["ids_restore"]
      if ir is not None:
        out[f"ids_restore_{br}"] = ir.to(torch.int16).clone()
      out[f"pre_logits_{br}"] = ex["aux"][f"out_{br}"]["pre_logits"].detach()[:, :32].clone()
  return out


def masking_vectors():
  g = torch.Generator().manual_seed(7)
  noise = torch.rand(6, 256, generator=g)
  noise[0, 5] = noise[0, 200]
  noise[1, 10:20] = noise[1, 30]
  noise[2, :] = 0.5
  noise[3] = torch.linspace(1, 0, 256)
  ids_shuffle = torch.argsort(noise, dim=1, stable=True)
  return {"noise": noise, "ids_shuffle": ids_shuffle.to(torch.int16),
          "ids_restore": torch.argsort(ids_shuffle, dim=1, stable=True).to(torch.int16)}


def main():
  gold = {"provenance": "oracle/umd_oracle.py fp32 on CPU (torch %s); not the JAX reference" % torch.__version__,
          "cases": {n: build_case(n) for n in CASES}, "masking": masking_vectors()}
  path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "umd_golden.pt")
  torch.save(gold, path)
  print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
  main()
