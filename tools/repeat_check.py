"""Run-to-run difference of the gradient arena for one full-size step (tests/test_properties_gpu.py), printed
instead of asserted; used to attribute nondeterminism (fp32 atomics) to kernels via their env switches."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import util as U
from tests.test_properties_gpu import _grads
from small_vision_b200.config import TrainConfig
from small_vision_b200.params import tree_from_arena

model, _ = U.make_models("B/4", adaln=True)
tcfg = TrainConfig(batch_size=16, total_steps=1000, warmup_steps=10)
params = U.perturb_init(model, 1, "cuda")
batch, rand = U.make_batch(model, 16, n_noise=8, seed=3)
gs = []
for i in range(int(os.environ.get("REPS", "4"))):
  g, l = _grads(model, tcfg, tree_from_arena(model.layout, params.arena.clone()), batch, rand)
  gs.append(g.clone())
rels = [float((gs[0] - g).double().norm() / gs[0].double().norm()) for g in gs[1:]]
print(" ".join(f"{r:.2e}" for r in rels))

# which leaves carry the difference (largest run-to-run delta against run 0)
worst = max(range(1, len(gs)), key=lambda i: rels[i - 1])
d = (gs[0] - gs[worst])
rows = []
for lf in model.layout.leaves:
  a, b = lf.view(gs[0]).double().reshape(-1), lf.view(d).double().reshape(-1)
  if float(b.abs().max()) > 0:
    rows.append((float(b.norm() / (a.norm() + 1e-30)), "/".join(lf.path), int((b != 0).sum()), lf.size))
rows.sort(reverse=True)
print(f"leaves that differ in run {worst} vs run 0: {len(rows)} of {len(model.layout.leaves)}")
for r in rows[:14]:
  print(f"  rel {r[0]:.2e}  {r[1]}  ({r[2]} of {r[3]} elements)")
