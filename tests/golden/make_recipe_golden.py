"""Generates tests/golden/reference_recipe_golden.json: the training recipes of the BASELINE.json configurations as THE
REFERENCE'S OWN CODE resolves them —
  * `big_vision/configs/ae_i1k.py::get_config(arg)` (with `configs/common.py::parse_arg`, `configs/common_fewshot.py`),
    imported unmodified over a stand-in for `ml_collections.ConfigDict` (tests/golden/refshim/ml_collections);
  * `big_vision/utils.py::steps` (lifted with `ast`) for `total_steps` exactly as train_ae.py:77 calls it;
  * the optimiser wiring block of train_ae.py:124-151 (the `else:` branch of the adafactor test, lifted with `ast`) run
    against a recording `optax` stub, which captures what the reference passes to `warmup_cosine_decay_schedule`, `adamw`,
    `clip_by_global_norm` and `chain`, and the reference's own `get_weight_decay_mask` evaluated on the parameter tree.
Nothing numerical of optax runs (it is absent): this pins the hyper-parameter wiring, not the AdamW arithmetic.

  python tests/golden/make_recipe_golden.py       (build container only: needs /root/reference)
"""
import ast
import json
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("UMD_REFERENCE_ROOT", "/root/reference")
sys.path.insert(0, ROOT)

NTRAIN = 1_268_355      # imagenet2012 train[:99%] (configs/ae_i1k.py:57-58): 1 281 167 * 0.99, tfds rounds the boundary up
RECIPES = {             # BASELINE.json configs -> the arg string a launcher would pass to ae_i1k.py
    "umd_b4_bs4096": "variant=B/4,batch_size=4096",
    "umd_s4_bs32": "variant=S/4,batch_size=32",
    "mae_b4": "variant=B/4,batch_size=4096,adaln=False",
    "dit_b4_labels": "variant=B/4,batch_size=256,use_labels=True,mask_ratio=0.0,no_noise_prob=0.0",
    "latent_umd_l2": "variant=L/2,batch_size=1024,size=256,latent_diffusion=True",
    "default": "",
}


def flatten_dict(d, prefix=()):
  out = {}
  for k, v in d.items():
    if isinstance(v, dict):
      out.update(flatten_dict(v, prefix + (k,)))
    else:
      out[prefix + (k,)] = v
  return out


def unflatten_dict(flat):
  out = {}
  for path, v in flat.items():
    d = out
    for k in path[:-1]:
      d = d.setdefault(k, {})
    d[path[-1]] = v
  return out


def init_log(model, engine_model, mkw):
  """One forward pass of the reference model over zero parameters of the engine layout's shapes (the stand-in asserts
  every shape the reference asks for) while the stand-in records which initialiser the reference hands to each leaf
  (None = the Flax layer's default)."""
  import numpy as np
  import jax
  import jax.numpy as jnp
  import flax.linen as nn
  tree = {}
  for lf in engine_model.layout.leaves:
    d = tree
    for k in lf.path[:-1]:
      d = d.setdefault(k, {})
    d[lf.path[-1]] = np.zeros(lf.shape)
  nn.INIT_LOG.clear()
  S, C = mkw["img_size"], mkw["channels"]
  L = engine_model.cfg.num_patches
  y = jnp.asarray(np.zeros((1,), np.int32)) if mkw.get("num_classes") else None
  model.apply({"params": tree}, jnp.asarray(np.zeros((1, S, S, C))), t=jnp.asarray(np.ones((1, 1), np.int32)), y=y, mask=0.375,
              rngs={"mae_noise": jax.Key({"uniform": np.linspace(0, 1, L)[None]})})
  assert set(nn.INIT_LOG) == {lf.path for lf in engine_model.layout.leaves}, set(nn.INIT_LOG) ^ {lf.path for lf in engine_model.layout.leaves}
  return {"/".join(k): {"layer": layer, "init": init} for k, (layer, init) in sorted(nn.INIT_LOG.items())}


def main():
  sys.path.insert(0, os.path.join(HERE, "refshim"))
  sys.path.insert(0, REF)
  for name in ("big_vision.utils", "big_vision.models.common"):   # imported by vit.py for checkpoint loading only
    sys.modules[name] = types.ModuleType(name)
  from big_vision.configs import ae_i1k
  from big_vision.models import ae
  assert ae_i1k.__file__.startswith(REF)
  # utils.steps
  upath = os.path.join(REF, "big_vision", "utils.py")
  utree = ast.parse(open(upath).read(), filename=upath)
  node = next(n for n in utree.body if isinstance(n, ast.FunctionDef) and n.name == "steps")
  uenv = {}
  exec(compile(ast.Module(body=[node], type_ignores=[]), upath, "exec"), uenv)
  # optimiser wiring block: the `else:` of `if 'adafactor' in config.optax_name` (train_ae.py:119-152)
  tpath = os.path.join(REF, "big_vision", "trainers", "train_ae.py")
  ttree = ast.parse(open(tpath).read(), filename=tpath)
  blk = next(n for n in ast.walk(ttree) if isinstance(n, ast.If) and "adafactor" in ast.unparse(n.test))
  body = [s for s in blk.orelse if "eval_shape" not in ast.unparse(s)]           # drop the jax.eval_shape(tx.init, ...) line
  code = compile(ast.Module(body=body, type_ignores=[]), tpath, "exec")
  lines = (blk.orelse[0].lineno, blk.orelse[-1].end_lineno)

  from small_vision_b200.model import Model as EngineModel
  gold = {"provenance": "configs/ae_i1k.py + utils.steps + train_ae.py:%d-%d executed over stand-ins" % lines, "ntrain_img": NTRAIN,
          "recipes": {}}
  for name, arg in RECIPES.items():
    config = ae_i1k.get_config(arg)
    batch_size = config.input.batch_size
    total_steps = uenv["steps"]("total", config, NTRAIN, batch_size)                 # train_ae.py:77
    calls = {}
    optax = types.SimpleNamespace(
        warmup_cosine_decay_schedule=lambda **kw: calls.setdefault("schedule", kw) and "lr",
        adamw=lambda **kw: calls.setdefault("adamw", kw) and "adamw",
        clip_by_global_norm=lambda c: calls.setdefault("clip", c) and "clip",
        chain=lambda *a: calls.setdefault("chain", list(a)))
    flax = types.SimpleNamespace(traverse_util=types.SimpleNamespace(flatten_dict=flatten_dict, unflatten_dict=unflatten_dict),
                                 core=types.SimpleNamespace(frozen_dict=types.SimpleNamespace(unfreeze=lambda x: x)))
    model = ae.Model(**config.model.to_dict())                                     # for model.no_decay_list
    env = dict(optax=optax, flax=flax, config=config, batch_size=batch_size, ntrain_img=NTRAIN, total_steps=total_steps,
               model=model)
    exec(code, env)
    mkw = config.model.to_dict()
    em = EngineModel(**{k: v for k, v in mkw.items()})
    tree = {}
    for lf in em.layout.leaves:
      d = tree
      for k in lf.path[:-1]:
        d = d.setdefault(k, {})
      d[lf.path[-1]] = 0
    mask = flatten_dict(calls["adamw"]["mask"](tree))
    inits = init_log(model, em, mkw)
    gold["recipes"][name] = {
        "arg": arg, "model": {k: (list(v) if isinstance(v, tuple) else v) for k, v in mkw.items()},
        "batch_size": batch_size, "total_epochs": config.total_epochs, "total_steps": total_steps,
        "schedule": calls["schedule"], "clip_norm": calls["clip"], "chain": calls["chain"],
        "adamw": {k: (list(v) if isinstance(v, tuple) else v) for k, v in calls["adamw"].items() if k not in ("mask", "learning_rate")},
        "adamw_learning_rate_is_schedule": calls["adamw"]["learning_rate"] == "lr",
        "decay_mask": {"/".join(k): bool(v) for k, v in mask.items()},
        "train": {k: config.get(k) for k in ("no_noise_prob", "mask_ratio", "mask_ratio_no_noise", "use_labels", "num_classes",
                                             "ema_decay", "latent_diffusion")},
        "diffusion_space": list(config.diffusion_space),
        "diff_schedule": config.diff_schedule.to_dict(),
        "fewshot": {k: config.evals.fewshot.get(k) for k in ("shots", "l2_reg", "num_seeds", "representation_layer", "pred")},
        "input_pp": config.input.pp,
        "init": inits,
    }
    r = gold["recipes"][name]
    print(name, "steps", total_steps, "schedule", r["schedule"], "adamw", r["adamw"])
  dst = os.path.join(HERE, "reference_recipe_golden.json")
  json.dump(gold, open(dst, "w"), indent=0, sort_keys=True, default=lambda o: list(o) if isinstance(o, tuple) else str(o))
  print(dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
  main()
