"""Generates tests/golden/reference_golden.pt by EXECUTING THE REFERENCE'S OWN SOURCE FILES.

PROVENANCE: unlike umd_golden.pt (oracle-generated), every number in this fixture is produced by the unmodified
reference code under /root/reference — `big_vision/models/{ae,vit,embeddings}.py`, `big_vision/gaussian_diffusion.py`
and the `loss_fn` closure of `big_vision/trainers/train_ae.py::update_fn` (train_ae.py:323-361, lifted out of the
enclosing function with `ast` at run time, not copied) — imported over `tests/golden/refshim/`, a numpy-float64
stand-in for the `jax` / `flax.linen` names those files use (JAX/Flax themselves are not installable here, SURVEY.md
F2).  What the shim restates (library layers) and what is the reference's own code (all wiring, masking, conditioning,
diffusion formulas, the loss) is listed in refshim/README.md.  Random draws are supplied through the shim's keys, so the
same draws can be given to the oracle and to the CUDA path.  Parameter gradients: the reference's loss is differenced
(central, fp64) along seeded parameter directions; the oracle's autograd gradient must reproduce those slopes.

Runs only in the build container (it needs /root/reference):
  python tests/golden/make_reference_golden.py            # reference_golden.pt (step cases + diffusion vectors)
  python tests/golden/make_reference_golden.py sampler    # reference_sampler_golden.pt (create_apply_fn + DDIM loop, evaluator fns)
"""
import ast
import hashlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("UMD_REFERENCE_ROOT", "/root/reference")
sys.path.insert(0, ROOT)

from tests import util as U  # noqa: E402  (inputs and parameters are drawn by the same helpers the tests use)

CASES = {
    # name: (model kwargs, train config, batch, n_noise)
    "umd_s4": (dict(variant="S/4", adaln=True, depth=2, dec_depth=1),
               dict(mask_ratio=0.375, mask_ratio_no_noise=0.75, no_noise_prob=0.5, use_labels=False), 4, 2),
    "mae_s4": (dict(variant="S/4", adaln=False, depth=2, dec_depth=1),
               dict(mask_ratio=0.375, mask_ratio_no_noise=0.75, no_noise_prob=0.5, use_labels=False), 4, 2),
    "dit_s4": (dict(variant="S/4", adaln=True, num_classes=10, depth=2, dec_depth=1),
               dict(mask_ratio=0.0, mask_ratio_no_noise=0.75, no_noise_prob=0.0, use_labels=True), 4, 4),
    "umd_lbl_s4": (dict(variant="S/4", adaln=True, num_classes=10, depth=1, dec_depth=1),
                   dict(mask_ratio=0.375, mask_ratio_no_noise=0.75, no_noise_prob=0.5, use_labels=True), 4, 2),
    # Latent-UMD recipe (configs/ae_i1k.py:45-51): 32x32x4 latents, patch 2, linear beta schedule
    "latent_s2": (dict(variant="S/2", adaln=True, depth=1, dec_depth=1, img_size=32, channels=4),
                  dict(mask_ratio=0.375, mask_ratio_no_noise=0.75, no_noise_prob=0.5, use_labels=False), 4, 2),
}
SCHEDULE = {"latent_s2": "linear"}     # beta schedule per case (default cosine)
PARAM_SEED, BATCH_SEED, DIR_SEED = 0, 100, 4242
FD_STEP = 1e-3
GROUP_DIRECTIONS = 3   # whole-tree directions; plus one direction per top-level parameter group


def load_reference():
  """Imports the reference's files over the shim and lifts loss_fn out of update_fn."""
  sys.path.insert(0, os.path.join(HERE, "refshim"))
  sys.path.insert(0, REF)
  for name in ("big_vision.utils", "big_vision.models.common"):   # imported by vit.py for checkpoint loading only
    sys.modules[name] = types.ModuleType(name)
  import jax  # noqa: F401  (the shim)
  assert "refshim" in jax.__file__
  from big_vision.models import ae
  from big_vision import gaussian_diffusion as gd
  assert ae.__file__.startswith(REF) and gd.__file__.startswith(REF)
  path = os.path.join(REF, "big_vision", "trainers", "train_ae.py")
  tree = ast.parse(open(path).read(), filename=path)
  def lift(fn_name):
    node = next(n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name == fn_name)
    return compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), (node.lineno, node.end_lineno)
  code, lines = lift("loss_fn")
  load_reference.lift = lift
  return ae, gd, code, lines


class Config(dict):
  __getattr__ = dict.__getitem__


def np64(t):
  return np.asarray(t.detach().cpu().double().numpy() if isinstance(t, torch.Tensor) else t, dtype=np.float64)


def tree64(tree):
  return {k: (tree64(v) if isinstance(v, dict) else np64(v)) for k, v in tree.items()}


def flatten(tree, prefix=()):
  out = {}
  for k in sorted(tree):
    v = tree[k]
    if isinstance(v, dict):
      out.update(flatten(v, prefix + (k,)))
    else:
      out[prefix + (k,)] = v
  return out


def directions(params, seed=DIR_SEED):
  """Seeded unit-norm parameter directions: GROUP_DIRECTIONS over the whole tree, then one per top-level group.
  Regenerated identically by the tests (torch CPU generator, sorted leaf order)."""
  flat = flatten(params)
  g = torch.Generator().manual_seed(seed)
  groups = [None] * GROUP_DIRECTIONS + sorted({k[0] for k in flat})
  out = []
  for grp in groups:
    d = {}
    for k, v in flat.items():
      r = torch.randn(tuple(np.shape(v)), generator=g, dtype=torch.float64)
      d[k] = r if (grp is None or k[0] == grp) else torch.zeros_like(r)
    nrm = float(torch.sqrt(sum((x ** 2).sum() for x in d.values())))
    out.append((grp or "*", {k: x / nrm for k, x in d.items()}))
  return out


def shifted(params, d, h):
  def rec(tree, prefix):
    return {k: (rec(v, prefix + (k,)) if isinstance(v, dict) else v + h * d[prefix + (k,)].numpy())
            for k, v in tree.items()}
  return rec(params, ())


def digest(t):
  return hashlib.sha256(t.contiguous().numpy().tobytes()).hexdigest()[:16]


def build_case(name, ae, gdm, loss_code):
  import jax
  import jax.numpy as jnp
  mkw, tkw, B, n_noise = CASES[name]
  engine_model, _ = U.make_models(**mkw)
  params_t = U.cpu_tree(U.perturb_init(engine_model, PARAM_SEED, "cpu"))
  batch, rand = U.make_batch(engine_model, B, n_noise=n_noise, seed=BATCH_SEED, use_labels=tkw["use_labels"],
                             device="cpu")
  params = tree64(params_t)
  model = ae.Model(**mkw)                                   # the reference's factory (ae.py:220-222)
  gd = gdm.create_gaussian_diffusion(SCHEDULE.get(name, "cosine"), 1000)        # gaussian_diffusion.py:32-66
  n_clean = B - n_noise
  assert n_clean == int(B * tkw["no_noise_prob"])
  images = jnp.asarray(np64(batch["image"]))
  x0_noise, x0_clean = images[:n_noise], images[n_noise:]
  t = jnp.asarray(rand["t"].numpy().astype(np.int32))
  noise = jnp.asarray(np64(rand["noise"]))
  x_t = gdm.q_sample(gd=gd, x_start=x0_noise, t=t, noise=noise)         # gaussian_diffusion.py:84-98
  labels = jnp.asarray(batch["label"][:n_noise].numpy().astype(np.int32)) if tkw["use_labels"] else None
  has_lbl = mkw.get("num_classes") is not None
  img_size, channels = mkw.get("img_size", 64), mkw.get("channels", 3)
  patch = int(mkw["variant"].split("/")[1])

  def keys():
    drop = rand.get("label_drop_noise")
    return dict(
        rng_model=jax.Key(), rng_model_noise=jax.Key(),
        mae_noise_rng=jax.Key({"uniform": np64(rand["mask_noise_clean"])}),
        mae_noise_rng_noise=jax.Key({"uniform": np64(rand["mask_noise_noise"])}),
        cfg_rng=jax.Key({"bernoulli": np.zeros((n_clean,))}),
        cfg_rng_noise=jax.Key({"bernoulli": np64(drop.double()) if drop is not None else np.zeros((n_noise,))}))

  def reference_loss(p):
    env = dict(jnp=jnp, model=model, config=Config(diffusion_space=(img_size, img_size, channels), **tkw), B=B, n_noise=n_noise,
               n_no_noise=n_clean, x_0_noise=x0_noise, x_0_no_noise=x0_clean, x_t_noise=x_t, batched_t=t,
               labels_t=labels, noise=noise, **keys())
    exec(loss_code, env)                                     # defines the reference's loss_fn in env
    return float(env["loss_fn"](p))

  out = {"input_digest": digest(batch["image"]) + digest(rand["noise"]) + digest(rand["mask_noise_clean"]),
         "param_digest": digest(torch.cat([v.reshape(-1) for _, v in sorted(flatten(params_t).items())])),
         "x_t": torch.from_numpy(np.asarray(x_t)).float(),
         "loss": reference_loss(params)}
  k = keys()
  if n_clean > 0:
    pred, o = model.apply({"params": params}, x0_clean, t=jnp.zeros((n_clean, 1), dtype=jnp.int32), train=True,
                          mask=tkw["mask_ratio_no_noise"],
                          rngs={"dropout": k["rng_model"], "cfg": k["cfg_rng"], "mae_noise": k["mae_noise_rng"]})
    out["clean"] = pack(pred, o, patch)
  if n_noise > 0:
    pred, o = model.apply({"params": params}, x_t, t=t + 1, y=labels, train=True, mask=tkw["mask_ratio"],
                          rngs={"dropout": k["rng_model_noise"], "cfg": k["cfg_rng_noise"],
                                "mae_noise": k["mae_noise_rng_noise"]})
    out["noise"] = pack(pred, o, patch)
  slopes = []
  for grp, d in directions(params_t):
    lp, lm = reference_loss(shifted(params, d, FD_STEP)), reference_loss(shifted(params, d, -FD_STEP))
    lp2, lm2 = reference_loss(shifted(params, d, FD_STEP / 2)), reference_loss(shifted(params, d, -FD_STEP / 2))
    d1, d2 = (lp - lm) / (2 * FD_STEP), (lp2 - lm2) / FD_STEP
    slopes.append((grp, (4 * d2 - d1) / 3, abs(d2 - d1)))   # Richardson-extrapolated slope, and its error scale
  out["slopes"] = slopes
  if has_lbl:   # classifier-free-guidance forward (ae.py:177-195), inference mode, no masking
    y = jnp.asarray(batch["label"][:2].numpy().astype(np.int32))
    tt = jnp.asarray(np.array([[500], [17]], dtype=np.int32))
    pred, o = model.apply({"params": params}, images[:2], t=tt, y=y, cfg_scale=1.5)
    out["cfg"] = {"pred": torch.from_numpy(np.asarray(pred)).float(), "t": torch.tensor([[500], [17]], dtype=torch.int32),
                  "cfg_scale": 1.5, "pre_logits": torch.from_numpy(np.asarray(o["pre_logits"])).double()}
  return out


def pack(pred, o, patch=4):
  pred = np.asarray(pred)
  d = {"pred0": torch.from_numpy(pred[0]).float(),                                  # sample 0 in full
       "pred_sample_means": torch.from_numpy(pred.mean(axis=(1, 2))).double(),      # [n, 2C]
       "pred_abs_mean": float(np.abs(pred).mean()),
       "pre_logits": torch.from_numpy(np.asarray(o["pre_logits"])).double()}
  if o["mask"] is not None:
    m = np.asarray(o["mask"])[:, ::patch, ::patch, 0]                               # one value per patch
    assert set(np.unique(m)) <= {0.0, 1.0}
    d["patch_mask"] = torch.from_numpy(m.reshape(m.shape[0], -1).astype(np.uint8))
  return d


def diffusion_vectors(gdm):
  """gaussian_diffusion.py run as is: tables, q_sample, x0/eps conversions, DDIM steps and the DDIM loop."""
  import jax
  import jax.numpy as jnp
  out = {}
  for sched in ("cosine", "linear"):
    gd = gdm.create_gaussian_diffusion(sched, 1000)
    out[f"tables_{sched}"] = {k: torch.from_numpy(np.asarray(v, dtype=np.float64)) for k, v in gd.items()}
  gd = gdm.create_gaussian_diffusion("cosine", 1000)
  g = torch.Generator().manual_seed(77)
  x = torch.randn(4, 8, 8, 3, generator=g, dtype=torch.float64)
  eps = torch.randn(4, 8, 8, 3, generator=g, dtype=torch.float64)
  nz = torch.randn(4, 8, 8, 3, generator=g, dtype=torch.float64)
  t = torch.tensor([[0], [1], [500], [999]], dtype=torch.int32)
  tn = torch.tensor([[0], [0], [492], [991]], dtype=torch.int32)
  J = lambda a: jnp.asarray(a.numpy())
  out["inputs"] = {"x": x, "eps": eps, "noise": nz, "t": t, "t_next": tn}
  out["q_sample"] = torch.from_numpy(np.asarray(gdm.q_sample(gd=gd, x_start=J(x), t=J(t), noise=J(nz))))
  out["xstart_from_eps"] = torch.from_numpy(np.asarray(gdm._predict_xstart_from_eps(gd, J(x), J(t), J(eps))))
  steps = []
  for eta, clip, use_next in ((0.0, False, True), (1.0, False, True), (0.5, True, True), (1.0, True, False)):
    r = gdm.ddim_sample(gd, lambda x_t, t, rng, **kw: J(eps), J(x), J(t), J(tn) if use_next else None,
                        jax.Key({"normal": nz.numpy()}), clip_denoised=clip, eta=eta)
    steps.append({"eta": eta, "clip": clip, "use_next": use_next,
                  "sample": torch.from_numpy(np.asarray(r["sample"])),
                  "pred_xstart": torch.from_numpy(np.asarray(r["pred_xstart"]))})
  out["ddim_steps"] = steps
  # whole loop with a closed-form model: eps(x_t, t, y) = 0.3 x_t + 0.01 (t/1000) + 0.05 y
  S = 5
  draws = [torch.randn(2, 8, 8, 3, generator=g, dtype=torch.float64) for _ in range(S + 2)]
  ys = torch.tensor([3, 7], dtype=torch.int32)

  def apply_fn(x_t, t, rng, y=None, cfg_scale=None):
    return 0.3 * x_t + 0.01 * (np.asarray(t, dtype=np.float64) / 1000.0)[:, :, None, None] \
        + 0.05 * np.asarray(y, dtype=np.float64)[:, None, None, None]
  ret, _ = gdm.ddim_sample_loop(gd, apply_fn, jax.Key({"normal": [d.numpy() for d in draws]}),
                                np.zeros((2, 8, 8, 3)), ys=J(ys), sampling_steps=S, eta=0.7)
  S_eff = None
  out["ddim_loop"] = {"draws": torch.stack(draws), "ys": ys, "sampling_steps": S, "eta": 0.7,
                      "sample": torch.from_numpy(np.asarray(ret["sample"]))}
  for n, s in ((1000, 250), (1000, 125), (1000, 7), (1000, 1000)):
    # the time-step grid exactly as ddim_sample_loop builds it (gaussian_diffusion.py:241-242)
    ts = jnp.append(jnp.arange(n - 1, 0, step=-n // s, dtype=jnp.int32), 0)
    out[f"timesteps_{n}_{s}"] = torch.from_numpy(np.asarray(ts).astype(np.int64))
  return out


SAMPLER = dict(case="dit_s4", n=2, steps=4, eta=0.3, ys=(3, 7), noise_seed=21,
               variants=((1.5, True), (2.0, False), (None, True)))   # (cfg_scale, eps_pred)


def sampler_vectors(ae, gdm):
  """The whole sampler as the reference wires it: `create_apply_fn` (train_ae.py:472-483, lifted like loss_fn: the model
  at t + 1 on `ema_params`, eps head or x0 head converted) under `ddim_sample_loop` (gaussian_diffusion.py:213-280) with
  classifier-free guidance, on the dit_s4 parameters."""
  import jax
  import jax.numpy as jnp
  code, lines = load_reference.lift("create_apply_fn")
  mkw = CASES[SAMPLER["case"]][0]
  engine_model, _ = U.make_models(**mkw)
  params = tree64(U.cpu_tree(U.perturb_init(engine_model, PARAM_SEED, "cpu")))
  model = ae.Model(**mkw)
  gd = gdm.create_gaussian_diffusion("cosine", 1000)
  env = dict(model=model, config=Config(diffusion_space=(64, 64, 3)), _predict_eps_from_xstart=gdm._predict_eps_from_xstart)
  exec(code, env)
  n, steps = SAMPLER["n"], SAMPLER["steps"]
  g = torch.Generator().manual_seed(SAMPLER["noise_seed"])
  noises = [torch.randn(n, 64, 64, 3, generator=g) for _ in range(steps + 2)]
  ys = jnp.asarray(np.array(SAMPLER["ys"], dtype=np.int32))
  out = {"lines": lines, "noise_digest": digest(torch.stack(noises)), "samples": []}
  for cfg_scale, eps_pred in SAMPLER["variants"]:
    apply_fn = env["create_apply_fn"]({"ema_params": params, "gd": gd}, eps_pred=eps_pred)
    ret, _ = gdm.ddim_sample_loop(gd, apply_fn, jax.Key({"normal": [np64(z) for z in noises]}), np.zeros((n, 64, 64, 3)),
                                  ys=ys if cfg_scale is not None else None, sampling_steps=steps, cfg_scale=cfg_scale,
                                  eta=SAMPLER["eta"])
    out["samples"].append(torch.from_numpy(np.asarray(ret["sample"])).float())
    print("sampler", cfg_scale, eps_pred, float(np.abs(np.asarray(ret["sample"])).mean()))
  return out


EVAL = dict(model=dict(variant="S/4", adaln=True, depth=2, dec_depth=1), param_seed=5, data_seed=9, n=4, t_noised=50,
            mask_ratio_no_noise=0.75)


def eval_inputs(num_patches=256):
  """The inputs of tests/test_model_gpu.py::test_evaluator_predict_functions_match_oracle (same generator order)."""
  g = torch.Generator().manual_seed(EVAL["data_seed"])
  n = EVAL["n"]
  image = torch.rand(n, 64, 64, 3, generator=g) * 2 - 1
  noise = torch.randn(n, 64, 64, 3, generator=g)
  t = torch.randint(0, 1000, (n, 1), generator=g, dtype=torch.int32)
  mn = torch.rand(n, num_patches, generator=g)
  return image, noise, t, mn


def evaluator_vectors(ae, gdm):
  """The evaluator closures of train_ae.py:384-470 (predict_fn, create_noised_pred_fn, eval_patch_fn, eval_loss_fn), lifted
  by ast and run on supplied draws."""
  import jax
  import jax.numpy as jnp
  engine_model, _ = U.make_models(**EVAL["model"])
  params = tree64(U.cpu_tree(U.perturb_init(engine_model, EVAL["param_seed"], "cpu")))
  model = ae.Model(**EVAL["model"])
  gd = gdm.create_gaussian_diffusion("cosine", 1000)
  image, noise, t, mn = eval_inputs()
  key = jax.Key({"normal": np64(noise), "randint": t.numpy().astype(np.int32), "uniform": np64(mn)})
  state = {"params": params, "gd": gd, "rng": key}
  batch = {"image": jnp.asarray(np64(image))}
  env = dict(jax=jax, jnp=jnp, model=model, q_sample=gdm.q_sample, _predict_xstart_from_eps=gdm._predict_xstart_from_eps,
             config=Config(diffusion_space=(64, 64, 3), mask_ratio_no_noise=EVAL["mask_ratio_no_noise"]))
  lines = {}
  for fn in ("predict_fn", "create_noised_pred_fn", "eval_patch_fn", "eval_loss_fn"):
    code, lines[fn] = load_reference.lift(fn)      # ast.walk is breadth-first: the outer predict_fn (:384) comes first
    exec(code, env)
  T = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float64))
  out = {"lines": lines, "input_digest": digest(image) + digest(noise) + digest(mn)}
  _, o = env["predict_fn"](state, batch)
  out["predict_pre_logits"] = T(o["pre_logits"])
  _, o = env["create_noised_pred_fn"](EVAL["t_noised"])(state, batch)
  out["noised_pre_logits"] = T(o["pre_logits"])
  px0, mask = env["eval_patch_fn"](state, batch)
  out["patch_pred_x0"] = T(px0).float()[:2].clone()                     # first two samples (fixture size)
  out["patch_mask"] = torch.from_numpy(np.asarray(mask)[:, ::4, ::4, 0].reshape(EVAL["n"], -1).astype(np.uint8))
  loss, x_t, pred_x0, pred_x0_eps = env["eval_loss_fn"](state, batch)
  out["loss"] = float(loss)
  out["x_t"], out["pred_x0"], out["pred_x0_eps"] = T(x_t).float()[:2].clone(), T(pred_x0).float()[:2].clone(), T(pred_x0_eps).float()[:2].clone()
  print("evaluators: loss", out["loss"], "lines", lines)
  return out


STEP = dict(case="umd_lbl_s4", ema_decay=0.01, update_seed=77, update_scale=1e-3)


def step_updates(params_t):
  """The stand-in for optax's output: a seeded small random update tree (sorted leaf order)."""
  g = torch.Generator().manual_seed(STEP["update_seed"])
  return {k: STEP["update_scale"] * torch.randn(tuple(v.shape), generator=g, dtype=torch.float64)
          for k, v in sorted(flatten(params_t).items())}


def update_fn_vectors(ae, gdm):
  """The reference's whole `update_fn` (train_ae.py:287-382; decorators dropped) on the umd_lbl_s4 case: its own RNG split
  tree, batch split `int(B * no_noise_prob)`, label slicing, q_sample call, loss_fn, measurements and EMA update.  Only the
  library calls are stand-ins: `jax.value_and_grad` evaluates the loss (gradients are pinned by the slopes of the step
  cases), `tx.update` returns a seeded update tree, `optax.apply_updates` / `incremental_update` are optax's two one-line
  formulas (p + u;  old + s * (new - old))."""
  import jax
  import jax.numpy as jnp
  path = os.path.join(REF, "big_vision", "trainers", "train_ae.py")
  tree = ast.parse(open(path).read(), filename=path)
  node = next(n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name == "update_fn")
  node.decorator_list = []
  code = compile(ast.Module(body=[node], type_ignores=[]), path, "exec")
  name = STEP["case"]
  mkw, tkw, B, n_noise = CASES[name]
  engine_model, _ = U.make_models(**mkw)
  params_t = U.cpu_tree(U.perturb_init(engine_model, PARAM_SEED, "cpu"))
  batch, rand = U.make_batch(engine_model, B, n_noise=n_noise, seed=BATCH_SEED, use_labels=tkw["use_labels"], device="cpu")
  params = tree64(params_t)
  upd = step_updates(params_t)
  n_clean = B - n_noise
  K = jax.Key
  second = K({"split": [K(), K(), K({"uniform": np64(rand["mask_noise_noise"])}),
                        K({"bernoulli": np64(rand["label_drop_noise"].double())})]})              # rng, model, mae, cfg (:303)
  root = K({"split": [second, K(), K({"randint": rand["t"].numpy().astype(np.int32)}), K({"normal": np64(rand["noise"])}),
                      K({"uniform": np64(rand["mask_noise_clean"])}), K({"bernoulli": np.zeros((n_clean,))})]})   # (:302)

  def add(a, b):
    return {k: (add(a[k], b[k]) if isinstance(a[k], dict) else a[k] + b[k]) for k in a}

  def unflat(flat_t):
    out = {}
    for path_, v in flat_t.items():
      d = out
      for k in path_[:-1]:
        d = d.setdefault(k, {})
      d[path_[-1]] = v.numpy()
    return out
  updates = unflat(upd)
  optax = types.SimpleNamespace(
      apply_updates=add,
      incremental_update=lambda new, old, s: {k: (optax.incremental_update(new[k], old[k], s) if isinstance(new[k], dict)
                                                  else old[k] + s * (new[k] - old[k])) for k in new})
  env = dict(jax=types.SimpleNamespace(random=jax.random, tree_util=jax.tree_util,
                                       value_and_grad=lambda f: (lambda p: (f(p), None))),
             jnp=jnp, optax=optax, q_sample=gdm.q_sample, model=ae.Model(**mkw),
             tx=types.SimpleNamespace(update=lambda grads, opt, p: (updates, opt)),
             config=Config(diffusion_space=(64, 64, 3), ema_decay=STEP["ema_decay"], **tkw))
  exec(code, env)
  state = {"params": params, "ema_params": jax.tree_map(lambda a: 0.9 * a, params), "opt": {"count": 0}, "rng": root,
           "gd": gdm.create_gaussian_diffusion("cosine", 1000)}
  new_state, meas = env["update_fn"](state, {"image": jnp.asarray(np64(batch["image"])),
                                             "label": jnp.asarray(batch["label"].numpy().astype(np.int32))})
  fl = flatten(new_state["params"])
  fe = flatten(new_state["ema_params"])
  probe = ("final_conv", "bias")
  out = {"lines": (node.lineno, node.end_lineno), "training_loss": float(meas["training_loss"]),
         "l2_params": float(meas["l2_params"]), "l2_updates": float(meas["l2_updates"]),
         "new_param_probe": torch.from_numpy(np.asarray(fl[probe])), "new_ema_probe": torch.from_numpy(np.asarray(fe[probe])),
         "state_keys": sorted(new_state)}
  print("update_fn:", {k: out[k] for k in ("lines", "training_loss", "l2_params", "l2_updates", "state_keys")})
  return out


def main():
  ae, gdm, loss_code, loss_lines = load_reference()
  if sys.argv[1:] == ["sampler"]:
    path = os.path.join(HERE, "reference_sampler_golden.pt")
    torch.save({"provenance": "create_apply_fn + ddim_sample_loop of the reference executed over tests/golden/refshim",
                "sampler": sampler_vectors(ae, gdm), "evaluators": evaluator_vectors(ae, gdm),
                "update_fn": update_fn_vectors(ae, gdm)}, path)
    print(path, os.path.getsize(path), "bytes")
    return
  gold = {"provenance": "reference source executed over tests/golden/refshim (numpy fp64); loss_fn = train_ae.py:%d-%d"
                        % loss_lines,
          "fd_step": FD_STEP, "cases": {}, "diffusion": diffusion_vectors(gdm)}
  for n in CASES:
    gold["cases"][n] = build_case(n, ae, gdm, loss_code)
    c = gold["cases"][n]
    print(n, "loss", c["loss"], "slopes", [(g, f"{s:.6e}", f"{e:.1e}") for g, s, e in c["slopes"]][:4])
  path = os.path.join(HERE, "reference_golden.pt")
  torch.save(gold, path)
  print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
  main()
