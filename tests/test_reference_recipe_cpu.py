"""small-vision_b200/config.py and the parameter layout against tests/golden/reference_recipe_golden.json: the recipes
of the BASELINE.json configurations as the reference's own `configs/ae_i1k.py`, `utils.steps` and the optimiser wiring of
train_ae.py:124-151 resolve them (executed over stand-ins, tests/golden/make_recipe_golden.py): step counts, learning-rate
schedule arguments, AdamW hyper-parameters, the weight-decay mask per leaf, model kwargs, masking / branch settings."""
import json
import os

import pytest

from small_vision_b200.config import TrainConfig
from small_vision_b200.model import Model
from tests.golden import make_recipe_golden as RG

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_recipe_golden.json")))


def _train_config(r):
  t, sched = r["train"], r["diff_schedule"]
  return TrainConfig(batch_size=r["batch_size"], no_noise_prob=t["no_noise_prob"], mask_ratio=t["mask_ratio"],
                     mask_ratio_no_noise=t["mask_ratio_no_noise"], use_labels=bool(t["use_labels"]),
                     beta_schedule=sched["beta_schedule"], timesteps=sched["timesteps"],
                     diffusion_space=tuple(r["diffusion_space"]), total_epochs=r["total_epochs"], ntrain_img=GOLD["ntrain_img"],
                     ema_decay=t["ema_decay"])


@pytest.mark.parametrize("name", sorted(RG.RECIPES))
def test_schedule_and_optimiser_hyperparameters_match_reference_source(name):
  r = GOLD["recipes"][name]
  c = _train_config(r).resolved()           # peak_lr, wd, betas, clip_norm, mu_dtype stay at TrainConfig's defaults
  s = r["schedule"]
  assert c.total_steps == r["total_steps"] == s["decay_steps"]           # utils.steps rounds (utils.py:1059-1061)
  assert c.warmup_steps == s["warmup_steps"]
  assert c.scaled_peak_lr == pytest.approx(s["peak_value"], rel=1e-12) and s["init_value"] == 0.0
  a = r["adamw"]
  assert (c.wd, c.betas[0], c.betas[1], c.mu_dtype) == (a["weight_decay"], a["b1"], a["b2"], a["mu_dtype"])
  assert c.clip_norm == r["clip_norm"] and r["chain"] == ["clip", "adamw"] and r["adamw_learning_rate_is_schedule"]


@pytest.mark.parametrize("name", sorted(RG.RECIPES))
def test_model_kwargs_and_decay_mask_match_reference_source(name):
  r = GOLD["recipes"][name]
  model = Model(**r["model"])                # the reference's own config.model dict, unchanged
  cfg = model.cfg
  assert cfg.img_size == r["diffusion_space"][0] and cfg.channels == r["diffusion_space"][2]
  assert cfg.num_classes == r["train"]["num_classes"]
  got = {"/".join(lf.path): model.layout.decay(lf) for lf in model.layout.leaves}
  assert got == r["decay_mask"]
  assert any(got.values()) and not all(got.values())


def test_defaults_are_the_reference_defaults():
  r = GOLD["recipes"]["default"]
  d = TrainConfig()
  t = r["train"]
  assert (d.batch_size, d.no_noise_prob, d.mask_ratio, d.mask_ratio_no_noise, d.use_labels) == \
      (r["batch_size"], t["no_noise_prob"], t["mask_ratio"], t["mask_ratio_no_noise"], bool(t["use_labels"]))
  assert d.total_epochs == r["total_epochs"] and tuple(d.diffusion_space) == tuple(r["diffusion_space"])
  assert d.peak_lr * d.batch_size / 256 == pytest.approx(r["schedule"]["peak_value"], rel=1e-12)
  # label fine-tuning sets the EMA rate from the batch size (configs/ae_i1k.py:31-33)
  assert GOLD["recipes"]["dit_b4_labels"]["train"]["ema_decay"] == pytest.approx(1e-4 * 256 / 256)
  # the few-shot probe and sampler settings the §8f components default to
  assert r["fewshot"]["l2_reg"] == 1024 and r["fewshot"]["representation_layer"] == "pre_logits"
  assert r["diff_schedule"]["sampling_timesteps"] == 125 and r["diff_schedule"]["eta"] == 1.0


@pytest.mark.parametrize("name", sorted(RG.RECIPES))
def test_parameter_tree_and_initialisers_match_reference_source(name):
  """The fixture's `init` table was recorded while the reference's own model ran over parameters of the engine layout's
  shapes (the stand-in asserts every shape the reference requests and that both leaf sets coincide — at full depth for
  B/4 and L/2).  Here: the initialiser the reference hands to each leaf (None = the Flax layer's default) against the
  engine's `Leaf.init` used by Model.init / init_arena."""
  import math
  r = GOLD["recipes"][name]
  model = Model(**r["model"])
  assert {"/".join(lf.path) for lf in model.layout.leaves} == set(r["init"])
  for lf in model.layout.leaves:
    rec = r["init"]["/".join(lf.path)]
    layer, init, leaf = rec["layer"], rec["init"], lf.path[-1]
    shape = lf.shape[1:] if len(lf.path) > 2 and lf.path[1].startswith("Scan") else lf.shape   # stacked [depth, ...]
    if init is None:                                                     # Flax defaults
      if layer == "LayerNorm":
        want = "ones" if leaf == "scale" else "zeros"
      elif layer == "Embed":
        want = ("normal", 1 / math.sqrt(shape[-1]))                      # variance_scaling(1, fan_in, normal, out_axis=0)
      elif leaf == "bias":
        want = "zeros"
      else:
        want = ("lecun", math.prod(shape[:-1]))                          # lecun_normal: fan_in = everything but the last axis
    elif init["name"] == "zeros":
      want = "zeros"
    elif init["name"] == "normal":
      want = ("normal", (init["args"] or [init["kwargs"]["stddev"]])[0])
    elif init["name"] == "xavier_uniform":
      if layer == "_Proj" and lf.path[-2] == "out":
        want = ("xavier", shape[0] * shape[1], shape[2])                 # DenseGeneral flattens [H, Dh] -> D
      elif layer == "_Proj":
        want = ("xavier", shape[0], shape[1] * shape[2])
      else:
        want = ("xavier", shape[0], shape[1])
    else:
      raise AssertionError(f"unexpected initialiser {init}")
    got = "zeros" if lf.init == "adaln" else lf.init                     # 'adaln' = zero-init unless nonzero_adaln is asked for
    if isinstance(want, tuple):
      assert isinstance(got, tuple) and got[0] == want[0] and got[1:] == pytest.approx(want[1:], rel=1e-12), (lf.path, got, want)
    else:
      assert got == want, (lf.path, got, want)


@pytest.mark.parametrize("name", sorted(RG.RECIPES))
def test_get_config_resolves_the_reference_option_strings(name):
  """config.get_config(arg) against what the reference's configs/ae_i1k.py::get_config resolved for the same string."""
  from small_vision_b200.config import get_config
  r = GOLD["recipes"][name]
  model_kw, tc = get_config(r["arg"])
  assert model_kw == r["model"]
  t = r["train"]
  assert (tc.batch_size, tc.no_noise_prob, tc.mask_ratio, tc.mask_ratio_no_noise, tc.use_labels, tc.ema_decay) == \
      (r["batch_size"], t["no_noise_prob"], t["mask_ratio"], t["mask_ratio_no_noise"], bool(t["use_labels"]), t["ema_decay"])
  assert list(tc.diffusion_space) == r["diffusion_space"] and tc.beta_schedule == r["diff_schedule"]["beta_schedule"]
  c = tc.resolved()
  s, a = r["schedule"], r["adamw"]
  assert (c.total_steps, c.warmup_steps) == (s["decay_steps"], s["warmup_steps"])
  assert c.scaled_peak_lr == pytest.approx(s["peak_value"], rel=1e-12)
  assert (c.wd, list(c.betas), c.clip_norm, c.mu_dtype) == (a["weight_decay"], [a["b1"], a["b2"]], r["clip_norm"], a["mu_dtype"])
  Model(**model_kw)


def test_get_config_option_string_errors():
  from small_vision_b200.config import get_config, parse_arg
  assert parse_arg("S/4")["variant"] == "S/4"                      # a lone value goes to the first option
  assert parse_arg("use_labels")["use_labels"] is True             # a bare flag
  assert parse_arg("adaln=False,beta2=0.99") ["beta2"] == 0.99
  with pytest.raises(ValueError):
    parse_arg("no_such_option=1")
  with pytest.raises(AssertionError):
    get_config("latent_diffusion=True")                            # needs size=256 (ae_i1k.py:17)
