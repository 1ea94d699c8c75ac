"""jax.sharding names used by big_vision/sharding.py (executed by make_sharding_golden.py) and imported at module level
by fewshot_lsr.py: plain value objects."""


class PartitionSpec(tuple):
  def __new__(cls, *parts):
    return super().__new__(cls, parts)


class NamedSharding:
  def __init__(self, mesh, spec):
    self.mesh, self.spec = mesh, spec


class Mesh:
  def __init__(self, devices, axis_names):
    import numpy as np
    self.devices, self.axis_names = np.asarray(devices), tuple(axis_names)
