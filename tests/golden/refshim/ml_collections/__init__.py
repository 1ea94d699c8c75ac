"""Stand-in for ml_collections.ConfigDict as the reference's config files use it: attribute / item access, nested dicts
promoted to ConfigDict on assignment, `get`, `in`, `keys`, `items`, `to_dict`."""


class ConfigDict:
  def __init__(self, initial=None, type_safe=True):
    object.__setattr__(self, "_d", {})
    for k, v in (initial or {}).items():
      self[k] = v

  @staticmethod
  def _wrap(v):
    return ConfigDict(v) if isinstance(v, dict) else v

  def __setitem__(self, k, v):
    self._d[k] = self._wrap(v)

  def __getitem__(self, k):
    return self._d[k]

  def __setattr__(self, k, v):
    self[k] = v

  def __getattr__(self, k):
    try:
      return object.__getattribute__(self, "_d")[k]
    except KeyError:
      raise AttributeError(k) from None

  def __contains__(self, k):
    return k in self._d

  def __iter__(self):
    return iter(self._d)

  def get(self, k, default=None):
    return self._d.get(k, default)

  def keys(self):
    return self._d.keys()

  def items(self):
    return self._d.items()

  def to_dict(self):
    return {k: (v.to_dict() if isinstance(v, ConfigDict) else v) for k, v in self._d.items()}
