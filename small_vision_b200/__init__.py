"""Import shim: the package directory mandated by the repo layout is `small-vision_b200/`
(not a valid Python identifier), so `import small_vision_b200` resolves to it from here."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "small-vision_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
  exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
