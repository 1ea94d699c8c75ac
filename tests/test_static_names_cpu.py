"""Undefined-name check (the part of pyflakes that matters for code which only runs on a GPU box): every name that
bench.py, __graft_entry__.py and the package's host modules load must be bound in an enclosing scope, at module level or be
a builtin.  The GPU arm of bench.py cannot execute here, so a misspelt variable in it would otherwise first show up on the
driver's box at round end."""
import ast
import builtins
import glob
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")] + \
    sorted(glob.glob(os.path.join(ROOT, "small-vision_b200", "*.py"))) + sorted(glob.glob(os.path.join(ROOT, "tools", "*.py"))) + \
    sorted(glob.glob(os.path.join(ROOT, "tests", "*.py"))) + sorted(glob.glob(os.path.join(ROOT, "oracle", "*.py")))


class Scope:
  def __init__(self, node, parent):
    self.node, self.parent, self.bound, self.loads = node, parent, set(), []


def _bind_target(scope, t):
  for n in ast.walk(t):
    if isinstance(n, ast.Name) and isinstance(n.ctx, (ast.Store, ast.Del)):
      scope.bound.add(n.id)


def _collect(node, scope, scopes):
  """Fills scope.bound / scope.loads for the body of `node`, opening child scopes for functions, lambdas, classes and
  comprehensions."""
  for child in ast.iter_child_nodes(node):
    if isinstance(child, (ast.FunctionDef, ast.AsyncFunctionDef)):
      scope.bound.add(child.name)
      for d in child.decorator_list + child.args.defaults + [x for x in child.args.kw_defaults if x is not None]:
        _collect(ast.Expression(body=d), scope, scopes)
      s = Scope(child, scope)
      scopes.append(s)
      a = child.args
      for arg in a.posonlyargs + a.args + a.kwonlyargs + ([a.vararg] if a.vararg else []) + ([a.kwarg] if a.kwarg else []):
        s.bound.add(arg.arg)
      for stmt in child.body:
        _collect(ast.Module(body=[stmt], type_ignores=[]), s, scopes)
    elif isinstance(child, ast.Lambda):
      for d in child.args.defaults + [x for x in child.args.kw_defaults if x is not None]:
        _collect(ast.Expression(body=d), scope, scopes)
      s = Scope(child, scope)
      scopes.append(s)
      a = child.args
      for arg in a.posonlyargs + a.args + a.kwonlyargs + ([a.vararg] if a.vararg else []) + ([a.kwarg] if a.kwarg else []):
        s.bound.add(arg.arg)
      _collect(ast.Expression(body=child.body), s, scopes)
    elif isinstance(child, ast.ClassDef):
      scope.bound.add(child.name)
      for d in child.decorator_list + child.bases:
        _collect(ast.Expression(body=d), scope, scopes)
      s = Scope(child, scope)
      s.is_class = True
      scopes.append(s)
      for stmt in child.body:
        _collect(ast.Module(body=[stmt], type_ignores=[]), s, scopes)
    elif isinstance(child, (ast.ListComp, ast.SetComp, ast.DictComp, ast.GeneratorExp)):
      s = Scope(child, scope)
      scopes.append(s)
      for g in child.generators:
        _bind_target(s, g.target)
      _collect(child, s, scopes)
    else:
      if isinstance(child, (ast.Import, ast.ImportFrom)):
        for al in child.names:
          scope.bound.add((al.asname or al.name).split(".")[0])
      elif isinstance(child, (ast.Global, ast.Nonlocal)):
        scope.bound.update(child.names)
      elif isinstance(child, ast.ExceptHandler) and child.name:
        scope.bound.add(child.name)
      elif isinstance(child, ast.Name):
        if isinstance(child.ctx, ast.Load):
          scope.loads.append((child.id, child.lineno))
        else:
          scope.bound.add(child.id)
      elif isinstance(child, ast.NamedExpr):
        _bind_target(scope, child.target)
      _collect(child, scope, scopes)


def undefined_names(path):
  tree = ast.parse(open(path).read(), path)
  top = Scope(tree, None)
  scopes = [top]
  _collect(tree, top, scopes)
  known = set(dir(builtins)) | {"__file__", "__name__", "__doc__"}
  bad = []
  for s in scopes:
    for name, line in s.loads:
      p, found = s, False
      while p is not None:
        # class bodies are not enclosing scopes for the functions inside them
        if name in p.bound and (p is s or not getattr(p, "is_class", False)):
          found = True
          break
        p = p.parent
      if not found and name not in known:
        bad.append((name, line))
  return bad


@pytest.mark.parametrize("path", FILES, ids=[os.path.relpath(p, ROOT) for p in FILES])
def test_no_undefined_names(path):
  assert undefined_names(path) == []


def test_the_checker_sees_a_misspelt_name(tmp_path):
  p = tmp_path / "m.py"
  p.write_text("import os\n\ndef f(a):\n  b = a + 1\n  def g():\n    return b + c_missing + os.sep\n  return [x for x in range(b)] + [y_missing]\n")
  assert sorted(n for n, _ in undefined_names(str(p))) == ["c_missing", "y_missing"]
