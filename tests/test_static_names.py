"""Every module-level name the package and bench.py load must be bound somewhere in the module (imports, defs,
assignments).  The product path cannot be imported-and-run without a B200, so this catches a missing import on CPU."""
import ast
import builtins
import glob
import os

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
PKG = os.path.join(ROOT, "boosting-neural-video-representation-via-online-structural-reparameteration_b200")


def unbound_names(path):
    tree = ast.parse(open(path).read())
    bound = set(dir(builtins)) | {"__file__", "__name__", "__doc__"}
    for n in ast.walk(tree):
        if isinstance(n, (ast.Import, ast.ImportFrom)):
            bound.update((a.asname or a.name).split('.')[0] for a in n.names)
        elif isinstance(n, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef)):
            bound.add(n.name)
        elif isinstance(n, ast.Name) and isinstance(n.ctx, (ast.Store, ast.Del)):
            bound.add(n.id)
        elif isinstance(n, ast.arg):
            bound.add(n.arg)
        elif isinstance(n, ast.ExceptHandler) and n.name:
            bound.add(n.name)
        elif isinstance(n, (ast.Global, ast.Nonlocal)):
            bound.update(n.names)
    return sorted({n.id for n in ast.walk(tree)
                   if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Load) and n.id not in bound})


def test_no_unbound_names():
    files = glob.glob(os.path.join(PKG, "*.py")) + [os.path.join(ROOT, f) for f in ("bench.py", "__graft_entry__.py")]
    assert files
    bad = {os.path.basename(f): unbound_names(f) for f in files}
    assert not any(bad.values()), {k: v for k, v in bad.items() if v}
