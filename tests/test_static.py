"""Static check (no GPU): every name a function of bagel_b200 / bench.py / __graft_entry__.py reads is bound somewhere
(argument, local, enclosing scope, module global or builtin). The GPU-only code paths cannot be executed in the build
container, so a typo there would otherwise first show up on the B200 box."""
import ast
import builtins
import glob
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bound_names(node):
    names = set()
    for n in ast.walk(node):
        if isinstance(n, ast.Name) and isinstance(n.ctx, (ast.Store, ast.Del)):
            names.add(n.id)
        elif isinstance(n, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef)):
            names.add(n.name)
        elif isinstance(n, ast.arg):
            names.add(n.arg)
        elif isinstance(n, (ast.Import, ast.ImportFrom)):
            for a in n.names:
                names.add((a.asname or a.name).split(".")[0])
        elif isinstance(n, ast.ExceptHandler) and n.name:
            names.add(n.name)
    return names


def _check(path):
    tree = ast.parse(open(path).read())
    module_names = _bound_names(tree) | set(dir(builtins)) | {"__file__", "__name__"}
    problems = []

    def visit(fn, outer):
        scope = outer | _bound_names(fn)
        for n in ast.walk(fn):
            if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Load) and n.id not in scope:
                problems.append(f"{os.path.relpath(path, ROOT)}:{n.lineno}: undefined name {n.id!r} in {fn.name}()")

    for node in ast.walk(tree):
        if isinstance(node, (ast.FunctionDef, ast.AsyncFunctionDef)):
            visit(node, module_names)
    return problems


def test_no_undefined_names():
    files = glob.glob(os.path.join(ROOT, "bagel_b200", "*.py")) + [os.path.join(ROOT, "bench.py"),
                                                                   os.path.join(ROOT, "__graft_entry__.py")]
    files += glob.glob(os.path.join(ROOT, "tools", "*.py"))
    problems = [p for f in sorted(files) for p in _check(f)]
    assert not problems, "\n".join(problems)
