#!/usr/bin/env python
"""Poor man's pyflakes (not in this image): reports names that are loaded but bound nowhere in the module / enclosing functions."""
import ast, builtins, sys


def bound_names(node):
    out = set()
    for n in ast.walk(node):
        if isinstance(n, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef)):
            out.add(n.name)
            if not isinstance(n, ast.ClassDef):
                a = n.args
                for x in a.posonlyargs + a.args + a.kwonlyargs:
                    out.add(x.arg)
                if a.vararg: out.add(a.vararg.arg)
                if a.kwarg: out.add(a.kwarg.arg)
        elif isinstance(n, ast.Lambda):
            for x in n.args.args: out.add(x.arg)
        elif isinstance(n, ast.Name) and isinstance(n.ctx, (ast.Store, ast.Del)):
            out.add(n.id)
        elif isinstance(n, (ast.Import, ast.ImportFrom)):
            for al in n.names:
                out.add((al.asname or al.name).split(".")[0])
        elif isinstance(n, ast.ExceptHandler) and n.name:
            out.add(n.name)
    return out


bad = 0
for path in sys.argv[1:]:
    tree = ast.parse(open(path).read())
    known = bound_names(tree) | set(dir(builtins)) | {"__file__", "__name__"}
    for n in ast.walk(tree):
        if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Load) and n.id not in known:
            print(f"{path}:{n.lineno}: undefined name {n.id}")
            bad += 1
sys.exit(1 if bad else 0)
