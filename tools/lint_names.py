#!/usr/bin/env python
"""Tiny offline lint (no pyflakes in the image): names that are loaded but bound nowhere in the file, and imports that
are never used.  Usage: python tools/lint_names.py [paths...]"""
import ast
import builtins
import os
import sys


def check(path):
    src = open(path).read()
    tree = ast.parse(src)
    noqa = {k + 1 for k, line in enumerate(src.splitlines()) if '# noqa' in line}
    bound, imported = set(dir(builtins)) | {'__file__', '__name__', '__doc__'}, {}
    for n in ast.walk(tree):
        if isinstance(n, (ast.Import, ast.ImportFrom)):
            for a in n.names:
                name = (a.asname or a.name).split('.')[0]
                bound.add(name)
                imported[name] = n.lineno
        elif isinstance(n, (ast.FunctionDef, ast.ClassDef, ast.AsyncFunctionDef)):
            bound.add(n.name)
        elif isinstance(n, ast.Name) and isinstance(n.ctx, (ast.Store, ast.Del)):
            bound.add(n.id)
        elif isinstance(n, ast.arg):
            bound.add(n.arg)
        elif isinstance(n, ast.ExceptHandler) and n.name:
            bound.add(n.name)
    loaded = {n.id for n in ast.walk(tree) if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Load)}
    problems = ['undefined name %r' % x for x in sorted(loaded - bound)]
    if not path.endswith('__init__.py'):
        problems += ['unused import %r (line %d)' % (k, v) for k, v in sorted(imported.items())
                     if k not in loaded and k != '*' and v not in noqa]
    return problems


if __name__ == '__main__':
    roots = sys.argv[1:] or [os.path.dirname(os.path.dirname(os.path.abspath(__file__)))]
    bad = 0
    for root in roots:
        files = [root] if root.endswith('.py') else [os.path.join(d, f) for d, _, fs in os.walk(root) for f in fs if f.endswith('.py')
                                                     and not any(x in d for x in ('.git', 'gpurun_out', '__pycache__', os.sep + 'build'))]
        for p in sorted(files):
            for msg in check(p):
                print('%s: %s' % (p, msg))
                bad += 1
    sys.exit(1 if bad else 0)
