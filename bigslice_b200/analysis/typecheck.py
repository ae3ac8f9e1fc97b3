"""Static analyzer for ``Session.run`` / ``Session.must`` call sites.

Walks python source with ``ast`` and, for every ``<x>.run(f, args...)`` /
``.must(...)`` whose first argument names a function decorated with
``@bs.Func`` (or assigned from ``bs.Func(...)``) in the same module, checks

* the number of arguments against the Func's signature, and
* literal argument types against the parameter annotations
  (``int``/``float``/``str``/``bool``), rejecting arguments that cannot be
  shipped between ranks (lambdas, generators),

reporting ``file:line: message`` diagnostics.

Capability parity: reference ``analysis/typecheck/typecheck.go:14-145`` and the
``cmd/slicetypecheck`` vet driver.
"""
from __future__ import annotations

import ast
import sys
from typing import Dict, List, Optional, Tuple

_LIT = {"int": int, "float": (int, float), "str": str, "bool": bool}


class _Func:
    def __init__(self, name: str, node: ast.AST):
        self.name = name
        self.node = node
        self.params: List[Tuple[str, Optional[str], bool]] = []
        self.variadic = False
        args = node.args
        ndef = len(args.defaults)
        pos = list(args.posonlyargs) + list(args.args)
        for i, a in enumerate(pos):
            ann = a.annotation.id if isinstance(a.annotation, ast.Name) else None
            self.params.append((a.arg, ann, i >= len(pos) - ndef))
        self.variadic = args.vararg is not None


def _is_func_decorator(d: ast.AST) -> bool:
    if isinstance(d, ast.Attribute) and d.attr in ("Func",):
        return True
    return isinstance(d, ast.Name) and d.id == "Func"


def _collect_funcs(tree: ast.AST) -> Dict[str, _Func]:
    out: Dict[str, _Func] = {}
    lambdas_or_defs: Dict[str, ast.AST] = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef):
            lambdas_or_defs[node.name] = node
            if any(_is_func_decorator(d) for d in node.decorator_list):
                out[node.name] = _Func(node.name, node)
    for node in ast.walk(tree):
        if isinstance(node, ast.Assign) and isinstance(node.value, ast.Call) \
                and _is_func_decorator(node.value.func) and node.value.args:
            arg = node.value.args[0]
            target = node.targets[0]
            if not isinstance(target, ast.Name):
                continue
            if isinstance(arg, ast.Lambda):
                out[target.id] = _Func(target.id, arg)
            elif isinstance(arg, ast.Name) and arg.id in lambdas_or_defs:
                out[target.id] = _Func(target.id, lambdas_or_defs[arg.id])
    return out


def check_source(src: str, filename: str = "<string>") -> List[str]:
    """Diagnostics for one module's source."""
    tree = ast.parse(src, filename)
    funcs = _collect_funcs(tree)
    diags: List[str] = []
    for node in ast.walk(tree):
        if not (isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute)
                and node.func.attr in ("run", "must", "Run", "Must") and node.args):
            continue
        first = node.args[0]
        if not (isinstance(first, ast.Name) and first.id in funcs):
            continue
        f = funcs[first.id]
        args = node.args[1:]
        if any(isinstance(a, ast.Starred) for a in args):
            continue
        required = sum(1 for _, _, has_def in f.params if not has_def)
        if (not f.variadic and len(args) > len(f.params)) or len(args) < required:
            diags.append("%s:%d: bigslice.Func %s takes %d arguments, but %d given"
                         % (filename, node.lineno, f.name, len(f.params), len(args)))
            continue
        for i, (a, (pname, ann, _)) in enumerate(zip(args, f.params)):
            if isinstance(a, (ast.Lambda, ast.GeneratorExp)):
                diags.append("%s:%d: argument %d (%s) of %s cannot be a %s: Func arguments must "
                             "be serialisable" % (filename, node.lineno, i, pname, f.name,
                                                  type(a).__name__.lower()))
                continue
            if ann in _LIT and isinstance(a, ast.Constant) and a.value is not None:
                ok = isinstance(a.value, _LIT[ann])
                if ann in ("int", "float") and isinstance(a.value, bool):
                    ok = False
                if not ok:
                    diags.append("%s:%d: argument %d (%s) of %s: expected %s, got %s"
                                 % (filename, node.lineno, i, pname, f.name, ann,
                                    type(a.value).__name__))
    return diags


def check_file(path: str) -> List[str]:
    with open(path) as fh:
        return check_source(fh.read(), path)


def main(argv=None) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        sys.stderr.write("usage: slicetypecheck file.py [...]\n")
        return 2
    n = 0
    for p in argv:
        for d in check_file(p):
            print(d)
            n += 1
    return 1 if n else 0


if __name__ == "__main__":
    sys.exit(main())
