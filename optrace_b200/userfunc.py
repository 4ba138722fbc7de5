"""User callables on the device: Python -> CUDA source translation for FunctionSurface (func / mask_func /
deriv_func), RefractionIndex("Function") and TransmissionSpectrum("Function").

The reference calls arbitrary numpy callables on whole arrays (function_surface_2d.py:133-156).  With no CPU
on the ray path, a callable must become a __device__ function: its source is parsed with `ast` and a numpy
expression subset (arithmetic, comparisons, np.<ufunc> calls, np.where, constants from globals / closures /
the *_args dict, straight-line assignments before the return) is re-emitted as C++ with every operation in
the same order, so results agree with numpy to the rounding of the elementary functions.  The generated
header is compiled into a scene-specialised copy of the engine (build.build_library with
-DOTB_USER_FUNCS_H) and cached in-tree under csrc/jit/ keyed by source hash.

Anything outside the subset raises NotImplementedError naming the construct (OTB_ERR_UNSUPPORTED in C terms).
"""
from __future__ import annotations

import ast
import hashlib
import inspect
import math
import pathlib
import textwrap

import numpy as np

from . import build

JIT_DIR = build.CSRC / "jit"

# numpy name -> (C function, arity)
_FUNCS = {"cos": "cos", "sin": "sin", "tan": "tan", "arccos": "acos", "arcsin": "asin", "arctan": "atan",
          "acos": "acos", "asin": "asin", "atan": "atan", "arctan2": "atan2", "atan2": "atan2",
          "exp": "exp", "log": "log", "log10": "log10", "log2": "log2", "sqrt": "sqrt", "cbrt": "cbrt",
          "abs": "fabs", "absolute": "fabs", "fabs": "fabs", "tanh": "tanh", "sinh": "sinh", "cosh": "cosh",
          "hypot": "hypot", "power": "pow", "pow": "pow", "minimum": "fmin", "maximum": "fmax",
          "fmin": "fmin", "fmax": "fmax", "floor": "floor", "ceil": "ceil", "expm1": "expm1", "log1p": "log1p",
          "erf": "erf", "deg2rad": None, "rad2deg": None, "radians": None, "degrees": None,
          "square": None, "sign": None, "where": None, "float64": None, "asarray": None, "array": None,
          "ones_like": None, "zeros_like": None, "full_like": None, "logical_and": None, "logical_or": None,
          "logical_not": None, "isfinite": None}
_CONSTS = {"pi": math.pi, "e": math.e, "inf": math.inf, "nan": math.nan}


def _lit(v) -> str:
    v = float(v)
    if math.isnan(v):
        return 'nan("")'
    if math.isinf(v):
        return "INFINITY" if v > 0 else "(-INFINITY)"
    return repr(v) if ("e" in repr(v) or "." in repr(v)) else repr(v) + ".0"


class _Emitter:
    def __init__(self, fn, params, kwargs):
        self.fn, self.params, self.kwargs = fn, params, dict(kwargs)
        self.locals = set()
        cv = inspect.getclosurevars(fn)
        self.env = {**cv.globals, **cv.nonlocals}

    def fail(self, node, what):
        raise NotImplementedError(f"user callable {getattr(self.fn, '__name__', self.fn)}: {what} "
                                  f"(line {getattr(node, 'lineno', '?')}) is outside the translatable numpy subset")

    def name_value(self, node):
        n = node.id
        if n in self.params or n in self.locals:
            return n
        if n in self.kwargs:
            return _lit(self.kwargs[n])
        if n in self.env and isinstance(self.env[n], (int, float, np.integer, np.floating)):
            return _lit(self.env[n])
        if n in ("True", "False"):
            return "1.0" if n == "True" else "0.0"
        self.fail(node, f"name '{n}'")

    def module_attr(self, node):
        """np.pi, math.pi, np.cos ... -> (kind, name)"""
        if isinstance(node, ast.Attribute) and isinstance(node.value, ast.Name):
            mod = self.env.get(node.value.id)
            if mod is np or mod is math or node.value.id in ("np", "numpy", "math"):
                return node.attr
        return None

    def expr(self, node) -> str:
        if isinstance(node, ast.Constant):
            if isinstance(node.value, bool):
                return "1.0" if node.value else "0.0"
            if isinstance(node.value, (int, float)):
                return _lit(node.value)
            self.fail(node, f"constant {node.value!r}")
        if isinstance(node, ast.Name):
            return self.name_value(node)
        if isinstance(node, ast.Attribute):
            a = self.module_attr(node)
            if a in _CONSTS:
                return _lit(_CONSTS[a])
            self.fail(node, "attribute access")
        if isinstance(node, ast.UnaryOp):
            v = self.expr(node.operand)
            if isinstance(node.op, ast.USub):
                return f"(-{v})"
            if isinstance(node.op, ast.UAdd):
                return v
            if isinstance(node.op, (ast.Invert, ast.Not)):
                return f"(({v}) == 0.0 ? 1.0 : 0.0)"
        if isinstance(node, ast.BinOp):
            a, b = self.expr(node.left), self.expr(node.right)
            op = node.op
            if isinstance(op, ast.Add):
                return f"({a} + {b})"
            if isinstance(op, ast.Sub):
                return f"({a} - {b})"
            if isinstance(op, ast.Mult):
                return f"({a}*{b})"
            if isinstance(op, ast.Div):
                return f"({a}/{b})"
            if isinstance(op, ast.Pow):
                # numpy squares exactly (x*x); every other exponent goes through pow()
                if isinstance(node.right, ast.Constant) and node.right.value == 2:
                    return f"otb_sq({a})"
                return f"pow({a}, {b})"
            if isinstance(op, ast.BitAnd):
                return f"((({a}) != 0.0 && ({b}) != 0.0) ? 1.0 : 0.0)"
            if isinstance(op, ast.BitOr):
                return f"((({a}) != 0.0 || ({b}) != 0.0) ? 1.0 : 0.0)"
            if isinstance(op, ast.Mod):
                return f"otb_pymod({a}, {b})"
            self.fail(node, f"operator {type(op).__name__}")
        if isinstance(node, ast.BoolOp):
            vals = [f"(({self.expr(v)}) != 0.0)" for v in node.values]
            j = " && " if isinstance(node.op, ast.And) else " || "
            return f"(({j.join(vals)}) ? 1.0 : 0.0)"
        if isinstance(node, ast.Compare):
            parts, left = [], self.expr(node.left)
            for op, right in zip(node.ops, node.comparators):
                r = self.expr(right)
                sym = {ast.Lt: "<", ast.LtE: "<=", ast.Gt: ">", ast.GtE: ">=", ast.Eq: "==", ast.NotEq: "!="}.get(type(op))
                if sym is None:
                    self.fail(node, "comparison operator")
                parts.append(f"({left} {sym} {r})")
                left = r
            return f"(({' && '.join(parts)}) ? 1.0 : 0.0)"
        if isinstance(node, ast.IfExp):
            return f"((({self.expr(node.test)}) != 0.0) ? {self.expr(node.body)} : {self.expr(node.orelse)})"
        if isinstance(node, ast.Call):
            return self.call(node)
        self.fail(node, type(node).__name__)

    def call(self, node) -> str:
        name = self.module_attr(node.func)
        if name is None and isinstance(node.func, ast.Name) and node.func.id in ("abs", "float", "min", "max", "pow"):
            name = {"abs": "abs", "float": "float64", "min": "minimum", "max": "maximum", "pow": "power"}[node.func.id]
        if name is None or name not in _FUNCS:
            self.fail(node, f"call to {ast.unparse(node.func)}")
        args = [self.expr(a) for a in node.args]
        c = _FUNCS[name]
        if c is not None:
            return f"{c}({', '.join(args)})"
        if name in ("float64", "asarray", "array"):
            return args[0]
        if name == "square":
            return f"otb_sq({args[0]})"
        if name == "sign":
            return f"otb_sign({args[0]})"
        if name in ("deg2rad", "radians"):
            return f"({args[0]}*{_lit(math.pi/180)})"
        if name in ("rad2deg", "degrees"):
            return f"({args[0]}*{_lit(180/math.pi)})"
        if name == "where":
            return f"((({args[0]}) != 0.0) ? {args[1]} : {args[2]})"
        if name == "ones_like":
            return "1.0"
        if name == "zeros_like":
            return "0.0"
        if name == "full_like":
            return args[1]
        if name == "logical_and":
            return f"((({args[0]}) != 0.0 && ({args[1]}) != 0.0) ? 1.0 : 0.0)"
        if name == "logical_or":
            return f"((({args[0]}) != 0.0 || ({args[1]}) != 0.0) ? 1.0 : 0.0)"
        if name == "logical_not":
            return f"((({args[0]}) == 0.0) ? 1.0 : 0.0)"
        if name == "isfinite":
            return f"(isfinite({args[0]}) ? 1.0 : 0.0)"
        self.fail(node, f"call to {name}")


def _find_def(fn):
    """(params, body statements, return expression node) of a def or lambda"""
    try:
        src = textwrap.dedent(inspect.getsource(fn))
    except (OSError, TypeError) as e:
        raise NotImplementedError(f"source of user callable {fn} is not available: {e}")
    names = fn.__code__.co_varnames[:fn.__code__.co_argcount]
    try:
        tree = ast.parse(src)
    except SyntaxError:
        # a lambda inside a larger (multi-line) expression: isolate the longest parseable "lambda ..." text
        tree = None
        start = src.find("lambda")
        while start >= 0 and tree is None:
            for end in range(len(src), start + 6, -1):
                try:
                    cand = ast.parse(src[start:end].strip(), mode="eval")
                except SyntaxError:
                    continue
                if isinstance(cand.body, ast.Lambda) and tuple(a.arg for a in cand.body.args.args) == names:
                    tree = cand
                    break
            start = src.find("lambda", start + 6)
        if tree is None:
            raise NotImplementedError(f"could not parse the source of user callable {fn}")
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == fn.__name__:
            return [a.arg for a in node.args.args], node.body
        if isinstance(node, ast.Lambda) and fn.__name__ == "<lambda>" and tuple(a.arg for a in node.args.args) == names:
            return [a.arg for a in node.args.args], [ast.Return(value=node.body)]
    raise NotImplementedError(f"could not locate the definition of user callable {fn}")


def translate(kind: str, fn, kwargs: dict, cname: str) -> str:
    """C++ device function for one user callable"""
    nin = 1 if kind in ("surf1d", "mask1d", "deriv1d", "wl") else 2
    all_params, body = _find_def(fn)
    params = all_params[:nin]
    if len(params) != nin:
        raise NotImplementedError(f"user callable {fn} must take {nin} positional array argument(s)")
    # defaults of further parameters act like keyword arguments
    sig = inspect.signature(fn)
    kw = {k: p.default for k, p in sig.parameters.items() if p.default is not inspect.Parameter.empty}
    kw.update(kwargs)
    em = _Emitter(fn, params, kw)
    lines = []
    ret = None
    for stmt in body:
        if isinstance(stmt, ast.Expr) and isinstance(stmt.value, ast.Constant) and isinstance(stmt.value.value, str):
            continue   # docstring
        if isinstance(stmt, ast.Assign) and len(stmt.targets) == 1 and isinstance(stmt.targets[0], ast.Name):
            nm = stmt.targets[0].id
            decl = "" if nm in em.locals or nm in params else "double "
            lines.append(f"    {decl}{nm} = {em.expr(stmt.value)};")
            em.locals.add(nm)
        elif isinstance(stmt, ast.Return):
            ret = stmt.value
            break
        else:
            em.fail(stmt, f"statement {type(stmt).__name__}")
    if ret is None:
        raise NotImplementedError(f"user callable {fn} has no return statement")
    args = ", ".join(f"double {p}" for p in params)
    if kind == "deriv2d":
        if not isinstance(ret, ast.Tuple) or len(ret.elts) != 2:
            raise NotImplementedError("a 2-D derivative function must return a tuple (dz/dx, dz/dy)")
        lines.append(f"    *otb_dx = {em.expr(ret.elts[0])};")
        lines.append(f"    *otb_dy = {em.expr(ret.elts[1])};")
        return (f"__device__ __forceinline__ void {cname}({args}, double* otb_dx, double* otb_dy)\n{{\n"
                + "\n".join(lines) + "\n}\n")
    lines.append(f"    return {em.expr(ret)};")
    return f"__device__ __forceinline__ double {cname}({args})\n{{\n" + "\n".join(lines) + "\n}\n"


def generate_header(user_funcs) -> str:
    """user_funcs: list of (kind, callable, kwargs) in func-id order (scene.FlatScene.user_funcs)"""
    out = ["// GENERATED by optrace_b200/userfunc.py — user callables translated to device functions", "#pragma once",
           "#include <math.h>",
           "__device__ __forceinline__ double otb_sq(double v) { return v*v; }",
           "__device__ __forceinline__ double otb_sign(double v) { return (v > 0.0) ? 1.0 : ((v < 0.0) ? -1.0 : v); }",
           "__device__ __forceinline__ double otb_pymod(double a, double b) { double r = fmod(a, b); "
           "return (r != 0.0 && ((r < 0.0) != (b < 0.0))) ? r + b : r; }", ""]
    f1, f2, d2 = [], [], []
    for i, (kind, fn, kwargs) in enumerate(user_funcs):
        out.append(f"// id {i}: {kind} {getattr(fn, '__qualname__', fn)} {kwargs}")
        out.append(translate(kind, fn, kwargs, f"otb_uf_{i}"))
        (d2 if kind == "deriv2d" else (f1 if kind in ("surf1d", "mask1d", "deriv1d", "wl") else f2)).append(i)
    out.append("__device__ __forceinline__ double otb_user_f1(int id, double a)\n{\n    switch (id) {")
    out += [f"    case {i}: return otb_uf_{i}(a);" for i in f1]
    out.append('    default: return nan("");\n    }\n}\n')
    out.append("__device__ __forceinline__ double otb_user_f2(int id, double a, double b)\n{\n    switch (id) {")
    out += [f"    case {i}: return otb_uf_{i}(a, b);" for i in f2]
    out.append('    default: return nan("");\n    }\n}\n')
    out.append("__device__ __forceinline__ void otb_user_d2(int id, double a, double b, double* dx, double* dy)\n{\n    switch (id) {")
    out += [f"    case {i}: otb_uf_{i}(a, b, dx, dy); break;" for i in d2]
    out.append('    default: *dx = nan(""); *dy = nan(""); break;\n    }\n}\n')
    return "\n".join(out)


def build_specialised_library(user_funcs, api_only: bool = False) -> pathlib.Path:
    """engine variant with the given user callables compiled in; cached by header + engine source hash.
    api_only: only the array entry points (Surface.find_hit / normals / values on a stand-alone surface) see the
    callables; the trace kernels are linked from the base build (much faster to compile)."""
    header = generate_header(user_funcs)
    key = hashlib.sha256((header + build.source_digest() + ("api" if api_only else "")).encode()).hexdigest()[:16]
    JIT_DIR.mkdir(parents=True, exist_ok=True)
    lib = JIT_DIR / f"libotb_{key}.so"
    if lib.exists():
        return lib
    hdr = JIT_DIR / f"user_{key}.cuh"
    hdr.write_text(header)
    objdir = JIT_DIR / f"obj_{key}"
    flags = [f'-DOTB_USER_FUNCS_H="{hdr}"']
    build.build_library()
    keep = ("otb_api.cu",) if api_only else ("otb_api.cu", "otb_trace.cu", "otb_render.cu")
    base_objs = [build.CSRC / "build" / f.replace(".cu", ".o") for f in build.SOURCES if f not in keep]
    build.build_library(lib, extra_flags=flags, force=True, objdir=objdir, sources=list(keep), extra_objects=base_objs)
    return lib
