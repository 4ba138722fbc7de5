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
import os
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
        self.vec_locals = {}        # name -> (c0, c1, c2): locals holding an (N, 3) array (orientation functions)
        self.pre = []               # statements to emit before the expression under construction (vector temporaries)
        self.ntmp = 0
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
        if isinstance(node, ast.Subscript):
            k = self._component_index(node)
            if k is not None:                       # v[:, k] / v[..., k] of an (N, 3) value
                v = self.vec(node.value)
                if v is None:
                    self.fail(node, "component subscript of a non-vector value")
                return v[k]
            if self._is_newaxis(node):              # x[:, None] / x[:, np.newaxis]: broadcasting helper, same scalar
                return self.expr(node.value)
            self.fail(node, "subscript")
        if isinstance(node, ast.Call):
            return self.call(node)
        self.fail(node, type(node).__name__)

    # ---- (N, 3) array values of orientation functions: per ray a 3-vector ------------------------------------
    def _is_none_like(self, n) -> bool:
        return (isinstance(n, ast.Constant) and n.value is None) or self.module_attr(n) == "newaxis"

    def _is_full_slice(self, n) -> bool:
        return (isinstance(n, ast.Slice) and n.lower is None and n.upper is None and n.step is None) or \
            (isinstance(n, ast.Constant) and n.value is Ellipsis)

    def _component_index(self, node):
        sl = node.slice
        if isinstance(sl, ast.Tuple) and len(sl.elts) == 2 and self._is_full_slice(sl.elts[0]) \
                and isinstance(sl.elts[1], ast.Constant) and isinstance(sl.elts[1].value, int) and not isinstance(sl.elts[1].value, bool):
            k = sl.elts[1].value
            return k % 3 if -3 <= k < 3 else None
        return None

    def _is_newaxis(self, node) -> bool:
        sl = node.slice
        return isinstance(sl, ast.Tuple) and len(sl.elts) == 2 and self._is_full_slice(sl.elts[0]) and self._is_none_like(sl.elts[1])

    def _tmp(self, value: str) -> str:
        self.ntmp += 1
        nm = f"otb_t{self.ntmp}"
        self.pre.append(f"    const double {nm} = {value};")
        return nm

    def _three(self, node):
        """the three element expressions of a tuple / list literal, else None"""
        if isinstance(node, (ast.Tuple, ast.List)) and len(node.elts) == 3:
            return tuple(self._tmp(self.expr(e)) for e in node.elts)
        return None

    def vec(self, node):
        """(c0, c1, c2) C expressions when `node` evaluates to an (N, 3) array, else None"""
        if isinstance(node, ast.Name):
            return self.vec_locals.get(node.id)
        if isinstance(node, ast.UnaryOp) and isinstance(node.op, (ast.USub, ast.UAdd)):
            v = self.vec(node.operand)
            if v is None:
                return None
            return tuple(f"(-{c})" for c in v) if isinstance(node.op, ast.USub) else v
        if isinstance(node, ast.Attribute) and node.attr == "T":            # np.array([a, b, c]).T / np.vstack(...).T
            inner = node.value
            if isinstance(inner, ast.Call) and self.module_attr(inner.func) in ("array", "asarray", "vstack", "stack") \
                    and inner.args and not inner.keywords:
                return self._three(inner.args[0])
            return None
        if isinstance(node, ast.Call):
            name = self.module_attr(node.func)
            fname = node.func.attr if isinstance(node.func, ast.Attribute) else (node.func.id if isinstance(node.func, ast.Name) else None)
            if name == "column_stack" and len(node.args) == 1:
                return self._three(node.args[0])
            if name == "stack" and len(node.args) >= 1:
                ax = node.args[1] if len(node.args) > 1 else next((k.value for k in node.keywords if k.arg == "axis"), None)
                if isinstance(ax, ast.UnaryOp) and isinstance(ax.op, ast.USub) and isinstance(ax.operand, ast.Constant):
                    axv = -ax.operand.value
                else:
                    axv = ax.value if isinstance(ax, ast.Constant) else None
                if axv in (1, -1):
                    return self._three(node.args[0])
                return None
            if name == "transpose" and len(node.args) == 1 and isinstance(node.args[0], ast.Call) \
                    and self.module_attr(node.args[0].func) in ("array", "asarray", "vstack"):
                return self._three(node.args[0].args[0])
            if fname == "normalize" and len(node.args) == 1:               # misc.normalize (misc.py:136-150)
                v = self.vec(node.args[0])
                if v is None:
                    return None
                l = self._tmp(f"sqrt((({v[0]}*{v[0]}) + ({v[1]}*{v[1]})) + ({v[2]}*{v[2]}))")
                return tuple(self._tmp(f"({c}/{l})") for c in v)
            if name == "where" and len(node.args) == 3:
                a, b = self.vec(node.args[1]), self.vec(node.args[2])
                if a is None and b is None:
                    return None
                c = self.expr(node.args[0])
                a = a or (self.expr(node.args[1]),)*3
                b = b or (self.expr(node.args[2]),)*3
                return tuple(f"((({c}) != 0.0) ? {x} : {y})" for x, y in zip(a, b))
            return None
        if isinstance(node, ast.BinOp):
            a, b = self.vec(node.left), self.vec(node.right)
            if a is None and b is None:
                return None
            sym = {ast.Add: "+", ast.Sub: "-", ast.Mult: "*", ast.Div: "/"}.get(type(node.op))
            if sym is None:
                self.fail(node, f"operator {type(node.op).__name__} on an (N, 3) value")
            if a is None:
                sa = self._tmp(self.expr(node.left))
                a = (sa, sa, sa)
            if b is None:
                sb = self._tmp(self.expr(node.right))
                b = (sb, sb, sb)
            return tuple(f"({x} {sym} {y})" for x, y in zip(a, b))
        return None

    def call(self, node) -> str:
        # np.linalg.norm(v, axis=1) of an (N, 3) value
        if isinstance(node.func, ast.Attribute) and node.func.attr == "norm" and node.args:
            v = self.vec(node.args[0])
            if v is not None:
                return f"sqrt((({v[0]}*{v[0]}) + ({v[1]}*{v[1]})) + ({v[2]}*{v[2]}))"
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
    def flush():
        lines.extend(em.pre)
        em.pre.clear()

    for stmt in body:
        if isinstance(stmt, ast.Expr) and isinstance(stmt.value, ast.Constant) and isinstance(stmt.value.value, str):
            continue   # docstring
        if isinstance(stmt, ast.Assign) and len(stmt.targets) == 1 and isinstance(stmt.targets[0], ast.Name):
            nm = stmt.targets[0].id
            v = em.vec(stmt.value) if kind == "orient" else None
            if v is not None:                   # local holding an (N, 3) array
                flush()
                em.ntmp += 1
                names = tuple(f"{nm}_{em.ntmp}_{c}" for c in "xyz")
                for n_, c_ in zip(names, v):
                    lines.append(f"    const double {n_} = {c_};")
                em.vec_locals[nm] = names
                em.locals.discard(nm)
                continue
            val = em.expr(stmt.value)
            flush()
            decl = "" if nm in em.locals or nm in params else "double "
            lines.append(f"    {decl}{nm} = {val};")
            em.locals.add(nm)
            em.vec_locals.pop(nm, None)
        elif isinstance(stmt, ast.Return):
            ret = stmt.value
            break
        else:
            em.fail(stmt, f"statement {type(stmt).__name__}")
    if ret is None:
        raise NotImplementedError(f"user callable {fn} has no return statement")
    args = ", ".join(f"double {p}" for p in params)
    if kind == "orient":
        v = em.vec(ret)
        if v is None:
            raise NotImplementedError(f"orientation function {fn}: the return value must be an (N, 3) array expression "
                                      "(np.column_stack / np.stack(axis=1) / np.array([...]).T, optionally normalised)")
        flush()
        for c_, o_ in zip(v, ("otb_x", "otb_y", "otb_z")):
            lines.append(f"    *{o_} = {c_};")
        return (f"__device__ __forceinline__ void {cname}({args}, double* otb_x, double* otb_y, double* otb_z)\n{{\n"
                + "\n".join(lines) + "\n}\n")
    if kind == "deriv2d":
        if not isinstance(ret, ast.Tuple) or len(ret.elts) != 2:
            raise NotImplementedError("a 2-D derivative function must return a tuple (dz/dx, dz/dy)")
        dxv, dyv = em.expr(ret.elts[0]), em.expr(ret.elts[1])
        flush()
        lines.append(f"    *otb_dx = {dxv};")
        lines.append(f"    *otb_dy = {dyv};")
        return (f"__device__ __forceinline__ void {cname}({args}, double* otb_dx, double* otb_dy)\n{{\n"
                + "\n".join(lines) + "\n}\n")
    rv = em.expr(ret)
    flush()
    lines.append(f"    return {rv};")
    return f"__device__ __forceinline__ double {cname}({args})\n{{\n" + "\n".join(lines) + "\n}\n"


def generate_header(user_funcs) -> str:
    """user_funcs: list of (kind, callable, kwargs) in func-id order (scene.FlatScene.user_funcs)"""
    out = ["// GENERATED by optrace_b200/userfunc.py — user callables translated to device functions", "#pragma once",
           "#include <math.h>",
           "__device__ __forceinline__ double otb_sq(double v) { return v*v; }",
           "__device__ __forceinline__ double otb_sign(double v) { return (v > 0.0) ? 1.0 : ((v < 0.0) ? -1.0 : v); }",
           "__device__ __forceinline__ double otb_pymod(double a, double b) { double r = fmod(a, b); "
           "return (r != 0.0 && ((r < 0.0) != (b < 0.0))) ? r + b : r; }", ""]
    f1, f2, d2, v3 = [], [], [], []
    for i, (kind, fn, kwargs) in enumerate(user_funcs):
        out.append(f"// id {i}: {kind} {getattr(fn, '__qualname__', fn)} {kwargs}")
        out.append(translate(kind, fn, kwargs, f"otb_uf_{i}"))
        (v3 if kind == "orient" else d2 if kind == "deriv2d" else
         (f1 if kind in ("surf1d", "mask1d", "deriv1d", "wl") else f2)).append(i)
    out.append("__device__ __forceinline__ double otb_user_f1(int id, double a)\n{\n    switch (id) {")
    out += [f"    case {i}: return otb_uf_{i}(a);" for i in f1]
    out.append('    default: return nan("");\n    }\n}\n')
    out.append("__device__ __forceinline__ double otb_user_f2(int id, double a, double b)\n{\n    switch (id) {")
    out += [f"    case {i}: return otb_uf_{i}(a, b);" for i in f2]
    out.append('    default: return nan("");\n    }\n}\n')
    out.append("__device__ __forceinline__ void otb_user_d2(int id, double a, double b, double* dx, double* dy)\n{\n    switch (id) {")
    out += [f"    case {i}: otb_uf_{i}(a, b, dx, dy); break;" for i in d2]
    out.append('    default: *dx = nan(""); *dy = nan(""); break;\n    }\n}\n')
    out.append("__device__ __forceinline__ void otb_user_v3(int id, double a, double b, double* x, double* y, double* z)\n{\n    switch (id) {")
    out += [f"    case {i}: otb_uf_{i}(a, b, x, y, z); break;" for i in v3]
    out.append('    default: *x = *y = *z = nan(""); break;\n    }\n}\n')
    return "\n".join(out)


def build_specialised_library(user_funcs, api_only: bool = False) -> pathlib.Path:
    """engine variant with the given user callables compiled in; cached by header + engine source hash.
    api_only: only the array entry points (Surface.find_hit / normals / values on a stand-alone surface) see the
    callables; the trace kernels are linked from the base build (much faster to compile)."""
    header = generate_header(user_funcs)
    # OTB_NVCC_EXTRA: extra nvcc flags for variant builds (kernel-tuning experiments, e.g. -DOTB_TRACE_THREADS_FULL=512)
    extra = os.environ.get("OTB_NVCC_EXTRA", "").split()
    key = hashlib.sha256((header + build.source_digest() + ("api" if api_only else "") + " ".join(extra)).encode()).hexdigest()[:16]
    JIT_DIR.mkdir(parents=True, exist_ok=True)
    lib = JIT_DIR / f"libotb_{key}.so"
    if lib.exists():
        return lib
    hdr = JIT_DIR / f"user_{key}.cuh"
    hdr.write_text(header)
    objdir = JIT_DIR / f"obj_{key}"
    flags = [f'-DOTB_USER_FUNCS_H="{hdr}"', *extra]
    build.build_library()
    keep = ("otb_api.cu",) if api_only else ("otb_api.cu", "otb_trace.cu", "otb_render.cu")
    if any(k == "orient" for k, _, _ in user_funcs):
        keep += ("otb_gen.cu",)          # the stand-alone generator calls orientation functions too
    base_objs = [build.CSRC / "build" / f.replace(".cu", ".o") for f in build.SOURCES if f not in keep]
    build.build_library(lib, extra_flags=flags, force=True, objdir=objdir, sources=list(keep), extra_objects=base_objs)
    return lib
