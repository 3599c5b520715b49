"""A tiny stand-in for the ``taichi`` package, just big enough to EXECUTE the reference's kernels as they are written
(test infrastructure for golden-vector generation; never imported by the product or by the test-suite itself).

The reference's rasteriser lives in ``@ti.kernel`` / ``@ti.func`` bodies that need the third-party Taichi compiler,
which cannot be installed here.  Taichi kernels are, syntactically, Python.  This module provides the handful of
``ti.*`` names those bodies use so that the unmodified reference source runs under CPython on small scenes:

* ``ti.func`` / ``ti.kernel`` re-compile the function from its source after two mechanical AST rewrites --
  ``ti.atomic_add(a[i], v)`` becomes ``a[i] += v`` and ``ti.simt.block.sync()`` becomes ``yield`` -- and wrap it:
  vector / matrix arguments of a ``ti.func`` are passed BY VALUE (copied) as in Taichi, torch tensors handed to a
  kernel become numpy views (kernels write their outputs in place);
* kernels that use ``ti.simt.block.SharedArray`` are run as SIMT: the outermost ``for ... in ti.ndrange(N)`` loop
  body becomes a generator per thread, the ``block_dim`` threads of a block are advanced from barrier to barrier in
  lock-step and share the block's ``SharedArray`` instances;
* ``ti.math.vecN / matN``, ``ti.Vector`` / ``ti.Matrix``, ``ti.types.vector / matrix`` build small float32 numpy
  arrays (class ``Tensor``) with ``.x .y .z .w``, ``transpose()``, ``@``, ``sum()``, ...; scalar math maps to numpy
  float32 functions, so arithmetic is IEEE float32 like Taichi's default ``f32`` (without its fast-math).

What it does not model: parallel execution order (atomics are applied in thread order), Taichi's LLVM fast-math, and
anything the rasteriser path does not touch.
"""
import ast
import builtins as _py
import functools
import inspect
import math
import sys
import textwrap
import types

import numpy as np


# ------------------------------------------------------------------------------------------------ dtypes
class _DType:
    def __init__(self, name, np_type, is_int):
        self.name, self.np_type, self.is_int = name, np_type, is_int

    def __call__(self, x):
        return cast(x, self)

    def __repr__(self):
        return f"ti.{self.name}"


f32, f64 = _DType("f32", np.float32, False), _DType("f64", np.float64, False)
i8, i32, i64 = _DType("i8", np.int8, True), _DType("i32", np.int32, True), _DType("i64", np.int64, True)
u8, u32 = _DType("u8", np.uint8, True), _DType("u32", np.uint32, True)
float = float32 = f32  # noqa: A001  (ti.float)
int32, int64, float64 = i32, i64, f64
cpu = cuda = gpu = "arch"


def cast(x, dtype):
    if isinstance(x, np.ndarray):
        return x.astype(dtype.np_type)
    if dtype.is_int:
        return int(x)  # truncation toward zero, like Taichi's float -> int cast; Python int never overflows
    return dtype.np_type(x)


def init(*args, **kwargs):
    return None


def static(x):
    return x


def template():
    return "template"


def loop_config(block_dim=None, **kwargs):
    _simt.pending_block_dim = block_dim


def ndrange(*extents):
    if len(extents) == 1:
        return range(int(extents[0]))
    import itertools
    return itertools.product(*[range(int(e)) for e in extents])


def grouped(x):
    return np.ndindex(*x.shape)


def random(dtype=None):
    raise NotImplementedError("ti.random is not needed on the rasteriser path")


# ------------------------------------------------------------------------------------------------ small tensors
class Tensor(np.ndarray):
    """float32 vector / matrix value with the Taichi conveniences the reference uses."""

    def __new__(cls, data):
        return np.asarray(data, dtype=np.float32).view(cls)

    x = property(lambda s: s[0], lambda s, v: s.__setitem__(0, v))
    y = property(lambda s: s[1], lambda s, v: s.__setitem__(1, v))
    z = property(lambda s: s[2], lambda s, v: s.__setitem__(2, v))
    w = property(lambda s: s[3], lambda s, v: s.__setitem__(3, v))

    def transpose(self):
        return np.ndarray.transpose(self).copy()

    def to_numpy(self):
        return np.asarray(self).copy()

    def _seq_sum(self, values):
        # Taichi unrolls Matrix.sum() as ((e0 + e1) + e2) + ... in float32
        acc = np.float32(values[0])
        for v in values[1:]:
            acc = np.float32(acc + v)
        return acc

    def norm_sqr(self):
        flat = np.asarray(self, dtype=np.float32).reshape(-1)
        return self._seq_sum(flat * flat)

    def norm(self):
        return np.sqrt(self.norm_sqr())

    def normalized(self):
        # taichi/lang/matrix.py: invlen = 1 / (self.norm() + eps); return invlen * self
        invlen = np.float32(1.0) / self.norm()
        return invlen * self

    def dot(self, other):
        a = np.asarray(self, dtype=np.float32).reshape(-1)
        b = np.asarray(other, dtype=np.float32).reshape(-1)
        return self._seq_sum(a * b)

    def determinant(self):
        a = np.asarray(self)
        if a.shape == (2, 2):
            return a[0, 0] * a[1, 1] - a[0, 1] * a[1, 0]
        return np.float32(np.linalg.det(a.astype(np.float64)))

    def inverse(self):
        a = np.asarray(self)
        if a.shape == (2, 2):
            det = a[0, 0] * a[1, 1] - a[0, 1] * a[1, 0]
            return Tensor([[a[1, 1] / det, -a[0, 1] / det], [-a[1, 0] / det, a[0, 0] / det]])
        return Tensor(np.linalg.inv(a.astype(np.float64)))

    def trace(self):
        return np.float32(np.trace(np.asarray(self)))

    def outer_product(self, other):
        return Tensor(np.outer(np.asarray(self), np.asarray(other)))

    def sum(self, *args, **kwargs):  # a Taichi vector sum is a scalar, accumulated in element order
        return self._seq_sum(np.asarray(self, dtype=np.float32).reshape(-1))

    def __matmul__(self, other):
        # explicit float32 accumulation in index order (no BLAS), like the unrolled Taichi code
        a, b = np.asarray(self, dtype=np.float32), np.asarray(other, dtype=np.float32)
        if a.ndim == 2 and b.ndim == 1:
            out = np.zeros(a.shape[0], dtype=np.float32)
            for i in range(a.shape[0]):
                acc = np.float32(0.0)
                for k in range(a.shape[1]):
                    acc = np.float32(acc + a[i, k] * b[k])
                out[i] = acc
            return Tensor(out)
        if a.ndim == 2 and b.ndim == 2:
            out = np.zeros((a.shape[0], b.shape[1]), dtype=np.float32)
            for i in range(a.shape[0]):
                for j in range(b.shape[1]):
                    acc = np.float32(0.0)
                    for k in range(a.shape[1]):
                        acc = np.float32(acc + a[i, k] * b[k, j])
                    out[i, j] = acc
            return Tensor(out)
        if a.ndim == 1 and b.ndim == 2:
            return Tensor(np.asarray(Tensor(b.T) @ Tensor(a)))
        return np.float32((a * b).sum(dtype=np.float32))


def _flatten(args):
    out = []
    for a in args:
        if isinstance(a, (list, tuple, np.ndarray)):
            out.extend(_flatten(list(a)))
        else:
            out.append(a)
    return out


class _Field:
    """A 0-d Taichi field of vectors / matrices, as the reference's unit tests use it: from_numpy, f[None], to_numpy."""

    def __init__(self, tensor_type, shape=()):
        assert tuple(shape) == (), "only 0-d fields are modelled"
        self.value = tensor_type(0.0)

    def from_numpy(self, arr):
        self.value = Tensor(np.asarray(arr, dtype=np.float32).reshape(self.value.shape))

    def to_numpy(self):
        return np.asarray(self.value).copy()

    def __getitem__(self, index):
        return self.value

    def __setitem__(self, index, v):
        self.value = Tensor(np.asarray(v, dtype=np.float32).reshape(self.value.shape))


class _TensorType:
    def __init__(self, n, m=None):
        self.n, self.m = n, m

    def field(self, shape=()):
        return _Field(self, shape)

    def __call__(self, *args):
        n, m = self.n, self.m
        if m is None:
            flat = _flatten(args)
            if len(flat) == 1:
                flat = flat * n
            assert len(flat) == n, (n, flat)
            return Tensor(flat)
        if len(args) == 1 and isinstance(args[0], (list, tuple, np.ndarray)) and len(args[0]) and \
                isinstance(args[0][0], (list, tuple, np.ndarray)):
            arr = np.asarray([[np.float32(v) for v in _flatten([row])] for row in args[0]], dtype=np.float32)
        else:
            flat = _flatten(args)
            if len(flat) == 1:
                flat = flat * (n * m)
            arr = np.asarray(flat, dtype=np.float32).reshape(n, m)
        assert arr.shape == (n, m), (arr.shape, n, m)
        return Tensor(arr)


def _vector_type(n):
    return _TensorType(n)


def _matrix_type(n, m):
    return _TensorType(n, m)


def Vector(values, dt=None):  # noqa: N802
    return Tensor(_flatten([values]))


def Matrix(rows, dt=None):  # noqa: N802
    return Tensor([[np.float32(v) for v in _flatten([row])] for row in rows])


Matrix.zero = lambda dt, n, m=None: Tensor(np.zeros((n, m) if m else (n,), dtype=np.float32))
Matrix.identity = lambda dt, n: Tensor(np.eye(n, dtype=np.float32))
Matrix.rows = lambda rows: Tensor(np.stack([np.asarray(r, dtype=np.float32) for r in rows]))
Matrix.cols = lambda cols: Tensor(np.stack([np.asarray(c, dtype=np.float32) for c in cols], axis=1))
Vector.zero = lambda dt, n: Tensor(np.zeros(n, dtype=np.float32))


def _unary(np_fn):
    """float32 -> float32 elementary function, evaluated in double and rounded once: the correctly rounded float32
    result, independent of the libm / SIMD routine numpy happens to use (the oracle makes the same choice)."""
    def fn(x):
        if isinstance(x, np.ndarray):
            out = np_fn(x.astype(np.float64)).astype(np.float32)
            return out.view(type(x)) if isinstance(x, Tensor) else out
        return np.float32(np_fn(np.float64(np.float32(x))))
    return fn


sqrt, exp, log, sin, cos, tanh = (_unary(f) for f in (np.sqrt, np.exp, np.log, np.sin, np.cos, np.tanh))
floor, ceil = _unary(np.floor), _unary(np.ceil)


def abs(x):  # noqa: A001
    return np.abs(x)


def _fold(np_fn, py_fn, args):
    if any(isinstance(a, np.ndarray) for a in args):
        out = args[0]
        for a in args[1:]:
            out = np_fn(out, a)
        return out
    return py_fn(args)


def min(*args):  # noqa: A001
    return _fold(np.minimum, _py.min, args)


def max(*args):  # noqa: A001
    return _fold(np.maximum, _py.max, args)


def select(cond, a, b):
    return a if cond else b


# ti.math
math_ns = types.ModuleType("taichi.math")
for _n in (2, 3, 4):
    setattr(math_ns, f"vec{_n}", _vector_type(_n))
    setattr(math_ns, f"ivec{_n}", _vector_type(_n))
    setattr(math_ns, f"mat{_n}", _matrix_type(_n, _n))
math_ns.exp, math_ns.sqrt, math_ns.log, math_ns.sin, math_ns.cos = exp, sqrt, log, sin, cos
math_ns.pi = math.pi
math_ns.dot = lambda a, b: Tensor(a).dot(b)
math_ns.normalize = lambda v: Tensor(v).normalized()
math_ns.length = lambda v: Tensor(v).norm()
math_ns.cross = lambda a, b: Tensor(np.cross(np.asarray(a), np.asarray(b)))
math_ns.clamp = lambda x, lo, hi: min(max(x, lo), hi)
math_ns.min, math_ns.max, math_ns.floor = min, max, floor
math_ns.inverse = lambda m: Tensor(m).inverse()
math_ns.determinant = lambda m: Tensor(m).determinant()

# ti.types
types_ns = types.SimpleNamespace(
    ndarray=lambda *a, **k: "ndarray", vector=lambda n, dtype=None: _vector_type(n),
    matrix=lambda n, m, dtype=None: _matrix_type(n, m), struct=lambda **k: dict)


# ------------------------------------------------------------------------------------------------ SIMT emulation
class _Simt:
    pending_block_dim = None
    block_arrays = None     # SharedArray instances of the block being executed
    thread_cursor = None    # per-thread index of the next SharedArray() call
    current_thread = 0


_simt = _Simt()


def _shared_array(shape, dtype=f32):
    k = _simt.thread_cursor[_simt.current_thread]
    _simt.thread_cursor[_simt.current_thread] = k + 1
    if k == len(_simt.block_arrays):
        _simt.block_arrays.append(np.zeros(shape, dtype=dtype.np_type))
    return _simt.block_arrays[k]


def _run_blocks(thread_fn, n_threads, block_dim):
    block_dim = int(block_dim or 1)
    n_threads = int(n_threads)
    for first in range(0, n_threads, block_dim):
        ids = list(range(first, _py.min(first + block_dim, n_threads)))
        _simt.block_arrays, _simt.thread_cursor = [], [0] * len(ids)
        gens = [thread_fn(i) for i in ids]
        alive = list(range(len(ids)))
        while alive:  # advance every live thread to its next barrier
            still = []
            for t in alive:
                _simt.current_thread = t
                try:
                    next(gens[t])
                    still.append(t)
                except StopIteration:
                    pass
            alive = still
    _simt.block_arrays = _simt.thread_cursor = None


simt = types.SimpleNamespace(block=types.SimpleNamespace(
    SharedArray=_shared_array, sync=lambda: None, sync_all_nonzero=lambda predicate: predicate))


# ------------------------------------------------------------------------------------------------ source rewriting
def _is_ti_call(node, dotted):
    if not isinstance(node, ast.Call):
        return False
    f, parts = node.func, []
    while isinstance(f, ast.Attribute):
        parts.append(f.attr)
        f = f.value
    if isinstance(f, ast.Name):
        parts.append(f.id)
    return ".".join(reversed(parts)) == dotted


class _Rewriter(ast.NodeTransformer):
    def __init__(self):
        self.has_sync = False

    def visit_Expr(self, node):
        self.generic_visit(node)
        v = node.value
        if _is_ti_call(v, "ti.atomic_add") and len(v.args) == 2:
            target = v.args[0]
            target.ctx = ast.Store()
            return ast.copy_location(ast.AugAssign(target=target, op=ast.Add(), value=v.args[1]), node)
        if _is_ti_call(v, "ti.simt.block.sync"):
            self.has_sync = True
            return ast.copy_location(ast.Expr(value=ast.Yield(value=None)), node)
        return node


def _recompile(fn, simt_kernel=False):
    source = textwrap.dedent(inspect.getsource(fn))
    tree = ast.parse(source)
    fdef = tree.body[0]
    fdef.decorator_list = []
    for arg in fdef.args.args + fdef.args.kwonlyargs:
        arg.annotation = None
    fdef.returns = None
    rewriter = _Rewriter()
    rewriter.visit(fdef)
    if simt_kernel:
        # every top-level  `for v in ti.ndrange(N): <body with barriers>`  becomes
        #   def __ti_thread_k(v): <body>; yield
        #   __ti_run_blocks(__ti_thread_k, N, <block_dim of the preceding ti.loop_config>)
        body = []
        for k, stmt in enumerate(fdef.body):
            is_simt_loop = (isinstance(stmt, ast.For) and _is_ti_call(stmt.iter, "ti.ndrange") and
                            any(isinstance(n, ast.Yield) for n in ast.walk(stmt)))
            if not is_simt_loop:
                body.append(stmt)
                continue
            assert isinstance(stmt.target, ast.Name) and len(stmt.iter.args) == 1
            name = f"__ti_thread_{k}"
            thread = ast.FunctionDef(
                name=name, args=ast.arguments(posonlyargs=[], args=[ast.arg(arg=stmt.target.id)], kwonlyargs=[],
                                              kw_defaults=[], defaults=[]),
                body=stmt.body + [ast.Expr(value=ast.Yield(value=None))], decorator_list=[], type_params=[])
            run = ast.Expr(value=ast.Call(
                func=ast.Name(id="__ti_run_blocks", ctx=ast.Load()),
                args=[ast.Name(id=name, ctx=ast.Load()), stmt.iter.args[0],
                      ast.Call(func=ast.Name(id="__ti_take_block_dim", ctx=ast.Load()), args=[], keywords=[])],
                keywords=[]))
            body += [thread, run]
        fdef.body = body
    ast.fix_missing_locations(tree)
    namespace = fn.__globals__
    namespace.setdefault("__ti_run_blocks", _run_blocks)
    namespace.setdefault("__ti_take_block_dim", _take_block_dim)
    if fn.__closure__:  # a kernel defined inside a test method: give it the enclosing variables it refers to
        namespace = dict(namespace)
        namespace.update({name: cell.cell_contents for name, cell in zip(fn.__code__.co_freevars, fn.__closure__)})
    local = {}
    exec(compile(tree, inspect.getsourcefile(fn) or "<taichi_shim>", "exec"), namespace, local)
    new_fn = local[fdef.name]
    return new_fn


def _take_block_dim():
    dim, _simt.pending_block_dim = _simt.pending_block_dim, None
    return dim


def _by_value(a):
    return a.copy() if isinstance(a, Tensor) else a


def func(fn):
    compiled = _recompile(fn)

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        return compiled(*[_by_value(a) for a in args], **{k: _by_value(v) for k, v in kwargs.items()})
    return wrapper


def _to_numpy(a):
    if type(a).__module__.startswith("torch") and hasattr(a, "numpy"):
        assert a.device.type == "cpu" and a.is_contiguous(), "shim kernels take contiguous CPU tensors"
        return a.detach().numpy()
    return a


def kernel(fn):
    compiled = _recompile(fn, simt_kernel="SharedArray" in inspect.getsource(fn))

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        return compiled(*[_to_numpy(a) for a in args], **{k: _to_numpy(v) for k, v in kwargs.items()})
    return wrapper


def dataclass(cls):
    fields = list(getattr(cls, "__annotations__", {}))

    def __init__(self, *args, **kwargs):
        kwargs.update(dict(zip(fields, args)))  # Taichi structs also take their members positionally
        for name in fields:
            setattr(self, name, _by_value(kwargs[name]) if name in kwargs else None)
    cls.__init__ = __init__
    return cls


math = math_ns  # noqa: A001
types = types_ns  # noqa: A001


def install():
    """Register this module as ``taichi`` (and ``taichi.math``) in ``sys.modules``."""
    me = sys.modules[__name__]
    sys.modules["taichi"] = me
    sys.modules["taichi.math"] = math_ns
    return me
