"""A fake ``taichi`` namespace that executes the reference's *unmodified* kernels as plain Python.

TEST INFRASTRUCTURE ONLY (see oracle/README.md).  Taichi is not installable in the build
container, so the reference's device code (``render.py:2389-3489``) cannot be
JIT-compiled there.  Every ``ti.*`` use in the reference lives inside ``TaichiRenderer`` methods,
so placing this module in ``sys.modules['taichi']`` before ``import render`` lets the reference's
own kernels run, one Python iteration per pixel/texel, with float32 scalar semantics:

* scalars are ``numpy.float32`` (IEEE add/sub/mul/div/sqrt are exact f32 operations; Python
  literals are "weak" under NEP 50 and are rounded to f32 when they meet an f32 value, which is
  what Taichi's ``default_fp = f32`` does);
* ``ti.cast(x, ti.i32)`` truncates toward zero, integer ``%`` is Python-style, ``ti.pow`` with an
  integer exponent lowers to multiplications;
* transcendental functions are evaluated in double and rounded to f32 (an ideal f32 libm);
* every top-level ``for`` of a kernel runs sequentially, in order.

It is slow (~10 ms per traced pixel) and is used only by ``oracle/make_golden.py`` to produce the
small fixtures in ``tests/golden/`` that pin the C oracle (``oracle/bhr_oracle.c``).
"""
import itertools
import math as _m
import sys
import types

import numpy as np

F = np.float32


class _DType:
    def __init__(self, name):
        self.name = name

    def __repr__(self):
        return "ti." + self.name


f32 = _DType("f32")
i32 = _DType("i32")
cpu = "cpu"
gpu = "gpu"


def init(*a, **k):
    return None


class _Template:
    pass


def template():
    return _Template()


def _is_float(x):
    return isinstance(x, (float, np.floating))


def _to_f(x):
    return x if type(x) is F else F(x)


class Vector:
    """ti.Vector: small fixed-size f32 vector with Taichi's operation order."""
    __slots__ = ("v",)
    __array_ufunc__ = None      # make np.float32 <op> Vector defer to Vector.__r<op>__

    def __init__(self, comps):
        self.v = [_to_f(c) for c in comps]

    # taichi allows ti.Vector.field(...)
    @staticmethod
    def field(n, dtype=f32, shape=()):
        return _Field(shape, n=n, dtype=dtype)

    def __len__(self):
        return len(self.v)

    def __getitem__(self, i):
        return self.v[i]

    def __setitem__(self, i, val):
        self.v[i] = _to_f(val)

    def __iter__(self):
        return iter(self.v)

    def _bin(self, o, op):
        if isinstance(o, Vector):
            return Vector([op(a, b) for a, b in zip(self.v, o.v)])
        return Vector([op(a, o) for a in self.v])

    def _rbin(self, o, op):
        return Vector([op(o, a) for a in self.v])

    def __add__(self, o): return self._bin(o, lambda a, b: a + b)
    def __radd__(self, o): return self._rbin(o, lambda a, b: a + b)
    def __sub__(self, o): return self._bin(o, lambda a, b: a - b)
    def __rsub__(self, o): return self._rbin(o, lambda a, b: a - b)
    def __mul__(self, o): return self._bin(o, lambda a, b: a * b)
    def __rmul__(self, o): return self._rbin(o, lambda a, b: a * b)
    def __truediv__(self, o): return self._bin(o, lambda a, b: a / _to_f(b))
    def __neg__(self): return Vector([-a for a in self.v])

    def dot(self, o):
        acc = self.v[0] * o.v[0]
        for a, b in zip(self.v[1:], o.v[1:]):
            acc = acc + a * b
        return acc

    def cross(self, o):
        a, b = self.v, o.v
        return Vector([a[1] * b[2] - a[2] * b[1],
                       a[2] * b[0] - a[0] * b[2],
                       a[0] * b[1] - a[1] * b[0]])

    def norm(self):
        return sqrt(self.dot(self))

    def normalized(self):
        inv = F(1.0) / self.norm()
        return Vector([inv * a for a in self.v])

    def __repr__(self):
        return "Vector(%s)" % (self.v,)


class _Field:
    """Scalar or vector field backed by a numpy array."""

    def __init__(self, shape, n=None, dtype=f32):
        if isinstance(shape, int):
            shape = (shape,)
        self.shape = tuple(shape)
        self.n = n
        self.dtype = dtype
        np_dt = np.float32 if dtype is f32 else np.int32
        full = self.shape + ((n,) if n is not None else ())
        self.a = np.zeros(full, dtype=np_dt)

    def from_numpy(self, arr):
        arr = np.asarray(arr)
        assert arr.shape == self.a.shape, (arr.shape, self.a.shape)
        self.a[...] = arr

    def to_numpy(self):
        return self.a.copy()

    def _idx(self, idx):
        if idx is None:
            return ()
        if not isinstance(idx, tuple):
            idx = (idx,)
        return tuple(int(i) for i in idx)

    def __getitem__(self, idx):
        idx = self._idx(idx)
        if self.n is not None:
            return Vector(self.a[idx])
        val = self.a[idx]
        return int(val) if self.dtype is i32 else val

    def __setitem__(self, idx, val):
        idx = self._idx(idx)
        if self.n is not None:
            self.a[idx] = [c for c in val]
        else:
            self.a[idx] = val

    def __iter__(self):
        if len(self.shape) == 1:
            return iter(range(self.shape[0]))
        return itertools.product(*[range(s) for s in self.shape])


def field(dtype=f32, shape=()):
    return _Field(shape, n=None, dtype=dtype)


def ndrange(*dims):
    return itertools.product(*[range(int(d)) for d in dims])


def func(fn):
    return fn


def kernel(fn):
    ann = fn.__annotations__
    names = fn.__code__.co_varnames[:fn.__code__.co_argcount]

    def run(*args):
        conv = []
        for name, a in zip(names, args):
            t = ann.get(name)
            if t is f32:
                conv.append(F(a))
            elif t is i32:
                conv.append(int(a))
            else:
                conv.append(a)
        return fn(*conv)

    run.__name__ = fn.__name__
    return run


def cast(x, dtype):
    if dtype is i32:
        return int(x)          # truncation toward zero
    return F(x)


def _num(r, *ops):
    if any(_is_float(o) for o in ops):
        return F(r)
    return r


def max(a, b):  # noqa: A001  (mirrors ti.max)
    return _num(a if a >= b else b, a, b)


def min(a, b):  # noqa: A001
    return _num(a if a <= b else b, a, b)


def abs(x):  # noqa: A001
    return _num(-x if x < 0 else x, x)


def _f1(fn):
    def g(x):
        return F(fn(float(x)))
    return g


sqrt = _f1(_m.sqrt)
exp = _f1(_m.exp)
log = _f1(_m.log)
sin = _f1(_m.sin)
cos = _f1(_m.cos)
tan = _f1(_m.tan)
acos = _f1(_m.acos)
floor = _f1(_m.floor)


def atan2(y, x):
    return F(_m.atan2(float(y), float(x)))


def pow(x, y):  # noqa: A001
    if isinstance(y, (int, np.integer)) and not isinstance(y, bool) and 0 <= y <= 8:
        x = _to_f(x)
        r = F(1.0)
        for _ in range(int(y)):
            r = r * x
        return r
    return F(_m.pow(float(x), float(y)))


def _clamp(x, lo, hi):
    if isinstance(x, Vector):
        return Vector([min(max(c, lo), hi) for c in x.v])
    return min(max(x, lo), hi)


math_ns = types.SimpleNamespace(pi=_m.pi, clamp=_clamp)


def install():
    """Register this module as ``taichi`` (and a stub ``imageio.v3``) in sys.modules."""
    me = sys.modules[__name__]
    me.math = math_ns
    sys.modules["taichi"] = me
    if "imageio" not in sys.modules:
        try:
            import imageio.v3  # noqa: F401
        except Exception:
            iio = types.ModuleType("imageio")
            v3 = types.ModuleType("imageio.v3")
            iio.v3 = v3
            sys.modules["imageio"] = iio
            sys.modules["imageio.v3"] = v3
    return me
