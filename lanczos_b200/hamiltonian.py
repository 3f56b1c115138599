"""Drop-in for Python/Regular/Hamiltonian.py: class Hamiltonian, without the N^3 Python loops.

The reference builds H = -T + V as scipy CSR matrices: `create_sparse_T` loops over N^3 rows in
Python (Hamiltonian.py:48-69, minutes at N = 160) and `create_sparse_V` calls the potential N^3
times (Hamiltonian.py:35-46).  Here both are descriptors of the matrix-free operator the Lanczos
kernels apply:

  * `T_sparse` is the 7- or 27-point stencil with the reference's weights (Hamiltonian.py:19-25,
    116-128) - no matrix is stored;
  * `V_sparse` is a diagonal whose values are computed ON THE DEVICE: the caller's unchanged NumPy
    potential function is called once with symbolic coordinates, which records it as a short postfix
    program; `lz_potential_eval` (csrc/potential.cu) runs that program per grid point;
  * `H = (-T_sparse + V_sparse); H.sort_indices()` (3Ddeuteron.py:80-81) yields a
    `lanczos_b200.StencilOperator`, which `Lanczos(H)` takes as it is.

`tocsr()` on any of them materialises the exact scipy matrix the reference would hold (for checks
on small grids).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _capi, engine
from .engine import Context, StencilOperator

# op codes of include/lanczos_b200.h
_OP = dict(X=0, Y=1, Z=2, CONST=3, ADD=4, SUB=5, MUL=6, DIV=7, POW=8, MIN=9, MAX=10, NEG=11, SQRT=12, EXP=13,
           LOG=14, ABS=15, SIN=16, COS=17, TANH=18, SQUARE=19)
_BINARY_UFUNC = {"add": "ADD", "subtract": "SUB", "multiply": "MUL", "divide": "DIV", "true_divide": "DIV",
                 "power": "POW", "float_power": "POW", "minimum": "MIN", "maximum": "MAX", "fmin": "MIN", "fmax": "MAX"}
_UNARY_UFUNC = {"negative": "NEG", "sqrt": "SQRT", "exp": "EXP", "log": "LOG", "absolute": "ABS", "fabs": "ABS",
                "sin": "SIN", "cos": "COS", "tanh": "TANH", "square": "SQUARE"}


class _Untraceable(Exception):
    pass


class _Sym:
    """A value of the potential expression while it is being traced: a postfix program."""
    __array_priority__ = 1000.0

    def __init__(self, tracer, ops):
        self.tracer, self.ops = tracer, ops

    # -- building blocks
    def _lift(self, other):
        if isinstance(other, _Sym):
            return other
        if isinstance(other, (int, float, np.integer, np.floating)):
            return _Sym(self.tracer, [_OP["CONST"] | (self.tracer.const(float(other)) << 8)])
        raise _Untraceable(f"operand of type {type(other).__name__}")

    def _bin(self, name, other, swap=False):
        o = self._lift(other)
        a, b = (o, self) if swap else (self, o)
        if name == "POW" and not swap and not isinstance(other, _Sym) and float(other) == 2.0:
            return _Sym(self.tracer, a.ops + [_OP["SQUARE"]])           # NumPy's own fast path for x**2
        return _Sym(self.tracer, a.ops + b.ops + [_OP[name]])

    def _un(self, name):
        return _Sym(self.tracer, self.ops + [_OP[name]])

    # -- Python operators
    def __add__(self, o): return self._bin("ADD", o)
    def __radd__(self, o): return self._bin("ADD", o, True)
    def __sub__(self, o): return self._bin("SUB", o)
    def __rsub__(self, o): return self._bin("SUB", o, True)
    def __mul__(self, o): return self._bin("MUL", o)
    def __rmul__(self, o): return self._bin("MUL", o, True)
    def __truediv__(self, o): return self._bin("DIV", o)
    def __rtruediv__(self, o): return self._bin("DIV", o, True)
    def __pow__(self, o): return self._bin("POW", o)
    def __rpow__(self, o): return self._bin("POW", o, True)
    def __neg__(self): return self._un("NEG")
    def __pos__(self): return self
    def __abs__(self): return self._un("ABS")

    def __bool__(self):
        raise _Untraceable("the potential branches on its arguments")

    __lt__ = __le__ = __gt__ = __ge__ = __eq__ = __ne__ = lambda self, o: (_ for _ in ()).throw(
        _Untraceable("the potential compares its arguments"))
    __hash__ = None

    # -- NumPy functions (np.sqrt(r), np.exp(...), ...) arrive here
    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        if method != "__call__" or kwargs:
            raise _Untraceable(f"np.{ufunc.__name__}.{method}")
        name = ufunc.__name__
        if name in _UNARY_UFUNC and len(inputs) == 1:
            return self._un(_UNARY_UFUNC[name])
        if name in _BINARY_UFUNC and len(inputs) == 2:
            a, b = inputs
            if isinstance(a, _Sym):
                return a._bin(_BINARY_UFUNC[name], b)
            return b._bin(_BINARY_UFUNC[name], a, True)
        raise _Untraceable(f"np.{name}")


class _Tracer:
    def __init__(self):
        self.consts = []

    def const(self, v):
        for i, c in enumerate(self.consts):
            if c == v and np.signbit(c) == np.signbit(v):
                return i
        self.consts.append(v)
        return len(self.consts) - 1

    def trace(self, potential):
        x, y, z = (_Sym(self, [_OP[k]]) for k in "XYZ")
        out = potential(x, y, z)
        if isinstance(out, (int, float, np.integer, np.floating)):
            out = x._lift(out)
        if not isinstance(out, _Sym):
            raise _Untraceable(f"the potential returned a {type(out).__name__}")
        return out.ops, self.consts


def evaluate_potential(potential, x, y, z, ctx: Context = None):
    """The diagonal potential on the grid x (fastest) x y x z as a CUDA tensor of len(x)*len(y)*len(z)
    doubles, index i + nx*(j + ny*k) (Hamiltonian.py:42,73-76).  Returns (tensor, how): how ==
    "device" when the function could be traced and ran in csrc/potential.cu; otherwise it is evaluated
    with NumPy on the host grid ("host": vectorised call, "host-scalar": np.vectorize) and uploaded."""
    torch = engine._torch()
    ctx = ctx or Context.default()
    x, y, z = (np.ascontiguousarray(a, dtype=np.float64) for a in (x, y, z))
    shape = (len(x), len(y), len(z))
    try:
        ops, consts = _Tracer().trace(potential)
        ops_a = np.asarray(ops, dtype=np.int32)
        consts_a = np.asarray(consts if consts else [0.0], dtype=np.float64)
        out = torch.empty(shape[0] * shape[1] * shape[2], dtype=torch.float64, device=ctx.torch_device)
        torch.cuda.current_stream(ctx.device).synchronize()
        _capi.check(ctx.lib.lz_potential_eval(
            ctx.handle, (C.c_int64 * 3)(*shape), x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p),
            z.ctypes.data_as(C.c_void_p), len(ops_a), ops_a.ctypes.data_as(C.c_void_p), len(consts),
            consts_a.ctypes.data_as(C.c_void_p), C.c_void_p(out.data_ptr())))
        return out, "device"
    except (_Untraceable, TypeError, ValueError):
        pass
    X, Y, Z = x[None, None, :], y[None, :, None], z[:, None, None]          # result[k, j, i] -> flat i + nx*(j + ny*k)
    try:
        vals = np.broadcast_to(np.asarray(potential(X, Y, Z), dtype=np.float64), (shape[2], shape[1], shape[0]))
        how = "host"
    except Exception:
        vals = np.vectorize(potential, otypes=[np.float64])(X, Y, Z)
        how = "host-scalar"
    return torch.from_numpy(np.array(vals, dtype=np.float64).reshape(-1)).to(ctx.torch_device), how


class MatrixFreeMatrix(StencilOperator):
    """A term of the structured-grid Hamiltonian that behaves like the scipy matrix the reference
    holds - `-T`, `T + V`, `a * T`, `H.sort_indices()`, `H.shape`, `H * vec`, `print(H)`, `H.tocsr()` -
    but stores only the stencil weights (centre, face, edge, corner) and an optional diagonal in HBM."""

    def __init__(self, N, weights, diag=None, bc="periodic"):
        self.weights = tuple(float(w) for w in weights)
        w0, w1, w2, w3 = self.weights
        if w2 == 0.0 and w3 == 0.0:                                   # 7-point family: the faster kernels
            super().__init__((N, N, N), w0, w1, bc=bc, diag=diag)
        else:
            super().__init__((N, N, N), 0.0, 0.0, bc=bc, diag=diag, weights27=self.weights)
        self.N = int(N)

    # -- the algebra 3Ddeuteron.py:80 uses
    def _diag_op(self, other, sign):
        a, b = self.diag, other.diag
        if a is None and b is None:
            return None
        torch = engine._torch()
        dev = Context.default().torch_device

        def dev_t(t):
            return t.to(dev) if isinstance(t, torch.Tensor) else torch.from_numpy(np.asarray(t, dtype=np.float64).reshape(-1)).to(dev)
        if a is None:
            return dev_t(b) * sign
        if b is None:
            return dev_t(a)
        return dev_t(a) + sign * dev_t(b)

    def __neg__(self):
        return self * -1.0

    def __add__(self, other):
        if not isinstance(other, MatrixFreeMatrix) or other.N != self.N or other.bc != self.bc:
            return NotImplemented
        return MatrixFreeMatrix(self.N, [a + b for a, b in zip(self.weights, other.weights)], self._diag_op(other, 1.0), self.bc)

    def __sub__(self, other):
        if not isinstance(other, MatrixFreeMatrix) or other.N != self.N or other.bc != self.bc:
            return NotImplemented
        return MatrixFreeMatrix(self.N, [a - b for a, b in zip(self.weights, other.weights)], self._diag_op(other, -1.0), self.bc)

    def __mul__(self, other):
        if isinstance(other, (int, float, np.integer, np.floating)):
            d = self.diag
            if d is not None:
                d = d * float(other)
            return MatrixFreeMatrix(self.N, [float(other) * w for w in self.weights], d, self.bc)
        return self.matvec(other)                                    # H * vec (Lanczos.py:108)

    def __rmul__(self, other):
        if isinstance(other, (int, float, np.integer, np.floating)):
            return self * other
        return NotImplemented

    def sort_indices(self):
        """scipy's in-place canonicalisation (3Ddeuteron.py:81): a stencil has no stored order."""
        return None

    @property
    def nnz(self):
        return (7 if self.weights27 is None else 27) * self.M

    def get(self):
        return self.tocsr()

    def __repr__(self):
        kind = "7-point" if self.weights27 is None else "27-point"
        return ("<%dx%d matrix-free %s %s stencil operator, weights (centre, face, edge, corner) = %s%s>"
                % (self.M, self.M, self.bc, kind, self.weights, "" if self.diag is None else ", diagonal potential in HBM"))

    __str__ = __repr__


class Hamiltonian:
    """Mirror of the reference class (Python/Regular/Hamiltonian.py:6-128): same constructor, attributes
    and method names; `create_sparse_T` / `create_sparse_V` take milliseconds instead of minutes."""

    def __init__(self, N, L, potential, T_factor):
        self.N = N
        self.L = L
        self.potential = potential
        self.T_factor = T_factor
        self.dx = float(L) / N
        self.x = np.linspace(-L / 2, L / 2, N)                       # Hamiltonian.py:15-17 (as is: SURVEY.md §9.4)
        self.y = np.linspace(-L / 2, L / 2, N)
        self.z = np.linspace(-L / 2, L / 2, N)
        self.neighbors_relative_7point = np.array([[0, 0, 0], [-1, 0, 0], [0, -1, 0], [0, 0, -1], [1, 0, 0], [0, 1, 0], [0, 0, 1]])
        self.weights_7point = np.ones(7)
        self.weights_7point[0] = -6
        self.neighbors_relative_27point = np.array([[i, j, k] for i in range(-1, 2) for j in range(-1, 2) for k in range(-1, 2)])
        self.weights_27point = self.get_weights_27point()
        self.potential_evaluated_on = None

    def create_sparse_Hamiltonian(self):
        pass

    # ---- Hamiltonian.py:35-46 ------------------------------------------------------------------
    def create_sparse_V(self):
        print("+++ Setting up sparse potential matrix V.")
        # the reference calls potential(self.x[i], self.y[j], self.z[k]) for idx = i + j*N + k*N^2
        diag, how = evaluate_potential(self.potential, self.x, self.y, self.z)
        self.potential_evaluated_on = how
        self.V_sparse = MatrixFreeMatrix(self.N, (0.0, 0.0, 0.0, 0.0), diag)

    # ---- Hamiltonian.py:48-69 ------------------------------------------------------------------
    def create_sparse_T(self, points="27", save_cache=False):
        """`save_cache=True` also writes T_matrices/T_N=<N>_Laplace=<points>.npz like the reference
        (the CSR is exported from the device operator); an existing cache file is honoured: its
        weights are read back and used (lanczos_b200.io.stencil_from_t_matrix)."""
        print("+++ Setting up sparse laplacian matrix T.")
        from . import io as lzio
        points = str(points)
        if points not in ("7", "27"):
            raise UnboundLocalError("local variable 'Laplacian' referenced before assignment")   # Hamiltonian.py:56-60
        cached = lzio.load_t_matrix(self.N, points) if self.N >= 3 else None
        if cached is not None:
            print("+++ Laplacian matrix T for N = %d and %s points already created. Extracting..." % (self.N, points))
            op = lzio.stencil_from_t_matrix(cached, self.N, sign=1.0)
            w = op.weights27 if op.weights27 is not None else (op.center, op.off[0], 0.0, 0.0)
            self.T_sparse = MatrixFreeMatrix(self.N, w)
            return
        print("+++ Laplacian matrix T for N = %d and %s does not exist. Creating..." % (self.N, points))
        if points == "7":
            w = (self.T_factor * -6.0, self.T_factor * 1.0, 0.0, 0.0)
        else:
            w27 = self.weights_27point                                # ordered like neighbors_relative_27point
            kinds = np.sum(self.neighbors_relative_27point != 0, axis=1)
            w = tuple(self.T_factor * float(w27[np.argmax(kinds == k)]) for k in range(4))
        self.T_sparse = MatrixFreeMatrix(self.N, w)
        if save_cache:
            lzio.save_t_matrix(self.T_sparse.tocsr(), self.N, points)

    # ---- index helpers (Hamiltonian.py:73-84) -----------------------------------------------------
    def unravel_xyz(self, x, y, z):
        return x + y * self.N + z * self.N ** 2

    def ravel_i(self, i):
        N = self.N
        return (i % N, (i // N) % N, i // N ** 2)

    def _wrapped_neighbors(self, i, relative):
        x, y, z = self.ravel_i(i)
        nb = (relative + np.array([x, y, z])) % self.N               # periodic wrap, Hamiltonian.py:92-97
        return [int(self.unravel_xyz(a, b, c)) for a, b, c in nb]

    def Laplacian_7point(self, i):
        return self._wrapped_neighbors(i, self.neighbors_relative_7point), self.weights_7point

    def Laplacian_27point(self, i):
        return self._wrapped_neighbors(i, self.neighbors_relative_27point), self.weights_27point

    def get_weights_27point(self):
        """Hamiltonian.py:116-128: 3/13 * (-44/3 centre, 1 face, 1/2 edge, 1/3 corner)."""
        kinds = np.sum(self.neighbors_relative_27point != 0, axis=1)
        return np.array([-44 / 3, 1.0, 1.0 / 2, 1.0 / 3])[kinds] * 3.0 / 13


__all__ = ["Hamiltonian", "MatrixFreeMatrix", "evaluate_potential"]
