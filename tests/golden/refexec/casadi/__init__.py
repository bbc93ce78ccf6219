"""Numeric stand-in for the `casadi` module.  TEST INFRASTRUCTURE ONLY.

The reference (Thieso/BoundMPC) builds its NLP symbolically with CasADi `SX`
(casadi_ocp_formulation.py:9-391).  CasADi is not installed in this image, so
to *execute the reference's own formulation code* we put this module on
PYTHONPATH ahead of it: every `SX.sym` returns numbers (with a trailing batch
axis, complex dtype so that complex-step differentiation works) and every
operator evaluates eagerly.  `nlpsol` only captures `prob`; the captured
callable can be pointed at any solver (tests plug the repo's solver in).

Semantics reproduced (those the reference relies on):
  * everything is a 2-D matrix; a single index is column-major linear indexing
    that keeps the row/column orientation of vectors;
  * `M[r, :]`, `M[a:b, :]`, `.T`, `.reshape((-1, 1))` (column-major);
  * scalar (1x1) broadcasting, `@`, comparisons on 1x1;
  * `if_else` selects by branch (never by arithmetic masking).
"""
import numpy as np

_state = {"provider": None}


def set_provider(p):
    _state["provider"] = p


class M:
    __array_priority__ = 1000

    def __init__(self, a):
        a = np.asarray(a)
        assert a.ndim == 3, a.shape
        self.a = a

    # -- shape ---------------------------------------------------------
    @property
    def shape(self):
        return self.a.shape[:2]

    def __len__(self):
        return self.a.shape[0]

    @property
    def T(self):
        return M(np.swapaxes(self.a, 0, 1))

    def reshape(self, shp):
        r, c = shp
        flat = np.swapaxes(self.a, 0, 1).reshape(-1, self.a.shape[2])  # col-major
        n = flat.shape[0]
        if r == -1:
            r = n // c
        if c == -1:
            c = n // r
        out = flat.reshape(c, r, -1)
        return M(np.swapaxes(out, 0, 1))

    def numel(self):
        return self.a.shape[0] * self.a.shape[1]

    # -- indexing ------------------------------------------------------
    def _lin(self):
        return np.swapaxes(self.a, 0, 1).reshape(-1, self.a.shape[2])

    def __getitem__(self, k):
        if isinstance(k, tuple):
            r, c = k
            if isinstance(r, (int, np.integer)):
                r = slice(r, r + 1) if r != -1 else slice(-1, None)
            if isinstance(c, (int, np.integer)):
                c = slice(c, c + 1) if c != -1 else slice(-1, None)
            return M(self.a[r, c, :])
        flat = self._lin()
        if isinstance(k, (int, np.integer)):
            return M(flat[k].reshape(1, 1, -1))
        sel = flat[k]
        if self.a.shape[0] == 1 and self.a.shape[1] != 1:
            return M(sel.reshape(1, sel.shape[0], -1))
        return M(sel.reshape(sel.shape[0], 1, -1))

    def __setitem__(self, k, v):
        v = _lift(v, self.a.shape[2])
        if v.a.dtype.kind == 'c' and self.a.dtype.kind != 'c':
            self.a = self.a.astype(complex)
        if v.a.shape[2] != self.a.shape[2]:
            if self.a.shape[2] == 1:
                self.a = np.repeat(self.a, v.a.shape[2], axis=2)
            else:
                v = M(np.repeat(v.a, self.a.shape[2], axis=2))
        if isinstance(k, tuple):
            r, c = k
            if isinstance(r, (int, np.integer)):
                r = slice(r, r + 1)
            if isinstance(c, (int, np.integer)):
                c = slice(c, c + 1)
            tgt = self.a[r, c, :]
            self.a[r, c, :] = v._lin().reshape(tgt.shape[1], tgt.shape[0], -1).swapaxes(0, 1)
            return
        rr, cc = self.a.shape[:2]
        idx = np.arange(rr * cc)[k]
        idx = np.atleast_1d(idx)
        vals = v._lin()
        if vals.shape[0] == 1 and idx.shape[0] > 1:
            vals = np.repeat(vals, idx.shape[0], axis=0)
        for n_, i_ in enumerate(idx):
            self.a[i_ % rr, i_ // rr, :] = vals[n_]

    # -- arithmetic ----------------------------------------------------
    def _bin(self, o, f, rev=False):
        o = _lift(o, self.a.shape[2])
        x, y = (o.a, self.a) if rev else (self.a, o.a)
        if x.shape[:2] != y.shape[:2]:
            if x.shape[:2] == (1, 1) or y.shape[:2] == (1, 1):
                pass
            elif x.shape[0] * x.shape[1] == y.shape[0] * y.shape[1] and 1 in x.shape[:2] and 1 in y.shape[:2]:
                raise ValueError(f"orientation mismatch {x.shape} vs {y.shape}")
            else:
                raise ValueError(f"shape mismatch {x.shape} vs {y.shape}")
        return M(f(x, y))

    def __add__(self, o): return self._bin(o, np.add)
    def __radd__(self, o): return self._bin(o, np.add, True)
    def __sub__(self, o): return self._bin(o, np.subtract)
    def __rsub__(self, o): return self._bin(o, np.subtract, True)
    def __mul__(self, o): return self._bin(o, np.multiply)
    def __rmul__(self, o): return self._bin(o, np.multiply, True)
    def __truediv__(self, o): return self._bin(o, np.divide)
    def __rtruediv__(self, o): return self._bin(o, np.divide, True)
    def __neg__(self): return M(-self.a)
    def __pos__(self): return self

    def __pow__(self, e):
        if isinstance(e, (int, np.integer)):
            out = np.ones_like(self.a)
            for _ in range(int(e)):
                out = out * self.a
            return M(out)
        return M(self.a ** e)

    def __matmul__(self, o):
        o = _lift(o, self.a.shape[2])
        return M(np.einsum('ikb,kjb->ijb', self.a, o.a))

    def __rmatmul__(self, o):
        o = _lift(o, self.a.shape[2])
        return M(np.einsum('ikb,kjb->ijb', o.a, self.a))

    # comparisons (1x1 only) -> boolean array over the batch axis
    def _cmp(self, o, f):
        o = _lift(o, self.a.shape[2])
        assert self.a.shape[:2] == (1, 1) and o.a.shape[:2] == (1, 1)
        return f(self.a.real[0, 0], o.a.real[0, 0])

    def __lt__(self, o): return self._cmp(o, np.less)
    def __le__(self, o): return self._cmp(o, np.less_equal)
    def __gt__(self, o): return self._cmp(o, np.greater)
    def __ge__(self, o): return self._cmp(o, np.greater_equal)

    def __float__(self):
        assert self.a.size == 1
        return float(self.a.real.ravel()[0])

    def __deepcopy__(self, memo):
        return M(self.a.copy())

    def __repr__(self):
        return f"M{self.a.shape}"


def _lift(v, nb=1):
    if isinstance(v, M):
        return v
    a = np.asarray(v)
    if a.ndim == 0:
        return M(a.reshape(1, 1, 1))
    if a.ndim == 1:
        return M(a.reshape(-1, 1, 1))
    if a.ndim == 2:
        return M(a.reshape(a.shape[0], a.shape[1], 1))
    raise ValueError(a.shape)


class _SXMeta(type):
    pass


class SX(metaclass=_SXMeta):
    @staticmethod
    def sym(name, *shape):
        if len(shape) == 0:
            shape = (1, 1)
        elif len(shape) == 1:
            if isinstance(shape[0], tuple):
                shape = shape[0]
            else:
                shape = (shape[0], 1)
        return _state["provider"].sym(name, shape)

    @staticmethod
    def zeros(*shape):
        if len(shape) == 1:
            shape = shape[0] if isinstance(shape[0], tuple) else (shape[0], 1)
        return M(np.zeros((shape[0], shape[1], 1)))

    @staticmethod
    def eye(n):
        return M(np.eye(n).reshape(n, n, 1))


MX = SX


class DM:
    pass


def _un(f):
    def g(x):
        if isinstance(x, M):
            return M(f(x.a))
        return f(x)
    return g


sin = _un(np.sin)
cos = _un(np.cos)
exp = _un(np.exp)
sqrt = _un(np.sqrt)
acos = _un(np.arccos)


def dot(a, b):
    if not isinstance(a, M) and not isinstance(b, M):
        return np.dot(np.asarray(a).ravel(), np.asarray(b).ravel())
    a = _lift(a)
    b = _lift(b)
    return M(np.sum(a._lin() * b._lin(), axis=0).reshape(1, 1, -1))


def sumsqr(a):
    if not isinstance(a, M):
        return np.sum(np.asarray(a) ** 2)
    return M(np.sum(a.a * a.a, axis=(0, 1)).reshape(1, 1, -1))


def norm_2(a):
    a = _lift(a)
    return M(np.sqrt(np.sum(a.a * a.a, axis=(0, 1))).reshape(1, 1, -1))


def if_else(c, a, b):
    """Branch selection; `c` is a boolean array over the batch axis."""
    if not isinstance(a, M) and not isinstance(b, M):
        return a if bool(np.all(c)) else b
    a = _lift(a)
    b = _lift(b)
    nb = max(a.a.shape[2], b.a.shape[2], np.size(c))
    c = np.broadcast_to(np.asarray(c).reshape(-1), (nb,))
    aa = np.broadcast_to(a.a, a.a.shape[:2] + (nb,))
    bb = np.broadcast_to(b.a, b.a.shape[:2] + (nb,))
    return M(np.where(c[None, None, :], aa, bb))


def vertcat(*xs):
    if len(xs) == 0:
        return M(np.zeros((0, 1, 1)))
    ms = [_lift(x) for x in xs]
    nb = max(m.a.shape[2] for m in ms)
    dt = complex if any(m.a.dtype.kind == 'c' for m in ms) else float
    arrs = []
    for m in ms:
        a = m.a
        if a.shape[2] != nb:
            a = np.repeat(a, nb, axis=2)
        arrs.append(a.astype(dt))
    return M(np.concatenate(arrs, axis=0))


class _Solver:
    """Captured NLP.  `backend(x0, lbx, ubx, lbg, ubg, p) -> (dict, stats)` may be
    installed by tests (see set_solver_backend)."""

    def __init__(self, name, plugin, prob, opts):
        self.name, self.plugin, self.prob, self.opts = name, plugin, prob, opts
        self._stats = {"iter_count": 0, "success": False, "return_status": "unset"}
        self.calls = []

    def generate_dependencies(self, *a, **k):
        return None

    def stats(self):
        return self._stats

    def __call__(self, **kw):
        self.calls.append(kw)
        be = _state.get("backend")
        if be is None:
            raise RuntimeError("refexec.casadi: no solver backend installed")
        sol, st = be(**kw)
        self._stats = st
        return sol


def set_solver_backend(fn):
    _state["backend"] = fn


def nlpsol(name, plugin, prob, opts=None):
    s = _Solver(name, plugin, prob, opts)
    _state["last_solver"] = s
    return s


def last_solver():
    return _state.get("last_solver")
