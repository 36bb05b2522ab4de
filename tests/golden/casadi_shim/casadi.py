"""Minimal stand-in for the `casadi` Python module -- TEST INFRASTRUCTURE ONLY.

Why it exists.  The reference's hot path (biped_pympc/casadi/srbd_centroidal_model.py,
srbd_constraints.py, sparse_pdipm_solver.py) is symbolic Python over CasADi `SX`, and CasADi
(pyproject.toml:17, unpinned; CusADi was built on 3.6.5) is neither installed nor installable in
this container.  This module implements just enough of CasADi's public API -- with the published
semantics of each call, nothing else -- for the reference's OWN, UNMODIFIED source files to be
imported from /root/reference and executed: they build their expression graphs here exactly as they
would under CasADi, and `Function.__call__` evaluates those graphs in float64.  The golden vectors
of tests/golden/make_golden_ref.py come out of that; the oracle restatement and the CUDA kernels are
then checked against them.  Nothing in the product imports this file.

What is restated (third-party semantics, not reference code):
  SX            sparse matrix of scalar expressions; structural zeros are tracked (entry None)
  SX.sym/zeros/ones, SX_eye, vertcat, horzcat, diagcat, diag, reshape (column-major), vec, skew, dot,
  norm_2, mmin, fmin, fmax, if_else, inv (adjugate; CasADi solves A X = I -- same value to round-off)
  jacobian / hessian   reverse-mode symbolic differentiation with the usual simplifications
                       (0*x = 0, 1*x = x, x+0 = x), which is what produces the output sparsity
  ldl / ldl_solve      LDL' WITHOUT pivoting of the symmetric matrix, A[p,p] = (I+LT)' D (I+LT);
                       CasADi pre-orders with AMD (amd=True is its default); AMD itself is not restated:
                       p is a greedy minimum-degree order (fill-reducing like AMD; 390 k instructions for
                       the 5-iteration solver).  CASADI_SHIM_ORDER=natural keeps [dx,ds,dz,dy] order
                       (1.0 M instructions) -- the ordering changes round-off only; the two fixtures differ
                       by up to 1e-9 (x) / 4e-7 (z, BASELINE config 1) after 5 iterations, and the 80-bit
                       oracle sides with the fill-reducing order
  Function      symbolic call (substitution) and numeric call (float64 evaluation); DM results with
                nonzeros() / sparsity().get_ccs() in CCS order
Element order, shapes and CCS conventions follow CasADi (column-major).
"""
import math
import os

import numpy as np

# ---------------------------------------------------------------------------------------------
# instruction op codes (names as in casadi/core/calculus.hpp; the numeric values are the shim's own --
# callers only ever compare them with these constants, e.g. cusadi/src/CusadiOperations.py:11-111)
# ---------------------------------------------------------------------------------------------
(OP_ASSIGN, OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_NEG, OP_EXP, OP_LOG, OP_POW, OP_CONSTPOW, OP_SQRT, OP_SQ, OP_TWICE, OP_SIN, OP_COS,
 OP_TAN, OP_ASIN, OP_ACOS, OP_ATAN, OP_LT, OP_LE, OP_EQ, OP_NE, OP_NOT, OP_AND, OP_OR, OP_FLOOR, OP_CEIL, OP_FMOD, OP_FABS, OP_SIGN,
 OP_COPYSIGN, OP_IF_ELSE_ZERO, OP_ERF, OP_FMIN, OP_FMAX, OP_INV, OP_SINH, OP_COSH, OP_TANH, OP_ASINH, OP_ACOSH, OP_ATANH, OP_ATAN2,
 OP_CONST, OP_INPUT, OP_OUTPUT, OP_PARAMETER, OP_CALL) = range(49)

# ---------------------------------------------------------------------------------------------
# scalar expression nodes
# ---------------------------------------------------------------------------------------------
_C, _S, _ADD, _SUB, _MUL, _DIV, _NEG, _POW, _LT, _IFE, _FMIN, _FMAX, _SQRT, _NOT = range(14)
_counter = [0]


class Node(object):
    __slots__ = ("op", "a", "b", "c", "id")

    def __init__(self, op, a=None, b=None, c=None):
        self.op = op
        self.a = a
        self.b = b
        self.c = c
        _counter[0] += 1
        self.id = _counter[0]


def const(v):
    return Node(_C, float(v))


ZERO = const(0.0)
ONE = const(1.0)


def _isc(n):
    return n.op == _C


def n_add(a, b):
    if a.op == _C:
        if b.op == _C:
            return const(a.a + b.a)
        if a.a == 0.0:
            return b
    elif b.op == _C and b.a == 0.0:
        return a
    return Node(_ADD, a, b)


def n_sub(a, b):
    if b.op == _C:
        if a.op == _C:
            return const(a.a - b.a)
        if b.a == 0.0:
            return a
    elif a.op == _C and a.a == 0.0:
        return n_neg(b)
    if a is b:
        return ZERO
    return Node(_SUB, a, b)


def n_mul(a, b):
    if a.op == _C:
        if b.op == _C:
            return const(a.a * b.a)
        if a.a == 0.0:
            return ZERO
        if a.a == 1.0:
            return b
        if a.a == -1.0:
            return n_neg(b)
    elif b.op == _C:
        if b.a == 0.0:
            return ZERO
        if b.a == 1.0:
            return a
        if b.a == -1.0:
            return n_neg(a)
    return Node(_MUL, a, b)


def n_div(a, b):
    if a.op == _C:
        if b.op == _C:
            return const(a.a / b.a)
        if a.a == 0.0:
            return ZERO
    elif b.op == _C and b.a == 1.0:
        return a
    return Node(_DIV, a, b)


def n_neg(a):
    if a.op == _C:
        return const(-a.a)
    if a.op == _NEG:
        return a.a
    return Node(_NEG, a)


def n_pow(a, b):
    if a.op == _C and b.op == _C:
        return const(math.pow(a.a, b.a))
    return Node(_POW, a, b)


def n_lt(a, b):
    if a.op == _C and b.op == _C:
        return const(1.0 if a.a < b.a else 0.0)
    return Node(_LT, a, b)


def n_ife(c, x, y):
    if c.op == _C:
        return x if c.a != 0.0 else y
    return Node(_IFE, c, x, y)


def n_fmin(a, b):
    if a.op == _C and b.op == _C:
        return const(min(a.a, b.a))
    return Node(_FMIN, a, b)


def n_fmax(a, b):
    if a.op == _C and b.op == _C:
        return const(max(a.a, b.a))
    return Node(_FMAX, a, b)


def n_sqrt(a):
    if a.op == _C:
        return const(math.sqrt(a.a))
    return Node(_SQRT, a)


def _nz(n):
    """structural-zero filter: a constant 0 in a matrix entry is dropped from the pattern"""
    if n is None or (n.op == _C and n.a == 0.0):
        return None
    return n


# ---------------------------------------------------------------------------------------------
# SX: matrix of nodes (object ndarray, None = structural zero)
# ---------------------------------------------------------------------------------------------
def _as_sx(v):
    if isinstance(v, SX):
        return v
    if isinstance(v, DM):
        return SX._from_dense(v.full())
    if isinstance(v, (int, float, np.floating, np.integer)):
        out = SX(1, 1)
        out.arr[0, 0] = _nz(const(v))
        return out
    return SX._from_dense(np.asarray(v, dtype=np.float64))


class SX(object):
    __array_priority__ = 1000
    __array_ufunc__ = None      # numpy scalars defer to __rmul__ / __radd__ ...

    def __init__(self, *args):
        if len(args) == 2:
            self.arr = np.full((int(args[0]), int(args[1])), None, dtype=object)
        elif len(args) == 1:
            if isinstance(args[0], SX):
                self.arr = args[0].arr.copy()
            else:
                self.arr = SX._from_dense(np.asarray(args[0], dtype=np.float64)).arr
        elif len(args) == 0:
            self.arr = np.full((0, 0), None, dtype=object)
        else:
            raise TypeError("SX(): unsupported arguments")

    # -- constructors ------------------------------------------------------------------------
    @staticmethod
    def _from_dense(a):
        a = np.asarray(a, dtype=np.float64)
        if a.ndim == 0:
            a = a.reshape(1, 1)
        elif a.ndim == 1:
            a = a.reshape(-1, 1)
        out = SX(a.shape[0], a.shape[1])
        for i in range(a.shape[0]):
            for j in range(a.shape[1]):
                out.arr[i, j] = const(a[i, j])      # dense: numeric zeros stay in the pattern
        return out

    @staticmethod
    def sym(name, n=1, m=1):
        out = SX(n, m)
        for j in range(m):          # column-major numbering like CasADi's name_k
            for i in range(n):
                out.arr[i, j] = Node(_S, "%s_%d" % (name, j * n + i))
        return out

    @staticmethod
    def zeros(n=1, m=1):
        return SX(n, m)

    @staticmethod
    def ones(n=1, m=1):
        out = SX(n, m)
        for i in range(n):
            for j in range(m):
                out.arr[i, j] = ONE
        return out

    @staticmethod
    def eye(n):
        out = SX(n, n)
        for i in range(n):
            out.arr[i, i] = ONE
        return out

    # -- shape -------------------------------------------------------------------------------
    @property
    def shape(self):
        return self.arr.shape

    def size(self):
        return self.arr.shape

    def size1(self):
        return self.arr.shape[0]

    def size2(self):
        return self.arr.shape[1]

    def numel(self):
        return self.arr.size

    @property
    def T(self):
        out = SX(0, 0)
        out.arr = self.arr.T.copy()
        return out

    def reshape(self, shp):
        return reshape(self, shp)

    def nnz(self):
        return int(np.count_nonzero(self.arr != None))  # noqa: E711

    def nonzeros(self):
        r, c = self.arr.shape
        return [self.arr[i, j] for j in range(c) for i in range(r) if self.arr[i, j] is not None]

    def sparsity(self):
        return Sparsity(self.arr != None)  # noqa: E711

    def _scalar(self):
        assert self.arr.shape == (1, 1), "expected a scalar SX, got %s" % (self.arr.shape,)
        n = self.arr[0, 0]
        return ZERO if n is None else n

    # -- indexing ----------------------------------------------------------------------------
    def _key(self, key):
        if not isinstance(key, tuple):
            if self.arr.shape[1] == 1:
                key = (key, 0)
            elif self.arr.shape[0] == 1:
                key = (0, key)
            else:
                raise IndexError("linear indexing of a matrix SX is not supported by the shim")
        k = []
        for q in key:
            if isinstance(q, (int, np.integer)):
                q = int(q)
                k.append(slice(q, q + 1) if q != -1 else slice(q, None))
            else:
                k.append(q)
        return tuple(k)

    def __getitem__(self, key):
        out = SX(0, 0)
        out.arr = self.arr[self._key(key)].copy()
        return out

    def __setitem__(self, key, value):
        v = _as_sx(value)
        tgt = self.arr[self._key(key)]
        if v.arr.shape == (1, 1) and tgt.shape != (1, 1):
            tgt[...] = v.arr[0, 0]
        else:
            assert tgt.shape == v.arr.shape, (tgt.shape, v.arr.shape)
            tgt[...] = v.arr

    # -- elementwise arithmetic ----------------------------------------------------------------
    def _binary(self, other, fn, zero_if_any_zero, swap=False):
        a, b = _as_sx(self), _as_sx(other)
        if swap:
            a, b = b, a
        A, Bm = a.arr, b.arr
        if A.shape != Bm.shape:
            if A.shape == (1, 1):
                A = np.broadcast_to(A, Bm.shape)
            elif Bm.shape == (1, 1):
                Bm = np.broadcast_to(Bm, A.shape)
            else:
                raise ValueError("shape mismatch %s vs %s" % (A.shape, Bm.shape))
        out = SX(A.shape[0], A.shape[1])
        ma, mb = (A != None), (Bm != None)  # noqa: E711
        mask = (ma & mb) if zero_if_any_zero else (ma | mb)
        for i, j in zip(*np.nonzero(mask)):
            x = A[i, j]
            y = Bm[i, j]
            # CasADi: the result pattern is the union / intersection of the operand patterns; an entry that
            # simplifies to the constant 0 (u - u) stays in the pattern as a numeric zero
            out.arr[i, j] = fn(ZERO if x is None else x, ZERO if y is None else y)
        return out

    def __add__(self, o):
        return self._binary(o, n_add, False)

    def __radd__(self, o):
        return self._binary(o, n_add, False, swap=True)

    def __sub__(self, o):
        return self._binary(o, n_sub, False)

    def __rsub__(self, o):
        return self._binary(o, n_sub, False, swap=True)

    def __mul__(self, o):
        return self._binary(o, n_mul, True)

    def __rmul__(self, o):
        return self._binary(o, n_mul, True, swap=True)

    def _dense_binary(self, other, fn, swap=False):
        a, b = _as_sx(self), _as_sx(other)
        if swap:
            a, b = b, a
        A, Bm = a.arr, b.arr
        if A.shape != Bm.shape:
            if A.shape == (1, 1):
                A = np.broadcast_to(A, Bm.shape)
            else:
                Bm = np.broadcast_to(Bm, A.shape)
        out = SX(A.shape[0], A.shape[1])
        for i in range(A.shape[0]):
            for j in range(A.shape[1]):
                x, y = A[i, j], Bm[i, j]
                if x is None and fn is n_div:
                    continue            # 0 / y = structural zero
                out.arr[i, j] = _nz(fn(ZERO if x is None else x, ZERO if y is None else y))
        return out

    def __truediv__(self, o):
        return self._dense_binary(o, n_div)

    def __rtruediv__(self, o):
        return self._dense_binary(o, n_div, swap=True)

    def __pow__(self, o):
        return self._dense_binary(o, n_pow)

    def __lt__(self, o):
        return self._dense_binary(o, n_lt)

    def __gt__(self, o):
        return _as_sx(o)._dense_binary(self, n_lt)

    def __neg__(self):
        out = SX(self.arr.shape[0], self.arr.shape[1])
        for i, j in zip(*np.nonzero(self.arr != None)):  # noqa: E711
            out.arr[i, j] = n_neg(self.arr[i, j])
        return out

    # -- matrix product (sparse aware; sums run over the inner index in ascending order) -------
    def __matmul__(self, o):
        a, b = self, _as_sx(o)
        assert a.arr.shape[1] == b.arr.shape[0], (a.arr.shape, b.arr.shape)
        out = SX(a.arr.shape[0], b.arr.shape[1])
        brows = [[(j, b.arr[k, j]) for j in np.nonzero(b.arr[k] != None)[0]]  # noqa: E711
                 for k in range(b.arr.shape[0])]
        for i in range(a.arr.shape[0]):
            acc = {}
            for k in np.nonzero(a.arr[i] != None)[0]:  # noqa: E711
                aik = a.arr[i, k]
                for j, bkj in brows[k]:
                    t = n_mul(aik, bkj)
                    acc[j] = t if j not in acc else n_add(acc[j], t)
            for j, v in acc.items():
                out.arr[i, j] = v           # structural product pattern
        return out

    def __rmatmul__(self, o):
        return _as_sx(o).__matmul__(self)

    def __repr__(self):
        return "SX(%dx%d, %d nnz)" % (self.arr.shape[0], self.arr.shape[1], self.nnz())


class Sparsity(object):
    def __init__(self, mask):
        self.mask = np.asarray(mask, dtype=bool)

    def get_ccs(self):
        r, c = self.mask.shape
        colind, row = [0], []
        for j in range(c):
            rows = np.nonzero(self.mask[:, j])[0]
            row.extend(int(i) for i in rows)
            colind.append(len(row))
        return colind, row

    @staticmethod
    def diag(n):
        return Sparsity(np.eye(n, dtype=bool))

    def nnz(self):
        return int(self.mask.sum())


class DM(object):
    """numeric result of a Function call: values on the structural pattern of the output"""

    def __init__(self, dense, mask):
        self._dense = np.asarray(dense, dtype=np.float64)
        self._mask = np.asarray(mask, dtype=bool)

    @property
    def shape(self):
        return self._dense.shape

    def full(self):
        return self._dense.copy()

    def __array__(self, dtype=None, copy=None):
        return self._dense.astype(dtype) if dtype is not None else self._dense.copy()

    def nonzeros(self):
        r, c = self._dense.shape
        return [float(self._dense[i, j]) for j in range(c) for i in range(r) if self._mask[i, j]]

    def sparsity(self):
        return Sparsity(self._mask)

    def nnz(self):
        return int(self._mask.sum())

    def __float__(self):
        assert self._dense.size == 1
        return float(self._dense.reshape(-1)[0])


# ---------------------------------------------------------------------------------------------
# free functions of the casadi namespace
# ---------------------------------------------------------------------------------------------
def SX_eye(n):
    return SX.eye(n)


def vertcat(*args):
    mats = [_as_sx(a) for a in args]
    mats = [m for m in mats if m.arr.size or m.arr.shape[1]]
    if not mats:
        return SX(0, 1)
    ncol = max(m.arr.shape[1] for m in mats)
    mats = [m for m in mats if m.arr.shape[0] > 0]
    if not mats:
        return SX(0, ncol)
    out = SX(0, 0)
    out.arr = np.concatenate([m.arr for m in mats], axis=0)
    return out


def horzcat(*args):
    mats = [_as_sx(a) for a in args]
    out = SX(0, 0)
    out.arr = np.concatenate([m.arr for m in mats], axis=1)
    return out


def diagcat(*args):
    mats = [_as_sx(a) for a in args]
    r = sum(m.arr.shape[0] for m in mats)
    c = sum(m.arr.shape[1] for m in mats)
    out = SX(r, c)
    i = j = 0
    for m in mats:
        out.arr[i:i + m.arr.shape[0], j:j + m.arr.shape[1]] = m.arr
        i += m.arr.shape[0]
        j += m.arr.shape[1]
    return out


def diag(v):
    v = _as_sx(v)
    if 1 in v.arr.shape:                    # vector -> diagonal matrix
        flat = v.arr.reshape(-1, order="F")
        out = SX(flat.size, flat.size)
        for i, n in enumerate(flat):
            out.arr[i, i] = n
        return out
    out = SX(v.arr.shape[0], 1)             # matrix -> its diagonal
    for i in range(v.arr.shape[0]):
        out.arr[i, 0] = v.arr[i, i]
    return out


def reshape(m, *shp):
    if len(shp) == 1:
        shp = tuple(shp[0])
    m = _as_sx(m)
    r, c = int(shp[0]), int(shp[1])
    if r == -1:
        r = m.arr.size // c
    if c == -1:
        c = m.arr.size // r
    out = SX(0, 0)
    out.arr = m.arr.reshape((r, c), order="F").copy()   # CasADi reshapes column-major
    return out


def vec(m):
    m = _as_sx(m)
    return reshape(m, (m.arr.size, 1))


def skew(v):
    v = _as_sx(v)
    a = [v.arr.reshape(-1, order="F")[k] for k in range(3)]
    neg = lambda n: None if n is None else n_neg(n)  # noqa: E731
    out = SX(3, 3)
    out.arr[0, 1], out.arr[0, 2] = neg(a[2]), a[1]
    out.arr[1, 0], out.arr[1, 2] = a[2], neg(a[0])
    out.arr[2, 0], out.arr[2, 1] = neg(a[1]), a[0]
    return out


def dot(a, b):
    a, b = _as_sx(a), _as_sx(b)
    fa, fb = a.arr.reshape(-1, order="F"), b.arr.reshape(-1, order="F")
    acc = None
    for x, y in zip(fa, fb):
        if x is None or y is None:
            continue
        t = n_mul(x, y)
        acc = t if acc is None else n_add(acc, t)
    out = SX(1, 1)
    out.arr[0, 0] = _nz(acc)
    return out


def norm_2(v):
    out = SX(1, 1)
    out.arr[0, 0] = _nz(n_sqrt(dot(v, v)._scalar()))
    return out


def mmin(v):
    v = _as_sx(v)
    flat = [ZERO if n is None else n for n in v.arr.reshape(-1, order="F")]
    acc = flat[0]
    for n in flat[1:]:
        acc = n_fmin(acc, n)
    out = SX(1, 1)
    out.arr[0, 0] = _nz(acc)
    return out


def _ew2(fn, a, b):
    return _as_sx(a)._dense_binary(b, fn)


def fmin(a, b):
    return _ew2(n_fmin, a, b)


def fmax(a, b):
    return _ew2(n_fmax, a, b)


def if_else(c, x, y, short_circuit=False):
    out = SX(1, 1)
    out.arr[0, 0] = _nz(n_ife(_as_sx(c)._scalar(), _as_sx(x)._scalar(), _as_sx(y)._scalar()))
    return out


def inv(m):
    """3x3 (or 2x2 / 1x1) inverse via the adjugate"""
    m = _as_sx(m)
    n = m.arr.shape[0]
    g = lambda i, j: ZERO if m.arr[i, j] is None else m.arr[i, j]  # noqa: E731
    out = SX(n, n)
    if n == 1:
        out.arr[0, 0] = n_div(ONE, g(0, 0))
        return out
    if n == 2:
        det = n_sub(n_mul(g(0, 0), g(1, 1)), n_mul(g(0, 1), g(1, 0)))
        out.arr[0, 0], out.arr[0, 1] = n_div(g(1, 1), det), n_div(n_neg(g(0, 1)), det)
        out.arr[1, 0], out.arr[1, 1] = n_div(n_neg(g(1, 0)), det), n_div(g(0, 0), det)
        return out
    assert n == 3, "inv: the shim handles up to 3x3"
    cof = [[None] * 3 for _ in range(3)]
    for i in range(3):
        for j in range(3):
            r = [k for k in range(3) if k != i]
            c = [k for k in range(3) if k != j]
            minor = n_sub(n_mul(g(r[0], c[0]), g(r[1], c[1])), n_mul(g(r[0], c[1]), g(r[1], c[0])))
            cof[i][j] = minor if (i + j) % 2 == 0 else n_neg(minor)
    det = n_add(n_add(n_mul(g(0, 0), cof[0][0]), n_mul(g(0, 1), cof[0][1])), n_mul(g(0, 2), cof[0][2]))
    for i in range(3):
        for j in range(3):
            out.arr[i, j] = _nz(n_div(cof[j][i], det))
    return out


# ---------------------------------------------------------------------------------------------
# differentiation (reverse mode, symbolic)
# ---------------------------------------------------------------------------------------------
def _reachable(roots):
    seen = {}
    stack = [r for r in roots if r is not None]
    while stack:
        n = stack.pop()
        if n.id in seen:
            continue
        seen[n.id] = n
        if n.op not in (_C, _S):
            for ch in (n.a, n.b, n.c):
                if ch is not None and ch.id not in seen:
                    stack.append(ch)
    return seen


def _gradient_of(root, wrt_ids):
    """adjoints of `root` w.r.t. the symbols in wrt_ids -> {symbol id: Node}"""
    nodes = _reachable([root])
    adj = {root.id: ONE}
    res = {}

    def acc(n, g):
        if g.op == _C and g.a == 0.0:
            return
        adj[n.id] = g if n.id not in adj else n_add(adj[n.id], g)

    for nid in sorted(nodes, reverse=True):
        g = adj.pop(nid, None)
        if g is None:
            continue
        n = nodes[nid]
        op = n.op
        if op == _S:
            if nid in wrt_ids:
                res[nid] = g
        elif op == _C:
            pass
        elif op == _ADD:
            acc(n.a, g)
            acc(n.b, g)
        elif op == _SUB:
            acc(n.a, g)
            acc(n.b, n_neg(g))
        elif op == _MUL:
            acc(n.a, n_mul(g, n.b))
            acc(n.b, n_mul(g, n.a))
        elif op == _DIV:
            acc(n.a, n_div(g, n.b))
            acc(n.b, n_neg(n_div(n_mul(g, n), n.b)))     # d(a/b)/db = -(a/b)/b
        elif op == _NEG:
            acc(n.a, n_neg(g))
        else:
            raise NotImplementedError("derivative of op %d is not needed by the reference path" % op)
    return res


def jacobian(expr, wrt):
    expr, wrt = _as_sx(expr), _as_sx(wrt)
    ev = expr.arr.reshape(-1, order="F")
    wv = wrt.arr.reshape(-1, order="F")
    wid = {n.id: k for k, n in enumerate(wv)}
    out = SX(ev.size, wv.size)
    for i, e in enumerate(ev):
        if e is None:
            continue
        for nid, g in _gradient_of(e, wid).items():
            out.arr[i, wid[nid]] = _nz(g)
    return out


def gradient(expr, wrt):
    return jacobian(expr, wrt).T


def hessian(expr, wrt):
    g = gradient(expr, wrt)
    return jacobian(g, wrt), g


# ---------------------------------------------------------------------------------------------
# LDL' without pivoting
# ---------------------------------------------------------------------------------------------
def _order(mask):
    n = mask.shape[0]
    if os.environ.get("CASADI_SHIM_ORDER", "mindeg") == "natural":
        return list(range(n))
    adj = [set(np.nonzero(mask[i] | mask[:, i])[0]) - {i} for i in range(n)]
    alive = set(range(n))
    p = []
    while alive:
        k = min(alive, key=lambda q: (len(adj[q]), q))
        nb = adj[k]
        for a in nb:
            adj[a] |= nb
            adj[a] -= {a, k}
        alive.remove(k)
        p.append(k)
    return p


def ldl(A, amd=True):
    """returns D (n x 1), LT (strictly upper triangular) and p with A[p,p] = (I+LT)' diag(D) (I+LT)"""
    A = _as_sx(A)
    n = A.arr.shape[0]
    p = _order(A.arr != None)  # noqa: E711
    # rows[i] = {j: node} for the lower triangle (j < i) and the diagonal of the permuted matrix
    low = [dict() for _ in range(n)]
    for i in range(n):
        for j in range(i + 1):
            e = A.arr[p[i], p[j]]
            if e is None:
                e = A.arr[p[j], p[i]]
            if e is not None:
                low[i][j] = e
    cols = [dict() for _ in range(n)]       # cols[k] = {i: L_ik} for i > k
    D = [None] * n
    for k in range(n):
        # L_kj (j < k) are final; D_k = A_kk - sum_j L_kj^2 D_j ; done lazily through `low`
        D[k] = low[k].get(k, ZERO)
        # column k below the diagonal: entries i > k with low[i][k]
        for i in range(k + 1, n):
            e = low[i].get(k)
            if e is None:
                continue
            cols[k][i] = n_div(e, D[k])
        # right-looking update of the trailing matrix: A_ij -= L_ik D_k L_jk  (= L_ik * low[j][k])
        idx = sorted(cols[k])
        for a_pos, i in enumerate(idx):
            lik = cols[k][i]
            for j in idx[:a_pos + 1]:
                t = n_mul(lik, low[j][k])
                cur = low[i].get(j)
                low[i][j] = n_neg(t) if cur is None else n_sub(cur, t)
    Dm = SX(n, 1)
    LT = SX(n, n)
    for k in range(n):
        Dm.arr[k, 0] = D[k]
        for i, v in cols[k].items():
            LT.arr[k, i] = v
    return Dm, LT, p


def ldl_solve(b, D, LT, p):
    b, D, LT = _as_sx(b), _as_sx(D), _as_sx(LT)
    n = D.arr.shape[0]
    assert b.arr.shape[1] == 1
    w = [b.arr[p[i], 0] for i in range(n)]
    rows = [[(int(i), LT.arr[k, i]) for i in np.nonzero(LT.arr[k] != None)[0]] for k in range(n)]  # noqa: E711
    # forward: (I+LT)' w = b   (column sweep)
    for k in range(n):
        if w[k] is None:
            continue
        for i, l in rows[k]:
            t = n_mul(l, w[k])
            w[i] = n_neg(t) if w[i] is None else n_sub(w[i], t)
    for k in range(n):
        if w[k] is not None:
            w[k] = n_div(w[k], D.arr[k, 0])
    # backward: (I+LT) x = w
    for k in range(n - 1, -1, -1):
        acc = w[k]
        for i, l in rows[k]:
            if w[i] is None:
                continue
            t = n_mul(l, w[i])
            acc = n_neg(t) if acc is None else n_sub(acc, t)
        w[k] = acc
    out = SX(n, 1)
    for i in range(n):
        out.arr[p[i], 0] = _nz(w[i]) if w[i] is not None else None
    return out


# ---------------------------------------------------------------------------------------------
# Function
# ---------------------------------------------------------------------------------------------
class Function(object):
    def __init__(self, name, ins, outs, *unused):
        self._name = name
        self._ins = [_as_sx(i) for i in ins]
        self._outs = [_as_sx(o) for o in outs]
        for m in self._ins:
            for n in m.arr.reshape(-1):
                assert n is not None and n.op == _S, "Function inputs must be purely symbolic"
        self._tape = None
        self._instr = None

    def name(self):
        return self._name

    def n_in(self):
        return len(self._ins)

    def n_out(self):
        return len(self._outs)

    def nnz_in(self, i):
        return self._ins[i].nnz()

    def nnz_out(self, i):
        return self._outs[i].nnz()

    def sparsity_out(self, i):
        return self._outs[i].sparsity()

    def _build_tape(self):
        roots = [n for o in self._outs for n in o.arr.reshape(-1) if n is not None]
        nodes = _reachable(roots)
        order = sorted(nodes)
        pos = {nid: k for k, nid in enumerate(order)}
        tape = []
        for nid in order:
            n = nodes[nid]
            if n.op == _C:
                tape.append((_C, n.a, 0, 0))
            elif n.op == _S:
                tape.append((_S, nid, 0, 0))
            else:
                tape.append((n.op, pos[n.a.id], pos[n.b.id] if n.b is not None else 0,
                             pos[n.c.id] if n.c is not None else 0))
        self._tape = tape
        self._pos = pos
        self._nodes = nodes
        self._order = order

    # -- instruction-level view (SXFunction's algorithm): what code generators walk -------------
    def _build_instructions(self):
        """Linearises the tape into CasADi-style instructions over a work vector whose slots are reused once a
        value is dead (CasADi allocates its work vector the same way; sz_w is the number of simultaneously live
        values, not the number of instructions).  if_else(c, x, y) is emitted the way CasADi defines it,
        if_else_zero(c, x) + if_else_zero(!c, y)."""
        if self._tape is None:
            self._build_tape()
        tape, n = self._tape, len(self._tape)
        sym_pos = {}
        for i, m in enumerate(self._ins):
            for e, node in enumerate(m.arr.reshape(-1, order="F")):
                sym_pos[node.id] = (i, e)
        out_use = []                                   # (tape position, output index, nonzero index)
        for o, m in enumerate(self._outs):
            e = 0
            r, c = m.arr.shape
            for j in range(c):
                for i in range(r):
                    if m.arr[i, j] is not None:
                        out_use.append((self._pos[m.arr[i, j].id], o, e))
                        e += 1
        last = list(range(n))                          # last instruction position that reads each value
        for k, (op, a, b, c) in enumerate(tape):
            if op not in (_C, _S):
                last[a] = k
                if op not in (_NEG, _SQRT):
                    last[b] = k
                if op == _IFE:
                    last[c] = k
        for pos, _, _ in out_use:
            last[pos] = n                              # outputs are written at the end: keep alive
        binop = {_ADD: OP_ADD, _SUB: OP_SUB, _MUL: OP_MUL, _DIV: OP_DIV, _POW: OP_POW, _LT: OP_LT, _FMIN: OP_FMIN, _FMAX: OP_FMAX}
        free, nslot, slot = [], 0, [0] * n
        ids, ins, outs, consts = [], [], [], []
        dying = {}
        for k in range(n):
            dying.setdefault(last[k], []).append(k)

        def alloc():
            nonlocal nslot
            if free:
                return free.pop()
            nslot += 1
            return nslot - 1

        def emit(op, i, o, cval=0.0):
            ids.append(op); ins.append(i); outs.append(o); consts.append(cval)

        for k, (op, a, b, c) in enumerate(tape):
            if op == _C:
                slot[k] = alloc(); emit(OP_CONST, [], [slot[k]], a)
            elif op == _S:
                slot[k] = alloc(); emit(OP_INPUT, list(sym_pos[a]), [slot[k]])
            elif op == _IFE:
                t1, t2 = alloc(), alloc()
                emit(OP_IF_ELSE_ZERO, [slot[a], slot[b]], [t1])
                emit(OP_NOT, [slot[a]], [t2])
                emit(OP_IF_ELSE_ZERO, [t2, slot[c]], [t2])
                slot[k] = alloc()
                emit(OP_ADD, [t1, t2], [slot[k]])
                free.extend((t1, t2))
            elif op == _NEG:
                slot[k] = alloc(); emit(OP_NEG, [slot[a]], [slot[k]])
            elif op == _SQRT:
                slot[k] = alloc(); emit(OP_SQRT, [slot[a]], [slot[k]])
            else:
                slot[k] = alloc(); emit(binop[op], [slot[a], slot[b]], [slot[k]])
            for d in dying.get(k, ()):                 # values whose last reader was this instruction
                if d != k or last[k] == k:
                    free.append(slot[d])
        for pos, o, e in out_use:
            emit(OP_OUTPUT, [slot[pos]], [o, e])
        self._instr = (ids, ins, outs, consts, nslot)

    def _instructions(self):
        if getattr(self, "_instr", None) is None:
            self._build_instructions()
        return self._instr

    def n_instructions(self):
        return len(self._instructions()[0])

    def instruction_id(self, k):
        return self._instructions()[0][k]

    def instruction_input(self, k):
        return self._instructions()[1][k]

    def instruction_output(self, k):
        return self._instructions()[2][k]

    def instruction_constant(self, k):
        return self._instructions()[3][k]

    def sz_w(self):
        return self._instructions()[4]

    def size1_out(self, i):
        return self._outs[i].arr.shape[0]

    def size2_out(self, i):
        return self._outs[i].arr.shape[1]

    # -- numeric ---------------------------------------------------------------------------
    def _call_numeric(self, args):
        if self._tape is None:
            self._build_tape()
        sym_val = {}
        for m, a in zip(self._ins, args):
            a = np.asarray(a.full() if isinstance(a, DM) else a, dtype=np.float64)
            flat = a.reshape(-1, order="F")
            assert flat.size == m.arr.size, "%s: argument of %d elements for an input of %d" % (
                self._name, flat.size, m.arr.size)
            for n, v in zip(m.arr.reshape(-1, order="F"), flat):
                sym_val[n.id] = float(v)
        w = [0.0] * len(self._tape)
        inf = float("inf")
        for k, (op, a, b, c) in enumerate(self._tape):
            if op == _MUL:
                w[k] = w[a] * w[b]
            elif op == _SUB:
                w[k] = w[a] - w[b]
            elif op == _ADD:
                w[k] = w[a] + w[b]
            elif op == _DIV:
                y = w[b]
                if y != 0.0:
                    w[k] = w[a] / y
                else:                       # IEEE semantics like the generated C
                    x = w[a]
                    w[k] = float("nan") if (x == 0.0 or x != x) else (inf if (x > 0) == (math.copysign(1.0, y) > 0) else -inf)
            elif op == _NEG:
                w[k] = -w[a]
            elif op == _C:
                w[k] = a
            elif op == _S:
                w[k] = sym_val[a]
            elif op == _LT:
                w[k] = 1.0 if w[a] < w[b] else 0.0
            elif op == _IFE:
                w[k] = w[b] if w[a] != 0.0 else w[c]
            elif op == _FMIN:
                w[k] = min(w[a], w[b])
            elif op == _FMAX:
                w[k] = max(w[a], w[b])
            elif op == _POW:
                w[k] = math.pow(w[a], w[b])
            elif op == _SQRT:
                w[k] = math.sqrt(w[a])
            else:
                raise NotImplementedError(op)
        res = []
        for o in self._outs:
            dense = np.zeros(o.arr.shape)
            mask = o.arr != None  # noqa: E711
            for i, j in zip(*np.nonzero(mask)):
                dense[i, j] = w[self._pos[o.arr[i, j].id]]
            res.append(DM(dense, mask))
        return res[0] if len(res) == 1 else tuple(res)

    # -- symbolic (substitution) -----------------------------------------------------------
    def _call_symbolic(self, args):
        if self._tape is None:
            self._build_tape()
        sub = {}
        for m, a in zip(self._ins, args):
            a = _as_sx(a)
            assert a.arr.size == m.arr.size, "%s: argument of %d elements for an input of %d" % (
                self._name, a.arr.size, m.arr.size)
            for n, v in zip(m.arr.reshape(-1, order="F"), a.arr.reshape(-1, order="F")):
                sub[n.id] = ZERO if v is None else v
        w = [None] * len(self._tape)
        for k, nid in enumerate(self._order):
            n = self._nodes[nid]
            op = n.op
            if op == _C:
                w[k] = n
            elif op == _S:
                w[k] = sub[nid]
            else:
                a = w[self._pos[n.a.id]]
                b = w[self._pos[n.b.id]] if n.b is not None else None
                if op == _ADD:
                    w[k] = n_add(a, b)
                elif op == _SUB:
                    w[k] = n_sub(a, b)
                elif op == _MUL:
                    w[k] = n_mul(a, b)
                elif op == _DIV:
                    w[k] = n_div(a, b)
                elif op == _NEG:
                    w[k] = n_neg(a)
                elif op == _POW:
                    w[k] = n_pow(a, b)
                elif op == _LT:
                    w[k] = n_lt(a, b)
                elif op == _IFE:
                    w[k] = n_ife(a, b, w[self._pos[n.c.id]])
                elif op == _FMIN:
                    w[k] = n_fmin(a, b)
                elif op == _FMAX:
                    w[k] = n_fmax(a, b)
                elif op == _SQRT:
                    w[k] = n_sqrt(a)
                else:
                    raise NotImplementedError(op)
        res = []
        for o in self._outs:
            out = SX(o.arr.shape[0], o.arr.shape[1])
            for i, j in zip(*np.nonzero(o.arr != None)):  # noqa: E711
                out.arr[i, j] = w[self._pos[o.arr[i, j].id]]      # the output pattern is the Function's
            res.append(out)
        return res[0] if len(res) == 1 else tuple(res)

    def __call__(self, *args):
        assert len(args) == len(self._ins), "%s expects %d arguments" % (self._name, len(self._ins))
        if any(isinstance(a, SX) for a in args):
            return self._call_symbolic(args)
        return self._call_numeric(args)

    def save(self, path):
        raise RuntimeError("the shim cannot serialise functions")

    @staticmethod
    def load(path):
        raise RuntimeError("the shim cannot load serialised functions (%s)" % path)
