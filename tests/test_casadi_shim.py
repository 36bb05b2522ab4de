"""Known-answer tests of the numeric CasADi stand-in (tests/golden/casadi_shim/casadi.py) that the reference's own
source is executed over when the golden fixtures are recorded.  Each test states the published CasADi semantics it
checks; none of them touches the reference tree, so they run everywhere."""
import importlib.util
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def ca():
    spec = importlib.util.spec_from_file_location("casadi_shim_under_test", os.path.join(HERE, "golden", "casadi_shim", "casadi.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_column_major_conventions(ca):
    """reshape / vec / nonzeros / get_ccs are column-major; SX.sym numbers its entries column by column."""
    A = ca.SX.sym("A", 2, 3)
    f = ca.Function("f", [A], [ca.reshape(A, (3, 2)), ca.vec(A), A.T])
    r, v, t = f(np.array([[1.0, 2, 3], [4, 5, 6]]))
    assert np.array_equal(np.array(r), [[1, 5], [4, 3], [2, 6]])
    assert np.array_equal(np.array(v).reshape(-1), [1, 4, 2, 5, 3, 6])
    assert t.nonzeros() == [1, 2, 3, 4, 5, 6]
    S = ca.SX(3, 3)
    S[0, 0] = A[0, 0]; S[2, 0] = A[1, 0]; S[1, 2] = A[0, 1]
    colind, row = S.sparsity().get_ccs()
    assert (colind, row) == ([0, 2, 2, 3], [0, 2, 1]) and S.nnz() == 3


def test_jacobian_and_hessian_with_structural_sparsity(ca):
    x = ca.SX.sym("x", 3)
    e = ca.vertcat(x[0] * x[1] + 2 * x[2], x[2] / x[0], x[1] - x[1] * 3)
    J = ca.jacobian(e, x)
    f = ca.Function("J", [x], [J])
    Jn = f(np.array([2.0, 3.0, 5.0]))
    assert np.allclose(np.array(Jn), [[3, 2, 2], [-5 / 4, 0, 1 / 2], [0, -2, 0]])
    assert J.sparsity().get_ccs() == ([0, 2, 4, 6], [0, 1, 0, 2, 0, 1])           # structural zeros are not stored
    q = 0.5 * (x.T @ ca.diag(ca.SX([1.0, 2.0, 3.0])) @ x)
    H, g = ca.hessian(q, x)
    Hn, gn = ca.Function("H", [x], [H, g])(np.array([1.0, 1.0, 1.0]))
    assert np.allclose(np.array(Hn), np.diag([1, 2, 3])) and H.nnz() == 3
    assert np.allclose(np.array(gn).reshape(-1), [1, 2, 3])


def test_subtraction_of_a_node_from_itself_keeps_the_entry(ca):
    """u - u simplifies to the constant 0 but stays in the pattern of a dense result (b = A z - eq of the reference
    relies on it: srbd_constraints.py:139-140 gives a dense 140-vector)."""
    u = ca.SX.sym("u", 2)
    d = u - u
    assert d.nnz() == 2
    assert ca.Function("d", [u], [d])(np.array([3.0, 4.0])).nonzeros() == [0.0, 0.0]


def test_if_else_fmin_fmax_mmin(ca):
    a, b = ca.SX.sym("a"), ca.SX.sym("b")
    r = ca.if_else(b < 0, -a / b, 1.0)
    f = ca.Function("r", [a, b], [r, ca.fmax(ca.fmin(1.0, 0.99 * ca.mmin(ca.vertcat(a, b, 2.0))), 1e-12)])
    assert float(f(2.0, -4.0)[0]) == 0.5
    assert float(f(2.0, 0.0)[0]) == 1.0            # the untaken branch divides by zero: if_else_zero(0, inf) = 0
    assert float(f(2.0, 3.0)[1]) == 1.0 and abs(float(f(0.5, 3.0)[1]) - 0.495) < 1e-16 and float(f(-1.0, 3.0)[1]) == 1e-12


@pytest.mark.parametrize("order", ["mindeg", "natural"])
def test_ldl_without_pivoting_solves_a_quasi_definite_system(ca, order, monkeypatch):
    """ldl / ldl_solve: A[p,p] = (I+LT)' D (I+LT) without numerical pivoting (quasi-definite KKT: any order works)."""
    monkeypatch.setenv("CASADI_SHIM_ORDER", order)
    rng = np.random.default_rng(0)
    n, m = 6, 3
    Hm = np.diag(rng.uniform(1, 2, n))
    Am = np.where(rng.uniform(size=(m, n)) < 0.5, rng.normal(size=(m, n)), 0.0)
    K = np.block([[Hm, Am.T], [Am, -1e-8 * np.eye(m)]])
    Ks = ca.SX.sym("K", n + m, n + m)
    rhs = ca.SX.sym("r", n + m)
    mask = K != 0
    Kmat = ca.SX(n + m, n + m)
    for i in range(n + m):
        for j in range(n + m):
            if mask[i, j]:
                Kmat[i, j] = Ks[i, j]
    D, LT, p = ca.ldl(Kmat)
    x = ca.ldl_solve(rhs, D, LT, p)
    f = ca.Function("s", [Ks, rhs], [x, D])
    r = rng.normal(size=n + m)
    xs, Dn = f(K, r)
    assert np.allclose(K @ np.array(xs).reshape(-1), r, atol=1e-9)
    assert sorted(p) == list(range(n + m)) and (np.array(Dn) > 0).sum() == n      # inertia (n, m)


def test_function_substitution_and_nested_calls(ca):
    x, y = ca.SX.sym("x", 2), ca.SX.sym("y")
    inner = ca.Function("inner", [x, y], [x * y + ca.skew(ca.vertcat(x, y))[0:2, 2]])
    z = ca.SX.sym("z", 2)
    outer = ca.Function("outer", [z], [inner(z + 1, z[0] * 2) + inner(z, 3.0)])
    got = np.array(outer(np.array([1.0, 2.0]))).reshape(-1)
    def ref_inner(xv, yv):
        return xv * yv + np.array([xv[1], -xv[0]])
    assert np.allclose(got, ref_inner(np.array([2.0, 3.0]), 2.0) + ref_inner(np.array([1.0, 2.0]), 3.0))


def test_instruction_stream_equals_graph_evaluation(ca):
    """The SXFunction view code generators walk (n_instructions, instruction_id / input / output / constant, sz_w with
    work-slot reuse) evaluates to exactly what the graph does."""
    x = ca.SX.sym("x", 4)
    s = ca.SX.sym("s", 4)
    e = x
    for _ in range(6):
        e = (e * s + ca.fmax(e, 0.5)) / (1.0 + s * s) - ca.if_else(e[0] < 0, e[1], e[2])
    f = ca.Function("f", [x, s], [e, ca.norm_2(e), ca.dot(e, s) ** 3])
    args = [np.array([0.3, -1.2, 2.0, 0.7]), np.array([0.9, 1.1, 0.4, 2.0])]
    want = [np.array(o.nonzeros()) for o in f(*args)]
    w = np.zeros(f.sz_w())
    outs = [np.zeros(f.nnz_out(i)) for i in range(3)]
    two = {ca.OP_ADD: np.add, ca.OP_SUB: np.subtract, ca.OP_MUL: np.multiply, ca.OP_DIV: np.divide, ca.OP_POW: np.power,
           ca.OP_FMIN: min, ca.OP_FMAX: max, ca.OP_LT: lambda a, b: float(a < b), ca.OP_IF_ELSE_ZERO: lambda c, v: 0.0 if c == 0 else v}
    for k in range(f.n_instructions()):
        op, i, o = f.instruction_id(k), f.instruction_input(k), f.instruction_output(k)
        if op == ca.OP_CONST: w[o[0]] = f.instruction_constant(k)
        elif op == ca.OP_INPUT: w[o[0]] = args[i[0]][i[1]]
        elif op == ca.OP_OUTPUT: outs[o[0]][o[1]] = w[i[0]]
        elif op == ca.OP_NEG: w[o[0]] = -w[i[0]]
        elif op == ca.OP_SQRT: w[o[0]] = np.sqrt(w[i[0]])
        elif op == ca.OP_NOT: w[o[0]] = float(not w[i[0]])
        else: w[o[0]] = two[op](w[i[0]], w[i[1]])
    for g, r in zip(outs, want):
        assert np.array_equal(g, r)
    assert f.sz_w() < f.n_instructions() / 3          # slots are reused
