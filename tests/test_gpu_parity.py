"""GPU parity tests (run with `-m gpu` on the B200 box): the CUDA path, called through the
C-ABI, against the oracle on the same seeded inputs and against the golden fixtures.

Tolerances (north_star): fp64 relative error <= 1e-9 (norm-wise per output vector and env)
on primal/dual iterates at identical iteration counts; assembly <= 1e-12 relative."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

import biped_pympc_b200 as bp  # noqa: E402
from helpers import rel_err  # noqa: E402
from oracle import c_oracle as co  # noqa: E402

TOL_IP = 1e-9
DEV = "cuda:0"


def _dev(arrs):
    return [torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=DEV) for a in arrs]


def _np(ts):
    return [t.detach().cpu().numpy() for t in ts]


@pytest.fixture(scope="module")
def batch64():
    inp, rot = bp.workload.make_batch(64, seed=4242)
    nom = bp.workload.nominal_standing_inputs()
    st, rot_s = bp.workload.make_batch(3, seed=5, standing=True)
    inp = [np.concatenate([a, b, c], 0) for a, b, c in zip(nom, inp, st)]
    rot = np.concatenate([np.eye(3)[None], rot, rot_s], 0)
    return inp, rot


def _check_qp(got, ref):
    names = "HfAbGd"
    for n, g, r in zip(names, got, ref):
        scale = max(1.0, np.abs(r).max())
        if n in "HGd":
            assert np.array_equal(g, r), f"{n} must be bit exact"
        elif n == "A":
            assert np.abs(g - r).max() <= 1e-12 * scale
            assert np.array_equal(g == 0, r == 0) or np.abs(g - r).max() < 1e-15
        else:   # b, f: differences of O(1) quantities that cancel -> absolute tolerance
            assert np.abs(g - r).max() <= 1e-12 * scale, (n, np.abs(g - r).max())


def test_qp_form_matches_oracle(batch64):
    inp, _ = batch64
    got = _np(bp.qp_form(_dev(inp)))
    _check_qp(got, co.qp_former(inp))


def test_qp_form_golden(golden):
    inputs = [golden[f"in_{n}"] for n in bp.workload.INPUT_NAMES]
    got = _np(bp.qp_form(_dev(inputs)))
    _check_qp(got, [golden[f"qp_{n}"] for n in "HfAbGd"])


@pytest.mark.parametrize("B", [1, 3, 13, 130])
def test_qp_form_ragged_batches(B):
    inp, _ = bp.workload.make_batch(B, seed=B)
    outs = bp.qp_form(_dev(inp))
    _check_qp(_np(outs), co.qp_former(inp))


def test_empty_batches_everywhere():
    e = lambda n, dt=torch.float64: torch.empty((0, n), dtype=dt, device=DEV)
    out = bp.pdipm([e(n) for n in bp._native.IP_NNZ_IN[:6]], num_iters=5, cold_start=True, stage_invariant=True)
    assert [tuple(o.shape) for o in out] == [(0, n) for n in bp._native.IP_NNZ_OUT]
    ctx = bp.Context(4)
    w = ctx.solve([e(n) for n in bp._native.QP_NNZ_IN], torch.empty((0, 3, 3), dtype=torch.float32, device=DEV))
    assert tuple(w.shape) == (0, 2, 6)
    ctx.close()


def test_qp_form_empty_batch():
    empty = [torch.empty((0, n), dtype=torch.float64, device=DEV) for n in bp._native.QP_NNZ_IN]
    out = bp.qp_form(empty)
    assert [tuple(o.shape) for o in out] == [(0, n) for n in bp._native.QP_NNZ_OUT]


@pytest.mark.parametrize("stage_invariant", [True, False])
def test_pdipm_5_iters_golden(golden, stage_invariant):
    qp = [golden[f"qp_{n}"] for n in "HfAbGd"]
    Hv, f, Av, b, Gv, d = _dev(qp)
    out = _np(bp.pdipm([Hv, Gv, Av, f, d, b], num_iters=5, cold_start=True, stage_invariant=stage_invariant))
    for i, n in enumerate(("x", "s", "z", "y")):
        for e in range(qp[0].shape[0]):
            assert rel_err(out[i][e], golden[f"ip5_{n}"][e]) < TOL_IP, (n, e)
    assert rel_err(out[4], golden["ip5_res"]) < 1e-8
    assert rel_err(out[5], golden["ip5_mu"]) < 1e-8


@pytest.mark.parametrize("iters", [1, 5])
def test_pdipm_matches_long_double_oracle(batch64, iters):
    inp, _ = batch64
    qp = co.qp_former(inp)
    ref = co.pdipm(qp, iters, long_double=True)
    Hv, f, Av, b, Gv, d = _dev(qp)
    out = _np(bp.pdipm([Hv, Gv, Av, f, d, b], num_iters=iters, cold_start=True, stage_invariant=True))
    worst = [0.0] * 4
    for i in range(4):
        for e in range(qp[0].shape[0]):
            worst[i] = max(worst[i], rel_err(out[i][e], ref[i][e]))
    # s, z, y: <= 1e-9 at every iteration count.  x: <= 1e-9 at the reference's iteration count
    # (5 per call).  After the *first* Newton step from the y = 1 cold start |dy| ~ 1e3 and the
    # two unconstrained yaw moments (Phi = R + beta = 1e-4) are recovered as
    # du = (r - b'dy) / 1e-4, which amplifies the last ~10 ulp of dy: 1-2e-9 of |x| on the worst
    # envs ("Accuracy").  The single-iteration diagnostic therefore allows 5e-9 on x.
    tol_x = TOL_IP if iters >= 5 else 5e-9
    assert worst[0] < tol_x and max(worst[1:]) < TOL_IP, worst


def test_config1_single_t1_standing_batch_of_one():
    """BASELINE configs[0]: one T1 biped, standing, horizon 10, 5 iterations -- and the reference's
    real tick, 4 chained 5-iteration calls (mpc_controller_cusadi.py:144-169)."""
    inp = bp.workload.config1_t1_standing()
    qp_ref = co.qp_former(inp)
    ins = _dev(inp)
    _check_qp(_np(bp.qp_form(ins)), qp_ref)
    ref5 = co.pdipm(qp_ref, 5, long_double=True)
    ctx = bp.Context(1)
    rot = torch.eye(3, dtype=torch.float32, device=DEV)[None]
    w, x, s, z, y, res = ctx.solve(ins, rot, num_calls=1, iters_per_call=5, want_iterates=True)
    for g, r in zip((x, s, z, y), ref5):
        assert rel_err(g.cpu().numpy(), r) < TOL_IP
    # 20 iterations: symmetric standing problem -> equal vertical forces approaching m g / 2 each
    w20, x20 = ctx.solve(ins, rot, num_calls=4, iters_per_call=5, want_iterates=True)[:2]
    ref20 = co.pdipm(qp_ref, 20, long_double=True)
    assert rel_err(x20.cpu().numpy(), ref20[0]) < TOL_IP   # standing: no path sensitivity, 1e-9 holds at 20
    u0 = x20[0, 120:132].cpu().numpy()
    # left / right symmetric; the first-stage vertical force is the weight plus the tracking correction
    assert abs(u0[2] - u0[5]) < 1e-4 * abs(u0[2])
    assert abs(u0[2] + u0[5] - 40.0 * 9.81) < 0.15 * 40.0 * 9.81
    assert tuple(w20.shape) == (1, 2, 6)
    ctx.close()


def test_pdipm_explicit_start_and_continuation(batch64):
    """10 inputs given explicitly (no cold-start flag) and 5+5 == 10 bit for bit."""
    inp, _ = batch64
    qp = co.qp_former([a[:16] for a in inp])
    Hv, f, Av, b, Gv, d = _dev(qp)
    B = 16
    x = torch.zeros(B, 240, dtype=torch.float64, device=DEV)
    s = torch.maximum(d, torch.ones_like(d)); z = torch.ones_like(d); y = torch.ones(B, 140, dtype=torch.float64, device=DEV)
    a = bp.pdipm([Hv, Gv, Av, f, d, b, x, s, z, y], num_iters=5, stage_invariant=True)
    c = bp.pdipm([Hv, Gv, Av, f, d, b], num_iters=5, cold_start=True, stage_invariant=True)
    for u, v in zip(a, c):
        assert torch.equal(u, v)
    a2 = bp.pdipm([Hv, Gv, Av, f, d, b, a[0], a[1], a[2], a[3]], num_iters=5, stage_invariant=True)
    ten = bp.pdipm([Hv, Gv, Av, f, d, b], num_iters=10, cold_start=True, stage_invariant=True)
    for u, v in zip(a2, ten):
        assert torch.equal(u, v)


def test_pdipm_general_path_random_values_on_pattern():
    """The reference's own GPU-vs-CPU check feeds torch.rand to every nonzero
    (cusadi/run_cusadi_function_test.py:15): stage-varying values on the frozen pattern."""
    rng = np.random.default_rng(4)
    B = 9
    Hv = rng.uniform(0.5, 2.0, (B, 240)); Gv = rng.uniform(-1, 1, (B, 280)); Av = rng.uniform(-1, 1, (B, 1196))
    f = rng.normal(size=(B, 240)); d = rng.uniform(0.5, 2, (B, 160)); b = rng.normal(size=(B, 140))
    x = rng.normal(size=(B, 240)); s = rng.uniform(0.5, 2, (B, 160)); z = rng.uniform(0.5, 2, (B, 160)); y = rng.normal(size=(B, 140))
    ref = co.pdipm([Hv, f, Av, b, Gv, d], 3, x, s, z, y, long_double=True)
    out = _np(bp.pdipm(_dev([Hv, Gv, Av, f, d, b, x, s, z, y]), num_iters=3, stage_invariant=False))
    for i in range(4):
        assert rel_err(out[i], ref[i]) < 1e-8


def test_tolerance_mode_and_iteration_counts(batch64):
    inp, _ = batch64
    qp = co.qp_former([a[:8] for a in inp])
    Hv, f, Av, b, Gv, d = _dev(qp)
    itd = torch.zeros(8, dtype=torch.int32, device=DEV)
    one = bp.pdipm([Hv, Gv, Av, f, d, b], num_iters=1, cold_start=True, stage_invariant=True)
    got = bp.pdipm([Hv, Gv, Av, f, d, b], num_iters=30, tol=1e30, cold_start=True, stage_invariant=True, iters_done=itd)
    assert itd.tolist() == [1] * 8
    for u, v in zip(one, got):
        assert torch.equal(u, v)
    bp.pdipm([Hv, Gv, Av, f, d, b], num_iters=30, tol=1e-8, cold_start=True, stage_invariant=True, iters_done=itd)
    assert all(1 <= v <= 30 for v in itd.tolist())


def test_optional_fp32_mode_tolerances_and_context_switch():
    """north_star: "an optional fp32 mode with a stated tolerance of <= 1e-4".  fp32 pivot blocks + two refinement steps
    (SRBD_F32_FACTOR | SRBD_F32_REFINE2) is that mode: <= 1e-4 on x, s, z, y against the 80-bit oracle after the
    reference's 5 iterations on 256 walking envs (measured 2e-6).  The unrefined and once-refined variants are measured
    too (6e-2, 3e-4): they document why two steps are needed.  The fp64 path must be untouched by the switch."""
    B = 256
    inp, rot = bp.workload.make_batch(B, seed=77)
    qp = co.qp_former(inp)
    ref = co.pdipm(qp, 5, long_double=True)
    Hv, f, Av, b, Gv, d = _dev(qp)
    worst = {}
    for prec in ("f64", "f32", "f32_refined", "f32_refined2"):
        out = _np(bp.pdipm([Hv, Gv, Av, f, d, b], num_iters=5, cold_start=True, stage_invariant=True, precision=prec))
        worst[prec] = max(rel_err(out[i][e], ref[i][e]) for i in range(4) for e in range(B))
    assert worst["f64"] < TOL_IP, worst
    assert worst["f32_refined2"] < 1e-4, worst
    assert worst["f32_refined2"] < worst["f32_refined"] < worst["f32"] < 0.5, worst
    with pytest.raises(bp._native.SrbdError):      # the general kernel has no reduced-precision blocks
        bp.pdipm([Hv, Gv, Av, f, d, b], num_iters=5, cold_start=True, stage_invariant=False, precision="f32")
    # the fused tick through a context: same switch, and switching back restores the parity path bit for bit
    ins = _dev(inp)
    rot_t = torch.as_tensor(rot, dtype=torch.float32, device=DEV)
    ctx = bp.Context(B, 0)
    a = ctx.solve(ins, rot_t, want_iterates=True)
    ctx.set_precision("f32_refined2")
    c = ctx.solve(ins, rot_t, want_iterates=True)
    ctx.set_precision("f64")
    a2 = ctx.solve(ins, rot_t, want_iterates=True)
    for u, v in zip(a, a2):
        assert torch.equal(u, v)
    xa, xc = a[1].cpu().numpy(), c[1].cpu().numpy()
    assert max(rel_err(xc[e], xa[e]) for e in range(B)) < 1e-4
    assert float((c[0] - a[0]).abs().max()) < 1e-2 * float(a[0].abs().max())     # wrench, fp32
    ctx.close()


def test_tolerance_mode_stops_each_env_at_the_oracles_iteration():
    """BASELINE configs[3] semantics: with tol > 0 an env stops after the first iteration whose pre-step residual
    norms and post-step mu are all < tol (generate_solver_function.py:187).  Standing envs with a loose tolerance stop
    at DIFFERENT intermediate iterations (21 ... 30 of 30): the kernel must stop every env at the oracle's iteration
    and return the oracle's iterate there.  After 21-30 iterations the reference formulation itself, evaluated in
    fp64, is 1e-7 ... 1e-6 away from its 80-bit evaluation on these envs; the kernel has to be at least as
    close as that, and within 5e-8 in absolute terms (measured: 9e-9)."""
    inp, _ = bp.workload.make_batch(48, seed=7, standing=True)
    qp = co.qp_former(inp)
    tol = 1e-1
    ref, it_ref = co.pdipm_tol(qp, 30, tol, long_double=True)
    dbl, it_d = co.pdipm_tol(qp, 30, tol)
    assert len(set(it_ref.tolist())) >= 4 and it_ref.min() < 30          # the test is about intermediate stops
    clear = it_ref == it_d                                                # 64-bit and 80-bit oracle agree on the count
    assert clear.mean() > 0.9
    Hv, f, Av, b, Gv, d = _dev(qp)
    for fast in (True, False):
        itd = torch.full((48,), -7, dtype=torch.int32, device=DEV)
        out = bp.pdipm([Hv, Gv, Av, f, d, b], num_iters=30, tol=tol, cold_start=True, stage_invariant=fast, iters_done=itd)
        got_it = itd.cpu().numpy()
        assert np.array_equal(got_it[clear], it_ref[clear]), (fast, got_it, it_ref)
        for i, n in enumerate("xszy"):
            g = out[i].cpu().numpy()
            e_k = max(rel_err(g[e:e + 1], ref[i][e:e + 1]) for e in np.nonzero(clear)[0])
            e_d = max(rel_err(dbl[i][e:e + 1], ref[i][e:e + 1]) for e in np.nonzero(clear)[0])
            assert e_k <= max(TOL_IP, e_d) and e_k < 5e-8, (fast, n, e_k, e_d)
        # the reported residuals are those of the stopping iteration
        assert np.allclose(out[4].cpu().numpy()[clear], ref[4][clear], rtol=1e-6, atol=1e-12)
    # through the fused tick as well
    ins = _dev(inp)
    ctx = bp.Context(48)
    itd = torch.zeros(48, dtype=torch.int32, device=DEV)
    x = ctx.solve(ins, torch.eye(3, device=DEV).repeat(48, 1, 1), num_calls=1, iters_per_call=30, tol=tol, want_iterates=True,
                  iters_done=itd)[1]
    assert np.array_equal(itd.cpu().numpy()[clear], it_ref[clear])
    assert rel_err(x.cpu().numpy()[clear], ref[0][clear]) < 5e-8
    ctx.close()


# per-iteration parity against the 80-bit oracle: (iterations, tolerance on x s z y).  DESIGN.md section 5: <= 1e-9 at
# the reference's 5 iterations (5e-9 on x after the very first step from the y = 1 cold start); beyond 10 iterations on
# walking contact tables the reference's own fp64 evaluation is only reproducible to 1e-8 ... 1e-4, which is what the
# second bound (10 x the 64-bit oracle's own distance from the 80-bit one) expresses
PARITY_CURVE = ((1, 5e-9), (2, 1e-9), (3, 1e-9), (5, 1e-9), (10, 1e-8), (15, 1e-7), (20, 2e-4), (30, 2e-3))


def test_parity_curve_1_to_30_iterations():
    inp, _ = bp.workload.make_batch(16, seed=99)
    qp = co.qp_former(inp)
    Hv, f, Av, b, Gv, d = _dev(qp)
    worst = {}
    for iters, tol in PARITY_CURVE:
        ref = co.pdipm(qp, iters, long_double=True)
        dbl = co.pdipm(qp, iters)
        out = _np(bp.pdipm([Hv, Gv, Av, f, d, b], num_iters=iters, cold_start=True, stage_invariant=True))
        for i, n in enumerate("xszy"):
            e_k = max(rel_err(out[i][e:e + 1], ref[i][e:e + 1]) for e in range(16))
            e_d = max(rel_err(dbl[i][e:e + 1], ref[i][e:e + 1]) for e in range(16))
            worst[(iters, n)] = (e_k, e_d)
            assert e_k <= tol, (iters, n, e_k)
            assert e_k <= max(1e-9, 10.0 * e_d) or iters == 1, (iters, n, e_k, e_d)
    # standing envs (all feet in contact): 1e-9 up to 10 iterations, and never farther from the 80-bit oracle than the
    # reference formulation evaluated in fp64 is
    st, _ = bp.workload.make_batch(8, seed=3, standing=True)
    qps = co.qp_former(st)
    dv = _dev(qps)
    for iters in (5, 10, 20, 30):
        ref = co.pdipm(qps, iters, long_double=True)
        dbl = co.pdipm(qps, iters)
        out = _np(bp.pdipm([dv[0], dv[4], dv[2], dv[1], dv[5], dv[3]], num_iters=iters, cold_start=True, stage_invariant=True))
        for i in range(4):
            assert rel_err(out[i], ref[i]) <= max(TOL_IP if iters <= 10 else 1e-8, rel_err(dbl[i], ref[i])), (iters, i)


@pytest.mark.parametrize("B,seed", [(4096, 20261018), (16384, 20261019)])
def test_full_baseline_batch_against_the_oracle(B, seed):
    """The whole BASELINE batch (configs[1]: 4 096 envs of the bench seed; and a 16 384-env batch, configs[3]'s size)
    through the fused tick, EVERY env against the 80-bit oracle at the reference's 5 iterations, assembly included.
    Bar: <= 1e-9 per env and output vector -- except on the few envs where the reference formulation itself, evaluated
    in fp64 (oracle/srbd_oracle.c, double build), is farther than 1e-9 from its own 80-bit evaluation (env 798 of the
    bench batch: 1.8e-9 on x): there the kernel must be at least as close to the 80-bit value as the reference's fp64
    evaluation is.  At most 0.1 % of the envs may be of that kind."""
    inp, rot = bp.workload.make_batch(B, seed=seed)
    qp = co.qp_former(inp)
    ref = co.pdipm(qp, 5, long_double=True)
    dbl = co.pdipm(qp, 5)
    ins = _dev(inp)
    _check_qp(_np(bp.qp_form(ins)), qp)
    ctx = bp.Context(B)
    out = ctx.solve(ins, torch.as_tensor(rot, dtype=torch.float32, device=DEV), num_calls=1, iters_per_call=5, want_iterates=True)
    over = np.zeros(B, bool)
    for i, n in enumerate("xszy"):
        g = out[1 + i].cpu().numpy()
        nrm = np.maximum(np.linalg.norm(ref[i], axis=1), 1e-300)
        err = np.linalg.norm(g - ref[i], axis=1) / nrm
        err_ref64 = np.linalg.norm(dbl[i] - ref[i], axis=1) / nrm
        bad = err > np.maximum(TOL_IP, err_ref64)
        assert not bad.any(), (n, int(bad.sum()), float(err.max()), int(err.argmax()), float(err_ref64[err.argmax()]))
        over |= err > TOL_IP
    assert over.mean() <= 1e-3, int(over.sum())
    ctx.close()


def test_cusadi_compatible_evaluate(golden):
    """lib<fn>.so::evaluate with DEVICE arrays of device pointers (CusadiFunction.py:84-92)."""
    import os
    libdir = os.path.join(os.path.dirname(bp._native.LIB_PATH))
    inputs = _dev([golden[f"in_{n}"] for n in bp.workload.INPUT_NAMES])
    B = inputs[0].shape[0]
    outs = [torch.zeros((B, n), dtype=torch.float64, device=DEV) for n in bp._native.QP_NNZ_OUT]
    in_ptrs = torch.tensor([t.data_ptr() for t in inputs], dtype=torch.int64, device=DEV)
    out_ptrs = torch.tensor([t.data_ptr() for t in outs], dtype=torch.int64, device=DEV)
    work = torch.zeros((B, 8), dtype=torch.float64, device=DEV)
    lib = ctypes.CDLL(os.path.join(libdir, "libqp_former.so"))
    lib.evaluate.restype = ctypes.c_float
    lib.evaluate.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    t = lib.evaluate(in_ptrs.data_ptr(), work.data_ptr(), out_ptrs.data_ptr(), B)
    assert t >= 0
    _check_qp(_np(outs), [golden[f"qp_{n}"] for n in "HfAbGd"])
    # solver: 5 iterations from the reference cold start
    Hv, f, Av, b, Gv, d = outs
    x = torch.zeros(B, 240, dtype=torch.float64, device=DEV)
    s = torch.maximum(d, torch.ones_like(d)); z = torch.ones_like(d); y = torch.ones(B, 140, dtype=torch.float64, device=DEV)
    ins2 = [Hv, Gv, Av, f, d, b, x, s, z, y]
    outs2 = [torch.zeros((B, n), dtype=torch.float64, device=DEV) for n in bp._native.IP_NNZ_OUT]
    in2 = torch.tensor([t_.data_ptr() for t_ in ins2], dtype=torch.int64, device=DEV)
    out2 = torch.tensor([t_.data_ptr() for t_ in outs2], dtype=torch.int64, device=DEV)
    lib2 = ctypes.CDLL(os.path.join(libdir, "libsparse_pdipm_multiple_iterations.so"))
    lib2.evaluate.restype = ctypes.c_float
    lib2.evaluate.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    assert lib2.evaluate(in2.data_ptr(), work.data_ptr(), out2.data_ptr(), B) >= 0
    for i, n in enumerate(("x", "s", "z", "y")):
        assert rel_err(outs2[i].cpu().numpy(), golden[f"ip5_{n}"]) < TOL_IP


def test_b200function_mirrors_cusadifunction(golden):
    inputs = _dev([golden[f"in_{n}"] for n in bp.workload.INPUT_NAMES])
    B = inputs[0].shape[0]
    former = bp.B200Function("qp_former", B)
    former.evaluate(inputs)
    A_dense = former.getDenseOutput(2)
    assert tuple(A_dense.shape) == (B, 140, 240)
    from oracle import srbd_oracle as so
    ref = so.dense_from_ccs(golden["qp_A"][1], so.A_COLPTR, so.A_ROW, (140, 240))
    assert np.abs(A_dense[1].cpu().numpy() - ref).max() < 1e-12
    f_dense = former.getDenseOutput(1)
    assert tuple(f_dense.shape) == (B, 240, 1)
    solver = bp.B200Function("sparse_pdipm_multiple_iterations", B, stage_invariant=True)
    Hs, fs, As, bs, Gs, ds = former.outputs_sparse
    x = torch.zeros(B, 240, dtype=torch.float64, device=DEV)
    s = torch.maximum(ds, torch.ones_like(ds)); z = torch.ones_like(ds); y = torch.ones(B, 140, dtype=torch.float64, device=DEV)
    for _ in range(4):      # the reference's MAX_ITER loop, mpc_controller_cusadi.py:144-169
        solver.evaluate([Hs, Gs, As, fs.reshape(B, -1, 1), ds.reshape(B, -1, 1), bs.reshape(B, -1, 1),
                         x.reshape(B, -1, 1), s.reshape(B, -1, 1), z.reshape(B, -1, 1), y.reshape(B, -1, 1)])
        x, s, z, y = [solver.getDenseOutput(i).reshape(B, -1).clone() for i in range(4)]
    assert rel_err(x.cpu().numpy(), golden["ip20_x"]) < 1e-6


class _Obj:
    pass


def _fake_controller(B, seed):
    """The attributes BaseMPCController prepares (base_controller.py:61-257), fp32 like the reference."""
    inp, rot = bp.workload.make_batch(B, seed=seed)
    c = bp.MPCControllerB200()
    c.num_envs, c.device, c.horizon_length = B, torch.device(DEV), 10
    f32 = lambda a: torch.as_tensor(a, dtype=torch.float32, device=DEV)
    c.x0 = f32(inp[0]); c.x_ref = f32(inp[3]).view(B, 10, 12)
    c.R_body = f32(rot).unsqueeze(1).repeat(1, 10, 1, 1)
    sed = _Obj(); sed.rotation_body = f32(rot); sed.root_position = f32(inp[9])
    sed.foot_position = torch.stack([f32(inp[10]), f32(inp[11])], 1)
    c.state_estimate_data = sed
    c.biped = _Obj(); c.biped.I_body = f32(bp.workload.I_BODY)
    c.contact_table = f32(inp[12]).view(B, 10, 2)
    c.dt_mpc = f32(inp[4]).view(B); c.mass, c.mu = 40.0, 1.0
    c.Q = torch.tensor([150, 150, 250, 100, 100, 250, 1, 1, 5, 10, 10, 1, 1], dtype=torch.float32, device=DEV)
    c.R = torch.tensor([1e-5] * 6 + [1e-4] * 6, dtype=torch.float32, device=DEV)
    c.residual_lin_accel = f32(inp[15]); c.residual_ang_accel = f32(inp[16])
    return c, inp, rot


@pytest.mark.parametrize("fix_layout,qbug", [(False, False), (True, False), (False, True)])
def test_input_builder_kernel_equals_torch_marshalling(fix_layout, qbug):
    """srbd_marshal_inputs (one launch) == the ~25 torch ops of mpc_controller_cusadi.py:54-95, bit for
    bit, including the layout quirks Q1-Q3 of SURVEY 8b."""
    c, _, _ = _fake_controller(37, seed=8)
    a = bp.mpc_controller_b200.marshal_inputs(c, fix_layout, qbug)
    b = bp.mpc_controller_b200.marshal_inputs_torch(c, fix_layout, qbug)
    for i, (u, v) in enumerate(zip(a, b)):
        assert u.dtype == torch.float64 and tuple(u.shape) == (37, bp._native.QP_NNZ_IN[i])
        if i == 8:      # I_world = R I_body R' computed inside the launch: fp32, within 2 ulp of torch's two matmuls
            ulp = torch.finfo(torch.float32).eps * v.reshape(37, -1).abs().amax(dim=1, keepdim=True)
            assert ((u - v.reshape(37, -1)).abs() <= 2.0 * ulp).all()
            assert torch.equal(u, u.float().double())                       # still fp32 values
        else:
            assert torch.equal(u, v.reshape(37, -1)), i
    t = bp.mpc_controller_b200.marshal_inputs(c, fix_layout, qbug, torch_i_world=True)   # torch's I_world passed in: bit for bit
    assert all(torch.equal(u, v.reshape(37, -1)) for u, v in zip(t, b))
    nox = bp.mpc_controller_b200.marshal_inputs(c, fix_layout, qbug, with_linearisation_point=False)
    assert nox[1] is None and nox[2] is None and torch.equal(nox[3], a[3])


def test_controller_run_matches_oracle_tick():
    B = 6
    c, inp, rot = _fake_controller(B, seed=31)
    c.init_solver()
    wrench, cost = c.run()
    assert tuple(wrench.shape) == (B, 2, 6) and wrench.dtype == torch.float32 and tuple(cost.shape) == (B,)
    # oracle: same marshalling (I_world recomputed in fp32 like the reference does); the controller's defaults are the
    # reference's three layout quirks (reference_quirks = True: Q1, Q2 and the stride-12 read of the 13-long Q)
    ins = [t.cpu().numpy() for t in bp.mpc_controller_b200.marshal_inputs_torch(c, reproduce_q_stride_bug=True)]
    qp = co.qp_former(ins)
    it = co.pdipm(qp, 20, long_double=True)
    from oracle import srbd_oracle as so
    for e in range(B):
        u0 = it[0][e][120:132]
        Rt = rot[e].astype(np.float32).T
        fl, fr, ml, mr = [u0[i:i + 3].astype(np.float32) for i in (0, 3, 6, 9)]
        ml[0] = 0; mr[0] = 0
        ref = np.concatenate([-(Rt @ fl), -(Rt @ ml), -(Rt @ fr), -(Rt @ mr)]).reshape(2, 6)
        got = wrench[e].cpu().numpy()
        assert np.abs(got - ref).max() <= 1e-4 * max(1.0, np.abs(ref).max())


def test_fused_tick_equals_two_step_path(batch64):
    inp, rot = batch64
    ins = _dev(inp)
    rot_t = torch.as_tensor(rot, dtype=torch.float32, device=DEV)
    ctx = bp.Context(len(rot))
    w, x, s, z, y, res = ctx.solve(ins, rot_t, num_calls=1, iters_per_call=5, want_iterates=True)
    Hv, f, Av, b, Gv, d = bp.qp_form(ins)
    out = bp.pdipm([Hv, Gv, Av, f, d, b], num_iters=5, cold_start=True, stage_invariant=True)
    assert torch.equal(out[0], x) and torch.equal(out[2], z) and torch.equal(out[4], res)
    assert torch.equal(bp.wrench(x, rot_t), w)
    # host entry point: pinned host buffers, H2D + kernels + D2H
    h_in = [torch.as_tensor(a).pin_memory() for a in inp]
    h_rot = torch.as_tensor(rot, dtype=torch.float32).pin_memory()
    h_w = torch.empty((len(rot), 2, 6), dtype=torch.float32).pin_memory()
    ctx.solve_host(h_in, h_rot, h_w, num_calls=1, iters_per_call=5)
    assert torch.equal(h_w, w.cpu())
    ctx.close()


@pytest.mark.parametrize("pinned,B", [(True, 2500), (False, 2500), (True, 9000)])
def test_host_entry_chunks_and_default_linearisation_point(pinned, B):
    """srbd_mpc_solve_host cuts the batch into chunks on two streams (pinned buffers: staged by the prefetch kernel /
    the copy engine, the wrench written in place over PCIe; pageable ones staged with cudaMemcpyAsync) and accepts NULL
    for the (dead) linearisation point x, u = ones: all of it must leave the result bit-identical.
    2 500 envs: the smaller round first (1 776 + 724; pageable: 3 chunks); 9 000: four chunks alternating streams."""
    inp, rot = bp.workload.make_batch(B, seed=77)
    ctx = bp.Context(B)
    rot_t = torch.as_tensor(rot, dtype=torch.float32, device=DEV)
    w_dev, x_dev = ctx.solve(_dev(inp), rot_t, num_calls=1, iters_per_call=5, want_iterates=True)[:2]
    pin = (lambda t: t.pin_memory()) if pinned else (lambda t: t)
    h_in = [pin(torch.as_tensor(a)) for a in inp]
    h_rot = pin(torch.as_tensor(rot, dtype=torch.float32))
    h_w = pin(torch.empty((B, 2, 6), dtype=torch.float32))
    h_x = pin(torch.empty((B, 240), dtype=torch.float64))
    ctx.solve_host(h_in, h_rot, h_w, num_calls=1, iters_per_call=5, x_out_host=h_x)
    assert torch.equal(h_w, w_dev.cpu()) and torch.equal(h_x, x_dev.cpu())
    h_in2 = list(h_in); h_in2[1] = None; h_in2[2] = None
    h_w2 = pin(torch.zeros_like(h_w))
    h_x2 = pin(torch.zeros_like(h_x))
    ctx.solve_host(h_in2, h_rot, h_w2, num_calls=1, iters_per_call=5, x_out_host=h_x2)
    assert torch.equal(h_w2, h_w) and torch.equal(h_x2, h_x)
    # bound call (buffers validated once): the tick a control loop runs after refilling the same buffers in place
    tick = ctx.bind_host(h_in2, h_rot, h_w2, x_out_host=h_x2)
    h_w2.zero_()
    assert tick(1, 5) is h_w2 and torch.equal(h_w2, h_w)
    inp3, rot3 = bp.workload.make_batch(B, seed=78)
    for t, a in zip(h_in2, inp3):
        if t is not None:
            t.copy_(torch.as_tensor(a))
    h_rot.copy_(torch.as_tensor(rot3, dtype=torch.float32))
    tick(1, 5)
    w3 = ctx.solve(_dev(inp3), torch.as_tensor(rot3, dtype=torch.float32, device=DEV), num_calls=1, iters_per_call=5)
    assert torch.equal(h_w2, w3.cpu())
    with pytest.raises(ValueError):
        ctx.bind_host([t if i else t[:5] for i, t in enumerate(h_in)], h_rot, h_w)      # inconsistent batch sizes
    ctx.close()


def test_stage_invariance_probe_and_auto_dispatch(golden):
    """The synchronous drop-in entries probe the batch and take the fast kernel for qp_former data."""
    qp = [golden[f"qp_{n}"] for n in "HfAbGd"]
    Hv, f, Av, b, Gv, d = _dev(qp)
    assert bp.is_stage_invariant([Hv, Gv, Av]) is True
    for arr, idx in ((Av, 36 + 40), (Gv, 28 * 3 + 5), (Hv, 12 * 7 + 2), (Av, 336 + 86 * 4 + 17)):
        t = arr.clone(); t[1, idx] += 1e-9
        trial = [t if a is arr else a for a in (Hv, Gv, Av)]
        assert bp.is_stage_invariant(trial) is False
    B = Hv.shape[0]
    fn = bp.B200Function("sparse_pdipm_multiple_iterations", B)          # default: probe
    x = torch.zeros(B, 240, dtype=torch.float64, device=DEV)
    s = torch.maximum(d, torch.ones_like(d)); z = torch.ones_like(d); y = torch.ones(B, 140, dtype=torch.float64, device=DEV)
    fn.evaluate([Hv, Gv, Av, f, d, b, x, s, z, y])
    fast = bp.pdipm([Hv, Gv, Av, f, d, b, x, s, z, y], num_iters=5, stage_invariant=True)
    assert torch.equal(fn.outputs_sparse[0], fast[0])


def test_fast_path_rejects_data_that_is_not_qp_former_shaped(golden):
    """SRBD_STAGE_INVARIANT promises unit coefficients on x_{k+1} / Mx rows; anything else must
    fail loudly (NaN iterates, iteration count -1), never silently solve another problem."""
    qp = [golden[f"qp_{n}"] for n in "HfAbGd"]
    Hv, f, Av, b, Gv, d = _dev(qp)
    Av = Av.clone(); Av[:, 36] = 2.0     # coefficient of x_2[0] in dyn_1 (a_xcol_start(1, 0), srbd_pattern.h)
    itd = torch.zeros(Av.shape[0], dtype=torch.int32, device=DEV)
    out = bp.pdipm([Hv, Gv, Av, f, d, b], num_iters=5, cold_start=True, stage_invariant=True, iters_done=itd)
    assert torch.isnan(out[0]).all() and (itd == -1).all()
    gen = bp.pdipm([Hv, Gv, Av, f, d, b], num_iters=5, cold_start=True, stage_invariant=False)
    assert torch.isfinite(gen[0]).all()


def test_compat_evaluate_falls_back_to_the_general_kernel_without_allocating(golden):
    """The drop-in `evaluate` launches the fast kernel, which verifies per env that the data is qp_former-shaped
    (PdipmArgs::inv_flag); a batch that is not goes to the general kernel in the same call.  The per-device scratch
    only grows: a second call of the same size allocates nothing (device memory in use is unchanged)."""
    import os
    qp = [golden[f"qp_{n}"] for n in "HfAbGd"]
    Hv, f, Av, b, Gv, d = _dev(qp)
    B = Hv.shape[0]
    lib = ctypes.CDLL(os.path.join(os.path.dirname(bp._native.LIB_PATH), "libsparse_pdipm_multiple_iterations.so"))
    lib.evaluate.restype = ctypes.c_float
    lib.evaluate.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    x = torch.zeros(B, 240, dtype=torch.float64, device=DEV)
    s = torch.maximum(d, torch.ones_like(d)); z = torch.ones_like(d); y = torch.ones(B, 140, dtype=torch.float64, device=DEV)

    def call(Avv):
        ins = [Hv, Gv, Avv, f, d, b, x, s, z, y]
        outs = [torch.zeros((B, n), dtype=torch.float64, device=DEV) for n in bp._native.IP_NNZ_OUT]
        ip = torch.tensor([t.data_ptr() for t in ins], dtype=torch.int64, device=DEV)
        op = torch.tensor([t.data_ptr() for t in outs], dtype=torch.int64, device=DEV)
        torch.cuda.synchronize()
        assert lib.evaluate(ip.data_ptr(), None, op.data_ptr(), B) >= 0
        return outs

    fast = call(Av)
    ref = bp.pdipm([Hv, Gv, Av, f, d, b, x, s, z, y], num_iters=5, stage_invariant=True)
    assert torch.equal(fast[0], ref[0])
    Av2 = Av.clone(); Av2[1, 36 + 40] *= 1.0 + 1e-6          # stage 2 differs from stage 1 in one env
    slow = call(Av2)
    gen = bp.pdipm([Hv, Gv, Av2, f, d, b, x, s, z, y], num_iters=5, stage_invariant=False)
    assert torch.isfinite(slow[0]).all() and torch.equal(slow[0], gen[0])
    free0 = torch.cuda.mem_get_info()[0]
    call(Av); call(Av2)
    assert torch.cuda.mem_get_info()[0] == free0


def test_operator_accepts_a_casadi_like_function_and_checks_its_buffers(golden):
    """B200Function(fn, B) with the object the reference passes (a casadi.Function: has name(), no triplet_out) resolves
    to the library's own patterns; a batch other than num_instances and short caller-provided buffers raise."""
    class FakeCasadiFunction:               # the slice of casadi.Function that identifies the function
        def name(self): return "qp_former"
    inputs = _dev([golden[f"in_{n}"] for n in bp.workload.INPUT_NAMES])
    B = inputs[0].shape[0]
    fn = bp.B200Function(FakeCasadiFunction(), B)
    fn.evaluate(inputs)
    assert tuple(fn.getDenseOutput(0).shape) == (B, 240, 240) and tuple(fn.getDenseOutput(2).shape) == (B, 140, 240)
    with pytest.raises(ValueError):
        bp.B200Function("qp_former", B - 1).evaluate(inputs)
    with pytest.raises(ValueError):
        bp.qp_form(inputs, out=[torch.empty((B - 1, n), dtype=torch.float64, device=DEV) for n in bp._native.QP_NNZ_OUT])
    Hv, f, Av, b, Gv, d = bp.qp_form(inputs)
    with pytest.raises(ValueError):
        bp.pdipm([Hv, Gv, Av, f, d, b], cold_start=True, stage_invariant=True,
                 work=torch.empty((B, 100), dtype=torch.float64, device=DEV))
    with pytest.raises(ValueError):
        bp.pdipm([Hv, Gv, Av, f, d, b], cold_start=True, stage_invariant=False,
                 work=torch.empty((B, bp._native.WORK_DOUBLES_FAST), dtype=torch.float64, device=DEV))
    with pytest.raises((ValueError, TypeError)):
        bp.pdipm([Hv, Gv, Av, f, d, b], cold_start=True, stage_invariant=True,
                 iters_done=torch.zeros(B - 1, dtype=torch.int32, device=DEV))
    # a misaligned view (odd double offset) is refused by the C-ABI instead of faulting in the kernel
    buf = torch.empty(B * 120 + 1, dtype=torch.float64, device=DEV)
    bad = list(inputs); bad[3] = buf[1:].view(B, 120); bad[3].copy_(inputs[3])
    with pytest.raises(bp.SrbdError):
        bp.qp_form(bad)


@pytest.mark.parametrize("B", [1, 13, 15, 16, 17, 147, 148, 149, 2073, 2367, 2368, 2369, 4145])
def test_fused_tick_ragged_batches(B):
    """Round boundaries of the persistent fused kernel (at most one CTA per SM, 16 warps; env = round * 16 G + warp * G +
    cta with G = min(SMs, B) CTAs): a batch gives the same per-env answer whatever its size, and the same as assembly +
    solve + wrench as three launches."""
    inp, rot = bp.workload.make_batch(B, seed=100 + B)
    ins = _dev(inp)
    rot_t = torch.as_tensor(rot, dtype=torch.float32, device=DEV)
    ctx = bp.Context(B)
    w, x, s, z, y, res = ctx.solve(ins, rot_t, num_calls=1, iters_per_call=5, want_iterates=True)
    Hv, f, Av, b, Gv, d = bp.qp_form(ins)
    out = bp.pdipm([Hv, Gv, Av, f, d, b], num_iters=5, cold_start=True, stage_invariant=True)
    for g, r in zip((x, s, z, y, res), out[:5]):
        assert torch.equal(g, r)
    assert torch.equal(w, bp.wrench(x, rot_t))
    ctx.close()


def test_65536_envs_in_one_launch():
    """BASELINE configs[2] total size on one GPU: finite, deterministic, independent of the batch position."""
    B = 65536
    inp, rot = bp.workload.make_batch(4096, seed=20261018)
    ins = [t.repeat(16, 1) for t in _dev(inp)]
    rot_t = torch.as_tensor(rot, dtype=torch.float32, device=DEV).repeat(16, 1, 1)
    ctx = bp.Context(B)
    w, x = ctx.solve(ins, rot_t, num_calls=1, iters_per_call=5, want_iterates=True)[:2]
    assert torch.isfinite(x).all() and torch.isfinite(w).all()
    assert torch.equal(x[:4096], x[4096 * 15:]) and torch.equal(w[:4096], w[4096 * 7:4096 * 8])
    ctx.close()


def test_full_size_properties():
    """BASELINE config 2 size (4096 envs): size-independent properties instead of the oracle."""
    B = 4096
    inp, rot = bp.workload.make_batch(B, seed=20261018)
    ins = _dev(inp)
    Hv, f, Av, b, Gv, d = bp.qp_form(ins)
    assert torch.equal(Hv[:, :12], ins[13]) and torch.equal(Hv[:, 228:], ins[14])
    dd = d.view(B, 10, 16)
    ct = ins[12].view(B, 2, 10).transpose(1, 2)                 # column-major (10,2)
    assert torch.equal(dd[:, :, 7], 500.0 * ct[:, :, 0]) and torch.equal(dd[:, :, 15], 500.0 * ct[:, :, 1])
    out5 = bp.pdipm([Hv, Gv, Av, f, d, b], num_iters=5, cold_start=True, stage_invariant=True)
    again = bp.pdipm([Hv, Gv, Av, f, d, b], num_iters=5, cold_start=True, stage_invariant=True)
    for u, v in zip(out5, again):
        assert torch.equal(u, v)                                  # deterministic
    assert all(torch.isfinite(t).all() for t in out5)
    assert (out5[1] >= 1e-8).all() and (out5[2] >= 1e-8).all()    # clamps
    # the residuals reported by iteration 5 are those of the iterate after 4 iterations
    x4, s4, z4, y4, _, _ = bp.pdipm([Hv, Gv, Av, f, d, b], num_iters=4, cold_start=True, stage_invariant=True)
    from oracle import srbd_oracle as so
    idx = torch.as_tensor(so.A_ROW.astype(np.int64) * 240 + so.A_COL, device=DEV)
    Ad = torch.zeros(B, 140 * 240, dtype=torch.float64, device=DEV); Ad[:, idx] = Av; Ad = Ad.view(B, 140, 240)
    idg = torch.as_tensor(so.G_ROW.astype(np.int64) * 240 + so.G_COL, device=DEV)
    Gd = torch.zeros(B, 160 * 240, dtype=torch.float64, device=DEV); Gd[:, idg] = Gv; Gd = Gd.view(B, 160, 240)
    mv = lambda M, v: torch.bmm(M, v.unsqueeze(2)).squeeze(2)
    rx = Hv * x4 + f + mv(Gd.transpose(1, 2), z4) + mv(Ad.transpose(1, 2), y4)
    re = mv(Ad, x4) - b
    rs = mv(Gd, x4) + s4 - d
    res = torch.stack([rx.norm(dim=1), rs.norm(dim=1), re.norm(dim=1)], 1)
    assert ((res - out5[4][:, :3]).abs() <= 1e-9 * (1 + res.abs())).all()
    mu = (out5[1] * out5[2]).sum(1) / 160
    assert ((mu - out5[4][:, 3]).abs() <= 1e-12 * (1 + mu.abs())).all()
    # env results do not depend on batch position / size
    sub = [t[1000:1013].contiguous() for t in (Hv, Gv, Av, f, d, b)]
    o13 = bp.pdipm(sub, num_iters=5, cold_start=True, stage_invariant=True)
    assert torch.equal(o13[0], out5[0][1000:1013])
    # the general (stage-varying) path gives the same answer to round-off
    gen = bp.pdipm([t[:256].contiguous() for t in (Hv, Gv, Av, f, d, b)], num_iters=5, cold_start=True, stage_invariant=False)
    assert rel_err(gen[0].cpu().numpy(), out5[0][:256].cpu().numpy()) < 1e-9


# ---- against the reference's own source executed over the CasADi stand-in ----------------------
# (tests/golden/srbd_ref_shim_v1.npz, generator tests/golden/make_golden_ref.py; see
#  tests/test_reference_pin.py for what the fixture is and for the CPU side of the same checks)
@pytest.fixture(scope="module")
def ref():
    import os
    return dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "srbd_ref_shim_v1.npz")))


def _worst_vs_ref(out, ref, it):
    return {n: max(rel_err(out[i][e], ref[f"ip{it}_{n}"][e]) for e in range(out[i].shape[0]))
            for i, n in enumerate("xszy")}


def test_qp_form_matches_the_reference_code(ref):
    got = _np(bp.qp_form(_dev([ref[f"in_{n}"] for n in bp.workload.INPUT_NAMES])))
    _check_qp(got, [ref[f"qp_{n}"] for n in "HfAbGd"])


@pytest.mark.parametrize("stage_invariant", [True, False])
def test_pdipm_matches_the_reference_code(ref, stage_invariant):
    """sparse_pdipm_multiple_iterations(num_iters=5) of the reference (sparse_pdipm_solver.py:357-534)
    on qp_former's outputs, from the controller's cold start: <= 1e-9 per env and vector."""
    Hv, f, Av, b, Gv, d = _dev([ref[f"qp_{n}"] for n in "HfAbGd"])
    out = _np(bp.pdipm([Hv, Gv, Av, f, d, b], num_iters=5, cold_start=True, stage_invariant=stage_invariant))
    w = _worst_vs_ref(out, ref, 5)
    assert max(w.values()) < TOL_IP, w
    assert rel_err(out[4], ref["ip5_res"]) < 1e-8
    assert rel_err(out[5].reshape(-1), ref["ip5_mu"].reshape(-1)) < 1e-8


def test_chained_calls_match_the_reference_code(ref):
    """The controller's four chained calls (mpc_controller_cusadi.py:144-169).  Beyond the first call
    the reference itself is only reproducible to the tolerances below on walking contact tables
    (tests/test_reference_pin.py); the standing problem of srbd_constraints.py:243-288 keeps 1e-9."""
    Hv, f, Av, b, Gv, d = _dev([ref[f"qp_{n}"] for n in "HfAbGd"])
    tol = {5: TOL_IP, 10: 1e-8, 15: 5e-7, 20: 5e-6}
    cur = None
    for it in (5, 10, 15, 20):
        ins = [Hv, Gv, Av, f, d, b] + ([] if cur is None else list(cur[:4]))
        cur = bp.pdipm(ins, num_iters=5, cold_start=cur is None, stage_invariant=True)
        w = _worst_vs_ref(_np(cur), ref, it)
        assert max(w.values()) < tol[it], (it, w)
    out = _np(cur)
    for i, n in enumerate("xszy"):
        assert rel_err(out[i][0], ref[f"ip20_{n}"][0]) < TOL_IP


def test_fused_tick_matches_the_reference_code(ref):
    ins = _dev([ref[f"in_{n}"] for n in bp.workload.INPUT_NAMES])
    B = ins[0].shape[0]
    rot = torch.eye(3, dtype=torch.float32, device=DEV)[None].repeat(B, 1, 1)
    ctx = bp.Context(B)
    w, x, s, z, y, res = ctx.solve(ins, rot, num_calls=1, iters_per_call=5, want_iterates=True)
    wst = _worst_vs_ref(_np([x, s, z, y]), ref, 5)
    assert max(wst.values()) < TOL_IP, wst
    # one launch of 20 iterations == the reference's 4 x 5, at the chained-call tolerance
    x20 = ctx.solve(ins, rot, num_calls=4, iters_per_call=5, want_iterates=True)[1]
    assert max(rel_err(x20[e].cpu().numpy(), ref["ip20_x"][e]) for e in range(B)) < 5e-6
    u0 = ref["ip20_x"][:, 120:132].astype(np.float32)
    u0[:, 6] = 0; u0[:, 9] = 0                                   # mpc_controller_cusadi.py:193-194
    want = -u0.reshape(B, 2, 2, 3).transpose(0, 2, 1, 3).reshape(B, 2, 6)   # [foot][f, m]; identity rotation
    w20 = ctx.solve(ins, rot, num_calls=4, iters_per_call=5)
    assert np.abs(w20.cpu().numpy() - want).max() <= 1e-3 * max(1.0, np.abs(want).max())
    ctx.close()


# ---- against the reference's own controller run (tests/golden/srbd_ref_controller_v1.npz) --------
@pytest.fixture(scope="module")
def refc():
    import os
    return dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "srbd_ref_controller_v1.npz")))


V_W_COLUMNS = [3, 4, 9, 10]     # entries that go through the body->world velocity (a torch.bmm in the reference)


@pytest.mark.parametrize("tick", [0, 1])
def test_state_builder_kernel_matches_the_reference_controller(refc, tick):
    """srbd_reference_trajectory (one launch) against x0 / x_ref / knot point as the reference's
    compute_knot_points + set_initial_state + compute_reference_trajectory left them
    (base_controller.py:164-257): bit-exact, except the four columns that depend on R @ v_b (the
    accumulation order of torch.bmm is the BLAS back-end's): 2 fp32 ulp there."""
    from test_reference_pin import fake_controller_from_fixture
    t = f"t{tick}_"
    c = fake_controller_from_fixture(refc, tick, DEV)
    x0, x_ref = c.build_state()
    assert np.array_equal(x0.cpu().numpy(), refc[t + "x0"])
    got, want = x_ref.cpu().numpy(), refc[t + "x_ref"]
    exact = [k for k in range(12) if k not in V_W_COLUMNS]
    assert np.array_equal(got[:, :, exact], want[:, :, exact])
    assert np.abs(got - want).max() <= 2.4e-7 * max(1.0, np.abs(want).max())
    assert not c.first_run.any()
    assert np.array_equal(c.yaw_desired.cpu().numpy(), refc[t + "yaw_desired_out"])
    assert np.array_equal(c.world_position_desired.cpu().numpy(), refc[t + "world_position_desired_out"])
    assert torch.equal(c.R_body[:, 3], c.state_estimate_data.rotation_body)


def test_controller_two_ticks_match_the_reference_controller(refc):
    """MPCControllerB200.run() -- state builder, input builder, fused tick -- over the two recorded ticks:
    the 17 marshalled inputs against what the reference handed to qp_former, the wrench against what its
    run() returned (mpc_controller_cusadi.py:43-205)."""
    from test_reference_pin import effective_in17, fake_controller_from_fixture
    c = fake_controller_from_fixture(refc, 0, DEV)
    c.init_solver()
    for tick in (0, 1):
        t = f"t{tick}_"
        nxt = fake_controller_from_fixture(refc, tick, DEV)       # this tick's estimator state / command
        for n in ("state_estimate_data", "desired_state_data", "contact_table", "dt_mpc", "residual_lin_accel",
                  "residual_ang_accel"):
            setattr(c, n, getattr(nxt, n))
        wrench, cost = c.run()                                     # knot-point state carries over from tick 0
        ins = bp.mpc_controller_b200.marshal_inputs(c, reproduce_q_stride_bug=True)
        want = effective_in17(refc, tick)
        for i, (n, g, w) in enumerate(zip(bp.workload.INPUT_NAMES, ins, want)):
            g = g.cpu().numpy()
            if n in ("x_ref", "I_world"):                          # fp32 matmuls upstream: 2 ulp
                assert np.abs(g - w).max() <= 2.4e-7 * max(1.0, np.abs(w).max()), n
            else:
                assert np.array_equal(g, w), n
        ref_w = refc[t + "wrench"]
        assert wrench.dtype == torch.float32 and tuple(wrench.shape) == ref_w.shape
        assert np.abs(wrench.cpu().numpy() - ref_w).max() <= 1e-3 * max(1.0, np.abs(ref_w).max())
        assert float(cost.abs().max()) == 0.0
    c._ctx.close()


# ---- gait generator kernel against a recording of the reference's GaitGenerator -----------------
def test_gait_kernel_matches_the_reference_gait_generator():
    """tests/golden/srbd_ref_gait_v1.npz (generator tests/golden/make_golden_gait.py): 32 envs x 400 control
    steps of the reference class on the CPU, four gaits, sampling time changed mid-run.  Sub-phases, MPC contact
    table and the phase itself must be bit-identical at every step, through the fused tick() and through the
    reference-style separate calls."""
    import os
    g = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "srbd_ref_gait_v1.npz")))
    B, H, T = g["ssp"].shape[0], int(g["horizon"]), g["table"].shape[0]
    mk = lambda: bp.GaitGeneratorB200(B, H, float(g["dt"]), torch.as_tensor(g["dt_mpc0"]), torch.as_tensor(g["dsp"]),  # noqa: E731
                                      torch.as_tensor(g["ssp"]), DEV)
    fused, sep = mk(), mk()
    for gen in (fused, sep):
        gen.gait_phase[:] = torch.as_tensor(g["phase0"], device=DEV)
    same = lambda a, b: np.array_equal(a, b, equal_nan=True)  # noqa: E731
    for t in range(T):
        if t == int(g["switch_step"]):
            fused.update_sampling_time(torch.as_tensor(g["dt_mpc1"]))
            sep.update_sampling_time(torch.as_tensor(g["dt_mpc1"]))
        assert same(fused.gait_phase.cpu().numpy(), g["phase"][t]), t
        c, s, tab = fused.tick()
        assert same(c.cpu().numpy(), g["contact_sub"][t]), t
        assert same(s.cpu().numpy(), g["swing_sub"][t]), t
        assert np.array_equal(tab.cpu().numpy(), g["table"][t].astype(np.float32)), t
        if t % 25 == 0:       # the reference's call sequence: three queries, then update_phase
            sep.gait_phase[:] = torch.as_tensor(g["phase"][t], device=DEV)
            assert same(sep.get_swing_sub_phase().cpu().numpy(), g["swing_sub"][t])
            assert same(sep.get_contact_sub_phase().cpu().numpy(), g["contact_sub"][t])
            assert sep.mpc_gait.dtype == torch.int32 and np.array_equal(sep.mpc_gait.cpu().numpy(), g["table"][t])
            assert np.array_equal(sep.contact_bool.cpu().numpy(), g["contact_sub"][t] != -1)
            sep.update_phase()
            assert same(sep.gait_phase.cpu().numpy(), g["phase"][t + 1])
    assert same(fused.gait_phase.cpu().numpy(), g["phase"][T])
    assert same(fused.swing_durations_sec.cpu().numpy(), g["swing_durations_sec"])
    fused.reset(torch.tensor([0, 3], device=DEV))
    assert fused.gait_phase[0] == 0 and fused.gait_phase[3] == 0


# ---- the whole control step against a recording of the reference stack ------------------------------
@pytest.mark.parametrize("fixture", ["srbd_ref_loop_v1.npz", "srbd_ref_loop_v1_world.npz", "srbd_ref_loop_v1_standing.npz",
                                     "srbd_ref_loop_v1_events.npz"])
def test_control_stack_matches_the_reference_stack(fixture):
    """tests/golden/srbd_ref_loop_v1.npz (generator tests/golden/make_golden_loop.py): 160 control steps x 6 envs
    of the reference's MPCController wrapper (gait generator, state estimator, HECTOR leg controller, swing-leg
    controller, MPC every 10th step) on the CPU.  BipedControllerB200 -- two launches per control step -- must
    reproduce every intermediate quantity: flags, phases and counters exactly, fp32 arithmetic to a few ulp
    (sin / cos / atan2 and matrix products are not bit-reproducible across libm / BLAS back-ends).

    The MPC wrench is compared at 5e-3 of its scale: after the reference's 4 x 5 cold-started iterations on a
    walking contact table the iterate is not a continuous function of the inputs at fp32 resolution -- moving
    x_ref by ONE fp32 ulp moves u0 by up to 0.95 N in the 80-bit oracle -- and x_ref carries a
    torch.bmm upstream (2 ulp between back-ends).  With bit-identical inputs the same solve agrees to 1e-3 /
    3e-5 N (test_controller_two_ticks_match_the_reference_controller, tests/test_reference_pin.py).  Downstream
    of the MPC the recorded wrench is injected so that the remaining quantities are compared tightly.
    `_world`: swing targets in the world frame (ControllerConf.swing_reference_frame); `_standing`: gait_id = 1,
    whose zero swing time makes the reference divide by zero (inf / nan in vDes: reproduced position by position);
    `_events`: an episode reset of two envs and the RL-side knobs (update_mpc_sampling_time, set_swing_parameters,
    set_srbd_accel: mpc_wrapper.py:44-67) changed mid-run."""
    import os
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, "tools"))
    import loop_errors
    g = dict(np.load(os.path.join(here, "golden", fixture)))
    err = loop_errors.run(g, inject=True, dev=DEV)
    scale = lambda n: max(1.0, float(np.abs(g[n][np.isfinite(g[n])]).max())) if np.isfinite(g[n]).any() else 1.0  # noqa: E731
    exact = ("leg_contact_bool", "leg_swing_bool", "contact_phase", "swing_phase", "mpc_contact_table", "swing_time_remaining",
             "first_swing", "cmd_kpjoint", "cmd_kdjoint", "gait_phase_after")
    tol = {"wrench": 5e-3, "cmd_vDes": 5e-6, "cmd_tau": 4e-6, "action": 1.5e-5}     # relative to the scale of the quantity
    for n, v in err.items():
        if n in exact:
            assert v == 0.0, (n, v)
        else:
            assert v <= tol.get(n, 2e-6) * scale(n), (n, v, scale(n))
    free = loop_errors.run(g, inject=False, dev=DEV)                # the closed chain, MPC wrench not injected
    assert free["action"] <= 5e-3 * scale("action") and free["first_swing"] == 0.0


def test_cuda_graph_step_is_bit_identical_to_the_eager_step():
    """BipedControllerB200.capture_graphs / step(): the control step replayed as a CUDA graph gives exactly the
    eager step's joint torques and leaves the same persistent state, over the recorded sensor stream."""
    import os
    g = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "srbd_ref_loop_v1.npz")))
    B, dec = g["state"].shape[1], int(g["decimation"])
    mk = lambda: bp.BipedControllerB200(B, DEV, dt=float(g["dt"]), dt_mpc=float(g["dt_mpc"]), decimation=dec)  # noqa: E731
    eager, graph = mk(), mk()
    graph.capture_graphs()
    for c in (eager, graph):
        c.gait_generator.gait_phase[:] = torch.as_tensor(g["phase0"], device=DEV)
    assert bool(graph.swing_leg_controller.first_swing.all()) and float(graph.swing_leg_controller.p0.abs().max()) == 0.0
    for t in range(60):
        st = torch.as_tensor(g["state"][t], device=DEV)
        for c in (eager, graph):
            c.set_command(torch.as_tensor(g["twist"][t], device=DEV), torch.as_tensor(g["height"][t], device=DEV))
        eager.update_state(st)
        if t % dec == 0:
            eager.run_mpc()
        eager.run_lowlevel()
        graph.step(st, t % dec == 0)
        assert torch.equal(eager.action, graph.action), t
    for a, b in ((eager.gait_generator.gait_phase, graph.gait_generator.gait_phase),
                 (eager.swing_leg_controller.first_swing, graph.swing_leg_controller.first_swing),
                 (eager.mpc_controller.world_position_desired, graph.mpc_controller.world_position_desired),
                 (eager.leg_controller.command.feedfowardforce, graph.leg_controller.command.feedfowardforce)):
        assert torch.equal(a, b)
    eager.close(); graph.close()


# ---- against the reference's CusADi GPU path (oracle/_ref, built by oracle/make_cusadi_ref.py) ------------
def _cusadi(name, B, suffix=""):
    import cusadi_ref
    if not cusadi_ref.available(name, suffix):
        pytest.skip(f"oracle/_ref/lib{name}{suffix}.so not built (python oracle/make_cusadi_ref.py)")
    return cusadi_ref.CusadiRefFunction(name, B, DEV, suffix)


def test_qp_form_matches_the_cusadi_generated_kernel(batch64):
    """The kernel the reference's generateCUDACode.py emits for qp_former (one thread per env running the whole
    expression tape, cusadi/src/generateCUDACode.py:62-190), on the GPU, against srbd_qp_form on the same inputs."""
    inp, _ = batch64
    ins = _dev(inp)
    ref = _cusadi("qp_former", len(inp[0]))
    assert ref.evaluate(ins) > 0.0
    _check_qp(_np(bp.qp_form(ins)), _np(ref.outputs_sparse))


def test_pdipm_matches_the_cusadi_generated_kernel(batch64):
    """... and for sparse_pdipm_multiple_iterations, chained like the controller does (mpc_controller_cusadi.py:144-169):
    GPU against GPU, <= 1e-9 after the reference's 5 iterations.  The 5-iteration kernel (400 k instructions per env) needs
    more than 58 GB in ptxas; where only the num_iters=1 build of the same reference function exists it is chained."""
    import cusadi_ref
    var = cusadi_ref.solver_variant()
    if var is None:
        pytest.skip("no CusADi solver kernel in oracle/_ref (python oracle/make_cusadi_ref.py --only solver [--iters 1])")
    suffix, per_call = var
    inp, _ = batch64
    B = len(inp[0])
    Hv, f, Av, b, Gv, d = bp.qp_form(_dev(inp))
    ref = _cusadi("sparse_pdipm_multiple_iterations", B, suffix)
    x = torch.zeros(B, 240, dtype=torch.float64, device=DEV)
    s = torch.maximum(d, torch.ones_like(d)); z = torch.ones_like(d); y = torch.ones(B, 140, dtype=torch.float64, device=DEV)
    cur = None
    for call, tol in ((1, TOL_IP), (2, 1e-8)):
        for _ in range(5 // per_call):
            assert ref.evaluate([Hv, Gv, Av, f, d, b, x, s, z, y]) > 0.0
            x, s, z, y = [t.clone() for t in ref.outputs_sparse[:4]]
        ins = [Hv, Gv, Av, f, d, b] + ([] if cur is None else list(cur[:4]))
        cur = bp.pdipm(ins, num_iters=5, cold_start=cur is None, stage_invariant=True)
        got, want = _np(cur), _np([x, s, z, y])
        worst = max(rel_err(got[i][e], want[i][e]) for i in range(4) for e in range(B))
        assert worst < tol, (call, worst)


def test_general_kernel_matches_the_cusadi_kernel_on_random_values():
    """The reference's own GPU check feeds torch.rand to every nonzero (cusadi/run_cusadi_function_test.py:15,32-41):
    stage-varying random values on the frozen pattern, one Newton iteration, general kernel against the CusADi-generated
    kernel of the reference function (num_iters=1 build)."""
    import cusadi_ref
    if not cusadi_ref.available(cusadi_ref.SOLVER, "_it1"):
        pytest.skip("oracle/_ref/libsparse_pdipm_multiple_iterations_it1.so not built")
    B = 33
    g = torch.Generator(device=DEV).manual_seed(5)
    r = lambda n, lo=0.0, hi=1.0: lo + (hi - lo) * torch.rand(B, n, dtype=torch.float64, device=DEV, generator=g)  # noqa: E731
    ins = [r(240, 0.5, 2.0), r(280, -1, 1), r(1196, -1, 1), r(240, -1, 1), r(160, 0.5, 2), r(140, -1, 1),
           r(240, -1, 1), r(160, 0.5, 2), r(160, 0.5, 2), r(140, -1, 1)]
    ref = cusadi_ref.CusadiRefFunction(cusadi_ref.SOLVER, B, DEV, "_it1")
    ref.evaluate(ins)
    out = bp.pdipm(ins, num_iters=1, stage_invariant=False)
    for i in range(4):
        got, want = out[i].cpu().numpy(), ref.outputs_sparse[i].cpu().numpy()
        assert max(rel_err(got[e], want[e]) for e in range(B)) < 1e-8, i


@pytest.mark.gpu
def test_two_contexts_on_two_streams_share_the_gpu():
    """Two contexts launching the persistent solve kernel on two streams at once: every CTA allocates all 512
    tensor-memory columns of its SM (and all of its shared memory), so the second kernel's CTAs can only become resident as
    the first one's leave -- nothing may deadlock, and both results equal the ones computed one after the other."""
    B = 3000
    batches = [bp.workload.make_batch(B, seed=500 + i) for i in range(2)]
    ctxs = [bp.Context(B) for _ in range(2)]
    devs = [(_dev(inp), torch.as_tensor(rot, dtype=torch.float32, device=DEV)) for inp, rot in batches]
    want = [c.solve(i, r, num_calls=1, iters_per_call=5, want_iterates=True)[:2] for c, (i, r) in zip(ctxs, devs)]
    want = [(w.clone(), x.clone()) for w, x in want]
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    got = [None, None]
    for rep in range(3):
        for k in range(2):
            with torch.cuda.stream(streams[k]):
                got[k] = ctxs[k].solve(devs[k][0], devs[k][1], num_calls=1, iters_per_call=5, want_iterates=True)[:2]
        torch.cuda.synchronize()
        for k in range(2):
            assert torch.equal(got[k][0], want[k][0]) and torch.equal(got[k][1], want[k][1])
    for c in ctxs:
        c.close()
