"""CPU tests that pin the oracle (oracle/cone_oracle.c) -- the checker the GPU tests rely on.

The reference (cvxpy/cvxpylayers) stores no golden vectors for this path and its arithmetic
(diffcp + SCS) is not importable here; what its own tests pin are analytic facts.
Those are restated below with the reference test they come from, next to solver-independent
checks (KKT certificates, HiGHS, SciPy's LSQR, finite differences).
"""
import numpy as np
import pytest
import scipy.optimize as sopt
import torch
from scipy.sparse.linalg import lsqr as scipy_lsqr

from cvxpylayers_b200 import problems as pr
from cvxpylayers_b200.structure import ConeSpec, Structure
from oracle import np_ref
from oracle import oracle as orc
from tests.util import GOLDEN_CASES, load_golden, ref_sdp_batch, ref_soc_batch, rel_err


# ----------------------------------------------------------------------------- forward
@pytest.mark.parametrize("name,B", [("C1", 4), ("C2", 3), ("C3", 4), ("C5", 3), ("C5S", 3)])
@pytest.mark.parametrize("eps", [1e-4, 1e-8])
def test_forward_kkt_certificate(name, B, eps):
    """Every returned point satisfies SCS's own termination criteria on the original data."""
    bt = pr.CONFIGS[name](B=B)
    x, y, s, status, iters = orc.solve_batch(bt.structure, bt.A_vals, bt.b, bt.c, bt.P_vals, eps=eps, max_iters=50000)
    assert (status == 1).all()
    for i in range(B):
        P = bt.P_dense(i) if bt.P_vals is not None else None
        r = np_ref.kkt_residuals(bt.A_dense(i), P, bt.b[i], bt.c[i], x[i], y[i], s[i])
        assert np_ref.is_converged(r, eps, eps, 1.0001)
        assert abs(s[i] @ y[i]) <= 1e-9 * max(1.0, np.abs(s[i]).max() * np.abs(y[i]).max() * bt.structure.m)  # exact complementarity
    if bt.x_star is not None and name != "C5S":   # (C5S: SURVEY's literal SDP, non-unique optimum)
        assert np.abs(x - bt.x_star).max() < 200 * eps  # planted optimum recovered


def test_lp_matches_highs():
    """Cross-solver check (the reference cross-checks against Clarabel, tests/test_dual_variables.py:14-42)."""
    # (plain operator splitting is slow on LPs -- no Anderson acceleration yet, DESIGN.md -- so the
    # comparison uses the instances of this seed that converge quickly)
    bt = pr.dense_lp(6, 8, 20, seed=4).select([1, 2, 4])
    x, y, s, status, _ = orc.solve_batch(bt.structure, bt.A_vals, bt.b, bt.c, eps=1e-9, max_iters=400000)
    assert (status == 1).all()
    for i in range(bt.B):
        A = bt.A_dense(i)
        res = sopt.linprog(bt.c[i], A_ub=A, b_ub=bt.b[i], bounds=[(None, None)] * bt.structure.n, method="highs")
        assert res.status == 0
        assert abs(bt.c[i] @ x[i] - res.fun) < 1e-6 * max(1, abs(res.fun))
        assert np.abs(x[i] - res.x).max() < 1e-5
        assert np.abs(y[i] + res.ineqlin.marginals).max() < 1e-5  # SCS dual sign convention: y >= 0


def test_equality_qp_known_answer():
    """min ||x||^2 s.t. x1 + x2 = 2 -> x* = [1, 1]  (reference tests/test_diffcp_optional_deps.py:29-57)."""
    st = Structure(2, 1, [0, 2], [0, 1], ConeSpec(z=1), [0, 1, 2], [0, 1])
    A = np.array([[1.0, 1.0]]); b = np.array([[2.0]]); c = np.zeros((1, 2)); P = np.array([[2.0, 2.0]])
    x, y, s, status, _ = orc.solve_batch(st, A, b, c, P, eps=1e-10)
    assert status[0] == 1 and np.abs(x[0] - 1.0).max() < 1e-7
    # d x / d b = [1/2, 1/2]: gradient of sum(x) wrt b is 1
    dA, dP, db, dc, _ = orc.vjp_batch(st, A, b, c, x, y, s, np.ones((1, 2)), np.zeros((1, 1)), P)
    assert abs(db[0, 0] - 1.0) < 1e-6


def test_dual_variables_known_answers():
    """The two closed-form dual-variable cases of the reference (tests/test_dual_variables.py:14-42 and :45-71), in solver form
    `A x + s = b, s in K` with the SCS sign convention (y = the multiplier of `A x + s - b`):
      min c'x s.t. sum(x) = b, x >= 0, c = [1, 2], b = 1   ->  x* = [1, 0], nu = -1, lambda = [0, 1];
      min c'x + ||x||^2 s.t. x >= 0, c = [1, -1]            ->  x* = [0, 1/2], lambda = c + 2 x* = [1, 0]."""
    st = Structure(2, 3, [0, 2, 3, 4], [0, 1, 0, 1], ConeSpec(z=1, l=2))   # rows: sum(x) = b | -x1 <= 0 | -x2 <= 0
    A = np.array([[1.0, 1.0, -1.0, -1.0]]); b = np.array([[1.0, 0.0, 0.0]]); c = np.array([[1.0, 2.0]])
    x, y, s, status, _ = orc.solve_batch(st, A, b, c, None, eps=1e-10, max_iters=200000)
    assert status[0] == 1
    assert np.abs(x[0] - [1.0, 0.0]).max() < 1e-6 and np.abs(y[0] - [-1.0, 0.0, 1.0]).max() < 1e-6
    # the dual's sensitivity (reference :177-206 differentiates through it): d(c'x*)/db = -nu = 1 here, and d x*/d b = e1
    dA, dP, db, dc, _ = orc.vjp_batch(st, A, b, c, x, y, s, np.array([[1.0, 0.0]]), np.zeros((1, 3)), None)
    assert abs(db[0, 0] - 1.0) < 1e-5
    st2 = Structure(2, 2, [0, 1, 2], [0, 1], ConeSpec(l=2), [0, 1, 2], [0, 1])
    A2 = np.array([[-1.0, -1.0]]); b2 = np.zeros((1, 2)); c2 = np.array([[1.0, -1.0]]); P2 = np.array([[2.0, 2.0]])
    x, y, s, status, _ = orc.solve_batch(st2, A2, b2, c2, P2, eps=1e-10)
    assert status[0] == 1 and np.abs(x[0] - [0.0, 0.5]).max() < 1e-7 and np.abs(y[0] - [1.0, 0.0]).max() < 1e-7


def test_ridge_closed_form_and_gradient():
    """Least squares with closed-form solution and gradient (reference tests/test_torch.py:90-118:
    x* = (A'A + I)^{-1} A'b, grads atol 1e-6 at eps 1e-10), written as a QP with P = 2(A'A + I)."""
    rng = np.random.default_rng(0)
    mA, n = 30, 8
    A_t = torch.tensor(rng.standard_normal((mA, n)), requires_grad=True)
    b_t = torch.tensor(rng.standard_normal(mA), requires_grad=True)
    x_cf = torch.linalg.solve(A_t.T @ A_t + torch.eye(n, dtype=torch.double), A_t.T @ b_t)
    x_cf.sum().backward()
    # QP data; one slack row 0*x + s = 1, s >= 0 keeps m >= 1
    iu = np.triu_indices(n)
    st = Structure(n, 1, [0, 0], np.zeros(0, np.int32), ConeSpec(l=1),
                   np.concatenate([[0], np.cumsum(np.arange(n, 0, -1))]), iu[1])
    A2 = A_t.detach().clone().requires_grad_(True); b2 = b_t.detach().clone().requires_grad_(True)
    P_full = 2 * (A2.T @ A2 + torch.eye(n, dtype=torch.double))
    P_up = P_full[iu[0], iu[1]]
    c_t = -2 * A2.T @ b2
    x, y, s, status, _ = orc.solve_batch(st, np.zeros((1, 0)), np.ones((1, 1)), c_t.detach().numpy()[None], P_up.detach().numpy()[None], eps=1e-10)
    assert status[0] == 1
    assert np.abs(x[0] - x_cf.detach().numpy()).max() < 1e-6
    dA, dP, db, dc, _ = orc.vjp_batch(st, np.zeros((1, 0)), np.ones((1, 1)), c_t.detach().numpy()[None], x, y, s,
                                      np.ones((1, n)), np.zeros((1, 1)), P_up.detach().numpy()[None])
    ((P_up * torch.tensor(dP[0])).sum() + (c_t * torch.tensor(dc[0])).sum()).backward()
    assert np.abs(A2.grad.numpy() - A_t.grad.numpy()).max() < 1e-6
    assert np.abs(b2.grad.numpy() - b_t.grad.numpy()).max() < 1e-6


def test_infeasible_and_unbounded_are_reported():
    """Status -> exception contract of the reference (tests/test_torch.py:299-316)."""
    st = Structure.dense(1, 2, ConeSpec(l=2))
    # x <= -1 and -x <= -1 (x >= 1): infeasible
    _, _, _, status, _ = orc.solve_batch(st, np.array([[1.0, -1.0]]), np.array([[-1.0, -1.0]]), np.array([[0.0]]))
    assert status[0] == -2
    # min x s.t. x <= 1: unbounded below
    _, _, _, status, _ = orc.solve_batch(st, np.array([[1.0, 0.0]]), np.array([[1.0, 1.0]]), np.array([[1.0]]))
    assert status[0] == -1


def test_max_iters_degrades_answer():
    """solver_args reach the solver: max_iters=1 must visibly degrade (reference tests/test_torch.py:705-752)."""
    bt = pr.CONFIGS["C1"](B=1)
    x1, _, _, st1, it1 = orc.solve_batch(bt.structure, bt.A_vals, bt.b, bt.c, bt.P_vals, max_iters=1)
    x2, _, _, st2, _ = orc.solve_batch(bt.structure, bt.A_vals, bt.b, bt.c, bt.P_vals, eps=1e-8)
    assert st1[0] == 2 and it1[0] == 1 and st2[0] == 1
    assert np.abs(x1 - bt.x_star).max() > 100 * np.abs(x2 - bt.x_star).max()


# ----------------------------------------------------------------------------- Anderson acceleration
@pytest.mark.parametrize("name,B", [("C2", 6), ("C3", 6), ("C5", 4), ("C5S", 4)])
def test_anderson_acceleration_same_solution_fewer_iterations(name, B):
    """SCS accelerates its iterate by default (lookback 10, every 10 iterations, memory filled first); the reference's
    tests switch it off with {"acceleration_lookback": 0} (tests/test_torch.py:401-405).  Off / type-I / type-II must
    reach the same certified optimum; at a tight tolerance the accelerated runs need fewer iterations; within the first
    100 iterations (window not full yet) the accelerated and the plain run are the same run."""
    bt = pr.CONFIGS[name](B=B)
    st = bt.structure
    runs = {}
    for lb in (0, 10, -10):
        x, y, s, status, iters = orc.solve_batch(st, bt.A_vals, bt.b, bt.c, bt.P_vals, eps=1e-9, max_iters=100000, acceleration_lookback=lb)
        assert (status == 1).all(), (lb, status)
        for i in range(B):
            P = bt.P_dense(i) if bt.P_vals is not None else None
            assert np_ref.is_converged(np_ref.kkt_residuals(bt.A_dense(i), P, bt.b[i], bt.c[i], x[i], y[i], s[i]), 1e-9, 1e-9, 1.0001)
        runs[lb] = (x, iters)
    for lb in (10, -10):
        if name == "C5S":   # the planted SDP optimum is not unique in x: compare the objective value
            assert np.abs((runs[lb][0] * bt.c).sum(1) - (runs[0][0] * bt.c).sum(1)).max() < 1e-6 * max(1.0, np.abs((runs[0][0] * bt.c).sum(1)).max())
        else:
            assert np.abs(runs[lb][0] - runs[0][0]).max() < 1e-6 * max(1.0, np.abs(runs[0][0]).max())
        if runs[0][1].mean() > 300:   # long runs: acceleration pays clearly; short ones have at most a few accelerated steps
            assert runs[lb][1].mean() < 0.9 * runs[0][1].mean(), (lb, runs[lb][1], runs[0][1])
        else:
            assert runs[lb][1].mean() <= 1.15 * runs[0][1].mean(), (lb, runs[lb][1], runs[0][1])
    short = {lb: orc.solve_batch(st, bt.A_vals, bt.b, bt.c, bt.P_vals, eps=1e-9, max_iters=100, acceleration_lookback=lb) for lb in (0, 10)}
    assert np.array_equal(short[0][0], short[10][0]) and np.array_equal(short[0][2], short[10][2])


def test_anderson_acceleration_safeguard_keeps_lps_convergent():
    """LP vertices are where unsafeguarded acceleration misbehaves; with the safeguard every variant still terminates
    with the certificate (and the every-iteration variant, acceleration_interval=1, too)."""
    bt = pr.dense_lp(4, 8, 20, seed=4)
    for lb, iv in ((10, 10), (-10, 10), (5, 1)):
        x, y, s, status, iters = orc.solve_batch(bt.structure, bt.A_vals, bt.b, bt.c, eps=1e-6, max_iters=400000,
                                                 acceleration_lookback=lb, acceleration_interval=iv)
        assert (status == 1).all(), (lb, iv, status, iters)
        for i in range(bt.B):
            assert np_ref.is_converged(np_ref.kkt_residuals(bt.A_dense(i), None, bt.b[i], bt.c[i], x[i], y[i], s[i]), 1e-6, 1e-6, 1.0001)


# ----------------------------------------------------------------------------- warm start
@pytest.mark.parametrize("name,B", [("C2", 4), ("C3", 4), ("C5", 3)])
def test_warm_start_is_a_fixed_point_and_speeds_up_nearby_problems(name, B):
    """SURVEY.md 8f.2 (the reference offers warm starts for one backend only, torch/cvxpylayer.py:464-487): started at its own
    solution the splitting stops at the first check with the same answer; started at the solution of a slightly perturbed
    problem (a training-loop step, examples/torch/algorithms.py:34-41) it needs fewer iterations than from cold and reaches
    the same certified optimum."""
    bt = pr.CONFIGS[name](B=B)
    st = bt.structure
    args = dict(eps=1e-8, max_iters=100000)
    x, y, s, status, it_cold = orc.solve_batch(st, bt.A_vals, bt.b, bt.c, bt.P_vals, **args)
    assert (status == 1).all()
    x2, y2, s2, status2, it_fix = orc.solve_batch(st, bt.A_vals, bt.b, bt.c, bt.P_vals, warm=(x, y, s), **args)
    assert (status2 == 1).all() and (it_fix <= 25).all() and np.abs(x2 - x).max() < 1e-6 * max(1.0, np.abs(x).max())
    rng = np.random.default_rng(0)
    b2 = bt.b + 1e-4 * rng.standard_normal(bt.b.shape) * (np.abs(bt.b) > 0)
    c2 = bt.c + 1e-4 * rng.standard_normal(bt.c.shape)
    args = dict(eps=1e-6, max_iters=100000)   # (the start is ~1e-4 from the new optimum: two decades to go instead of six)
    xc, yc, sc, stc, it_c = orc.solve_batch(st, bt.A_vals, b2, c2, bt.P_vals, **args)
    xw, yw, sw, stw, it_w = orc.solve_batch(st, bt.A_vals, b2, c2, bt.P_vals, warm=(x, y, s), **args)
    assert (stc == 1).all() and (stw == 1).all()
    assert it_w.mean() < 0.7 * it_c.mean(), (it_w, it_c)
    for i in range(B):
        P = bt.P_dense(i) if bt.P_vals is not None else None
        assert np_ref.is_converged(np_ref.kkt_residuals(bt.A_dense(i), P, b2[i], c2[i], xw[i], yw[i], sw[i]), 1e-6, 1e-6, 1.0001)
    assert np.abs((c2 * xw).sum(1) - (c2 * xc).sum(1)).max() < 1e-5 * max(1.0, np.abs((c2 * xc).sum(1)).max())


# ----------------------------------------------------------------------------- cones
def test_cone_projection_and_jacobian_against_numpy():
    cones = ConeSpec(z=2, l=3, q=[4, 1, 5], s=[3, 2])
    m = cones.m
    st = Structure(1, m, np.arange(m + 1), np.zeros(m, np.int32), cones)
    rng = np.random.default_rng(3)
    for _ in range(20):
        v = rng.standard_normal(m) * 2
        assert np.abs(orc.proj_dual_cone(st, v) - pr.proj_dual_cone(v, cones)).max() < 1e-12
        D = np_ref.dproj_matrix(v, cones)
        dv = rng.standard_normal(m)
        assert np.abs(orc.dproj_dual_cone(st, v, dv) - D @ dv).max() < 1e-10
        # Jacobian is the derivative of the projection (finite differences)
        h = 1e-6
        fd = (pr.proj_dual_cone(v + h * dv, cones) - pr.proj_dual_cone(v - h * dv, cones)) / (2 * h)
        assert np.abs(fd - D @ dv).max() < 1e-5


# ----------------------------------------------------------------------------- LSQR and the adjoint
def test_lsqr_matches_scipy():
    """diffcp's lsqr.cpp ports SciPy's LSQR; the oracle's must reproduce SciPy's iterates."""
    rng = np.random.default_rng(1)
    for (r, c) in [(30, 30), (40, 25), (25, 40)]:
        M = rng.standard_normal((r, c)); rhs = rng.standard_normal(r)
        sol, its = orc.lsqr_dense(M, rhs)
        ref = scipy_lsqr(M, rhs, atol=1e-8, btol=1e-8, conlim=1e8, iter_lim=2 * c)
        assert its == ref[2]
        assert np.abs(sol - ref[0]).max() < 1e-6 * max(1, np.abs(ref[0]).max())


@pytest.mark.parametrize("name,B", [("C1", 3), ("C3", 3)])
@pytest.mark.parametrize("precond", [0, 1])
def test_vjp_matches_scipy_restatement(name, B, precond):
    """Oracle adjoint vs the NumPy/SciPy restatement of diffcp's adjoint_derivative (explicit M, scipy lsqr)."""
    bt = pr.CONFIGS[name](B=B)
    st = bt.structure
    x, y, s, status, _ = orc.solve_batch(st, bt.A_vals, bt.b, bt.c, bt.P_vals, eps=1e-10, max_iters=100000)
    rng = np.random.default_rng(2)
    dx, dy = rng.standard_normal(x.shape), rng.standard_normal(y.shape)
    dA, dP, db, dc, _ = orc.vjp_batch(st, bt.A_vals, bt.b, bt.c, x, y, s, dx, dy, bt.P_vals, lsqr_precond=precond, lsqr_iter_lim=50000)
    rows = np.repeat(np.arange(st.m), np.diff(st.A_indptr))
    for i in range(B):
        P = bt.P_dense(i) if bt.P_vals is not None else None
        rA, rP, rb, rc, _ = np_ref.vjp_dense(bt.A_dense(i), P, bt.b[i], bt.c[i], x[i], y[i], s[i], dx[i], dy[i], st.cones, exact=True)
        tol = 5e-5 if precond else 1e-4   # north_star: 1e-4 relative; plain LSQR stops at atol = btol = 1e-8 on a worse-conditioned system
        assert rel_err(db[i], rb) < tol and rel_err(dc[i], rc) < tol
        assert rel_err(dA[i], rA[rows, st.A_indices]) < tol
        if P is not None:
            prow = np.repeat(np.arange(st.n), np.diff(st.P_indptr)); pcol = st.P_indices
            rPv = np.where(prow == pcol, rP[prow, pcol], rP[prow, pcol] + rP[pcol, prow])
            assert rel_err(dP[i], rPv) < tol


def test_vjp_matches_finite_differences():
    """Central differences through a tight forward solve (the reference uses torch.autograd.gradcheck
    with atol 1e-4 / rtol 1e-3, tests/test_torch.py:415-426, tests/test_dual_variables.py:209-313)."""
    bt = pr.dense_qp(1, 6, 10, 2, seed=21)
    st = bt.structure
    args = dict(eps=1e-12, max_iters=400000)
    x, y, s, status, _ = orc.solve_batch(st, bt.A_vals, bt.b, bt.c, bt.P_vals, **args)
    rng = np.random.default_rng(8)
    dx, dy = rng.standard_normal(x.shape), rng.standard_normal(y.shape)
    dA, dP, db, dc, _ = orc.vjp_batch(st, bt.A_vals, bt.b, bt.c, x, y, s, dx, dy, bt.P_vals, lsqr_precond=1, lsqr_iter_lim=10000)

    def loss(Av, Pv, b, c):
        xx, yy, _, stt, _ = orc.solve_batch(st, Av, b, c, Pv, **args)
        assert stt[0] == 1
        return float(xx[0] @ dx[0] + yy[0] @ dy[0])

    h = 1e-6
    for arr, grad, which in ((bt.b, db, "b"), (bt.c, dc, "c"), (bt.A_vals, dA, "A"), (bt.P_vals, dP, "P")):
        idxs = rng.choice(arr.shape[1], size=min(6, arr.shape[1]), replace=False)
        for k in idxs:
            vals = {"A": bt.A_vals.copy(), "P": bt.P_vals.copy(), "b": bt.b.copy(), "c": bt.c.copy()}
            vals[which][0, k] += h
            lp = loss(vals["A"], vals["P"], vals["b"], vals["c"])
            vals[which][0, k] -= 2 * h
            lm = loss(vals["A"], vals["P"], vals["b"], vals["c"])
            fd = (lp - lm) / (2 * h)
            assert abs(fd - grad[0, k]) <= 1e-4 + 1e-3 * abs(fd), (which, k, fd, grad[0, k])


def _fd_check(make, params, which, args, dx, dy, grads, atol=1e-4, rtol=1e-3, h=1e-6):
    """Central differences of loss = <x, dx> + <y, dy> in every entry of `params` (a flat vector the builder `make`
    turns into a one-instance batch) against the adjoint's gradient wrt that entry -- the reference's gradcheck
    tolerances (atol 1e-4, rtol 1e-3)."""
    for k in range(params.size):
        vals = []
        for sgn in (+1, -1):
            p = params.copy(); p[k] += sgn * h
            bt = make(p)
            x, y, _, stt, _ = orc.solve_batch(bt.structure, bt.A_vals, bt.b, bt.c, bt.P_vals, **args)
            assert stt[0] == 1
            vals.append(float(x[0] @ dx + y[0] @ dy))
        fd = (vals[0] - vals[1]) / (2 * h)
        assert abs(fd - grads[k]) <= atol + rtol * abs(fd), (which, k, fd, grads[k])


@pytest.mark.parametrize("precond", [0, 1])
def test_psd_adjoint_matches_finite_differences(precond):
    """The reference's PSD gradcheck (tests/test_torch.py:233-248: min tr(CX), tr X = 1, X >> 0 at a well-conditioned
    C, atol 1e-4 / rtol 1e-3) restated on the oracle: this pins the PSD projection Jacobian *inside* the adjoint
    independently of any other implementation."""
    C0 = np.array([[2.0, 0.5, 0.1], [0.5, 3.0, 0.2], [0.1, 0.2, 1.5]])
    iu = np.triu_indices(3)

    def make(p):   # p = the 6 free entries of the symmetric parameter
        C = np.zeros((3, 3)); C[iu] = p; C = C + C.T - np.diag(np.diag(C))
        return ref_sdp_batch([C])

    p0 = C0[iu].copy()
    bt = make(p0)
    st = bt.structure
    args = dict(eps=1e-12, max_iters=400000)
    x, y, s, status, _ = orc.solve_batch(st, bt.A_vals, bt.b, bt.c, **args)
    assert status[0] == 1
    lam, V = np.linalg.eigh(C0)
    assert np.abs(pr.svec_to_mat(x[0], 3) - np.outer(V[:, 0], V[:, 0])).max() < 1e-8   # analytic optimum: projector on the min eigenvector
    rng = np.random.default_rng(4)
    dx, dy = rng.standard_normal(st.n), rng.standard_normal(st.m)
    dA, dP, db, dc, _ = orc.vjp_batch(st, bt.A_vals, bt.b, bt.c, x, y, s, dx[None], dy[None], lsqr_precond=precond, lsqr_iter_lim=20000,
                                      lsqr_atol=1e-12, lsqr_btol=1e-12)
    # chain rule svec(C) -> free entries of C: diagonal entries map 1:1, an off-diagonal C_ij feeds sqrt2 * C_ij
    svec_pos = {(0, 0): 0, (1, 0): 1, (2, 0): 2, (1, 1): 3, (2, 1): 4, (2, 2): 5}
    grads = np.array([dc[0, svec_pos[(max(i, j), min(i, j))]] * (1.0 if i == j else np.sqrt(2.0)) for i, j in zip(*iu)])
    _fd_check(make, p0, "C", args, dx, dy, grads)
    # and the data the reference never perturbs (b: the trace level)
    def make_b(p):
        bt2 = make(p0); bt2.b[0, 0] = p[0]; return bt2
    _fd_check(make_b, np.array([1.0]), "b", args, dx, dy, np.array([db[0, 0]]))


@pytest.mark.parametrize("precond", [0, 1])
def test_soc_adjoint_matches_finite_differences(precond):
    """The reference's SOC gradcheck (tests/test_dual_variables.py:346-369: min c'x + 0.1||x||^2, ||x|| <= t, output =
    sum of the SOC dual, parameters c and t, atol 1e-4 / rtol 1e-3) restated on the oracle."""
    p0 = np.array([0.5, 0.3, -0.2, 2.0])
    make = lambda p: ref_soc_batch([p[:3]], [p[3]])  # noqa: E731
    bt = make(p0)
    st = bt.structure
    args = dict(eps=1e-12, max_iters=400000)
    x, y, s, status, _ = orc.solve_batch(st, bt.A_vals, bt.b, bt.c, bt.P_vals, **args)
    assert status[0] == 1
    assert abs(np.linalg.norm(x[0]) - 2.0) < 1e-8 and np.abs(x[0] / np.linalg.norm(x[0]) + p0[:3] / np.linalg.norm(p0[:3])).max() < 1e-8
    for dx, dy in ((np.zeros(3), np.ones(4)), (np.array([1.0, -2.0, 0.5]), np.array([0.3, -1.0, 2.0, 0.7]))):
        dA, dP, db, dc, _ = orc.vjp_batch(st, bt.A_vals, bt.b, bt.c, x, y, s, dx[None], dy[None], bt.P_vals, lsqr_precond=precond,
                                          lsqr_iter_lim=20000, lsqr_atol=1e-12, lsqr_btol=1e-12)
        _fd_check(make, p0, "c,t", args, dx, dy, np.concatenate([dc[0], db[0, :1]]))


def test_qp_in_the_reference_soc_form_matches_the_native_quadratic_form():
    """The reference's DIFFCP path never sees P: cvxpy turns 1/2 x'Px into an epigraph variable and one SOC of size
    n + 2 (_quad_form_dpp.py:29-32, tests/test_torch.py:1005-1020).  Same optimum, same derivative of the solution with
    respect to the data both forms share (b and c of the original rows), through two different programs."""
    bt = pr.dense_qp(3, 8, 14, 3, seed=6)
    bs = pr.qp_as_socp(bt)
    args = dict(eps=1e-11, max_iters=400000)
    x, y, s, st1, _ = orc.solve_batch(bt.structure, bt.A_vals, bt.b, bt.c, bt.P_vals, **args)
    xs, ys, ss, st2, _ = orc.solve_batch(bs.structure, bs.A_vals, bs.b, bs.c, None, **args)
    assert (st1 == 1).all() and (st2 == 1).all()
    n, m = bt.structure.n, bt.structure.m
    assert np.abs(xs[:, :n] - x).max() < 1e-8 and np.abs(ys[:, :m] - y).max() < 1e-7
    assert np.abs(xs[:, n] - 0.5 * np.einsum("bi,bi->b", x, np.stack([bt.P_dense(i) @ x[i] for i in range(bt.B)]))).max() < 1e-8   # t* = 1/2 x'Px
    rng = np.random.default_rng(3)
    dx, dy = rng.standard_normal(x.shape), rng.standard_normal(y.shape)
    g1 = orc.vjp_batch(bt.structure, bt.A_vals, bt.b, bt.c, x, y, s, dx, dy, bt.P_vals, lsqr_precond=1, lsqr_iter_lim=50000, lsqr_atol=1e-12, lsqr_btol=1e-12)
    dxs = np.concatenate([dx, np.zeros((bt.B, 1))], axis=1)
    dys = np.concatenate([dy, np.zeros((bt.B, n + 2))], axis=1)
    g2 = orc.vjp_batch(bs.structure, bs.A_vals, bs.b, bs.c, xs, ys, ss, dxs, dys, None, lsqr_precond=1, lsqr_iter_lim=50000, lsqr_atol=1e-12, lsqr_btol=1e-12)
    assert rel_err(g2[2][:, :m], g1[2]) < 1e-5 and rel_err(g2[3][:, :n], g1[3]) < 1e-5   # db, dc of the shared rows / columns


# ----------------------------------------------------------------------------- golden fixtures
@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_reproduces_golden_fixtures(name):
    bt, g = load_golden(name)
    st = bt.structure
    x, y, s, status, _ = orc.solve_batch(st, bt.A_vals, bt.b, bt.c, bt.P_vals, eps=1e-9, max_iters=200000, acceleration_lookback=0)
    assert (status == 1).all()
    assert np.abs(x - g["x"]).max() < 1e-9 and np.abs(y - g["y"]).max() < 1e-9 and np.abs(s - g["s"]).max() < 1e-9
    # Anderson acceleration (SCS default) changes the path, not the destination
    x2, y2, s2, status2, _ = orc.solve_batch(st, bt.A_vals, bt.b, bt.c, bt.P_vals, eps=1e-9, max_iters=200000)
    assert (status2 == 1).all() and np.abs(x2 - g["x"]).max() < 2e-6 * max(1.0, np.abs(g["x"]).max())
    dA, dP, db, dc, _ = orc.vjp_batch(st, bt.A_vals, bt.b, bt.c, g["x"], g["y"], g["s"], g["dx"], g["dy"], bt.P_vals,
                                      lsqr_precond=1, lsqr_iter_lim=100000)
    assert rel_err(dA, g["dA"]) < 1e-9 and rel_err(db, g["db"]) < 1e-9 and rel_err(dc, g["dc"]) < 1e-9


# ----------------------------------------------------------------------------- exponential cone
def test_exp_cone_projection_certificate_and_jacobian():
    """Moreau certificate (p in K, v - p in K^o, p'(v - p) = 0) + finite-difference Jacobian."""
    st_ed = Structure(1, 3, np.arange(4), np.zeros(3, np.int32), ConeSpec(ed=1))   # K* = K_exp
    st_ep = Structure(1, 3, np.arange(4), np.zeros(3, np.int32), ConeSpec(ep=1))   # K* = dual exp cone

    def in_K(v, tol):
        r, s_, t_ = v
        return (s_ > 0 and s_ * np.exp(min(r / s_, 700)) <= t_ + tol) or (r <= tol and abs(s_) <= tol and t_ >= -tol)

    def in_Kstar(u, tol):
        a, b_, c_ = u
        return (a < 0 and -a * np.exp(min(b_ / a, 700)) <= np.e * c_ + tol) or (abs(a) <= tol and b_ >= -tol and c_ >= -tol)

    rng = np.random.default_rng(2)
    jac_bad = 0
    for _ in range(1500):
        v = rng.standard_normal(3) * rng.choice([0.1, 1, 10, 100])
        sc = max(1.0, np.abs(v).max())
        p = orc.proj_dual_cone(st_ed, v)
        assert in_K(p, 1e-6 * sc) and in_Kstar(-(v - p), 1e-6 * sc) and abs(p @ (v - p)) <= 1e-6 * sc * sc
        q = orc.proj_dual_cone(st_ep, v)
        assert in_Kstar(q, 1e-6 * sc) and in_K(-(v - q), 1e-6 * sc) and abs(q @ (v - q)) <= 1e-6 * sc * sc
        dv = rng.standard_normal(3)
        h = 1e-6 * sc
        fd = (orc.proj_dual_cone(st_ed, v + h * dv) - orc.proj_dual_cone(st_ed, v - h * dv)) / (2 * h)
        jac_bad += np.abs(fd - orc.dproj_dual_cone(st_ed, v, dv)).max() > 1e-4 * max(1, np.abs(fd).max())
    assert jac_bad <= 3  # central differences straddle a kink of the projection now and then


def test_exp_cone_program_matches_scipy_minimize():
    """min sum exp(a_i'x + d_i) + c'x + lam/2||x||^2 solved as a cone program vs a smooth solver."""
    bt = pr.exp_sum(3, p=4, k=7, seed=3)
    st, aux = bt.structure, bt.aux
    x, y, s, status, _ = orc.solve_batch(st, bt.A_vals, bt.b, bt.c, bt.P_vals, eps=1e-9, max_iters=200000)
    assert (status == 1).all()
    for i in range(bt.B):
        a, d, cx = aux["a"][i], aux["d"][i], bt.c[i, : aux["p"]]
        f = lambda z: np.exp(a @ z + d).sum() + cx @ z + 0.5 * aux["lam"] * z @ z  # noqa: E731
        g = lambda z: a.T @ np.exp(a @ z + d) + cx + aux["lam"] * z  # noqa: E731
        res = sopt.minimize(f, np.zeros(aux["p"]), jac=g, method="BFGS", options={"gtol": 1e-12})
        assert np.abs(x[i, : aux["p"]] - res.x).max() < 1e-6
        assert np.abs(x[i, aux["p"]:] - np.exp(a @ res.x + d)).max() < 1e-6
    # adjoint against finite differences on b (the d_i shifts)
    rng = np.random.default_rng(0)
    dx, dy = rng.standard_normal(x.shape), np.zeros_like(y)
    dA, dP, db, dc, _ = orc.vjp_batch(st, bt.A_vals, bt.b, bt.c, x, y, s, dx, dy, bt.P_vals, lsqr_precond=1, lsqr_iter_lim=20000)
    h = 1e-6
    for k in (0, 3, 6):
        bp, bm = bt.b.copy(), bt.b.copy()
        bp[:, k] += h; bm[:, k] -= h
        xp = orc.solve_batch(st, bt.A_vals, bp, bt.c, bt.P_vals, eps=1e-11, max_iters=400000)[0]
        xm = orc.solve_batch(st, bt.A_vals, bm, bt.c, bt.P_vals, eps=1e-11, max_iters=400000)[0]
        fd = ((xp - xm) * dx).sum(1) / (2 * h)
        assert np.abs(fd - db[:, k]).max() <= 1e-4 + 1e-3 * np.abs(fd).max(), (k, fd, db[:, k])
