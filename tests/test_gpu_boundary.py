"""GPU tests through the reference-facing boundary and the C-ABI pack kernels, plus the committed
golden fixtures and size-independent properties at the headline batch size."""
import warnings
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from cvxpylayers_b200 import problems as pr
from cvxpylayers_b200.engine import Engine, make_settings
from cvxpylayers_b200.interface import B200_ctx, SolverError, _CvxpyLayer
from cvxpylayers_b200.structure import ConeSpec, Structure
from oracle import oracle as orc
from tests.util import GOLDEN_CASES, load_golden, rel_err

pytestmark = pytest.mark.gpu


def _t(a, dev):
    return None if a is None else torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)


def _layer(bt, **opts):
    st = bt.structure
    bd = pr.to_boundary(bt)
    pstruct = (st.P_indices, st.P_indptr, (st.n, st.n)) if st.P_indptr is not None else None
    ctx = B200_ctx(pstruct, (bd.con_indices, bd.con_ptr, bd.shape), bd.dims, options=opts)
    return ctx, bd, SimpleNamespace(solver_ctx=ctx)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_golden_fixtures_through_c_abi(name, cuda_device):
    bt, g = load_golden(name)
    st, dev = bt.structure, cuda_device
    eng = Engine(st, dev)
    sol = eng.solve(_t(bt.A_vals, dev), _t(bt.b, dev), _t(bt.c, dev), _t(bt.P_vals, dev), make_settings({"eps": 1e-9, "max_iters": 200000}))
    assert (sol.status.cpu().numpy() == 1).all()
    scale = max(1.0, np.abs(g["x"]).max())
    assert np.abs(sol.x.cpu().numpy() - g["x"]).max() < 1e-6 * scale
    assert np.abs(sol.y.cpu().numpy() - g["y"]).max() < 1e-6 * max(1.0, np.abs(g["y"]).max())
    dA, dP, db, dc, _ = eng.vjp(_t(bt.A_vals, dev), _t(bt.b, dev), _t(bt.c, dev), _t(g["x"], dev), _t(g["y"], dev), _t(g["s"], dev),
                                _t(g["dx"], dev), _t(g["dy"], dev), _t(bt.P_vals, dev),
                                make_settings({"lsqr_precond": 1, "lsqr_iter_lim": 100000}))
    tol = 1e-4
    assert rel_err(dA.cpu().numpy(), g["dA"]) < tol and rel_err(db.cpu().numpy(), g["db"]) < tol and rel_err(dc.cpu().numpy(), g["dc"]) < tol
    if dP is not None:
        assert rel_err(dP.cpu().numpy(), g["dP"]) < tol


@pytest.mark.parametrize("name,B", [("C1", 5), ("C3", 4), ("C5", 3)])
def test_pack_kernels_roundtrip(name, B, cuda_device):
    """bcone_ingest == the reference's per-instance re-packing (diffcp_if.py:57-68), bcone_emit == :88-94."""
    bt = pr.CONFIGS[name](B=B)
    ctx, bd, _ = _layer(bt)
    eng = ctx.engine(cuda_device)
    A_vals, P_vals, b, c = eng.ingest(_t(bd.A_eval, cuda_device), _t(bd.q_eval, cuda_device), _t(bd.P_eval, cuda_device))
    assert np.array_equal(A_vals.cpu().numpy(), bt.A_vals) and np.array_equal(b.cpu().numpy(), bt.b) and np.array_equal(c.cpu().numpy(), bt.c)
    if bt.P_vals is not None:
        assert np.array_equal(P_vals.cpu().numpy(), bt.P_vals)
    rng = np.random.default_rng(0)
    gA, gb, gc = rng.standard_normal(bt.A_vals.shape), rng.standard_normal(bt.b.shape), rng.standard_normal(bt.c.shape)
    dA_eval, dq_eval, _ = eng.emit(_t(gA, cuda_device), None, _t(gb, cuda_device), _t(gc, cuda_device))
    dA_eval, dq_eval = dA_eval.cpu().numpy(), dq_eval.cpu().numpy()
    st = bt.structure
    assert np.array_equal(-dA_eval[ctx.gather].T, gA)               # con_grad = [-dA.data ; db[b_idx]]
    assert np.array_equal(dA_eval[st.nnzA:].T, gb[:, np.asarray(ctx.b_idx)])
    assert np.array_equal(dq_eval[:-1].T, gc) and not dq_eval[-1].any()  # lin_grad = [dc ; 0]


@pytest.mark.parametrize("host_inputs", [False, True])
def test_layer_forward_backward_matches_oracle(host_inputs, cuda_device):
    bt = pr.dense_qp(6, 10, 20, 3, seed=5)
    st = bt.structure
    args = {"eps": 1e-9, "max_iters": 100000, "lsqr_precond": 1}
    ctx, bd, cl = _layer(bt, **args)
    dev = torch.device("cpu") if host_inputs else cuda_device
    A = torch.tensor(bd.A_eval, device=dev, requires_grad=True)
    q = torch.tensor(bd.q_eval, device=dev, requires_grad=True)
    P = torch.tensor(bd.P_eval, device=dev, requires_grad=True)
    if host_inputs:
        ctx.device = cuda_device
    primal, dual, _, _ = _CvxpyLayer.apply(P, q, A, cl, {}, True, None)
    assert primal.device.type == dev.type and primal.shape == (6, st.n) and dual.shape == (6, st.m)  # outputs live where inputs live
    rng = np.random.default_rng(1)
    dx, dy = rng.standard_normal(primal.shape), rng.standard_normal(dual.shape)
    ((primal * torch.tensor(dx, device=dev)).sum() + (dual * torch.tensor(dy, device=dev)).sum()).backward()
    xo, yo, so, sto, _ = orc.solve_batch(st, bt.A_vals, bt.b, bt.c, bt.P_vals, **args)
    assert np.abs(primal.detach().cpu().numpy() - xo).max() < 1e-6 and np.abs(dual.detach().cpu().numpy() - yo).max() < 1e-6
    gA, gP, gb, gc, _ = orc.vjp_batch(st, bt.A_vals, bt.b, bt.c, xo, yo, so, dx, dy, bt.P_vals, **args)
    dAe, dqe, dPe = A.grad.cpu().numpy(), q.grad.cpu().numpy(), P.grad.cpu().numpy()
    assert rel_err(-dAe[ctx.gather].T, gA) < 1e-4 and rel_err(dAe[st.nnzA:].T, gb) < 1e-4
    assert rel_err(dqe[:-1].T, gc) < 1e-4 and not dqe[-1].any() and rel_err(dPe.T, gP) < 1e-4


def test_unbatched_inputs_keep_the_reference_shapes(cuda_device):
    """1-D inputs = one instance; outputs stay 2-D [1, n] (torch/cvxpylayer.py:247-250), grads are 1-D."""
    bt = pr.dense_qp(1, 6, 9, 2, seed=6)
    ctx, bd, cl = _layer(bt, eps=1e-8)
    A = torch.tensor(bd.A_eval[:, 0], device=cuda_device, requires_grad=True)
    q = torch.tensor(bd.q_eval[:, 0], device=cuda_device, requires_grad=True)
    P = torch.tensor(bd.P_eval[:, 0], device=cuda_device, requires_grad=True)
    primal, dual, _, _ = _CvxpyLayer.apply(P, q, A, cl, {}, True, None)
    assert primal.shape == (1, 6) and dual.shape == (1, 9)
    primal.sum().backward()
    assert A.grad.shape == A.shape and q.grad.shape == q.shape and P.grad.shape == P.shape


def test_no_grad_path_and_solver_args_override(cuda_device):
    bt = pr.dense_qp(2, 6, 9, 2, seed=7)
    ctx, bd, cl = _layer(bt, eps=1e-8)
    A, q, P = _t(bd.A_eval, cuda_device), _t(bd.q_eval, cuda_device), _t(bd.P_eval, cuda_device)
    primal, dual, saved, _ = _CvxpyLayer.apply(P, q, A, cl, {}, False, None)
    assert saved is None and np.abs(primal.cpu().numpy() - bt.x_star).max() < 1e-5
    with warnings.catch_warnings(record=True) as w:  # per-call override reaches the solver (tests/test_torch.py:705-752)
        warnings.simplefilter("always")
        p1, _, _, _ = _CvxpyLayer.apply(P, q, A, cl, {"max_iters": 1}, False, None)
    assert any("Inaccurate" in str(x.message) for x in w)
    assert np.abs(p1.cpu().numpy() - bt.x_star).max() > 1e-2


def test_infeasible_raises_solver_error(cuda_device):
    """tests/test_torch.py:299-316: infeasible / unbounded problems raise."""
    st = Structure.dense(1, 2, ConeSpec(l=2))
    bt = pr.Batch(st, np.array([[1.0, -1.0]]), np.array([[-1.0, -1.0]]), np.array([[0.0]]))
    ctx, bd, cl = _layer(bt)
    with pytest.raises(SolverError):
        _CvxpyLayer.apply(None, _t(bd.q_eval, cuda_device), _t(bd.A_eval, cuda_device), cl, {}, False, None)


def test_headline_batch_properties(cuda_device):
    """BASELINE.json configs[1] at full size (B=4096): every instance satisfies the termination
    criteria on the original data (evaluated with torch on the device), the adjoint is linear in
    (dx, dy) and vanishes for dz = 0."""
    B = 4096
    bt = pr.config_c2(B=B, seed=1)
    st, dev = bt.structure, cuda_device
    eng = Engine(st, dev)
    eps = 1e-4
    A, b, c, P = _t(bt.A_vals, dev), _t(bt.b, dev), _t(bt.c, dev), _t(bt.P_vals, dev)
    sol = eng.solve(A, b, c, P, make_settings({"eps": eps, "max_iters": 10000}))
    assert int((sol.status == 1).sum()) == B
    Am = A.view(B, st.m, st.n)
    iu = torch.triu_indices(st.n, st.n, device=dev)
    Pm = torch.zeros((B, st.n, st.n), dtype=torch.float64, device=dev)
    Pm[:, iu[0], iu[1]] = P
    Pm = Pm + Pm.transpose(1, 2) - torch.diag_embed(torch.diagonal(Pm, dim1=1, dim2=2))
    Ax = torch.bmm(Am, sol.x.unsqueeze(2)).squeeze(2)
    Px = torch.bmm(Pm, sol.x.unsqueeze(2)).squeeze(2)
    ATy = torch.bmm(Am.transpose(1, 2), sol.y.unsqueeze(2)).squeeze(2)
    mx = lambda t: t.abs().amax(dim=1)  # noqa: E731
    rp, rd = mx(Ax + sol.s - b), mx(Px + ATy + c)
    xPx, ctx_, bty = (sol.x * Px).sum(1), (c * sol.x).sum(1), (b * sol.y).sum(1)
    tp = eps + eps * torch.maximum(torch.maximum(mx(Ax), mx(sol.s)), mx(b))
    td = eps + eps * torch.maximum(torch.maximum(mx(Px), mx(ATy)), mx(c))
    tg = eps + eps * torch.maximum(torch.maximum(xPx.abs(), ctx_.abs()), bty.abs())
    assert bool((rp <= 1.001 * tp).all()) and bool((rd <= 1.001 * td).all()) and bool(((xPx + ctx_ + bty).abs() <= 1.001 * tg).all())
    assert bool((sol.s[:, st.cones.z:] >= 0).all()) and bool((sol.y[:, st.cones.z:] >= 0).all())  # cone membership
    # adjoint: linearity and the zero shortcut, on a slice re-solved tightly (at eps = 1e-4 the
    # derivative system is numerically singular and LSQR's early stopping is not a linear map)
    k = 64
    sl = lambda t: t[:k].contiguous()  # noqa: E731
    sol = eng.solve(sl(A), sl(b), sl(c), sl(P), make_settings({"eps": 1e-10, "max_iters": 100000}))
    assert int((sol.status == 1).sum()) == k
    g = torch.Generator(device="cpu").manual_seed(0)
    d1x, d1y = torch.randn((k, st.n), dtype=torch.float64, generator=g).to(dev), torch.randn((k, st.m), dtype=torch.float64, generator=g).to(dev)
    d2x, d2y = torch.randn((k, st.n), dtype=torch.float64, generator=g).to(dev), torch.randn((k, st.m), dtype=torch.float64, generator=g).to(dev)
    tight = make_settings({"lsqr_precond": 1, "lsqr_atol": 1e-13, "lsqr_btol": 1e-13, "lsqr_iter_lim": 5000})
    run = lambda dx, dy: eng.vjp(sl(A), sl(b), sl(c), sl(sol.x), sl(sol.y), sl(sol.s), dx, dy, sl(P), tight)  # noqa: E731
    g1, g2, g12 = run(d1x, d1y), run(d2x, d2y), run(d1x + 2 * d2x, d1y + 2 * d2y)
    for a1, a2, a12 in zip(g1[:4], g2[:4], g12[:4]):
        ref = a1 + 2 * a2
        assert float((a12 - ref).abs().max() / ref.abs().max()) < 1e-5
    z = run(torch.zeros_like(d1x), torch.zeros_like(d1y))
    assert not z[0].any() and not z[2].any() and not z[3].any() and not z[4].any()


def test_pipelined_host_path_equals_plain_path(cuda_device, monkeypatch):
    """Pinned host inputs with a batch worth splitting run as batch slices on two streams (pitched
    H2D/D2H copies overlapped with the kernels); results must equal the single-shot path."""
    import cvxpylayers_b200.interface as itf

    bt = pr.dense_qp(13, 10, 20, 3, seed=9)
    args = {"eps": 1e-9, "max_iters": 100000, "lsqr_precond": 2}
    outs = []
    for chunk in (4, 10**9):
        monkeypatch.setattr(itf, "PIPE_CHUNK", chunk)
        ctx, bd, cl = _layer(bt, **args)
        ctx.device = cuda_device
        A = torch.tensor(bd.A_eval).pin_memory().requires_grad_(True)
        q = torch.tensor(bd.q_eval).pin_memory().requires_grad_(True)
        P = torch.tensor(bd.P_eval).pin_memory().requires_grad_(True)
        primal, dual, saved, _ = _CvxpyLayer.apply(P, q, A, cl, {}, True, None)
        assert saved.items[9] == (chunk == 4)   # the pipelined path was / was not taken
        (primal.sum() + (dual * dual).sum()).backward()
        outs.append([primal.detach().clone(), dual.detach().clone(), A.grad.clone(), q.grad.clone(), P.grad.clone()])
    for a_, b_ in zip(*outs):
        assert a_.device.type == "cpu" and torch.allclose(a_, b_, rtol=1e-9, atol=1e-11)


# ----------------------------------------------------------------------------- fused parameter -> matrix map
def _random_param_maps(bt, bd, rng, P1):
    """Random affine maps p_stack -> (A_eval, q_eval, P_eval) with the three kinds of parameter columns the reference
    produces: exclusive (one matrix entry IS the parameter), shared (a parameter feeds several entries, possibly of
    different tensors) and the constant-1 column (last)."""
    import scipy.sparse as sp

    st = bt.structure
    maps = []
    for rows in (bd.A_eval.shape[0], st.n + 1, st.nnzP):
        if rows == 0:
            maps.append(None)
            continue
        nz_rows, nz_cols, nz_vals = [], [], []
        for r in range(rows):
            k = rng.integers(0, 4)          # 0..3 entries per row (empty rows = structural zeros of the map)
            cols = rng.choice(P1, size=k, replace=False)
            nz_rows += [r] * k
            nz_cols += list(cols)
            nz_vals += list(rng.standard_normal(k))
        maps.append(sp.csr_matrix((nz_vals, (nz_rows, nz_cols)), shape=(rows, P1)))
    return maps


@pytest.mark.parametrize("name,B", [("C1", 70), ("C3", 37), ("C2", 130)])
def test_fused_parameter_map_equals_spmm_then_ingest(name, B, cuda_device):
    """bcone_ingest_params == (A_param @ p_stack, q_param @ p_stack, P_param @ p_stack) followed by bcone_ingest, and
    bcone_emit_params == the transposed products applied to bcone_emit's output (torch/cvxpylayer.py:443-451, :33-37)."""
    bt = pr.CONFIGS[name](B=B)
    st, dev = bt.structure, cuda_device
    ctx, bd, _ = _layer(bt)
    eng = ctx.engine(dev)
    rng = np.random.default_rng(3)
    P1 = 41
    Am, qm, Pm = _random_param_maps(bt, bd, rng, P1)
    # make a few columns exclusive on purpose (single entry in total) and keep the last column as the constant
    eng.set_param_maps(Am, qm, Pm)
    p_stack = rng.standard_normal((P1, B)); p_stack[-1] = 1.0
    A_eval, q_eval = Am @ p_stack, qm @ p_stack
    P_eval = Pm @ p_stack if Pm is not None else None
    ref = eng.ingest(_t(A_eval, dev), _t(q_eval, dev), _t(P_eval, dev))
    got = eng.ingest_params(_t(p_stack, dev))
    for r_, g_ in zip(ref, got):
        if r_ is not None:
            assert float((r_ - g_).abs().max()) <= 1e-13 * max(1.0, float(r_.abs().max()))
    gA, gb, gc = rng.standard_normal(bt.A_vals.shape), rng.standard_normal(bt.b.shape), rng.standard_normal(bt.c.shape)
    gP = rng.standard_normal(bt.P_vals.shape) if bt.P_vals is not None else None
    dA_eval, dq_eval, dP_eval = eng.emit(_t(gA, dev), _t(gP, dev), _t(gb, dev), _t(gc, dev))
    want = Am.T @ dA_eval.cpu().numpy() + qm.T @ dq_eval.cpu().numpy()
    if Pm is not None:
        want = want + Pm.T @ dP_eval.cpu().numpy()
    want[-1] = 0.0   # the constant's row is not a parameter
    dp = eng.emit_params(_t(gA, dev), _t(gP, dev), _t(gb, dev), _t(gc, dev)).cpu().numpy()
    assert np.abs(dp - want).max() <= 1e-12 * max(1.0, np.abs(want).max())


def test_identity_parameter_map_is_the_plain_boundary(cuda_device):
    """Every matrix entry is its own parameter (the headline workload): exclusive columns, plain stores on the way back."""
    import scipy.sparse as sp

    bt = pr.dense_qp(33, 10, 20, 3, seed=8)
    st, dev = bt.structure, cuda_device
    ctx, bd, _ = _layer(bt)
    eng = ctx.engine(dev)
    na, nq, nP = bd.A_eval.shape[0], st.n + 1, st.nnzP
    P1 = na + nq + nP + 1
    eye = lambda rows, off: sp.csr_matrix((np.ones(rows), (np.arange(rows), off + np.arange(rows))), shape=(rows, P1))  # noqa: E731
    eng.set_param_maps(eye(na, 0), eye(nq, na), eye(nP, na + nq))
    p_stack = np.concatenate([bd.A_eval, bd.q_eval, bd.P_eval, np.ones((1, bt.B))])
    got = eng.ingest_params(_t(p_stack, dev))
    assert np.array_equal(got[0].cpu().numpy(), bt.A_vals) and np.array_equal(got[2].cpu().numpy(), bt.b) and np.array_equal(got[3].cpu().numpy(), bt.c)
    assert np.array_equal(got[1].cpu().numpy(), bt.P_vals)
    rng = np.random.default_rng(0)
    gA, gP, gb, gc = (rng.standard_normal(a.shape) for a in (bt.A_vals, bt.P_vals, bt.b, bt.c))
    dA_eval, dq_eval, dP_eval = eng.emit(_t(gA, dev), _t(gP, dev), _t(gb, dev), _t(gc, dev))
    dp = eng.emit_params(_t(gA, dev), _t(gP, dev), _t(gb, dev), _t(gc, dev)).cpu().numpy()
    assert np.array_equal(dp[:na], dA_eval.cpu().numpy()) and np.array_equal(dp[na:na + nq - 1], dq_eval.cpu().numpy()[:-1])
    assert np.array_equal(dp[na + nq:na + nq + nP], dP_eval.cpu().numpy()) and not dp[-1].any()


@pytest.mark.parametrize("fuse", [False, True])
def test_registered_layer_end_to_end_with_quadratic_term(fuse, cuda_device, monkeypatch):
    """CvxpyLayer(problem, ..., solver="B200") through the registered wrappers: construction via the wrapped
    parse_args, forward / backward via the reference's own sequence (fuse=False: sparse products + _CvxpyLayer.apply) or
    the fused path (fuse=True: p_stack straight into the engine), native P passed through from a FULL symmetric pattern.
    Solutions and parameter gradients against the oracle."""
    from cvxpylayers_b200 import interface as itf
    from tests.util import fake_param_prob, install_fake_cvxpylayers

    fake = install_fake_cvxpylayers(monkeypatch)
    bt = pr.dense_qp(5, 8, 14, 3, seed=4)
    st = bt.structure
    problem, params = fake_param_prob(bt)
    itf.register(fuse=fuse)
    args = {"eps": 1e-9, "max_iters": 100000, "lsqr_precond": 1}
    layer = fake.tl.CvxpyLayer(problem, [], [], solver="B200", solver_args=args)
    layer.ctx.solver_ctx.device = cuda_device
    th = [torch.tensor(p, device=cuda_device, requires_grad=True) for p in params]
    primal, dual = layer(*th)
    xo, yo, so, sto, _ = orc.solve_batch(st, bt.A_vals, bt.b, bt.c, bt.P_vals, **args)
    assert (sto == 1).all()
    assert np.abs(primal.detach().cpu().numpy() - xo).max() < 1e-6 and np.abs(dual.detach().cpu().numpy() - yo).max() < 1e-6
    rng = np.random.default_rng(2)
    dx, dy = rng.standard_normal(xo.shape), rng.standard_normal(yo.shape)
    ((primal * torch.tensor(dx, device=cuda_device)).sum() + (dual * torch.tensor(dy, device=cuda_device)).sum()).backward()
    gA, gP, gb, gc, _ = orc.vjp_batch(st, bt.A_vals, bt.b, bt.c, xo, yo, so, dx, dy, bt.P_vals, **args)
    sc = layer.ctx.solver_ctx
    dAe = th[0].grad.cpu().numpy().T            # [nnz_aug, B]: [-dA (boundary order) ; db]
    assert rel_err(-dAe[sc.gather].T, gA) < 1e-4 and rel_err(dAe[st.nnzA:].T, gb) < 1e-4
    assert rel_err(th[1].grad.cpu().numpy(), gc) < 1e-4
    dPe = th[2].grad.cpu().numpy().T            # [n*n, B] rows of the full pattern: upper entries carry the gradient, mirrors 0
    assert rel_err(dPe[sc.gatherP].T, gP) < 1e-4
    mirror = np.setdiff1d(np.arange(st.n * st.n), sc.gatherP)
    assert not dPe[mirror].any()


def test_column_slice_ingest_and_emit_equal_the_whole_tensor_calls(cuda_device):
    """bcone_ingest_pitched / bcone_emit_pitched on column slices of the full boundary tensors (what a sharded or
    pipelined caller uses: no staging copy, no concatenation) == bcone_ingest / bcone_emit on the whole batch."""
    bt = pr.dense_qp(37, 10, 20, 3, seed=10)
    st, dev = bt.structure, cuda_device
    ctx, bd, _ = _layer(bt)
    eng = ctx.engine(dev)
    A_eval, q_eval, P_eval = _t(bd.A_eval, dev), _t(bd.q_eval, dev), _t(bd.P_eval, dev)
    ref = eng.ingest(A_eval, q_eval, P_eval)
    B = bt.B
    f64 = torch.float64
    out = (torch.zeros((B, st.nnzA), dtype=f64, device=dev), torch.zeros((B, st.nnzP), dtype=f64, device=dev),
           torch.zeros((B, st.m), dtype=f64, device=dev), torch.zeros((B, st.n), dtype=f64, device=dev))
    cuts = [(0, 9), (9, 10), (10, 26), (26, 37)]   # odd offsets: the 128-bit fast path must fall back where it is misaligned
    for lo, hi in cuts:
        eng.ingest_cols(A_eval, q_eval, P_eval, lo, hi, out=tuple(o[lo:hi] for o in out))
    for r_, g_ in zip(ref, out):
        assert torch.equal(r_, g_)
    rng = np.random.default_rng(1)
    gA, gP, gb, gc = (_t(rng.standard_normal(a.shape), dev) for a in (bt.A_vals, bt.P_vals, bt.b, bt.c))
    want = eng.emit(gA, gP, gb, gc)
    got = (torch.full_like(want[0], 7.0), torch.full_like(want[1], 7.0), torch.full_like(want[2], 7.0))
    for lo, hi in cuts:
        eng.emit_cols(gA[lo:hi], gP[lo:hi], gb[lo:hi], gc[lo:hi], lo, hi, out=got)
    for w_, g_ in zip(want, got):
        assert torch.equal(w_, g_)


def test_pipelined_host_path_on_the_workspace_backed_kernels(cuda_device, monkeypatch):
    """Two chunks in flight on two streams for a structure whose forward keeps its iterate vectors in a global-memory slab
    (indirect / CG path) and whose backward keeps the LSQR vectors there too: the slabs are per stream (ADVICE round 1:
    a single per-handle slab was shared by concurrent launches), so the pipelined result equals the single-shot one."""
    import cvxpylayers_b200.interface as itf

    bt = pr.sparse_qp(B=6, seed=2)
    args = {"eps": 1e-8, "max_iters": 100000, "lsqr_precond": 1, "lsqr_iter_lim": 20000}
    outs = []
    for chunk in (2, 10**9):
        monkeypatch.setattr(itf, "PIPE_CHUNK", chunk)
        ctx, bd, cl = _layer(bt, **args)
        ctx.device = cuda_device
        info = ctx.engine(cuda_device).path_info()
        assert "slab" in info["fwd"] or "indirect" in info["fwd"]   # a workspace-backed forward path
        A = torch.tensor(bd.A_eval).pin_memory().requires_grad_(True)
        q = torch.tensor(bd.q_eval).pin_memory().requires_grad_(True)
        P = torch.tensor(bd.P_eval).pin_memory().requires_grad_(True)
        primal, dual, saved, _ = _CvxpyLayer.apply(P, q, A, cl, {}, True, None)
        assert saved.items[9] == (chunk == 2)
        (primal.sum() + (dual * dual).sum()).backward()
        outs.append([primal.detach().clone(), dual.detach().clone(), A.grad.clone(), q.grad.clone(), P.grad.clone()])
    # (K = rho I + P + A'R^{-1}A of a sparse A is accumulated with shared-memory / global atomics: the summation order, hence
    #  the last bits of the factor, differ from launch to launch; the solutions agree to the solver tolerance, not bitwise)
    for a_, b_ in zip(*outs):
        assert torch.allclose(a_, b_, rtol=1e-6, atol=1e-8)


def test_warm_start_through_the_layer_matches_the_oracle(cuda_device):
    """{"warm_start": True} re-uses the previous call's solution as the starting point (the rule the
    reference applies for its one warm-startable backend, torch/cvxpylayer.py:464-487).  A training-loop step changes the
    data a little: the warm-started solve takes the oracle's (smaller) iteration count and reaches the same optimum;
    started at its own solution it stops at the first check."""
    bt = pr.dense_qp(40, 100, 200, 50, seed=13)
    st, dev = bt.structure, cuda_device
    args = {"eps": 1e-6, "max_iters": 100000, "warm_start": True}
    ctx, bd, cl = _layer(bt, **args)
    A, q, P = _t(bd.A_eval, dev), _t(bd.q_eval, dev), _t(bd.P_eval, dev)
    eng = ctx.engine(dev)
    p1, d1, _, _ = _CvxpyLayer.apply(P, q, A, cl, {}, False, None)              # cold: nothing cached yet
    it_cold = None
    x0, y0, s0 = (t_.clone() for t_ in ctx._last_solution[(dev, bt.B)])
    rng = np.random.default_rng(1)
    b2 = bt.b + 1e-4 * rng.standard_normal(bt.b.shape)
    c2 = bt.c + 1e-4 * rng.standard_normal(bt.c.shape)
    bt2 = pr.Batch(st, bt.A_vals, b2, c2, bt.P_vals)
    bd2 = pr.to_boundary(bt2)
    p2, d2, _, _ = _CvxpyLayer.apply(P, _t(bd2.q_eval, dev), _t(bd2.A_eval, dev), cl, {}, False, None)   # warm from call 1
    o_args = dict(eps=1e-6, max_iters=100000)
    xw, yw, sw, stw, it_w = orc.solve_batch(st, bt.A_vals, b2, c2, bt.P_vals, warm=(x0.cpu().numpy(), y0.cpu().numpy(), s0.cpu().numpy()), **o_args)
    xc, yc, sc_, stc, it_c = orc.solve_batch(st, bt.A_vals, b2, c2, bt.P_vals, **o_args)
    assert (stw == 1).all() and np.abs(p2.cpu().numpy() - xw).max() < 1e-5 and np.abs(p2.cpu().numpy() - xc).max() < 1e-4
    # iteration counts through the engine directly (the layer does not return them)
    sol_w = eng.solve(_t(bt.A_vals, dev), _t(b2, dev), _t(c2, dev), _t(bt.P_vals, dev), make_settings(o_args), warm=(x0, y0, s0))
    sol_c = eng.solve(_t(bt.A_vals, dev), _t(b2, dev), _t(c2, dev), _t(bt.P_vals, dev), make_settings(o_args))
    iw, ic = sol_w.iters.cpu().numpy(), sol_c.iters.cpu().numpy()
    assert np.abs(iw - it_w).max() <= 25 and np.abs(ic - it_c).max() <= 25
    assert iw.mean() < 0.7 * ic.mean(), (iw.mean(), ic.mean())
    sol_f = eng.solve(_t(bt.A_vals, dev), _t(b2, dev), _t(c2, dev), _t(bt.P_vals, dev), make_settings(o_args), warm=sol_w)
    assert int(sol_f.iters.max()) <= 25 and float((sol_f.x - sol_w.x).abs().max()) < 1e-4   # (two points inside the 1e-6 termination ball)


# ----------------------------------------------------------------------------- layer prologue / epilogue on the device
def _reshape_fortran(array, shape):
    """The reference's helper, restated (torch/cvxpylayer.py:40-56): reshape in column-major order via permutes."""
    x = array.permute(*reversed(range(len(array.shape))))
    return x.reshape(*reversed(shape)).permute(*reversed(range(len(shape))))


def test_device_flatten_and_recover_equal_the_reference_tensor_chains(cuda_device):
    """cvxpylayers_b200.layer_io vs the reference's own tensor-op chains restated here: _flatten_and_batch_params
    (torch/cvxpylayer.py:84-141: expand, Fortran reshape, cat in column order, ones row, transpose; GP log :58-81) and
    _recover_results (:225-282: slices, svec unpacking :143-222, Fortran reshape, GP exp) -- values and gradients."""
    from types import SimpleNamespace

    from cvxpylayers_b200 import layer_io

    dev, B = cuda_device, 7
    g = torch.Generator(device="cpu").manual_seed(3)
    mk = lambda *shape: (torch.rand(shape, dtype=torch.float64, generator=g) + 0.5).to(dev).requires_grad_(True)  # noqa: E731
    params = (mk(B, 3, 2), mk(4), mk(B, 5), mk(2, 2))                   # batched matrix, unbatched vector, batched vector, unbatched matrix
    lctx = SimpleNamespace(batch_sizes=[B, 0, B, 0], user_order_to_col_order=(2, 0, 3, 1), gp=True, gp_log_mask=(False, True, False, True))
    p_dev = layer_io.flatten_and_batch_params(params, lctx, (B,))
    # reference chain
    flat = [None] * 5
    for i, p in enumerate(params):
        q = torch.log(p) if lctx.gp_log_mask[i] else p
        if lctx.batch_sizes[i] == 0:
            q = q.unsqueeze(0).expand((B,) + q.shape)
        flat[lctx.user_order_to_col_order[i]] = _reshape_fortran(q, (B, -1))
    flat[-1] = torch.ones((B, 1), dtype=torch.float64, device=dev)
    p_ref = torch.cat(flat, -1).T
    assert p_dev.shape == p_ref.shape and torch.allclose(p_dev, p_ref, rtol=0, atol=1e-15)
    w = torch.randn(p_ref.shape, dtype=torch.float64, generator=g).to(dev)
    g_dev = torch.autograd.grad((p_dev * w).sum(), params)
    g_ref = torch.autograd.grad((p_ref * w).sum(), params)
    for a_, b_ in zip(g_dev, g_ref):
        assert torch.allclose(a_, b_, rtol=1e-13, atol=1e-13)
    # unbatched call: 1-D p_stack
    p1 = layer_io.flatten_and_batch_params((params[1], params[3]), SimpleNamespace(batch_sizes=[0, 0], user_order_to_col_order=(1, 0), gp=False), ())
    ref1 = torch.cat([_reshape_fortran(params[3].unsqueeze(0), (1, -1)), _reshape_fortran(params[1].unsqueeze(0), (1, -1)), torch.ones((1, 1), dtype=torch.float64, device=dev)], -1).T.reshape(-1)
    assert p1.shape == ref1.shape and torch.equal(p1, ref1)

    # ---- recover ----
    n, m = 20, 15
    primal = torch.randn((B, n), dtype=torch.float64, generator=g).to(dev).requires_grad_(True)
    dual = torch.randn((B, m), dtype=torch.float64, generator=g).to(dev).requires_grad_(True)
    V = lambda **kw: SimpleNamespace(**{"primal": None, "dual": None, **kw})  # noqa: E731
    rctx = SimpleNamespace(gp=False, var_recover=[
        V(primal=slice(2, 8), shape=(2, 3), source="primal", unpack_fn="reshape"),
        V(primal=slice(8, 14), shape=(3, 3), source="primal", unpack_fn="svec_primal"),
        V(dual=slice(4, 10), shape=(3, 3), source="dual", unpack_fn="svec_dual"),
        V(dual=slice(0, 4), shape=(4,), source="dual", unpack_fn="reshape")])
    outs = layer_io.recover_results(primal, dual, rctx, (B,))

    def ref_recover():
        res = []
        for var in rctx.var_recover:
            data = primal[..., var.primal] if var.source == "primal" else dual[..., var.dual]
            if var.unpack_fn == "reshape":
                r = _reshape_fortran(data, (B,) + var.shape)
            else:
                k = var.shape[0]
                if var.unpack_fn == "svec_primal":
                    rows, cols = np.triu_indices(k); sc = None
                else:
                    rr, cc = np.tril_indices(k); o = np.lexsort((rr, cc)); rows, cols = rr[o], cc[o]
                    sc = torch.tensor(np.where(rows == cols, 1.0, 1.0 / np.sqrt(2.0)), device=dev)
                d = data * sc if sc is not None else data
                r = torch.zeros((B, k, k), dtype=torch.float64, device=dev)
                r[..., torch.tensor(rows, device=dev), torch.tensor(cols, device=dev)] = d
                r[..., torch.tensor(cols, device=dev), torch.tensor(rows, device=dev)] = d
            res.append(r)
        return res

    refs = ref_recover()
    ws = [torch.randn(r.shape, dtype=torch.float64, generator=g).to(dev) for r in refs]
    for o_, r_ in zip(outs, refs):
        assert o_.shape == r_.shape and torch.allclose(o_, r_, rtol=0, atol=1e-15)
    gd = torch.autograd.grad(sum((o_ * w_).sum() for o_, w_ in zip(outs, ws)), (primal, dual))
    gr = torch.autograd.grad(sum((r_ * w_).sum() for r_, w_ in zip(refs, ws)), (primal, dual))
    for a_, b_ in zip(gd, gr):
        assert torch.allclose(a_, b_, rtol=1e-13, atol=1e-13)
    # GP: exp on primal variables only
    rctx.gp = True
    og = layer_io.recover_results(primal, dual, rctx, (B,))
    assert torch.allclose(og[0], torch.exp(refs[0])) and torch.allclose(og[2], refs[2])
    gg = torch.autograd.grad((og[0] * ws[0]).sum(), primal)[0]
    assert torch.allclose(gg, torch.autograd.grad((torch.exp(_reshape_fortran(primal[..., 2:8], (B, 2, 3))) * ws[0]).sum(), primal)[0], rtol=1e-13, atol=1e-13)


def test_pageable_host_inputs_take_the_staged_pipeline(cuda_device, monkeypatch):
    """What the reference's CPU path hands over is pageable memory (torch.from_numpy, torch/cvxpylayer.py:21-24).  Those inputs go
    through a ring of pinned staging buffers filled by a background thread; results and gradients must be those of the plain path
    (device-resident inputs), for a batch that is not a multiple of the slice and over several calls (the ring is re-used)."""
    from cvxpylayers_b200 import interface as itf

    monkeypatch.setattr(itf, "PIPE_CHUNK", 8)
    dev = cuda_device
    bt = pr.dense_qp(46, 20, 40, 10, seed=6)
    ctx, bd, cl = _layer(bt, eps=1e-8, max_iters=100000, lsqr_precond=2)
    rng = np.random.default_rng(0)
    for call in range(3):
        A_np = bd.A_eval * (1.0 + 1e-3 * call)
        q_np = bd.q_eval + 1e-3 * call * rng.standard_normal(bd.q_eval.shape)
        outs = []
        for where in ("cpu", "cuda"):
            mk = (lambda a: torch.tensor(a)) if where == "cpu" else (lambda a: _t(a, dev))
            A, q, P = mk(A_np).requires_grad_(True), mk(q_np).requires_grad_(True), mk(bd.P_eval).requires_grad_(True)
            assert where == "cuda" or (not A.is_pinned() and itf._stage_ok(bt.B, A, q, P))
            primal, dual, _, _ = _CvxpyLayer.apply(P, q, A, cl, {}, True, None)
            assert primal.device.type == where
            (primal.square().sum() + dual.sum()).backward()
            outs.append([t_.detach().cpu() for t_ in (primal, dual, A.grad, q.grad, P.grad)])
        for a_, b_ in zip(*outs):
            assert float((a_ - b_).abs().max()) <= 1e-9 * max(1.0, float(b_.abs().max()))
    assert len(ctx.engine(dev)._stager.bufs) == 4 * 3   # slices 10, 9, 9, 9, 9 on 3 slots: (slot 0, 10), (1, 9), (2, 9), (0, 9); x (A, q, P)
