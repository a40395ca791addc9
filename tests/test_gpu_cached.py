"""Cached set-up (second half): equilibration + factorisation kept across calls while A and P do not change
(the reference's template: the one-time ``setup()`` under ``PA_is_constant``, interfaces/moreau_if.py:233-256,316-320).
The cached path must be the SAME algorithm minus the recomputation: bit-identical to the uncached warm solve whenever the
cached scale is the initial one, and equal to the oracle started at the cached scale otherwise."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from cvxpylayers_b200 import problems as pr
from cvxpylayers_b200.engine import Engine, make_settings
from cvxpylayers_b200.interface import B200_ctx, _CvxpyLayer
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def _t(a, dev):
    return torch.tensor(np.ascontiguousarray(a), device=dev)


def _same(a, b):
    return all(torch.equal(getattr(a, k), getattr(b, k)) for k in ("x", "y", "s", "status", "iters"))


@pytest.mark.parametrize("shape", [(100, 200, 50), (80, 200, 40), (75, 190, 30)])   # the compile-time geometry and two runtime ones
def test_cached_setup_is_the_uncached_algorithm(cuda_device, shape):
    n, m, z = shape
    B, dev = 96, cuda_device
    bt = pr.dense_qp(B, n, m, z, seed=21)
    st = bt.structure
    eng = Engine(st, dev)
    assert "register-tiled" in eng.path_info()["fwd"]
    A, P, b, c = _t(bt.A_vals, dev), _t(bt.P_vals, dev), _t(bt.b, dev), _t(bt.c, dev)
    args = dict(eps=1e-6, max_iters=100000)
    S = make_settings(args)
    cache = eng.new_cache(B)
    assert cache is not None and cache.numel() * 8 == eng.cache_bytes(B)
    stride = cache.numel() // B
    plain = eng.solve(A, b, c, P, S)
    fill = eng.solve(A, b, c, P, S, cache=cache, reuse=False)
    assert _same(plain, fill)                                        # writing the cache does not touch the algorithm
    hdr = cache.view(B, stride)[:, :3].cpu().numpy()
    assert (hdr[:, 1] == 1.0).all() and (hdr[:, 2] == S.rho_x).all()
    fresh = eng.solve(A, b, c, P, S, cache=eng.new_cache(B), reuse=True)
    assert _same(plain, fresh)                                       # reuse on a zero-filled cache rebuilds every record

    rng = np.random.default_rng(3)
    b2 = bt.b + 1e-3 * rng.standard_normal(bt.b.shape)
    c2 = bt.c + 1e-3 * rng.standard_normal(bt.c.shape)
    b2t, c2t = _t(b2, dev), _t(c2, dev)
    warm = (plain.x.clone(), plain.y.clone(), plain.s.clone())
    ref_w = eng.solve(A, b2t, c2t, P, S, warm=warm)                   # uncached, warm
    got = eng.solve(A, b2t, c2t, P, S, warm=warm, cache=cache, reuse=True)
    assert int((got.status == 1).sum()) == B
    kept = hdr[:, 0] == S.scale                                        # records whose factorisation is at the initial scale
    assert kept.any()
    ki = torch.tensor(np.nonzero(kept)[0], device=dev)
    for k in ("x", "y", "s", "iters"):
        assert torch.equal(getattr(got, k)[ki], getattr(ref_w, k)[ki]), k   # bit-identical: same Kinv, same E / D, same loop
    # re-scaled records: the solve starts at the cached scale; the oracle told the same thing agrees
    moved = np.nonzero(~kept)[0][:6]
    for i in moved:
        xo, yo, so, sto, ito = orc.solve_batch(st, bt.A_vals[i:i + 1], b2[i:i + 1], c2[i:i + 1], bt.P_vals[i:i + 1],
                                               warm=tuple(w[i:i + 1].cpu().numpy() for w in warm), scale=float(hdr[i, 0]), **args)
        assert sto[0] == 1 and abs(int(got.iters[i]) - int(ito[0])) <= 25
        assert np.abs(got.x[i].cpu().numpy() - xo[0]).max() < 1e-5
    # whatever the route, the answer is the optimum of the new data (cold oracle, tight eps): as close to it as the uncached solve
    xc, yc, sc, stc, _ = orc.solve_batch(st, bt.A_vals, b2, c2, bt.P_vals, eps=1e-9, max_iters=200000)
    assert (stc == 1).all()
    for got_, ref_, opt in ((got.x, ref_w.x, xc), (got.y, ref_w.y, yc)):
        e_got, e_ref = np.abs(got_.cpu().numpy() - opt).max(), np.abs(ref_.cpu().numpy() - opt).max()
        assert e_got < 1e-3 and e_got <= 2.0 * e_ref + 1e-6, (e_got, e_ref)
    # a third call: records refreshed by re-scalings of the second call are valid too
    third = eng.solve(A, b2t, c2t, P, S, warm=got, cache=cache, reuse=True)
    assert int((third.status == 1).sum()) == B and int(third.iters.max()) <= 25


def test_cached_setup_without_quadratic_term(cuda_device):
    """An LP of the register-tiled shape (no P): the cached path has nothing to load but Kinv."""
    B, dev = 48, cuda_device
    bt = pr.dense_qp(B, 60, 160, 20, seed=4, with_P=False)
    st = bt.structure
    eng = Engine(st, dev)
    A, b, c = _t(bt.A_vals, dev), _t(bt.b, dev), _t(bt.c, dev)
    S = make_settings(dict(eps=1e-6, max_iters=200000))
    cache = eng.new_cache(B)
    if cache is None:
        pytest.skip("this LP shape does not run the register-tiled kernel")
    one = eng.solve(A, b, c, None, S, cache=cache, reuse=False)
    two = eng.solve(A, b, c, None, S, warm=one, cache=cache, reuse=True)
    ref = eng.solve(A, b, c, None, S, warm=one)
    hdr = cache.view(B, -1)[:, 0].cpu().numpy()
    ki = torch.tensor(np.nonzero(hdr == S.scale)[0], device=dev)
    assert torch.equal(two.x[ki], ref.x[ki]) and torch.equal(two.iters[ki], ref.iters[ki])


def test_structures_without_a_cached_path_say_so(cuda_device):
    bt = pr.socp_portfolio(4, seed=1)
    eng = Engine(bt.structure, cuda_device)
    assert eng.cache_bytes(4) == 0 and eng.new_cache(4) is None
    dev = cuda_device
    with pytest.raises(Exception, match="cache"):
        eng.solve(_t(bt.A_vals, dev), _t(bt.b, dev), _t(bt.c, dev), _t(bt.P_vals, dev) if bt.P_vals is not None else None,
                  make_settings({}), cache=torch.zeros(64, dtype=torch.float64, device=dev), reuse=True)


def test_layer_reuses_the_setup_when_asked(cuda_device):
    """{"reuse_setup": True, "warm_start": True}: the training-loop configuration.  Same results as the plain layer; the cache
    is filled by the first call and valid afterwards; forward + backward still agree with the oracle's gradient."""
    dev = cuda_device
    bt = pr.dense_qp(64, 100, 200, 50, seed=8)
    st = bt.structure
    bd = pr.to_boundary(bt)

    def layer(**opt):
        ctx = B200_ctx((st.P_indices, st.P_indptr, (st.n, st.n)), (bd.con_indices, bd.con_ptr, bd.shape), bd.dims, options=opt)
        return ctx, SimpleNamespace(solver_ctx=ctx)

    base = dict(eps=1e-7, max_iters=100000, lsqr_precond=2)
    ctx_c, cl_c = layer(**base, reuse_setup=True, warm_start=True)
    ctx_p, cl_p = layer(**base, warm_start=True)
    A, P = _t(bd.A_eval, dev), _t(bd.P_eval, dev)
    rng = np.random.default_rng(0)
    q = bd.q_eval.copy()
    for step in range(3):
        outs = []
        for cl in (cl_c, cl_p):
            qt = _t(q, dev).requires_grad_(True)
            primal, dual, _, _ = _CvxpyLayer.apply(P, qt, A, cl, {}, True, None)
            (primal.square().sum() + dual.sum()).backward()
            outs.append((primal.detach(), dual.detach(), qt.grad.clone()))
        for a_, b_, tol in zip(outs[0], outs[1], (1e-5, 1e-4, 1e-4)):
            assert float((a_ - b_).abs().max()) < tol * max(1.0, float(b_.abs().max()))
        q[:-1] += 1e-3 * rng.standard_normal(q[:-1].shape)
    cache = ctx_c._setup_cache[(dev, bt.B)]
    assert cache is not None and bool((cache.view(bt.B, -1)[:, 1] == 1.0).all())
    assert not hasattr(ctx_p, "_setup_cache")
