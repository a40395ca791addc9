"""Generates tests/golden/*.npz -- small fixed instances with the oracle's forward solution and
adjoint, each cross-checked AT GENERATION TIME against solver-independent mathematics (KKT
certificate on the original data; exact dense least-squares solve of diffcp's adjoint system
with NumPy).  The reference itself holds no golden vectors for this path and diffcp/SCS cannot
be imported here, so these fixtures pin the oracle against regressions and give
the GPU tests a data set that does not depend on the generator code.

    python -m oracle.make_golden
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from cvxpylayers_b200 import problems as pr  # noqa: E402
from oracle import np_ref  # noqa: E402
from oracle import oracle as orc  # noqa: E402

CASES = {
    "qp_c1": lambda: pr.dense_qp(3, 10, 20, 0, seed=11),
    "qp_eq": lambda: pr.dense_qp(3, 12, 18, 5, seed=12),
    "lp_dense": lambda: pr.dense_lp(2, 8, 20, seed=13),
    "socp": lambda: pr.socp_portfolio(3, n_assets=8, n_soc=2, k=3, seed=14),
    "sdp": lambda: pr.sdp(2, k=4, n_eq=3, seed=15),
}
# (the fixtures were generated before Anderson acceleration existed in the oracle: plain operator splitting)
ARGS = dict(eps=1e-9, max_iters=200000, acceleration_lookback=0)


def main():
    out = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out, exist_ok=True)
    rng = np.random.default_rng(99)
    for name, make in CASES.items():
        bt = make()
        st = bt.structure
        x, y, s, status, iters = orc.solve_batch(st, bt.A_vals, bt.b, bt.c, bt.P_vals, **ARGS)
        assert (status == 1).all(), (name, status)
        dx, dy = rng.standard_normal(x.shape), rng.standard_normal(y.shape)
        dA, dP, db, dc, its = orc.vjp_batch(st, bt.A_vals, bt.b, bt.c, x, y, s, dx, dy, bt.P_vals, lsqr_precond=1,
                                            lsqr_iter_lim=100000)
        for i in range(bt.B):
            A = bt.A_dense(i)
            P = bt.P_dense(i) if bt.P_vals is not None else None
            r = np_ref.kkt_residuals(A, P, bt.b[i], bt.c[i], x[i], y[i], s[i])
            assert np_ref.is_converged(r, 1e-9, 1e-9, 1.001), (name, i, r)
            if name != "sdp":  # SDP optima here are rank deficient: min-norm solution depends on scaling
                rA, rP, rb, rc, _ = np_ref.vjp_dense(A, P, bt.b[i], bt.c[i], x[i], y[i], s[i], dx[i], dy[i], st.cones, exact=True)
                assert np.abs(db[i] - rb).max() <= 1e-5 * max(1, np.abs(rb).max()), (name, i)
                assert np.abs(dc[i] - rc).max() <= 1e-5 * max(1, np.abs(rc).max()), (name, i)
        np.savez_compressed(
            os.path.join(out, f"{name}.npz"), n=st.n, m=st.m, A_indptr=st.A_indptr, A_indices=st.A_indices,
            P_indptr=st.P_indptr if st.P_indptr is not None else np.zeros(0, np.int32),
            P_indices=st.P_indices if st.P_indices is not None else np.zeros(0, np.int32),
            cone_z=st.cones.z, cone_l=st.cones.l, cone_q=np.asarray(st.cones.q, np.int32), cone_s=np.asarray(st.cones.s, np.int32),
            A_vals=bt.A_vals, P_vals=bt.P_vals if bt.P_vals is not None else np.zeros((bt.B, 0)), b=bt.b, c=bt.c,
            x=x, y=y, s=s, dx=dx, dy=dy, dA=dA, dP=dP if dP is not None else np.zeros((bt.B, 0)), db=db, dc=dc)
        print(name, "ok: iters", iters, "lsqr", its)


if __name__ == "__main__":
    main()
