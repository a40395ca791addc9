"""Independent NumPy/SciPy cross-checks for the C oracle -- TEST INFRASTRUCTURE.

* ``kkt_residuals``: solver-independent optimality certificate = SCS's termination
  quantities on the original data.
* ``vjp_dense``: diffcp's adjoint_derivative restated with an explicit dense M and
  ``scipy.sparse.linalg.lsqr`` -- the routine diffcp's lsqr.cpp ports --
  for zero / nonneg / SOC / PSD cones, with the quadratic-objective extension.
Used only to pin oracle/cone_oracle.c; the reference call sites being restated are
``src/cvxpylayers/interfaces/diffcp_if.py:365`` (forward) and ``:86`` (adjoint).
"""
from __future__ import annotations

import numpy as np
from scipy.sparse.linalg import lsqr

from cvxpylayers_b200.problems import mat_to_svec, proj_dual_cone, svec_to_mat
from cvxpylayers_b200.structure import ConeSpec


def kkt_residuals(A, P, b, c, x, y, s):
    Ax = A @ x
    Px = P @ x if P is not None else np.zeros_like(x)
    ATy = A.T @ y
    rp = np.abs(Ax + s - b).max()
    rd = np.abs(Px + ATy + c).max()
    xPx, ctx, bty = x @ Px, c @ x, b @ y
    gap = abs(xPx + ctx + bty)
    return dict(rp=rp, rd=rd, gap=gap,
                tp=max(np.abs(Ax).max(), np.abs(s).max(), np.abs(b).max()),
                td=max(np.abs(Px).max(), np.abs(ATy).max(), np.abs(c).max()),
                tg=max(abs(xPx), abs(ctx), abs(bty)))


def is_converged(r, eps_abs, eps_rel, slack=1.0):
    return (r["rp"] <= slack * (eps_abs + eps_rel * r["tp"]) and r["rd"] <= slack * (eps_abs + eps_rel * r["td"])
            and r["gap"] <= slack * (eps_abs + eps_rel * r["tg"]))


def dproj_matrix(v: np.ndarray, cones: ConeSpec) -> np.ndarray:
    """Dense Jacobian of Pi_{K*} at v."""
    m = v.size
    D = np.zeros((m, m))
    off = 0
    for i in range(cones.z):
        D[i, i] = 1.0
    off = cones.z
    for i in range(off, off + cones.l):
        D[i, i] = 1.0 if v[i] > 0 else 0.0
    off += cones.l
    for q in cones.q:
        t, x = v[off], v[off + 1 : off + q]
        nx = np.linalg.norm(x)
        if nx <= t:
            J = np.eye(q)
        elif nx <= -t:
            J = np.zeros((q, q))
        else:
            J = np.zeros((q, q))
            J[0, 0] = nx
            J[0, 1:] = x
            J[1:, 0] = x
            J[1:, 1:] = (t + nx) * np.eye(q - 1) - t * np.outer(x, x) / nx**2
            J /= 2 * nx
        D[off : off + q, off : off + q] = J
        off += q
    for k in cones.s:
        sz = k * (k + 1) // 2
        X = svec_to_mat(v[off : off + sz], k)
        lam, V = np.linalg.eigh(X)
        Bm = np.zeros((k, k))
        for i in range(k):
            for j in range(k):
                li, lj = lam[i], lam[j]
                if li > 0 and lj > 0:
                    Bm[i, j] = 1.0
                elif li <= 0 and lj <= 0:
                    Bm[i, j] = 0.0
                else:
                    lp, ln = (li, lj) if li > 0 else (lj, li)
                    Bm[i, j] = lp / (lp - ln)
        for e in range(sz):
            ev = np.zeros(sz)
            ev[e] = 1.0
            dX = svec_to_mat(ev, k)
            D[off : off + sz, off + e] = mat_to_svec(V @ (Bm * (V.T @ dX @ V)) @ V.T)
        off += sz
    return D


def vjp_dense(A, P, b, c, x, y, s, dx, dy, cones: ConeSpec, atol=1e-8, btol=1e-8, conlim=1e8, exact=False):
    """-> dA (dense m x n), dP (dense n x n, gradient wrt a full unsymmetric P), db, dc, r"""
    m, n = A.shape
    N = n + m + 1
    v = y - s
    piy = proj_dual_cone(v, cones)
    D = dproj_matrix(v, cones)
    Pm = np.zeros((n, n)) if P is None else P
    Px = Pm @ x
    DQ = np.zeros((N, N))
    DQ[:n, :n] = Pm
    DQ[:n, n : n + m] = A.T
    DQ[:n, -1] = c
    DQ[n : n + m, :n] = -A
    DQ[n : n + m, -1] = b
    DQ[-1, :n] = -(2 * Px + c)
    DQ[-1, n : n + m] = -b
    DQ[-1, -1] = x @ Px
    Dpi = np.eye(N)
    Dpi[n : n + m, n : n + m] = D
    M = (DQ - np.eye(N)) @ Dpi + np.eye(N)
    dz = np.concatenate([dx, D.T @ dy, [-(x @ dx + y @ dy)]])
    if np.allclose(dz, 0):
        r = np.zeros(N)
    elif exact:
        r = np.linalg.lstsq(M.T, dz, rcond=None)[0]
    else:
        r = lsqr(M.T, dz, atol=atol, btol=btol, conlim=conlim, iter_lim=2 * N)[0]
    rx, ry, rt = r[:n], r[n : n + m], r[-1]
    dA = np.outer(ry, x) - np.outer(piy, rx)
    db = piy * rt - ry
    dc = x * rt - rx
    dP = np.outer(rt * x - rx, x)
    return dA, dP, db, dc, r
