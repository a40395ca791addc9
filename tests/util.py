"""Shared helpers for the test-suite."""
import os

import numpy as np

from cvxpylayers_b200.problems import Batch
from cvxpylayers_b200.structure import ConeSpec, Structure

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_CASES = ["qp_c1", "qp_eq", "lp_dense", "socp", "sdp"]


def load_golden(name: str):
    z = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    cones = ConeSpec(z=int(z["cone_z"]), l=int(z["cone_l"]), q=[int(v) for v in z["cone_q"]], s=[int(v) for v in z["cone_s"]])
    hasP = z["P_indices"].size > 0
    st = Structure(int(z["n"]), int(z["m"]), z["A_indptr"], z["A_indices"], cones,
                   z["P_indptr"] if hasP else None, z["P_indices"] if hasP else None)
    bt = Batch(st, z["A_vals"], z["b"], z["c"], z["P_vals"] if hasP else None, name=name)
    return bt, {k: z[k] for k in ("x", "y", "s", "dx", "dy", "dA", "dP", "db", "dc")}


def rel_err(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


# ----------------------------------------------------------------------------- the reference's own gradcheck programs
def ref_sdp_batch(C_list):
    """``min tr(C X) s.t. tr(X) = 1, X >> 0`` for 3 x 3 symmetric C -- the program of the reference's PSD
    gradcheck (``/root/reference/tests/test_torch.py:233-248``) written directly in solver form with x = svec(X)
    (lower triangle, column-major, off-diagonals * sqrt 2): one zero-cone row for the trace, -x + s = 0 with s in the
    PSD cone.  The optimum is the rank-one projector on the smallest eigenvector of C and is strictly complementary
    when that eigenvalue is simple, so the solution map is differentiable.  One instance per C in ``C_list``."""
    from cvxpylayers_b200.problems import mat_to_svec

    k, n = 3, 6
    diag_idx = [0, 3, 5]   # svec positions of X_00, X_11, X_22
    indptr = [0, 3] + [3 + i + 1 for i in range(n)]
    indices = diag_idx + list(range(n))
    st = Structure(n, 1 + n, np.asarray(indptr, np.int32), np.asarray(indices, np.int32), ConeSpec(z=1, s=[k]))
    B = len(C_list)
    A_vals = np.tile(np.concatenate([np.ones(3), -np.ones(n)]), (B, 1))
    b = np.zeros((B, 1 + n)); b[:, 0] = 1.0
    c = np.stack([mat_to_svec(np.asarray(C, dtype=float)) for C in C_list])
    return Batch(st, A_vals, b, c, None, name="ref_sdp")


def ref_soc_batch(c_list, t_list):
    """``min c'x + 0.1 ||x||^2 s.t. ||x|| <= t`` (n = 3) -- the program of the reference's SOC gradcheck
    (``/root/reference/tests/test_dual_variables.py:346-369``) in solver form: P = 0.2 I, one SOC of size 4 with
    s = (t, x).  Outputs of the reference's check: the SOC dual (sum), parameters c and t."""
    n = 3
    indptr = [0, 0, 1, 2, 3]
    st = Structure(n, 4, np.asarray(indptr, np.int32), np.arange(3, dtype=np.int32), ConeSpec(q=[4]),
                   np.arange(n + 1, dtype=np.int32), np.arange(n, dtype=np.int32))
    B = len(c_list)
    A_vals = -np.ones((B, 3))
    b = np.zeros((B, 4)); b[:, 0] = np.asarray(t_list, dtype=float)
    return Batch(st, A_vals, b, np.asarray(c_list, dtype=float), np.full((B, n), 0.2), name="ref_soc")
