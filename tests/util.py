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
