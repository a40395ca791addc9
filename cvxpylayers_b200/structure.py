"""Problem structure shared by a whole batch.

One ``CvxpyLayer`` fixes one sparsity pattern and one cone spec at construction
(reference: ``src/cvxpylayers/interfaces/diffcp_if.py:114-119`` stores
``A_structure``/``A_shape``/``b_idx``/``dims``; ``moreau_if.py:206-222`` stores the
CSR twin); only the *values* change per call.  This module holds that structure
in the engine's own form: CSR ``A`` (m x n), optional CSR upper-triangular ``P``
(n x n), cones in SCS row order ``z, l, q, s, ep, ed``.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


@dataclass
class ConeSpec:
    """Cone dimensions, keys as produced by cvxpy's ``dims_to_solver_dict``
    (reference ``diffcp_if.py:8,361``; key names visible at ``moreau_if.py:148-154``)."""

    z: int = 0
    l: int = 0  # noqa: E741
    q: list[int] = field(default_factory=list)
    s: list[int] = field(default_factory=list)
    ep: int = 0
    ed: int = 0

    @staticmethod
    def from_dict(d: dict) -> "ConeSpec":
        known = {"z", "l", "q", "s", "ep", "ed", "p", "f"}
        for k in d:
            if k not in known:
                raise ValueError(f"unknown cone key {k!r}")
        if d.get("p"):
            raise NotImplementedError(
                "power cones are out of scope (the reference's own backward is broken for them: "
                "tests/test_dual_variables.py:513-520)"
            )
        return ConeSpec(
            z=int(d.get("z", d.get("f", 0)) or 0),
            l=int(d.get("l", 0) or 0),
            q=[int(v) for v in d.get("q", []) or []],
            s=[int(v) for v in d.get("s", []) or []],
            ep=int(d.get("ep", 0) or 0),
            ed=int(d.get("ed", 0) or 0),
        )

    def to_dict(self) -> dict:
        return {"z": self.z, "l": self.l, "q": list(self.q), "s": list(self.s), "ep": self.ep, "ed": self.ed}

    @property
    def m(self) -> int:
        return self.z + self.l + sum(self.q) + sum(k * (k + 1) // 2 for k in self.s) + 3 * (self.ep + self.ed)


@dataclass
class Structure:
    """Sparsity pattern + cone spec of one batch of problems
    ``min 1/2 x'Px + c'x  s.t.  Ax + s = b, s in K``."""

    n: int
    m: int
    A_indptr: np.ndarray  # int32 [m+1]
    A_indices: np.ndarray  # int32 [nnzA]
    cones: ConeSpec
    P_indptr: np.ndarray | None = None  # int32 [n+1], upper triangle incl. diagonal
    P_indices: np.ndarray | None = None

    def __post_init__(self) -> None:
        self.A_indptr = np.ascontiguousarray(self.A_indptr, dtype=np.int32)
        self.A_indices = np.ascontiguousarray(self.A_indices, dtype=np.int32)
        if self.A_indptr.shape != (self.m + 1,):
            raise ValueError("A_indptr must have m+1 entries")
        if self.P_indptr is not None:
            self.P_indptr = np.ascontiguousarray(self.P_indptr, dtype=np.int32)
            self.P_indices = np.ascontiguousarray(self.P_indices, dtype=np.int32)
            rows = np.repeat(np.arange(self.n), np.diff(self.P_indptr))
            if np.any(self.P_indices < rows):
                raise ValueError("P must be given as its upper triangle (col >= row)")
        if self.cones.m != self.m:
            raise ValueError(f"cone spec covers {self.cones.m} rows, A has {self.m}")
        if self.nnzA and (self.A_indices.min() < 0 or self.A_indices.max() >= self.n):
            raise ValueError("A column index out of range")

    @property
    def nnzA(self) -> int:
        return int(self.A_indices.shape[0])

    @property
    def nnzP(self) -> int:
        return 0 if self.P_indices is None else int(self.P_indices.shape[0])

    @property
    def is_dense_A(self) -> bool:
        """True when the pattern is the full m x n rectangle in row-major order."""
        if self.nnzA != self.m * self.n:
            return False
        return bool(np.array_equal(self.A_indices, np.tile(np.arange(self.n, dtype=np.int32), self.m)))

    @staticmethod
    def dense(n: int, m: int, cones: ConeSpec, with_P: bool = False) -> "Structure":
        indptr = np.arange(0, (m + 1) * n, n, dtype=np.int32)
        indices = np.tile(np.arange(n, dtype=np.int32), m)
        if with_P:
            pptr = np.zeros(n + 1, dtype=np.int32)
            pptr[1:] = np.cumsum(np.arange(n, 0, -1))
            pidx = np.concatenate([np.arange(i, n, dtype=np.int32) for i in range(n)])
            return Structure(n, m, indptr, indices, cones, pptr, pidx)
        return Structure(n, m, indptr, indices, cones)
