"""CPU tests of the host-side logic: the C-ABI library loads and exports every symbol declared in
include/bcone.h (no compute calls without a GPU), settings mapping, structure checks, the
boundary re-packing maps, and the world_size-2 gather path over gloo."""
import os
import re
import sys

import numpy as np
import pytest
import scipy.sparse as sp
import torch

from cvxpylayers_b200 import _lib, dist as bdist, problems as pr
from cvxpylayers_b200.engine import make_settings
from cvxpylayers_b200.interface import B200_ctx, dims_to_solver_dict
from cvxpylayers_b200.structure import ConeSpec, Structure

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "bcone.h")).read()
    declared = set(re.findall(r"\b(bcone_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name


def test_settings_struct_matches_header_defaults():
    st = _lib.default_settings()
    assert (st.eps_abs, st.eps_rel, st.eps_infeas) == (1e-4, 1e-4, 1e-7)  # SCS defaults
    assert (st.alpha, st.rho_x, st.scale) == (1.5, 1e-6, 0.1)
    assert (st.lsqr_atol, st.lsqr_btol, st.lsqr_conlim) == (1e-8, 1e-8, 1e8)  # diffcp / SciPy LSQR rules
    assert st.max_iters == 100000 and st.check_interval == 25 and st.lsqr_iter_lim == -1


def test_solver_args_mapping_follows_the_reference_keys():
    st = make_settings({"eps": 1e-10, "max_iters": 10000, "acceleration_lookback": 0, "verbose": True})  # tests/test_torch.py:401-405
    assert st.eps_abs == 1e-10 and st.eps_rel == 1e-10 and st.max_iters == 10000
    with pytest.raises(ValueError):
        make_settings({"not_a_solver_arg": 1})
    with pytest.raises(ValueError):
        make_settings({"mode": "dense"})


def test_no_gpu_means_loud_failure():
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from cvxpylayers_b200.engine import Engine

    with pytest.raises(_lib.EngineUnavailable):
        Engine(Structure.dense(4, 6, ConeSpec(l=6)), "cuda")


def test_structure_validation():
    with pytest.raises(ValueError):
        Structure(3, 2, [0, 1, 2], [0, 5], ConeSpec(l=2))  # column out of range
    with pytest.raises(ValueError):
        Structure(3, 2, [0, 1, 2], [0, 1], ConeSpec(l=3))  # cone rows != m
    with pytest.raises(ValueError):
        Structure(2, 1, [0, 1], [0], ConeSpec(l=1), [0, 1, 2], [1, 0])  # P not upper triangular
    with pytest.raises(NotImplementedError):
        ConeSpec.from_dict({"l": 1, "p": [0.5]})
    assert Structure.dense(5, 7, ConeSpec(z=2, l=5)).is_dense_A
    assert ConeSpec(z=1, l=2, q=[3, 4], s=[3]).m == 1 + 2 + 7 + 6


def test_dims_conversion_accepts_cvxpy_like_objects():
    class Dims:  # attribute names of cvxpy's ConeDims
        zero, nonneg, soc, psd, exp, p3d = 2, 3, [4], [3], 0, []

    assert dims_to_solver_dict(Dims()) == {"z": 2, "l": 3, "q": [4], "s": [3], "ep": 0, "ed": 0}


@pytest.mark.parametrize("name", ["C1", "C3", "C5"])
def test_boundary_roundtrip(name):
    """to_boundary() builds the [A_cvx | b] CSC tensors of diffcp_if.py:46-70; B200_ctx must recover the
    engine's CSR structure and the gather map that inverts the re-packing."""
    bt = pr.CONFIGS[name](B=3)
    st = bt.structure
    bd = pr.to_boundary(bt)
    ctx = B200_ctx(None, (bd.con_indices, bd.con_ptr, bd.shape), bd.dims)
    assert np.array_equal(ctx.structure.A_indptr, st.A_indptr) and np.array_equal(ctx.structure.A_indices, st.A_indices)
    # reference semantics: A = -A_aug[:, :-1], b = A_aug[:, -1]
    for i in range(bt.B):
        A_aug = sp.csc_matrix((bd.A_eval[:, i], bd.con_indices, bd.con_ptr), shape=bd.shape)
        assert np.allclose((-A_aug[:, :-1]).toarray(), bt.A_dense(i))
        assert np.allclose(A_aug[:, -1].toarray().ravel(), bt.b[i])
        assert np.allclose(-bd.A_eval[ctx.gather, i], bt.A_vals[i])
    assert np.array_equal(ctx.b_idx, np.arange(st.m))


def test_shard_ranges_partition_the_batch():
    for B, w in [(4096, 8), (10, 3), (5, 8), (7, 1)]:
        seen = []
        for r in range(w):
            lo, hi = bdist.shard_range(B, r, w)
            seen += list(range(lo, hi))
        assert seen == list(range(B))
        assert sum(bdist.shard_sizes(B, w)) == B


def _gloo_worker(rank, world, port, B, out):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = bdist.shard_range(B, rank, world)
    full = torch.arange(B * 3, dtype=torch.float64).reshape(B, 3)
    local = full[lo:hi].clone()
    g0 = bdist.gather_rows(local, B, dst=0)
    ga = bdist.gather_rows(local, B, dst=None)
    ok = bool(torch.equal(ga, full)) and ((rank != 0 and g0 is None) or (rank == 0 and torch.equal(g0, full)))
    out[rank] = ok
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [8, 7])
def test_gather_rows_world_size_2_gloo(B):
    import torch.multiprocessing as mp

    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 500) + B
    mp.spawn(_gloo_worker, args=(2, port, B, out), nprocs=2, join=True)
    assert out[0] and out[1]


def test_register_makes_the_solver_name_constructible(monkeypatch):
    """after ``register()`` a layer can be CONSTRUCTED with solver="B200" -- parse_args canonicalises for a
    solver cvxpy knows while the context carries the new name and a B200_ctx with the quadratic term passed through (full
    symmetric CSC pattern -> upper-triangular engine structure) and the parameter maps attached; other names still go to
    the reference's own dispatch.  cvxpy is absent here, so the reference package is a stand-in (tests/util.py)."""
    from cvxpylayers_b200 import interface as itf
    from tests.util import fake_param_prob, install_fake_cvxpylayers

    fake = install_fake_cvxpylayers(monkeypatch)
    bt = pr.dense_qp(3, 6, 9, 2, seed=1)
    problem, params = fake_param_prob(bt)
    with pytest.raises(ValueError):   # before registration cvxpy's canonicalisation rejects the unknown solver
        fake.pa.parse_args(problem, [], [], "B200")
    itf.register()
    ctx = fake.pa.parse_args(problem, [], [], "B200", solver_args={"eps": 1e-8})
    assert ctx.solver == "B200" and isinstance(ctx.solver_ctx, itf.B200_ctx)
    sc = ctx.solver_ctx
    st = bt.structure
    assert np.array_equal(sc.structure.A_indptr, st.A_indptr) and np.array_equal(sc.structure.A_indices, st.A_indices)
    assert np.array_equal(sc.structure.P_indptr, st.P_indptr) and np.array_equal(sc.structure.P_indices, st.P_indices)   # upper triangle of the full pattern
    assert sc.nnzP == st.nnzP and sc.nnzP_boundary == st.n * st.n and sc.options == {"eps": 1e-8}
    # gatherP picks the true upper entry (row i, col j), i <= j, out of the CSC-ordered full pattern
    pr_rows = problem["param_prob"].reduced_P.problem_data_index[0]
    pr_cols = np.repeat(np.arange(st.n), st.n)
    assert all(pr_rows[g] <= pr_cols[g] for g in sc.gatherP)
    assert sc._param_maps is not None and sc._param_maps[0].shape[0] == st.nnzA + st.m
    assert fake.ifs.get_torch_cvxpylayer("B200") is itf._CvxpyLayer
    with pytest.raises(RuntimeError):
        fake.ifs.get_torch_cvxpylayer("NOPE")
    with pytest.raises(ValueError):
        fake.pa.parse_args(problem, [], [], "NOPE")


def test_constant_matrices_are_detected_from_the_parameter_maps():
    """the reference's ``PA_is_constant`` rule (interfaces/moreau_if.py:233-241) -- no entry of A or P depends on
    a parameter -- switches the cached set-up on by default; rows of the constraint map that feed b may depend on parameters."""
    import scipy.sparse as sp

    from cvxpylayers_b200 import interface as itf
    from tests.util import fake_param_prob

    bt = pr.dense_qp(2, 6, 9, 2, seed=3)
    problem, _ = fake_param_prob(bt)
    pp = problem["param_prob"]
    ctx = itf.get_solver_ctx("B200", pp, problem["dims"], {}, None)
    assert ctx.PA_is_constant is False                     # every entry of A, b, c, P is a parameter in that stand-in
    assert ctx.setup_cache(None, "cpu", 2, {}) is None      # ... so no cache unless asked for
    A_map, q_map, P_map = pp.reduced_A.reduced_mat, pp.q, pp.reduced_P.reduced_mat
    nA, P1 = bt.structure.nnzA, A_map.shape[1]
    const = lambda rows: sp.csr_matrix((np.ones(rows), (np.arange(rows), np.full(rows, P1 - 1))), shape=(rows, P1))  # noqa: E731
    A_const_b_param = sp.vstack([const(nA), A_map.tocsr()[nA:]]).tocsr()       # A entries constant, b entries still parameters
    ctx.set_param_maps(A_const_b_param, q_map, const(P_map.shape[0]))
    assert ctx.PA_is_constant is True
    ctx.set_param_maps(A_const_b_param, q_map, P_map)      # P still parametrised
    assert ctx.PA_is_constant is False
    ctx.set_param_maps(A_map, q_map, const(P_map.shape[0]))
    assert ctx.PA_is_constant is False


def test_layer_epilogue_index_maps_restate_the_reference_unpacking():
    """SURVEY.md 8f.3 on the host: the index / scale maps that drive `bcone_gather_cols` are the composition of the reference's
    slice, symmetric unpacking and Fortran reshape (`torch/cvxpylayer.py:143-222, 225-282`), restated here with NumPy scatter
    exactly as the reference writes them, for vectors, matrices, 3-D arrays and both symmetric packings."""
    from types import SimpleNamespace

    from cvxpylayers_b200.layer_io import _var_map, fortran_map

    rng = np.random.default_rng(0)
    data = rng.standard_normal(200)

    def svec_to_symmetric(v, n, rows, cols, scale=None):   # the reference's _svec_to_symmetric for one instance
        out = np.zeros((n, n))
        d = v * scale if scale is not None else v
        out[rows, cols] = d
        out[cols, rows] = d
        return out

    def apply(var):
        imap, scale = _var_map(var)
        got = data[imap] * (scale if scale is not None else 1.0)
        return got.reshape(var.shape)

    for shape in [(7,), (3, 4), (2, 3, 4), ()]:
        size = int(np.prod(shape)) if shape else 1
        var = SimpleNamespace(source="primal", primal=slice(11, 11 + size), dual=None, shape=shape, unpack_fn="reshape")
        want = data[11:11 + size].reshape(shape, order="F") if shape else data[11:12].reshape(())
        assert np.array_equal(apply(var), want), shape
    n = 5
    k = n * (n + 1) // 2
    var = SimpleNamespace(source="primal", primal=slice(3, 3 + k), dual=None, shape=(n, n), unpack_fn="svec_primal")
    rows, cols = np.triu_indices(n)
    assert np.array_equal(apply(var), svec_to_symmetric(data[3:3 + k], n, rows, cols))
    var = SimpleNamespace(source="dual", primal=None, dual=slice(40, 40 + k), shape=(n, n), unpack_fn="svec_dual")
    rows_rm, cols_rm = np.tril_indices(n)
    order = np.lexsort((rows_rm, cols_rm))
    rows, cols = rows_rm[order], cols_rm[order]
    scale = np.where(rows == cols, 1.0, 1.0 / np.sqrt(2.0))
    assert np.allclose(apply(var), svec_to_symmetric(data[40:40 + k], n, rows, cols, scale), rtol=0, atol=0)
    # the prologue's map: Fortran-order flattening of a parameter (torch/cvxpylayer.py:40-56, 84-141)
    for shape in [(6,), (3, 5), (2, 3, 4)]:
        x = rng.standard_normal(shape)
        assert np.array_equal(x.reshape(-1)[fortran_map(shape)], x.reshape(-1, order="F")), shape


def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver times next to ours): exactly one JSON line on stdout carrying the same
    metric / unit / config block as our arm, `impl`, a `cpu_baseline` describing the run and an `e2e` with no PCIe traffic."""
    import json
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--config", "C1"],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-800:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "problems/s" and d["n_gpus"] == 1
    assert d["steps"] == 1 and d["warmup"] == 0 and d["value"] > 0 and d["dtype"] == "f64"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert "workload" in d["config"] and "model" not in d["config"]
