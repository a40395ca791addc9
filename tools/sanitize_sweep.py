"""One tiny forward + backward of every kernel path, meant to run under compute-sanitizer (racecheck / memcheck / initcheck):
    compute-sanitizer --tool racecheck python tools/sanitize_sweep.py
Tight tolerance on purpose: the runs are long enough for Anderson acceleration, rescaling and warm PSD / exp projections."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cvxpylayers_b200 import problems as pr
from cvxpylayers_b200.engine import Engine, make_settings

dev = torch.device("cuda", 0)
t = lambda a: None if a is None else torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
only = sys.argv[1:] or None
cases = [("C1", pr.CONFIGS["C1"](B=3), 2), ("C2", pr.CONFIGS["C2"](B=2), 2), ("C2p1", pr.CONFIGS["C2"](B=2), 1), ("C3", pr.CONFIGS["C3"](B=3), 1),
         ("C5", pr.CONFIGS["C5"](B=2), 1), ("EXP", pr.CONFIGS["EXP"](B=3), 1), ("C2SOC", pr.qp_as_socp(pr.dense_qp(1, 40, 80, 20, seed=1)), 1),
         ("sparse_qp", pr.sparse_qp(B=1, n=120, m=240, seed=2), 1)]
for name, bt, precond in cases:
    if only and name not in only:
        continue
    eng = Engine(bt.structure, dev)
    A, b, c, P = t(bt.A_vals), t(bt.b), t(bt.c), t(bt.P_vals)
    sol = eng.solve(A, b, c, P, make_settings({"eps": 1e-9, "max_iters": 3000}))
    g = torch.Generator(device="cpu").manual_seed(1)
    dx = torch.randn(sol.x.shape, dtype=torch.float64, generator=g).to(dev); dy = torch.randn(sol.y.shape, dtype=torch.float64, generator=g).to(dev)
    out = eng.vjp(A, b, c, sol.x, sol.y, sol.s, dx, dy, P, make_settings({"lsqr_precond": precond, "lsqr_iter_lim": 400}))
    torch.cuda.synchronize()
    print(name, eng.path_info(), "status", sol.status.tolist(), "iters", sol.iters.tolist(), "lsqr", out[4].tolist(), flush=True)
    if name == "C2":   # the cached set-up: fill, then reuse with a warm start (Kinv + E + D come back through the cache; P by TMA)
        S = make_settings({"eps": 1e-9, "max_iters": 3000})
        cache = eng.new_cache(bt.B)
        one = eng.solve(A, b, c, P, S, cache=cache, reuse=False)
        two = eng.solve(A, b * 1.001, c, P, S, warm=one, cache=cache, reuse=True)
        torch.cuda.synchronize()
        print("C2 cached", "status", two.status.tolist(), "iters", two.iters.tolist(), "valid", cache.view(bt.B, -1)[:, 1].tolist(), flush=True)
