"""Register-tiled forward (fwd_fast.cu) against the generic forward kernel on the same batch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cvxpylayers_b200 import problems as pr
from cvxpylayers_b200.engine import Engine, make_settings

dev = torch.device("cuda", 0)
t = lambda a: None if a is None else torch.as_tensor(a, dtype=torch.float64, device=dev)
for name, bt in [("C2", pr.config_c2(B=256)), ("LP", pr.dense_lp(n=60, m=150, B=64, seed=3) if hasattr(pr, "dense_lp") else None)]:
    if bt is None:
        continue
    st = bt.structure
    os.environ.pop("BCONE_NO_FAST_FWD", None)
    e_fast = Engine(st, dev)
    os.environ["BCONE_NO_FAST_FWD"] = "1"
    e_gen = Engine(st, dev)
    os.environ.pop("BCONE_NO_FAST_FWD", None)
    print(name, "fast:", e_fast.kernel_info(), "generic:", e_gen.kernel_info())
    A, b, c, P = t(bt.A_vals), t(bt.b), t(bt.c), t(bt.P_vals)
    for eps in (1e-4, 1e-9):
        args = make_settings({"eps": eps, "max_iters": 20000, "adaptive_check": 1})
        s1 = e_fast.solve(A, b, c, P, args); s2 = e_gen.solve(A, b, c, P, args)
        torch.cuda.synchronize()
        dx = (s1.x - s2.x).abs().max().item(); dy = (s1.y - s2.y).abs().max().item(); ds = (s1.s - s2.s).abs().max().item()
        print(f"  eps={eps:g}: status fast {s1.status.unique().tolist()} gen {s2.status.unique().tolist()}  iters mean {s1.iters.float().mean():.2f} / {s2.iters.float().mean():.2f}"
              f"  max|dx| {dx:.2e} |dy| {dy:.2e} |ds| {ds:.2e}  iters differ in {(s1.iters != s2.iters).sum().item()} instances")
