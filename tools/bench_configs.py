"""Times forward + backward of the non-headline BASELINE configs (parity-test cases, not bench lines)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cvxpylayers_b200 import problems as pr
from cvxpylayers_b200.engine import Engine, make_settings

dev = torch.device("cuda", 0)
t = lambda a: None if a is None else torch.as_tensor(a, dtype=torch.float64, device=dev)
out = []
for name, B in [("C1", 4096), ("C3", 2048), ("C5", 256), ("EXP", 1024)]:
    bt = pr.CONFIGS[name](B=B)
    st = bt.structure
    eng = Engine(st, dev)
    args = make_settings({"eps": 1e-4, "max_iters": 20000, "lsqr_precond": 1})
    A, b, c, P = t(bt.A_vals), t(bt.b), t(bt.c), t(bt.P_vals)
    g = torch.Generator(device="cpu").manual_seed(1)
    dx = torch.randn((B, st.n), dtype=torch.float64, generator=g).to(dev)
    dy = torch.randn((B, st.m), dtype=torch.float64, generator=g).to(dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    for rep in range(3):
        ev[0].record(); sol = eng.solve(A, b, c, P, args); ev[1].record()
        res = eng.vjp(A, b, c, sol.x, sol.y, sol.s, dx, dy, P, args); ev[2].record()
        torch.cuda.synchronize()
    f, bw = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
    out.append({"config": name, "B": B, "n": st.n, "m": st.m, "nnzA": st.nnzA, "fwd_ms": round(f, 3), "bwd_ms": round(bw, 3),
                "problems_per_s": round(B / ((f + bw) * 1e-3)), "solved": int((sol.status == 1).sum()),
                "fwd_iters_mean": float(sol.iters.float().mean()), "lsqr_iters_mean": float(res[4].float().mean()), **eng.kernel_info()})
    print(json.dumps(out[-1]))
