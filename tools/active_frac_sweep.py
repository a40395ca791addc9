"""C2 at other shares of active inequality rows: the headline plants active_frac = 0.2 (z + active rows < n, a
well-defined derivative); SURVEY.md's literal z ~ N(0,1) is 0.5.  For each share: forward / backward time of a 4096-batch, how many
instances the block-preconditioned backward handed to the equilibrated LSQR, iteration counts.  One JSON line per share."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cvxpylayers_b200 import problems as pr
from cvxpylayers_b200.engine import Engine, make_settings

dev = torch.device("cuda", 0)
B = 4096
args = make_settings({"eps": 1e-4, "max_iters": 10000, "lsqr_precond": 2, "adaptive_check": 1})
t = lambda a: None if a is None else torch.as_tensor(a, dtype=torch.float64, device=dev)
for af in (0.2, 0.3, 0.4, 0.5):
    bt = pr.dense_qp(B, 100, 200, 50, seed=0, active_frac=af)
    eng = Engine(bt.structure, dev)
    A, b, c, P = t(bt.A_vals), t(bt.b), t(bt.c), t(bt.P_vals)
    g = torch.Generator(device="cpu").manual_seed(1)
    dx = torch.randn((B, 100), dtype=torch.float64, generator=g).to(dev); dy = torch.randn((B, 200), dtype=torch.float64, generator=g).to(dev)
    for _ in range(2):
        sol = eng.solve(A, b, c, P, args); eng.vjp(A, b, c, sol.x, sol.y, sol.s, dx, dy, P, args)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    torch.cuda.synchronize(); e[0].record()
    sol = eng.solve(A, b, c, P, args); e[1].record()
    out = eng.vjp(A, b, c, sol.x, sol.y, sol.s, dx, dy, P, args); e[2].record(); torch.cuda.synchronize()
    fb = eng.fallback_count()
    act = float(((sol.y[:, 50:] > 1e-9).sum(1)).double().mean())
    print(json.dumps({"active_frac": af, "active_inequalities_mean": round(act, 1), "zero_plus_active_vs_n": f"{50 + act:.0f} vs 100",
                      "fwd_ms": round(e[0].elapsed_time(e[1]), 3), "bwd_ms": round(e[1].elapsed_time(e[2]), 3),
                      "problems_per_s": round(B / (e[0].elapsed_time(e[2]) * 1e-3)), "solved": int((sol.status == 1).sum()),
                      "fwd_iters_mean": round(float(sol.iters.float().mean()), 1), "lsqr_iters_mean": round(float(out[4].float().mean()), 1),
                      "lsqr_iters_max": int(out[4].max()), "block_solver_fallbacks": fb, "of": B}))
