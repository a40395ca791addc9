"""Small driver for ncu captures: one forward + one backward launch on a C2 batch."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cvxpylayers_b200 import problems as pr
from cvxpylayers_b200.engine import Engine, make_settings

B = int(sys.argv[1]) if len(sys.argv) > 1 else 296
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
name = sys.argv[3] if len(sys.argv) > 3 else "C2"
dev = torch.device("cuda", 0)
bt = pr.CONFIGS[name](B=B)
st = bt.structure
t = lambda a: None if a is None else torch.as_tensor(a, dtype=torch.float64, device=dev)
eng = Engine(st, dev)
args = make_settings({"eps": 1e-4, "max_iters": 10000, "lsqr_precond": 2})
A, b, c, P = t(bt.A_vals), t(bt.b), t(bt.c), t(bt.P_vals)
g = torch.Generator(device="cpu").manual_seed(1)
dx = torch.randn((B, st.n), dtype=torch.float64, generator=g).to(dev)
dy = torch.randn((B, st.m), dtype=torch.float64, generator=g).to(dev)
for _ in range(reps):
    sol = eng.solve(A, b, c, P, args)
    out = eng.vjp(A, b, c, sol.x, sol.y, sol.s, dx, dy, P, args)
torch.cuda.synchronize()
print("iters", sol.iters.float().mean().item(), "lsqr", out[4].float().mean().item())
