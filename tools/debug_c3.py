import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cvxpylayers_b200 import problems as pr
from cvxpylayers_b200.engine import Engine, make_settings
dev = torch.device("cuda", 0)
bt = pr.CONFIGS["C3"](B=2048)
st = bt.structure
t = lambda a: None if a is None else torch.as_tensor(a, dtype=torch.float64, device=dev)
A, b, c = t(bt.A_vals), t(bt.b), t(bt.c)
for args in ({"eps": 1e-4, "max_iters": 10000, "adaptive_check": 1}, {"eps": 1e-4, "max_iters": 10000}, {"eps": 1e-4, "max_iters": 10000, "adaptive_check": 1, "acceleration_lookback": 0}):
    eng = Engine(st, dev)
    for rep in range(3):
        sol = eng.solve(A, b, c, None, make_settings(args))
        torch.cuda.synchronize()
        stt = sol.status.cpu().numpy(); it = sol.iters.cpu().numpy()
        bad = np.nonzero(stt != 1)[0]
        print(args, "rep", rep, "bad", bad[:10], stt[bad[:10]], it[bad[:10]], "iters mean", it.mean(), "max", it.max(), eng.kernel_info()["fwd_ctas_per_sm"])
    # two streams concurrently, chunks of 1024
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    outs = []
    for k, (lo, hi) in enumerate([(0, 1024), (1024, 2048)]):
        with torch.cuda.stream([s1, s2][k]):
            outs.append(eng.solve(A[lo:hi], b[lo:hi], c[lo:hi], None, make_settings(args)))
    torch.cuda.synchronize()
    for k, o in enumerate(outs):
        stt = o.status.cpu().numpy(); bad = np.nonzero(stt != 1)[0]
        print("  concurrent chunk", k, "bad", bad[:10], stt[bad[:10]], o.iters.cpu().numpy()[bad[:10]])
