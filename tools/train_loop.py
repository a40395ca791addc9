"""Training-loop micro-benchmark for warm starts (reference loop: examples/torch/algorithms.py:34-41):
a batch of C2-sized QP layers whose linear cost is a trainable parameter, 20 gradient steps of a quadratic loss on the
solution; every step is one forward + backward through `_CvxpyLayer.apply`.  Cold vs {"warm_start": True} vs {"warm_start": True, "reuse_setup": True} (A and P are not trained here, so their
equilibration and factorisation stay valid from step to step).

    python tools/train_loop.py [B]        -> one JSON line
"""
import json, os, sys, time
from types import SimpleNamespace

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from cvxpylayers_b200 import problems as pr
from cvxpylayers_b200.interface import B200_ctx, _CvxpyLayer

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
dev = torch.device("cuda", 0)
bt = pr.config_c2(B=B, seed=5)
st = bt.structure
bd = pr.to_boundary(bt)
out = {}
for mode in ("cold", "warm", "warm_cached"):
    args = {"eps": 1e-4, "max_iters": 10000, "lsqr_precond": 2, "adaptive_check": 1, "warm_start": mode != "cold", "reuse_setup": mode == "warm_cached"}
    ctx = B200_ctx((st.P_indices, st.P_indptr, (st.n, st.n)), (bd.con_indices, bd.con_ptr, bd.shape), bd.dims, options=args)
    cl = SimpleNamespace(solver_ctx=ctx)
    A = torch.tensor(bd.A_eval, device=dev)
    P = torch.tensor(bd.P_eval, device=dev)
    q = torch.tensor(bd.q_eval, device=dev, requires_grad=True)
    target = torch.tensor(bt.x_star * 0.9, device=dev)
    eng = ctx.engine(dev)
    times, losses, iters = [], [], []
    for step in range(22):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        primal, dual, _, _ = _CvxpyLayer.apply(P, q, A, cl, {}, True, None)
        loss = ((primal - target) ** 2).sum() / B
        loss.backward()
        with torch.no_grad():
            q -= 0.05 * q.grad
            q.grad = None
        torch.cuda.synchronize()
        if step >= 2:
            times.append(1e3 * (time.perf_counter() - t0)); losses.append(float(loss))
    # iteration count of one more forward from the cached start (engine call, the layer does not expose it)
    A_vals, P_vals, b, c = eng.ingest(A, q.detach(), P)
    from cvxpylayers_b200.engine import make_settings
    warm = ctx._last_solution.get((dev, B)) if mode != "cold" else None
    sol = eng.solve(A_vals, b, c, P_vals, make_settings(args), warm=warm, cache=ctx.setup_cache(eng, dev, B, args), reuse=True)
    out[mode] = {"ms_per_step_mean": float(np.mean(times)), "ms_per_step_min": float(np.min(times)), "loss_first": losses[0], "loss_last": losses[-1],
                 "fwd_iters_mean_next_step": float(sol.iters.float().mean()), "solved": int((sol.status == 1).sum())}
out["speedup"] = out["cold"]["ms_per_step_mean"] / out["warm"]["ms_per_step_mean"]
out["speedup_cached"] = out["cold"]["ms_per_step_mean"] / out["warm_cached"]["ms_per_step_mean"]
out["config"] = {"workload": f"C2 layers, B={B}, 20 SGD steps on q (lr 0.05), fwd+bwd per step through _CvxpyLayer.apply, device-resident"}
print(json.dumps(out))
