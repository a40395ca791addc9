import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from types import SimpleNamespace
import numpy as np, torch
from cvxpylayers_b200 import problems as pr
import cvxpylayers_b200.interface as itf
from cvxpylayers_b200.interface import B200_ctx, _CvxpyLayer
dev = torch.device("cuda", 0)
bt = pr.CONFIGS["C3"](B=2048); st = bt.structure; bd = pr.to_boundary(bt)
args = {"eps": 1e-4, "max_iters": 10000, "lsqr_precond": 2, "adaptive_check": 1}
for chunk in (1024, 10**9):
    itf.PIPE_CHUNK = chunk
    ctx = B200_ctx(None, (bd.con_indices, bd.con_ptr, bd.shape), bd.dims, options=args)
    ctx.device = dev
    cl = SimpleNamespace(solver_ctx=ctx)
    hA = torch.from_numpy(bd.A_eval).pin_memory(); hq = torch.from_numpy(bd.q_eval).pin_memory()
    for rep in range(8):
        A = hA.detach().requires_grad_(True); q = hq.detach().requires_grad_(True)
        try:
            primal, dual, saved, _ = _CvxpyLayer.apply(None, q, A, cl, {}, True, None)
            sol_status = "ok"
            g = torch.Generator(device="cpu").manual_seed(7)
            dxh = torch.randn(primal.shape, dtype=torch.float64, generator=g); dyh = torch.randn(dual.shape, dtype=torch.float64, generator=g)
            ((primal * dxh).sum() + (dual * dyh).sum()).backward()
            sol_status += " bwd ok, grad finite %s" % bool(torch.isfinite(A.grad).all())
        except Exception as e:
            sol_status = str(e)
        print("chunk", chunk, "rep", rep, sol_status)
    # inspect with direct pipelined call
    eng = ctx.engine(dev)
    from cvxpylayers_b200.engine import make_settings
    out = itf._forward_pipelined(eng, dev, hA, hq, None, make_settings(args), False) if chunk == 1024 else None
    if out is not None:
        torch.cuda.synchronize()
        sol = out[4]; stt = sol.status.cpu().numpy(); bad = np.nonzero(stt != 1)[0]
        print("  direct pipelined: bad", bad[:8], stt[bad[:8]], sol.iters.cpu().numpy()[bad[:8]], sol.resid.cpu().numpy()[bad[:3]])
        A_vals = out[0]; print("  A_vals finite", bool(torch.isfinite(A_vals).all()), "b finite", bool(torch.isfinite(out[2]).all()), "equal to reference ingest", bool(torch.equal(A_vals.cpu(), torch.tensor(bt.A_vals))))
