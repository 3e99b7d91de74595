"""run_level (the real loop with its check schedule, sigma rule and rescalings) from the same state with the aligned and the
haloed k_mult: KKT history and final state must agree bit for bit.    python tools/diag_runlevel.py c4 [maxit]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import dotsocp_b200 as dp
from dotsocp_b200 import solver
wl = sys.argv[1] if len(sys.argv) > 1 else "c4"
maxit = int(sys.argv[2]) if len(sys.argv) > 2 else 60
nt, nx, ny = bench.WORKLOADS[wl]
var, model = bench.make_problem(nt, nx, ny)
opts = {"tol": 1e-30, "maxit": maxit, "tau": 1.9, "sigma": 1.0, "ifCheckStepByStep": False, "scaling": True}
def run(al):
    os.environ["DOTSOCP_KM_AL"] = al
    o = solver.make_level_opts("dot2d", "inPALM", var, dict(opts), model)
    with dp.Session("dot2d", nt, nx, ny) as s:
        s.upload(var.phi, var.q, None, var.alpha, var.beta, model.c)
        hb, res = s.run(o)
        out = s.download()
    return hb.kkt[:res.hist_len].copy(), hb.iter[:res.hist_len].copy(), out
ref = None
for rep, al in enumerate(["0", "1", "1", "0"]):
    kkt, its, out = run(al)
    if ref is None:
        ref = (kkt, its, out)
        print(f"{wl} maxit={maxit}: reference AL=0: {len(its)} checks at {its.astype(int).tolist()}", flush=True)
        continue
    n = min(len(kkt), len(ref[0]))
    d = np.abs(kkt[:n] - ref[0][:n]).max(axis=1)
    bad = np.nonzero(d > 0)[0]
    st = [nm for nm, a, b in zip(["phi", "q", "z", "alpha", "beta"], out, ref[2]) if not np.array_equal(a, b, equal_nan=True)]
    print(f"  run {rep} AL={al}: checks {len(its)}; first differing KKT row {bad[0] if bad.size else None}"
          f"{' (iteration %d, diff %.3e)' % (int(its[bad[0]]), d[bad[0]]) if bad.size else ''}; state differs in {st or 'nothing'}", flush=True)
