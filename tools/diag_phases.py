"""Host time of every Session call inside a 3-level solve (where does a small multilevel solve spend its wall time?)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import dotsocp_b200 as dp
from dotsocp_b200 import solver as S
n = int(sys.argv[1]) if len(sys.argv) > 1 else 129
nt = (n - 1) // 2 + 1
log = []
def wrap(cls, name):
    orig = getattr(cls, name)
    def f(*a, **k):
        t0 = time.perf_counter()
        r = orig(*a, **k)
        log.append((name, time.perf_counter() - t0))
        return r
    setattr(cls, name, f)
for nm in ("__init__", "upload", "run", "prolong_from", "recover", "download", "close"):
    wrap(S.Session, nm)
orig_ref = S.Session.refined.__func__
def refined(cls, coarse):
    t0 = time.perf_counter()
    r = orig_ref(cls, coarse)
    log.append(("refined", time.perf_counter() - t0))
    return r
S.Session.refined = classmethod(refined)
wrap(S.OutputBuffers, "get")
r0, r1 = bench.densities_matlab(n, n)
for rep in range(2):
    log.clear()
    t0 = time.perf_counter()
    o, tml, ML, rh = dp.solver_dotsocp2d(r0, r1, nt, 3, {"tol": 1e-4, "maxit": 3000}, "inPALM")
    tot = time.perf_counter() - t0
    print(f"n={n} rep {rep}: total {tot:.3f} s, iters {[int(v) for v in o.level_iters]}; " + ", ".join(f"{k} {1e3*v:.0f}" for k, v in log) + " (ms)", flush=True)
