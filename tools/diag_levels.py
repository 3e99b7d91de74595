"""3-level solve with DOTSOCP_KM_AL switched per level (pattern argument, e.g. 110 = aligned k_mult on levels 1 and 2 only)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import dotsocp_b200 as dp
from dotsocp_b200 import solver as S
wl, pattern = sys.argv[1], sys.argv[2]
nt, nx, ny = bench.WORKLOADS[wl]
r0, r1 = bench.densities_matlab(nx, ny)
level = [0]
orig = S.Session.run
def run(self, lo):
    os.environ["DOTSOCP_KM_AL"] = pattern[level[0]]
    level[0] += 1
    return orig(self, lo)
S.Session.run = run
o, tml, ML, rh = dp.solver_dotsocp2d(r0, r1, nt, 3, {"tol": 1e-4, "maxit": 3000}, "inPALM")
print(f"{wl} AL pattern {pattern} poison={os.environ.get('DOTSOCP_POISON','0')}: level_iters={[int(v) for v in o.level_iters]} final kkt max={np.max(ML.kkt[-1][[0,2,5,6]]):.3e} nan rows={int(np.isnan(ML.kkt).any(axis=1).sum())}", flush=True)
