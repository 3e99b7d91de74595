"""Multilevel solve with the time axis cut into emulated slabs against the single-slab run: level iterations and the first KKT
row that differs.    python tools/diag_slabs.py c3 1 2 8"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import dotsocp_b200 as dp
wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
worlds = [int(v) for v in sys.argv[2:]] or [1, 2, 8]
nt, nx, ny = bench.WORKLOADS[wl]
r0, r1 = bench.densities_matlab(nx, ny)
ref = None
for w in worlds:
    o, tml, ML, rh = dp.solver_dotsocp2d(r0, r1, nt, 3, {"tol": 1e-4, "maxit": 3000, "slabs": (w if w > 1 else None)}, "inPALM")
    print(f"{wl} slabs={w}: level_iters={[int(v) for v in o.level_iters]} checks={ML.len} final kkt max={np.max(ML.kkt[-1][[0,2,5,6]]):.3e} "
          f"objective={rh.priVal[-1]:.12f}", flush=True)
    if ref is None:
        ref = ML
    else:
        n = min(ML.kkt.shape[0], ref.kkt.shape[0])
        d = np.abs(ML.kkt[:n] - ref.kkt[:n]).max(axis=1)
        bad = np.nonzero(d > 1e-12)[0]
        print(f"   rows compared {n}, max diff {d.max():.3e}, first row with diff > 1e-12: {bad[0] if bad.size else None}"
              f" (iteration {int(ML.iter[bad[0]]) if bad.size else '-'})", flush=True)
