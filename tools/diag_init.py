"""Small mixed run (check iterations = fused KKT kernels, the others = aligned k_mult) for compute-sanitizer."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import dotsocp_b200 as dp
n = int(sys.argv[1]) if len(sys.argv) > 1 else 33
nt = (n - 1) // 2 + 1
r0, r1 = bench.densities_matlab(n, n)
o, tml, ML, rh = dp.solver_dotsocp2d(r0, r1, nt, 2, {"tol": 1e-4, "maxit": 12}, "inPALM")
print("level_iters", [int(v) for v in o.level_iters], "kkt", ML.kkt[-1])
