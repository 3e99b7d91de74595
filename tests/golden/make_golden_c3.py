"""Golden history of the CPU oracle at a BASELINE grid: 3-level inPALM solve of example1 at 256x256x128 cells
(nodes 129 x 257 x 257), tol 1e-4, reference defaults.  Takes about 5 minutes on 8 cores; writes solver_c3.json.

    python tests/golden/make_golden_c3.py
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import dotsocp_oracle as O  # noqa: E402


def densities(nx, ny):
    """examples/dot2d/gene_example1.m:5-25 on an (ny, nx) grid, mean 1"""
    xs = np.linspace(0, 1, nx).reshape(1, nx)
    ys = np.linspace(0, 1, ny).reshape(ny, 1)
    r0 = np.exp(-0.5 * ((xs - 0.25) ** 2 + (ys - 0.75) ** 2) / 0.05)
    r1 = np.exp(-0.5 * ((xs - 0.75) ** 2 + (ys - 0.25) ** 2) / 0.05)
    return r0 * (r0.size / r0.sum()), r1 * (r1.size / r1.sum())


if __name__ == "__main__":
    nt, nx, ny = 129, 257, 257
    r0, r1 = densities(nx, ny)
    t0 = time.perf_counter()
    out, _, ML, rh = O.solver_dotsocp2d(r0, r1, nt, 3, {"tol": 1e-4, "maxit": 3000}, "inPALM", workers=os.cpu_count() or 1)
    res = {"level_iters": [int(v) for v in out.level_iters], "final_kkt": [float(v) for v in ML.kkt[-1]],
           "final_kkt_max": float(np.max(ML.kkt[-1][[0, 2, 5, 6]])), "hist_iter": [int(v) for v in ML.iter],
           "kkt": ML.kkt.tolist(), "priVal": float(rh.priVal[-1]), "seconds": time.perf_counter() - t0}
    with open(os.path.join(HERE, "solver_c3.json"), "w") as f:
        json.dump(res, f)
    print(res["level_iters"], res["final_kkt_max"], res["priVal"], res["seconds"])
