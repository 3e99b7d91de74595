"""Oracle goldens of two more BASELINE.json configurations (run on the CPU, ~15 minutes on 8 cores):

  configs[0]  demo_dot1d.m defaults: 1-D Gaussian instance, nt = 33, nx = 1025, 3 levels, tol 1e-5, maxit 3000, inPALM
  configs[2]  demo_wdot2d.m-sized weighted instance: example1 densities, weight = gene_weight_circle, 256x256x128 cells
              (nodes 129 x 257 x 257), 3 levels, tol 1e-3, maxit 1e4, inPALM                       

    python tests/golden/make_golden_baseline_configs.py [dot1d] [wdot2d]
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import dotsocp_oracle as O  # noqa: E402


def config_dot1d():
    rho0, rho1 = O.get_example1d("gaussian", 1025)
    return rho0, rho1, 33, 3, {"tol": 1e-5, "maxit": 3000}


def config_wdot2d():
    nt, n = 129, 257
    rho0, rho1 = O.get_example2d("example1", n, n)
    return rho0, rho1, nt, 3, {"tol": 1e-3, "maxit": 10000, "weight": O.gene_weight_circle(nt, n, n)}


def record(out, ML, rh, t0):
    return {"level_iters": [int(v) for v in out.level_iters], "hist_iter": [int(v) for v in ML.iter], "kkt": ML.kkt.tolist(),
            "priVal": float(rh.priVal[-1]), "w2": float(O.w2_cost(out, 1 if not hasattr(out, "Ey") else 2)),
            "seconds": time.perf_counter() - t0}


if __name__ == "__main__":
    which = sys.argv[1:] or ["dot1d", "wdot2d"]
    path = os.path.join(HERE, "solver_baseline_configs.json")
    res = json.load(open(path)) if os.path.exists(path) else {}
    if "dot1d" in which:
        rho0, rho1, nt, levelN, opts = config_dot1d()
        t0 = time.perf_counter()
        out, _, ML, rh = O.solver_dotsocp1d(rho0, rho1, nt, levelN, opts, "inPALM")
        res["dot1d_demo_default"] = record(out, ML, rh, t0)
        print("dot1d", res["dot1d_demo_default"]["level_iters"], res["dot1d_demo_default"]["seconds"], flush=True)
        json.dump(res, open(path, "w"))
    if "wdot2d" in which:
        rho0, rho1, nt, levelN, opts = config_wdot2d()
        t0 = time.perf_counter()
        out, _, ML, rh = O.solver_wdotsocp2d(rho0, rho1, nt, levelN, opts, "inPALM")
        res["wdot2d_circle_256x256x128"] = record(out, ML, rh, t0)
        print("wdot2d", res["wdot2d_circle_256x256x128"]["level_iters"], res["wdot2d_circle_256x256x128"]["seconds"], flush=True)
        json.dump(res, open(path, "w"))
