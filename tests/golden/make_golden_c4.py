"""Golden history of the CPU oracle at BASELINE configs[3]: example2 (Gaussian -> 4-Gaussian mixture, gene_example2.m) at
512x512x256 cells (nodes 257 x 513 x 513), ONE level, inPALM from the reference's initial state, a fixed 12 iterations with
ifCheckStepByStep = true so that every iteration leaves a KKT row.  About 40 GB of host memory and 20-30 minutes on 8 cores;
writes solver_c4.json (KKT rows, objective values, and fingerprints of the final iterates: scaled 2-norms and 64 sampled
entries per array).

    python tests/golden/make_golden_c4.py [nt nx ny iters]
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

ITERS = 12


def problem(nx, ny):
    """examples/dot2d/gene_example2.m:5-18 on an (ny, nx) grid, mean 1 (same arithmetic as oracle.get_example2d)"""
    from oracle import dotsocp_oracle as O
    return O.get_example2d("example2", nx, ny)


def sample_index(n, k=64):
    """k deterministic positions of an n-vector (no RNG: a fixed odd stride)"""
    return [(int(i) * 2654435761 + 12345) % n for i in range(k)]


def fingerprint(a):
    v = np.ravel(a, order="K" if a.ndim == 1 else "F")
    return {"norm2": float(np.sqrt(np.dot(v, v))), "samples": [float(v[i]) for i in sample_index(v.size)]}


if __name__ == "__main__":
    from oracle import dotsocp_oracle as O
    args = [int(v) for v in sys.argv[1:]]
    nt, nx, ny = (args + [257, 513, 513])[:3] if len(args) >= 3 else (257, 513, 513)
    iters = args[3] if len(args) > 3 else ITERS
    rho0, rho1 = problem(nx, ny)
    t0 = time.perf_counter()
    var, model = O.initialize2d(rho0, rho1, nt)
    O.InitialScaling(var, model, True, None, "dot2d")
    opts = {"tol": 1e-4, "maxit": iters, "tau": 1.9, "sigma": 1.0, "ifCheckStepByStep": True, "scaling": True}
    rh, sigma = O.solver_socp_inPALM(var, opts, model, workers=os.cpu_count() or 1)
    res = {"grid_nodes": [nt, nx, ny], "iters": int(var.time["Iters"]), "hist_iter": [int(v) for v in rh.iter], "kkt": rh.kkt.tolist(),
           "priVal": rh.priVal.tolist(), "dualVal": rh.dualVal.tolist(), "pdGap": rh.pdGap.tolist(), "sigma": float(sigma),
           "scal": [float(var.cScale), float(var.dScale), float(var.D), float(var.E)],
           "final": {k: fingerprint(getattr(var, k)) for k in ("phi", "q", "z", "alpha", "beta")},
           "step_seconds": {k: float(v) for k, v in var.time.items()}, "seconds": time.perf_counter() - t0}
    name = "solver_c4.json" if (nt, nx, ny) == (257, 513, 513) else f"solver_c4_{nt}x{nx}x{ny}.json"
    with open(os.path.join(HERE, name), "w") as f:
        json.dump(res, f)
    print(res["hist_iter"], res["kkt"][-1], res["seconds"], flush=True)
