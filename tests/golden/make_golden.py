#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ (run in the build container, where /root/reference exists).

  kernels.npz   seeded inputs + the outputs of the reference's GENUINE MEX binaries (oracle/_ref, built by oracle/Makefile
                from /root/reference/socp/*/utils/*.mexa64) for mexBFd, mexBFdConj, mexProjSoc, mexBFd1d, mexBFdConj1d.
                These pin the oracle's C / numpy restatements (and the CUDA kernels) bit for bit.
  solver.json   iteration counts, KKT history and objectives of the ORACLE's restatement of the MATLAB loops on small
                analytic instances.  The reference has no tests or logs of its own and MATLAB is unavailable, so these
                are self-generated regression goldens ("parity unpinned by the reference" at the solver level).
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import dotsocp_oracle as O  # noqa: E402
from oracle import kernels as K  # noqa: E402
from oracle import refmex  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def kernels():
    assert refmex.available(), "oracle/_ref missing: run `make -C oracle` where /root/reference exists"
    rng = np.random.default_rng(20261018)
    out = {}
    for i, (nt, nx, ny) in enumerate([(3, 3, 4), (5, 4, 3), (4, 6, 5), (2, 2, 2)]):
        L, nbx, nby = K.sizes2d(nt, nx, ny)
        q = rng.standard_normal(L + nbx + nby)
        S, DF = 0.5 + rng.random(), rng.standard_normal()
        z2 = np.full((L, 10), 0.25, order="F")
        K.mexBFd(z2, q, nt, nx, ny, S, DF, backend="ref")
        z = np.asfortranarray(rng.standard_normal((L, 10)))
        q2 = np.zeros(L + nbx + nby)
        K.mexBFdConj(q2, z, nt, nx, ny, S, backend="ref")
        out.update({f"g{i}_dims": np.array([nt, nx, ny]), f"g{i}_S": S, f"g{i}_DF": DF, f"g{i}_q": q, f"g{i}_z2": z2,
                    f"g{i}_z": z, f"g{i}_q2": q2})
    for i, (M, N) in enumerate([(7, 10), (8, 10), (9, 6), (6, 6), (5, 3), (4, 12)]):
        v = np.asfortranarray(rng.standard_normal((M, N)))
        v[0, 0] = 30.0
        v[1, 0] = -30.0
        v[2, 1:] = 0.0
        v[3, :] = 0.0
        p = np.zeros((M, N), order="F")
        K.mexProjSoc(p, v, backend="ref")
        out.update({f"p{i}_in": v, f"p{i}_out": p})
    for i, (nt, nx) in enumerate([(3, 4), (5, 7), (2, 3)]):
        L = (nt - 1) * nx
        Q = L + nt * (nx - 1)
        q = rng.standard_normal(Q)
        S, DF = 0.5 + rng.random(), rng.standard_normal()
        z = np.full((L, 6), -0.5, order="F")
        K.mexBFd1d(z, q, nt, nx, S, DF, backend="ref")
        zin = np.asfortranarray(rng.standard_normal((L, 6)))
        q2 = np.zeros(Q)
        K.mexBFdConj1d(q2, zin, nt, nx, S, backend="ref")
        out.update({f"h{i}_dims": np.array([nt, nx]), f"h{i}_S": S, f"h{i}_DF": DF, f"h{i}_q": q, f"h{i}_z": z,
                    f"h{i}_zin": zin, f"h{i}_q2": q2})
    np.savez_compressed(os.path.join(HERE, "kernels.npz"), **out)
    print("kernels.npz:", len(out), "arrays")


def solver():
    cases = []

    def record(name, out, ML, rh, extra=None):
        d = {"name": name, "level_iters": [int(v) for v in out.level_iters], "hist_iter": ML.iter.tolist(),
             "kkt": ML.kkt.tolist(), "pdGap": ML.pdGap.tolist(), "priVal": float(rh.priVal[-1]),
             "dualVal": float(rh.dualVal[-1]), "sigma": float(out.sigma), "mass_min": float(out.sumRho.min()),
             "mass_max": float(out.sumRho.max()), "w2": O.w2_cost(out, 1 if name.startswith("dot1d") else 2)}
        d.update(extra or {})
        cases.append(d)
        print(name, d["level_iters"], d["priVal"])

    rho0, rho1 = O.get_example2d("example1", 17, 17)
    out, _, ML, rh = O.solver_dotsocp2d(rho0, rho1, 9, 2, {"tol": 1e-4, "maxit": 3000}, "inPALM")
    record("dot2d_example1_17x17x9_L2_inPALM", out, ML, rh)
    rho0, rho1 = O.get_example2d("example2", 33, 33)
    out, _, ML, rh = O.solver_dotsocp2d(rho0, rho1, 17, 2, {"tol": 1e-4, "maxit": 3000}, "ALG2")
    record("dot2d_example2_33x33x17_L2_ALG2", out, ML, rh)
    rho0, rho1 = O.get_example2d("example1", 17, 17)
    out, _, ML, rh = O.solver_dotsocp2d(rho0, rho1, 9, 1, {"tol": 1e-4, "maxit": 3000}, "acc-ADMM")
    record("dot2d_example1_17x17x9_L1_accADMM", out, ML, rh)
    out, _, ML, rh = O.solver_dotsocp2d(rho0, rho1, 9, 1, {"tol": 1e-4, "maxit": 3000}, "PALM")
    record("dot2d_example1_17x17x9_L1_PALM", out, ML, rh)
    w = O.gene_weight_circle(9, 17, 17)
    out, _, ML, rh = O.solver_wdotsocp2d(rho0, rho1, 9, 2, {"tol": 1e-3, "maxit": 10000, "weight": w}, "inPALM")
    record("wdot2d_example1_circle_17x17x9_L2_inPALM", out, ML, rh)
    rho0, rho1 = O.get_example1d("gaussian", 129)
    out, _, ML, rh = O.solver_dotsocp1d(rho0, rho1, 9, 2, {"tol": 1e-5, "maxit": 3000}, "inPALM")
    record("dot1d_gaussian_129x9_L2_inPALM", out, ML, rh)
    with open(os.path.join(HERE, "solver.json"), "w") as f:
        json.dump(cases, f)
    print("solver.json:", len(cases), "cases")


if __name__ == "__main__":
    kernels()
    solver()
