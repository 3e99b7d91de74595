"""Small workload for compute-sanitizer (memcheck / racecheck / initcheck, one tool per gpurun call): every kernel family of the
hot path on grids small enough for the sanitizer -- register-FFT and shared-memory DCT kernels (129 and 65 points), k_qstep,
k_mult in its three schedules (register prefetch, plain, fused KKT), the stand-alone KKT kernels, the pipelined and transposed
slab solves with three emulated slabs, the level transfer and the output recovery on slabs, PALM / acc-ADMM, the 1-D variant
and the sGS loop.

    compute-sanitizer --tool racecheck python tools/sanitize_workload.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dotsocp_b200 as dp  # noqa: E402
from dotsocp_b200 import driver, solver  # noqa: E402

IT = int(os.environ.get("SAN_ITERS", "7"))


def densities(nx, ny):
    xs = np.linspace(0, 1, nx).reshape(1, nx)
    ys = np.linspace(0, 1, ny).reshape(ny, 1)
    r0 = np.exp(-0.5 * ((xs - 0.25) ** 2 + (ys - 0.75) ** 2) / 0.05)
    r1 = np.exp(-0.5 * ((xs - 0.75) ** 2 + (ys - 0.25) ** 2) / 0.05)
    return r0 * (r0.size / r0.sum()), r1 * (r1.size / r1.sum())


def level(variant, nt, nx, ny, method="inPALM", world=1, weight=None, iters=IT):
    r0, r1 = densities(nx, ny) if ny > 1 else (np.full(nx, 1.0) + 0.3 * np.cos(np.linspace(0, 3, nx)), np.full(nx, 1.0))
    var, model = driver.initialize(r0, r1, nt)
    if weight is not None:
        model.weight = weight
    driver.InitialScaling(var, model, True, None, variant)
    opts = {"tol": 1e-12, "maxit": iters, "tau": 1.9, "sigma": 1.0, "ifCheckStepByStep": False, "scaling": True}
    o = solver.make_level_opts(variant, method, var, opts, model)
    with dp.Session(variant, nt, nx, ny, world=world) as s:
        s.upload(var.phi, var.q, var.z, var.alpha, var.beta, model.c, weight)
        hb, res = s.run(o)
        s.download()
    print(f"  {variant} {method} {nt}x{nx}x{ny} world={world}: {res.iters} iterations, {res.hist_len} checks", flush=True)


if __name__ == "__main__":
    nt, nx, ny = 17, 129, 65
    level("dot2d", nt, nx, ny)
    os.environ["DOTSOCP_KM_PF"] = "0"
    level("dot2d", nt, nx, ny, iters=4)
    del os.environ["DOTSOCP_KM_PF"]
    os.environ["DOTSOCP_KKT"] = "separate"
    level("dot2d", nt, nx, ny, iters=4)
    del os.environ["DOTSOCP_KKT"]
    os.environ["DOTSOCP_TCHUNKS"] = "2"
    level("dot2d", nt, nx, ny, world=3)
    del os.environ["DOTSOCP_TCHUNKS"]
    os.environ["DOTSOCP_TSOLVE"] = "transpose"
    level("dot2d", nt, nx, ny, world=3, iters=4)
    del os.environ["DOTSOCP_TSOLVE"]
    Q = (nt - 1) * nx * ny + nt * (nx - 1) * ny + nt * nx * (ny - 1)
    level("wdot2d", nt, nx, ny, weight=1.0 + 0.5 * np.cos(np.arange(Q) * 0.01), world=2)
    level("dot2d", 9, 33, 33, method="PALM", iters=4)
    level("dot2d", 9, 33, 33, method="acc-ADMM", iters=4)
    level("dot2d", 9, 33, 33, method="sGS-inPALM", iters=6)
    level("dot2d", 9, 33, 33, method="sGS-inPALM", iters=6, world=2)
    level("dot1d", 9, 129, 1, iters=5)
    # multilevel on slabs: level transfer + output recovery on the device
    r0, r1 = densities(65, 65)
    out, _, ML, _ = dp.solver_dotsocp2d(r0, r1, 17, 2, {"tol": 1e-2, "maxit": 12, "slabs": 2}, "inPALM")
    print("  multilevel on 2 slabs:", list(out.level_iters), "mass ok", out.massOK, flush=True)
    print("sanitize workload done", flush=True)
