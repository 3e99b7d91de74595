"""Run under torchrun on >= 2 GPUs (one process per GPU): parity at the BASELINE grids on time slabs over NCCL.

  1. configs[3]: example2 (mixture) at 512x512x256, one level, 12 checked inPALM iterations -- every KKT row and objective value
     against the CPU oracle's golden (tests/golden/solver_c4.json), each rank building only its own slab.
  2. the multilevel driver on slabs: example1 at 256x256x128, 3 levels to tol 1e-4 -- iterations per level, check schedule and KKT
     history against the CPU oracle's golden (tests/golden/solver_c3.json); mass conservation and the transport cost from the
     device-side output recovery.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/dist_parity_big.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
import dotsocp_b200 as dp  # noqa: E402
from dotsocp_b200 import _lib, slab, solver  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    _lib.check(_lib.lib().dotsocp_set_device(local))
    dist.init_process_group("nccl")
    ident = slab.broadcast_unique_id(dist, rank)
    small = os.environ.get("DOTSOCP_DIST_SMALL") == "1"      # CPU-sized stand-ins for a quick functional run

    # ---- 1. configs[3], 12 checked iterations
    with open(os.path.join(ROOT, "tests", "golden", "solver_c4.json")) as f:
        gold = json.load(f)
    nt, nx, ny = gold["grid_nodes"]
    var, model = bench.make_problem(nt, nx, ny, rank, world, problem="example2")
    opts = {"tol": 1e-4, "maxit": gold["iters"], "tau": 1.9, "sigma": 1.0, "ifCheckStepByStep": True, "scaling": True}
    o = solver.make_level_opts("dot2d", "inPALM", var, opts, model)
    with dp.Session("dot2d", nt, nx, ny, rank=rank, world=world, nccl_id=ident) as s:
        s.upload(var.phi, var.q, None, var.alpha, var.beta, model.c)
        del var
        hb, res = s.run(o)
    n = res.hist_len
    assert res.iters == gold["iters"] and [int(v) for v in hb.iter[:n]] == gold["hist_iter"]
    d = float(np.abs(hb.kkt[:n] - np.array(gold["kkt"])).max())
    assert d < 1e-8, d
    assert np.abs(hb.priVal[:n] - np.array(gold["priVal"])).max() <= 1e-6 * np.abs(gold["priVal"]).max()
    assert np.abs(hb.dualVal[:n] - np.array(gold["dualVal"])).max() <= 1e-6 * np.abs(gold["dualVal"]).max()
    assert abs(res.sigma - gold["sigma"]) <= 1e-12 * gold["sigma"]
    assert np.allclose([res.cScale, res.dScale, res.D, res.E], gold["scal"], rtol=1e-12, atol=0)     # after the in-loop rescalings
    if rank == 0:
        print(f"dist parity ok: configs[3] 512x512x256 world={world} max|kkt - oracle| = {d:.2e}", flush=True)

    # ---- 2. multilevel driver on slabs against the 256x256x128 golden
    with open(os.path.join(ROOT, "tests", "golden", "solver_c3.json")) as f:
        g3 = json.load(f)
    r0, r1 = bench.densities_matlab(257, 257)
    out, _, ML, rh = dp.solver_dotsocp2d(r0, r1, 129, 3, {"tol": 1e-4, "maxit": 3000,
                                                          "slabs": {"rank": rank, "world": world, "nccl_id": slab.REUSE_COMM}}, "inPALM")
    assert [int(v) for v in out.level_iters] == g3["level_iters"], (list(out.level_iters), g3["level_iters"])
    assert [int(v) for v in ML.iter] == g3["hist_iter"]
    d3 = float(np.abs(ML.kkt - np.array(g3["kkt"])).max())
    assert d3 < 1e-8, d3
    assert abs(rh.priVal[-1] - g3["priVal"]) <= 1e-6 * abs(g3["priVal"])
    assert out.massOK and out.rho.shape[0] == out.slab[1] - out.slab[0]
    # every time level of rho carries unit mass (check_massConservation) -- on every rank's own levels, from the device sums
    assert np.abs(out.sumRho - 1).max() < 1e-4
    if rank == 0:
        print(f"dist parity ok: 3-level 256x256x128 world={world} iters={list(out.level_iters)} max|kkt - oracle| = {d3:.2e} "
              f"w2={out.w2:.9f}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
