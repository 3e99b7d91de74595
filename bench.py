#!/usr/bin/env python
"""bench.py -- ADMM/inPALM iterations per second of the DOT-SOCP hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W                 # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W  # the reference's CPU path (oracle), rank 0 only

A "step" is ONE inPALM iteration (Poisson solve, q-step, projection + multiplier step) on the named grid with the
state resident in HBM.  `value` = iterations/s over all N GPUs (time slabs => one job, strong scaling);
`e2e` = the same metric through the reference-facing C-ABI call dotsocp_solve_level() with HOST buffers: upload of
(phi,q,z,alpha,beta,c), K iterations, download, all inside the timed region.
Prints exactly one JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {   # name -> (nt, nx, ny) nodes ; BASELINE.json configs
    "c2": (33, 65, 65),        # 64x64x32
    "c3": (129, 257, 257),     # 256x256x128
    "c4": (257, 513, 513),     # 512x512x256
    "c5": (513, 1025, 1025),   # 1024x1024x512  <- the grid the metric is quoted on
}
WL_LABEL = {"c2": "64x64x32", "c3": "256x256x128", "c4": "512x512x256", "c5": "1024x1024x512"}


def sizes(nt, nx, ny):
    N = nt * nx * ny
    L = (nt - 1) * nx * ny
    Q = L + nt * (nx - 1) * ny + nt * nx * (ny - 1)
    return N, L, Q


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic(kernel, wl, world):
    """DRAM bytes of one launch of the dominant kernel from the committed ncu --set full capture (profiles/traffic.json);
    only the single-GPU captures exist, so multi-GPU lines report null."""
    if world != 1:
        return None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return float(json.load(f)[kernel][wl]["bytes"])
    except (OSError, KeyError, ValueError):
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_problem(nt, nx, ny, rank=0, world=1, problem="example1"):
    """Synthetic Gaussian densities (examples/dot2d/gene_example1.m) on the named grid and the reference's initial state
    (initialize.m) after InitialScaling (solver_dotsocp2d.m:304-365) -- built slab-locally with the product's own host
    code (no oracle on this path): every rank only materialises its time slab; q, z, alpha, beta start as zeros."""
    from dotsocp_b200 import slab
    x = np.linspace(0, 1, nx).reshape(nx, 1)
    y = np.linspace(0, 1, ny).reshape(1, ny)
    s = 0.05
    if problem == "example1":
        rho0 = np.exp(-0.5 * ((x - 0.25) ** 2 + (y - 0.75) ** 2) / s)
        rho1 = np.exp(-0.5 * ((x - 0.75) ** 2 + (y - 0.25) ** 2) / s)
        rho0 = rho0 * (nx * ny / rho0.sum())
        rho1 = rho1 * (nx * ny / rho1.sum())
    else:  # examples/dot2d/gene_example2.m:4-18 + get_example.m:45-46, built MATLAB-shaped (ny, nx) like the reference so that
        # every rounding (term order, normalising sum) is the one of the golden rows; then viewed as C order (x, y)
        Y, X = np.meshgrid(np.linspace(0, 1, nx), np.linspace(0, 1, ny))
        g = lambda a, b, sg: np.exp(-((X - a) ** 2 + (Y - b) ** 2) / (2 * sg ** 2))
        r0 = g(.25, .25, .1)
        r1 = g(.25, .25, .05) + g(.25, .75, .05) + g(.75, .25, .05) + g(.75, .75, .05)
        rho0 = np.ascontiguousarray(((nx * ny / r0.sum()) * r0).T)
        rho1 = np.ascontiguousarray(((nx * ny / r1.sum()) * r1).T)
    N = nt * nx * ny
    ht, hx, hy = 1 / (nt - 1), 1 / (nx - 1), 1 / (ny - 1)
    c_first, c_last = (-rho0 / ht).ravel(), (rho1 / ht).ravel()           # initialize.m:41-44, C order (x, y)
    h = 1.0 / N
    hMean = h ** (1 / 3)
    norm_c = np.sqrt(h) * np.sqrt(np.dot(c_first, c_first) + np.dot(c_last, c_last)) * np.sqrt(nt)
    D = np.sqrt(2) * np.sqrt(hMean)
    E = D / np.sqrt(2)
    cScale = max(1.0, norm_c * np.sqrt(hMean))
    dScale = E * np.sqrt(2)
    tc0, tc1, tn0, tn1 = slab.partition(nt, world)[rank]
    P = nx * ny
    phi_plane = (0.5 * ((np.arange(nx) * hx)[:, None] ** 2 + (np.arange(ny) * hy)[None, :] ** 2)).ravel() * (1 / dScale)
    phi = np.tile(phi_plane, tn1 - tn0)
    c = np.zeros((tn1 - tn0) * P)
    if tn0 == 0:
        c[:P] = (1.0 / cScale) * c_first
    if tn1 == nt:
        c[-P:] = (1.0 / cScale) * c_last
    sz = slab.local_sizes(rank, world, nt, nx, ny)
    from types import SimpleNamespace
    var = SimpleNamespace(phi=phi, q=np.zeros(sz["Q"]), alpha=np.zeros(sz["Q"]), z=np.zeros((sz["L"], 10), order="F"),
                          beta=np.zeros((sz["L"], 10), order="F"), cScale=cScale, dScale=dScale, D=D, E=E)
    model = SimpleNamespace(nt=nt, nx=nx, ny=ny, dim=2, c=c, normc=norm_c / cScale, normd=np.sqrt(2) * E / dScale,
                            grad=(D / ht, D / hx, D / hy))
    return var, model


def touch_pages(arrays, threads=8):
    """Write every page of the (zero-initialised, lazily backed) host arrays once, in parallel, keeping their contents:
    the e2e leg then measures transfers into resident memory, as a caller's live arrays are, not first-touch page faults."""
    from concurrent.futures import ThreadPoolExecutor
    jobs = []
    for a in arrays:
        v = a.reshape(-1, order="A")
        step = max(1, v.size // 64)
        jobs += [v[i:i + step] for i in range(0, v.size, step)]
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(lambda piece: np.multiply(piece, 1.0, out=piece), jobs))


def level_opts(var, model, maxit, tol=1e-4):
    from dotsocp_b200 import solver
    opts = {"tol": tol, "maxit": maxit, "tau": 1.9, "sigma": 1.0, "ifCheckStepByStep": False, "scaling": True}
    return solver.make_level_opts("dot2d", "inPALM", var, opts, model)


def pick_workload(requested, world=1):
    if requested != "auto":
        return requested
    if world > 1:
        return "c5"    # 147 GB of state / world GPUs and 121 GB of host buffers / world ranks
    # the metric's grid needs ~147 GB of HBM (34 N doubles) and 121 GB of host buffers for the e2e leg
    try:
        out = subprocess.check_output(["nvidia-smi", "--query-gpu=memory.total,memory.used", "--format=csv,noheader,nounits", "-i", "0"], text=True)
        tot, used = [float(v) for v in out.strip().split(",")]
        gpu_free_gb = (tot - used) / 1024
    except Exception:
        gpu_free_gb = 0
    try:
        with open("/proc/meminfo") as f:
            mem = {l.split(":")[0]: float(l.split()[1]) / 1048576 for l in f}
        host_gb = mem.get("MemAvailable", 0)
    except Exception:
        host_gb = 0
    return "c5" if (gpu_free_gb >= 160 and host_gb >= 150) else "c4"


def host_mem_gb():
    try:
        with open("/proc/meminfo") as f:
            mem = {l.split(":")[0]: float(l.split()[1]) / 1048576 for l in f}
        return mem.get("MemAvailable", 0.0)
    except Exception:
        return 0.0


def cpu_reference_leg(steps, sample_grid, problem="example1"):
    """The reference's own CPU implementation of the path: genuine reference MEX binaries (oracle/_ref, single-threaded by
    construction) when present, else the bit-identical C restatement, driven by the numpy/scipy restatement of the MATLAB
    glue (scipy DCT with all host threads).  MATLAB/Octave are not available offline.  Times exactly `steps` REAL iterations
    on `sample_grid` (nodes) with the loop's own clock, after a small warm-up run that starts the thread pools."""
    from oracle import dotsocp_oracle as O
    from oracle import kernels as K
    cores = os.cpu_count() or 1

    def run(grid, k):
        nt, nx, ny = grid
        rho0, rho1 = O.get_example2d(problem, nx, ny)
        var, model = O.initialize2d(rho0, rho1, nt)
        O.InitialScaling(var, model, True, None, "dot2d")
        opts = {"tol": 1e-30, "maxit": k, "tau": 1.9, "sigma": 1.0, "ifCheckStepByStep": False, "scaling": True}
        O.solver_socp_inPALM(var, opts, model, workers=cores)
        # iterations only, like this repo's arm: the KKT checks the loop schedules (its own KKT timer, solver_socp_inPALM.m
        # :135,218,327) are not counted
        return float(var.time["Total_Time"]) - float(var.time["KKT"]), dict(var.time)
    run((17, 33, 33), 2)
    dt, tbl = run(sample_grid, steps)
    dt = max(dt, 1e-9)
    return steps / dt, cores, K.default_backend(), dt, tbl


def cpu_sample_grid(budget="reference"):
    """largest BASELINE grid whose CPU run fits the host: the oracle needs ~750 B per node (state + MATLAB-style temporaries +
    the sparse gradient), i.e. 51 GB at 512x512x256"""
    avail = host_mem_gb()
    if budget == "reference" and avail >= 80:
        return "c4"
    return "c3" if avail >= 12 else "c2"


def densities_matlab(nx, ny, problem="example1"):
    """examples/dot2d/gene_example1.m on a MATLAB-shaped (ny, nx) grid, mean 1 (input of the driver mirror)"""
    xs = np.linspace(0, 1, nx).reshape(1, nx)
    ys = np.linspace(0, 1, ny).reshape(ny, 1)
    r0 = np.exp(-0.5 * ((xs - 0.25) ** 2 + (ys - 0.75) ** 2) / 0.05)
    r1 = np.exp(-0.5 * ((xs - 0.75) ** 2 + (ys - 0.25) ** 2) / 0.05)
    r0 *= r0.size / r0.sum()
    r1 *= r1.size / r1.sum()
    return r0, r1


PARITY_ITERS = 12


def parity_rows(dp, slab, wl, rank, world, barrier):
    """12 inPALM iterations with ifCheckStepByStep from the reference's initial state of the mixture instance (example2) on
    the benchmarked grid; KKT rows against tests/golden/bench_rows.json.  Returns the comparison (rank 0 decides)."""
    nt, nx, ny = WORKLOADS[wl]
    var, model = make_problem(nt, nx, ny, rank, world, problem="example2")
    from dotsocp_b200 import solver
    opts = {"tol": 1e-4, "maxit": PARITY_ITERS, "tau": 1.9, "sigma": 1.0, "ifCheckStepByStep": True, "scaling": True}
    o = solver.make_level_opts("dot2d", "inPALM", var, opts, model)
    with dp.Session("dot2d", nt, nx, ny, rank=rank, world=world, nccl_id=slab.REUSE_COMM if world > 1 else None) as s:
        s.upload(var.phi, var.q, None, var.alpha, var.beta, model.c)
        barrier()
        hb, res = s.run(o)
    rows = hb.kkt[:res.hist_len]
    out = {"grid": WL_LABEL[wl], "instance": "example2 mixture, one level, 12 checked iterations", "rows": int(res.hist_len),
           "priVal": [float(v) for v in hb.priVal[:res.hist_len]]}
    path = os.path.join(ROOT, "tests", "golden", "bench_rows.json")
    gold = {}
    if os.path.exists(path):
        with open(path) as f:
            gold = json.load(f)
    g = gold.get(wl)
    if g is None:
        out.update(status="no golden for this grid", kkt=rows.tolist())
        return out
    ref = np.array(g["kkt"])
    diff = float(np.abs(rows - ref).max()) if ref.shape == rows.shape else float("inf")
    out.update(source=g["source"], max_abs_diff=diff, bitwise=bool(ref.shape == rows.shape and np.array_equal(rows, ref)),
               objective_rel_diff=float(abs(hb.priVal[res.hist_len - 1] - g["priVal"][-1]) / abs(g["priVal"][-1])),
               status="ok" if diff < (1e-8 if g["source"].startswith("cpu oracle") else 1e-10) else "MISMATCH")
    if out["status"] != "ok":
        raise SystemExit(f"parity check failed on {WL_LABEL[wl]}: max |kkt - golden| = {diff:.3e} ({g['source']})")
    return out


def same_grid_pair(dp):
    """One ratio that is a measurement end to end: the same 3-level solve of 128x128x64 to tol 1e-4 on the GPU (driver mirror,
    everything a caller waits for) and on the CPU path (oracle = reference MEX kernels + restated glue, all host threads)."""
    from oracle import dotsocp_oracle as O
    nt, n = 65, 129
    r0, r1 = densities_matlab(n, n)
    opts = {"tol": 1e-4, "maxit": 3000}
    tgs = []
    for _ in range(2):      # the first call after the large-grid legs also pays for the driver giving their memory back
        t0 = time.perf_counter()
        og, _, MLg, rhg = dp.solver_dotsocp2d(r0, r1, nt, 3, dict(opts), "inPALM")
        tgs.append(time.perf_counter() - t0)
    tg = min(tgs)
    t0 = time.perf_counter()
    oc, _, MLc, rhc = O.solver_dotsocp2d(r0, r1, nt, 3, dict(opts), "inPALM", workers=os.cpu_count() or 1)
    tc = time.perf_counter() - t0
    return {"workload": "128x128x64 example1, 3 levels, inPALM, tol 1e-4 (whole solver call)", "gpu_seconds": tg, "gpu_seconds_by_call": tgs, "cpu_seconds": tc,
            "cpu_over_gpu": tc / tg, "level_iters_gpu": [int(v) for v in og.level_iters], "level_iters_cpu": [int(v) for v in oc.level_iters],
            "objective_gpu": float(rhg.priVal[-1]), "objective_cpu": float(rhc.priVal[-1]),
            "kkt_history_max_abs_diff": float(np.abs(MLg.kkt - MLc.kkt).max()) if MLg.kkt.shape == MLc.kkt.shape else None}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="dotsocp_b200", choices=["dotsocp_b200", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto"] + list(WORKLOADS))
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-ttt", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    K_, W_ = args.steps, max(args.warmup, 3)
    hbm_peak, peak_src = measured_peaks()

    wl = pick_workload(args.workload, world) if rank == 0 or world == 1 else args.workload
    dist = None
    if world > 1 and args.impl == "reference":
        if rank != 0:
            return                                            # the CPU arm runs on rank 0 alone
        world = 1
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl")
        obj = [wl]
        dist.broadcast_object_list(obj, src=0)
        wl = obj[0]
    nt, nx, ny = WORKLOADS[wl]
    N, L, Q = sizes(nt, nx, ny)
    W_iter = (14 * N + 6 * Q + 30 * L) * 8.0          # algorithmic bytes per inPALM iteration
    config = {"workload": f"2D DOT {WL_LABEL[wl]} cells = ({nt},{nx},{ny}) nodes, example1 Gaussian->Gaussian, inPALM tau=1.9",
              "grid_nodes": [nt, nx, ny], "algorithm": "inPALM", "partition": f"time-slab x{world}",
              "l2": "state (>= 15 GB) far larger than the 126 MB L2; no flush needed"}

    # ---------------------------------------------------------------------------------------------- reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        # The CPU path cannot hold the named grid (>= 400 GB, ~4 min per iteration): it is MEASURED on the largest BASELINE grid
        # that fits this host -- 512x512x256 where there is memory for it, ~30 s per iteration -- for min(K, 3) real iterations,
        # and the line's value is that measurement scaled by the node count to the named grid (flagged, both numbers given).
        sw = cpu_sample_grid("reference")
        sample = WORKLOADS[sw]
        k_timed = max(1, min(K_, 3))
        its, cores, backend, dt, tbl = cpu_reference_leg(k_timed, sample)
        Ns = sample[0] * sample[1] * sample[2]
        value = its * Ns / N
        config["workload"] += f" | CPU arm TIMED on {WL_LABEL[sw]} cells ({k_timed} iterations), scaled by node count x{N / Ns:.2f}"
        config["timed_grid_nodes"] = list(sample)
        measured = {"grid": WL_LABEL[sw], "grid_nodes": list(sample), "iterations": k_timed, "seconds": dt, "value": its,
                    "unit": "iterations/s", "ms_per_step": 1e3 / its,
                    "step_seconds": {k: float(v) for k, v in tbl.items() if k != "Iters"}}
        line = {"impl": "reference", "metric": "ADMM iters/sec", "value": value, "unit": "iterations/s", "n_gpus": args.gpus,
                "steps": k_timed, "steps_requested": K_, "warmup": W_, "ms_per_step": 1e3 / value, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "extrapolated": sw != wl, "measured": measured,
                "cpu_baseline": {"value": value, "unit": "iterations/s", "cores": cores,
                                 "kind": "reference" if backend == "ref" else "port",
                                 "sample": f"{k_timed} real inPALM iterations on the {WL_LABEL[sw]} grid ({its:.4f} it/s = {1e3 / its:.0f} ms per "
                                           f"iteration MEASURED), scaled by node count {Ns}/{N} to the named grid; native kernels = "
                                           f"{'genuine reference MEX binaries (1 thread each)' if backend == 'ref' else backend + ' restatement'}, "
                                           f"MATLAB glue restated in numpy/scipy (DCT on {cores} threads); MATLAB/Octave unavailable offline"},
                "e2e": {"value": value, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return

    # ---------------------------------------------------------------------------------------------- this repo's arm
    import dotsocp_b200 as dp
    from dotsocp_b200 import _lib
    _lib.check(_lib.lib().dotsocp_set_device(local_rank))
    ident = None
    from dotsocp_b200 import slab
    if world > 1:
        import torch
        torch.cuda.set_device(local_rank)
        ident = slab.broadcast_unique_id(dist, rank)

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(v):
        if world == 1:
            return v
        import torch
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    var, model = make_problem(nt, nx, ny, rank, world)
    o = level_opts(var, model, K_)

    sess = dp.Session("dot2d", nt, nx, ny, rank=rank, world=world, nccl_id=ident)
    sess.upload(var.phi, var.q, var.z, var.alpha, var.beta, model.c)
    sess.iter_begin(o)
    sess.iterate(W_)                                         # warm-up (untimed)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    launches0 = sess.launches
    t_wall0 = time.perf_counter()
    ms, per_kernel = sess.iterate(K_, per_kernel=True)       # CUDA events on the launching stream
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = sess.launches - launches0
    clocks = sampler.stop()
    sess.iter_end()
    sess.close()
    ms = max_over_ranks(ms)
    per_kernel = [max_over_ranks(v) for v in per_kernel]
    ms_per_step = ms / K_
    value = 1e3 / ms_per_step

    # roofline of the dominant kernel (k_mult: projection + multiplier step + next rhs/q2), measured live with events
    mult_ms = per_kernel[2] / K_
    mult_bytes = (N + 4 * Q + 20 * L) * 8.0 / world          # per GPU: reads q_old,q_new,alpha,beta ; writes beta,q2,rhs
    ach = mult_bytes / (mult_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "k_mult", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                "traffic": measured_traffic("k_mult", wl, world), "peak_source": peak_src,
                "per_kernel_ms": {"poisson": per_kernel[0] / K_, "k_qstep": per_kernel[1] / K_, "k_mult": mult_ms}}
    it_ach = W_iter / (ms_per_step * 1e-3) / 1e9
    roofline_iter = {"bound": "hbm", "achieved": it_ach, "peak": hbm_peak * world, "unit": "GB/s", "frac": it_ach / (hbm_peak * world),
                     "bytes_per_iteration": W_iter, "note": "W_iter = (14N+6Q+30L)*8 B, SURVEY.md §8(d)"}

    # e2e: the reference-facing call with HOST buffers (upload + K iterations incl. the KKT checks the loop schedules +
    # download inside the timed region); every rank moves its own slab
    e2e = None
    if not args.no_e2e:
        o2 = level_opts(var, model, K_, tol=1e-30)            # run exactly K iterations
        ident2 = slab.REUSE_COMM if world > 1 else None      # re-use the (warm) process-wide communicator
        f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64) if a.ndim == 1 else np.asfortranarray(a, dtype=np.float64)
        var.phi, var.q, var.z, var.alpha, var.beta = (f64(np.asarray(a)) for a in (var.phi, var.q, var.z, var.alpha, var.beta))
        e2e_out = (var.phi, var.q, var.z, var.alpha, var.beta)
        touch_pages(e2e_out)   # np.zeros pages are not backed until written: make the buffers resident before timing
        barrier()
        t0 = time.perf_counter()
        ph = {}
        s2 = dp.Session("dot2d", nt, nx, ny, rank=rank, world=world, nccl_id=ident2)
        ph["create"] = time.perf_counter() - t0
        t1 = time.perf_counter()
        s2.upload(var.phi, var.q, None, var.alpha, var.beta, model.c)   # inPALM never reads the incoming z (as solve_level)
        ph["upload"] = time.perf_counter() - t1
        t1 = time.perf_counter()
        _, res2 = s2.run(o2)
        ph["run"] = time.perf_counter() - t1
        t1 = time.perf_counter()
        out_state = s2.download(out=e2e_out)   # in place, like the reference's MEX calls / dotsocp_solve_level
        ph["download"] = time.perf_counter() - t1
        t1 = time.perf_counter()
        s2.close()
        ph["destroy"] = time.perf_counter() - t1
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0)
        del out_state
        h2d = (2 * N + 2 * Q + 10 * L) * 8.0
        d2h = (N + 2 * Q + 20 * L) * 8.0
        e2e = {"value": K_ / dt, "unit": "iterations/s", "h2d_bytes_per_step": h2d / K_, "d2h_bytes_per_step": d2h / K_,
               "seconds": dt, "seconds_by_phase_rank0": {k: round(v, 4) for k, v in ph.items()},
               "run_device_seconds_by_step": [round(float(x), 4) for x in res2.times], "call": "dotsocp_create+upload+run+download (= dotsocp_solve_level / solver_socp_inPALM), "
                                     "pageable host buffers updated in place, all ranks"}
    # time-to-tolerance (second half of the BASELINE metric): one inPALM level solve of the 256x256x128 instance from the
    # reference's initial state to opts.tol = 1e-4, KKT checks on the reference schedule, state resident in HBM
    ttt = None
    if not args.no_ttt:
        tnt, tnx, tny = WORKLOADS["c3"]
        tv, tm = make_problem(tnt, tnx, tny, rank, world)
        to = level_opts(tv, tm, 3000, tol=1e-4)
        identt = slab.REUSE_COMM if world > 1 else None
        with dp.Session("dot2d", tnt, tnx, tny, rank=rank, world=world, nccl_id=identt) as s3:
            s3.upload(tv.phi, tv.q, tv.z, tv.alpha, tv.beta, tm.c)
            barrier()
            t0 = time.perf_counter()
            hb3, r3 = s3.run(to)
            barrier()
            dt3 = max_over_ranks(time.perf_counter() - t0)
        ttt = {"workload": "256x256x128 example1, single level, tol 1e-4", "iterations": int(r3.iters), "seconds": dt3,
               "iters_per_sec_incl_kkt": r3.iters / dt3, "kkt_checks": int(r3.hist_len),
               "final_kkt_max": float(np.max(hb3.kkt[r3.hist_len - 1][[0, 2, 5, 6]])),
               "objective": float(hb3.priVal[r3.hist_len - 1]),
               "device_seconds_by_step": {"FFT": r3.times[0], "Q_Step": r3.times[2], "ProjSOC+Multiplier": r3.times[3], "KKT": r3.times[4]}}
        # whole multilevel solves through the driver mirror (solver_dotsocp2d, 3 levels, reference defaults, tol 1e-4): state
        # resident in HBM across the levels (transitions and output recovery on the device, time slabs on `world` GPUs); the
        # clock covers everything a caller waits for: set-up of the coarsest level, upload, the three level solves with their
        # KKT checks, the transitions, and the download of the six output fields.  512x512x256 on any number of GPUs, the
        # metric's own 1024x1024x512 when the grid fits (8 GPUs; one GPU needs 9/8 of 147 GB during the last transition).
        slabs_opt = {"rank": rank, "world": world, "nccl_id": slab.REUSE_COMM} if world > 1 else None

        def solve_to_tol(wname):
            snt, snx, sny = WORKLOADS[wname]
            r0, r1 = densities_matlab(snx, sny)
            barrier()
            t0 = time.perf_counter()
            o, tml, ML_, rh_ = dp.solver_dotsocp2d(r0, r1, snt, 3, {"tol": 1e-4, "maxit": 3000, "slabs": slabs_opt}, "inPALM")
            barrier()
            dt_ = max_over_ranks(time.perf_counter() - t0)
            return {"workload": f"{WL_LABEL[wname]} example1, 3 levels, inPALM, tol 1e-4", "seconds": dt_,
                    "level_iters": [int(v) for v in o.level_iters], "kkt_checks": int(ML_.len),
                    "final_kkt": [float(v) for v in ML_.kkt[-1]], "final_kkt_max": float(np.max(ML_.kkt[-1][[0, 2, 5, 6]])),
                    "objective": float(rh_.priVal[-1]), "w2_cost": float(o.w2), "mass_ok": bool(o.massOK),
                    "level_device_seconds": [float(t["Total_Time"]) for t in tml[:3]]}
        ttt["multilevel_512x512x256"] = solve_to_tol("c4")
        if world >= 4:
            ttt["multilevel_1024x1024x512"] = solve_to_tol("c5")
        elif world == 1 and wl == "c5":
            try:
                ttt["multilevel_1024x1024x512"] = solve_to_tol("c5")
            except dp.DotsocpError as e:       # the last transition keeps both levels alive: may not fit one GPU
                ttt["multilevel_1024x1024x512"] = {"unavailable": str(e)[:200]}
        if world == 1:
            r0, r1 = densities_matlab(tnx, tny)
            ml = {}
            for mode in ("resident", "host"):
                t0 = time.perf_counter()
                out_ml, _, ML_ml, _ = dp.solver_dotsocp2d(r0, r1, tnt, 3, {"tol": 1e-4, "maxit": 3000, "resident": mode == "resident"},
                                                          "inPALM")
                ml[mode] = {"seconds": time.perf_counter() - t0, "level_iters": [int(v) for v in out_ml.level_iters],
                            "final_kkt_max": float(np.max(ML_ml.kkt[-1][[0, 2, 5, 6]]))}
            ttt["multilevel_3_levels"] = ml

    # iterations/s with the KKT checks on the reference schedule: 200 iterations of the loop itself (check at
    # iteration 1, 4, ... every 15 by the end: 21 checks, sigma updates and rescalings included), device time of the whole call
    sched = None
    parity = None
    if not args.no_ttt:
        o200 = level_opts(var, model, 200, tol=1e-30)
        with dp.Session("dot2d", nt, nx, ny, rank=rank, world=world, nccl_id=slab.REUSE_COMM if world > 1 else None) as s4:
            s4.upload(var.phi, var.q, None, var.alpha, var.beta, model.c)
            barrier()
            hb4, r4 = s4.run(o200)
            sec = max_over_ranks(float(r4.times[5]))
            sched = {"iterations": int(r4.iters), "kkt_checks": int(r4.hist_len), "device_seconds": sec, "value": r4.iters / sec,
                     "unit": "iterations/s", "kkt_seconds": float(r4.times[4]),
                     "frac_of_unchecked_rate": (r4.iters / sec) / value}
        # parity at the benchmarked grid: 12 checked iterations from the reference's initial state against the committed rows
        # (tests/golden/bench_rows.json: the CPU oracle's rows at 512x512x256, the single-GPU rows at 1024x1024x512 -- which
        # no CPU here can hold); the reductions are independent of the GPU count, so the rows must agree bit for bit
        parity = parity_rows(dp, slab, wl, rank, world, barrier)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cpu = None
    same_grid = None
    if not args.no_cpu and world == 1:
        # CPU path measured (not extrapolated) on a grid it finishes in ~20 s: 3 iterations of 256x256x128
        same_grid = same_grid_pair(dp)      # (before the CPU-only leg: its thread pools keep spinning for a while)
        sw = cpu_sample_grid("baseline")
        sg = WORKLOADS[sw]
        its, cores, backend, dt, tbl = cpu_reference_leg(3, sg)
        Ns = sg[0] * sg[1] * sg[2]
        cpu = {"value": its * Ns / N, "unit": "iterations/s", "cores": cores, "kind": "reference" if backend == "ref" else "port",
               "extrapolated": sw != wl, "measured": {"grid": WL_LABEL[sw], "iterations": 3, "seconds": dt, "value": its, "unit": "iterations/s"},
               "sample": f"3 real inPALM iterations on the {WL_LABEL[sw]} grid ({its:.4f} it/s MEASURED), value = that scaled by node count "
                         f"{Ns}/{N} to the named grid; native kernels = {'genuine reference MEX binaries' if backend == 'ref' else backend}; "
                         f"glue = numpy/scipy restatement (MATLAB unavailable offline)"}

    line = {"metric": "ADMM iters/sec", "value": value, "unit": "iterations/s", "n_gpus": world, "steps": K_, "warmup": W_,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config, "clocks": clocks, "gpu_launches": launches, "wall_ms_per_step": 1e3 * t_wall / K_,
            "roofline": roofline, "roofline_iteration": roofline_iter, "e2e": e2e, "time_to_tol": ttt,
            "value_with_kkt_schedule": sched, "parity": parity, "cpu_baseline": cpu, "same_grid_pair": same_grid}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
