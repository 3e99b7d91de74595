"""Solver-level mirror of the reference's algorithms/*.m entry points, running on the GPU.

    [runHist, sigma] = solver_socp_inPALM(var, opts, model)      socp/dot2d/algorithms/solver_socp_inPALM.m:1
                       solver_socp_PALM / solver_socp_accADMM     socp/dot2d/algorithms/*.m
                       solver_wsocp_inPALM / solver_wsocp_accADMM socp/wdot2d/algorithms/*.m
                       (1-D) solver_socp_inPALM                   socp/dot1d/algorithms/solver_socp_inPALM.m:1

Same argument meaning and side effects as the reference: ``var`` (phi, q, z, alpha, beta, cScale, dScale, D, E) is
mutated in place, ``var.alpha``/``var.beta`` come back multiplied by sigma (:335-336), ``var.time`` carries the
step-time table (:339-341), the returned sigma is un-rescaled (:357) and ``runHist`` has kkt/time/iter/pdGap/len
(+ priVal/dualVal as extras).  ``var``/``model`` are any attribute containers (``types.SimpleNamespace`` works).

``Session`` keeps the state resident in HBM (what the multilevel driver and the benchmark use).
"""
from __future__ import annotations

import ctypes as C
from types import SimpleNamespace

import numpy as np

from ._lib import METHOD, VARIANT, HistBuffers, LevelOpts, LevelResult, ProlongScal, RecoverScal, check, lib, ptr

TIME_NAMES = {
    0: ["Step_1_1_FFT", "Step_1_2_ProjSOC", "Step_2_Q_Step", "Step_3_Multiplier", "KKT", "Total_Time"],
    1: ["Step_1_Q_Step", "Step_2_1_FFT", "Step_2_2_ProjSOC", "Step_3_Q_Step", "Step_4_Multiplier", "KKT", "Total_Time"],
    2: ["Step_1_Q_Step", "Step_2_Multiplier", "Step_3_1_FFT", "Step_3_2_ProjSOC", "KKT", "Interp", "Total_Time"],
    3: ["Step_1_1_sGS", "Step_1_2_ProjSOC", "Step_2_Q_Step", "Step_3_Multiplier", "KKT", "Total_Time"],
    4: ["Step_1_1_sGS", "Step_1_2_ProjSOC", "Step_2_Multiplier", "Step_3_Q_Step", "Step_4_Interp", "KKT", "Total_Time"],
}
METHOD_NAMES = {0: "Inexact Proximal ALM", 1: "Proximal ALM", 2: "Accelerated ADMM", 3: "Symmetric Gauss-seidel based inPALM",
                4: "Accelerated symmetric Gauss-Seidel based ADMM"}


def _get(opts, name, default=None):
    if isinstance(opts, dict):
        return opts.get(name, default)
    return getattr(opts, name, default)


def grad_scalars(model):
    """The three magnitudes of model.grad (already multiplied by D in InitialScaling): D*(1/ht), D*(1/hx), D*(1/hy).
    The reference stores them in a sparse matrix (initialize.m:35-39,67-87); this path only needs the values."""
    g = getattr(model, "grad", None)
    if isinstance(g, (tuple, list)) and len(g) == 3:
        return tuple(float(v) for v in g)
    if g is not None and hasattr(g, "tocsr"):  # a scipy sparse matrix built like the reference's
        nt, nx, ny = model.nt, model.nx, getattr(model, "ny", 1) or 1
        L = (nt - 1) * nx * ny
        nbx = nt * (nx - 1) * ny
        g = g.tocsr()
        gt = abs(g[0].data).max()
        gx = abs(g[L].data).max()
        gy = abs(g[L + nbx].data).max() if ny > 1 else 0.0
        return float(gt), float(gx), float(gy)
    raise ValueError("model.grad must be (grad_t, grad_x, grad_y) or the reference's sparse gradient")


def make_level_opts(variant, method, var, opts, model):
    nt, nx = int(model.nt), int(model.nx)
    ny = 1 if variant == "dot1d" else int(model.ny)
    gt, gx, gy = grad_scalars(model)
    o = LevelOpts()
    o.variant = VARIANT[variant]
    o.method = METHOD[method]
    o.nt, o.nx, o.ny = nt, nx, ny
    o.maxit = int(_get(opts, "maxit"))
    o.ifCheckStepByStep = 1 if _get(opts, "ifCheckStepByStep", False) else 0
    o.scaling = 1 if _get(opts, "scaling", False) else 0
    cpd = _get(opts, "checkPrimDualFeas", None)
    o.checkPrimDualFeas = -1 if cpd is None else (1 if cpd else 0)
    o.restart = int(_get(opts, "restart", 0) or 0)
    o.tau = float(_get(opts, "tau", 1.0) or 1.0)
    o.sigma = float(_get(opts, "sigma"))
    o.tol = float(_get(opts, "tol"))
    tl = _get(opts, "time_limit", None)          # NaN = absent (3600 s); <= 0 = budget already spent (one iteration + check)
    o.time_limit = float("nan") if tl is None else float(tl)
    o.rho = float(_get(opts, "rho", 0) or 0)
    o.theta = float(_get(opts, "theta", 0) or 0)
    o.cScale, o.dScale, o.D, o.E = float(var.cScale), float(var.dScale), float(var.D), float(var.E)
    o.normc = float(model.normc)
    o.normd = float(getattr(model, "normd", 0.0) or 0.0)
    o.grad_t, o.grad_x, o.grad_y = gt, gx, gy
    return o


class Weights:
    """Pyramid of level weights resident on the device (dotsocp_weights_*): what the weighted drivers do with the weight
    before the first level starts -- the generators' time replication (gene_weight_circle.m:24-27, get_weight_by_barrier.m:
    30-33), the restriction chain downSample_q.m / downSample_barrier.m and mean(log10(weight + 1e-10)) of
    solver_wdotsocp2d.m:312-316 -- without a Q-sized host array per level.  Level 0 is the finest grid."""

    def __init__(self, nt, nx, ny, levels):
        self._h = C.c_void_p()
        check(lib().dotsocp_weights_create(C.byref(self._h), int(nt), int(nx), int(ny), int(levels)))
        self.levels = int(levels)

    def dims(self, level):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        check(lib().dotsocp_weights_dims(self._h, int(level), C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def size(self, level):
        nt, nx, ny = self.dims(level)
        return (nt - 1) * nx * ny + nt * (nx - 1) * ny + nt * nx * (ny - 1)

    def set(self, weight):
        """finest level from a host array (Q doubles, [q0 | bx | by])"""
        w = np.ascontiguousarray(weight, dtype=np.float64).reshape(-1)
        assert w.size == self.size(0), f"weight has {w.size} entries, expected {self.size(0)}"
        check(lib().dotsocp_weights_set(self._h, ptr(w)))
        return self

    def set_planes(self, weightX, weightY):
        """finest level from the generators' two planes, MATLAB-shaped weightX (ny, nx-1) and weightY (ny-1, nx)"""
        nt, nx, ny = self.dims(0)
        wx = np.ascontiguousarray(np.asarray(weightX, dtype=np.float64).T).reshape(-1)     # C order (x, y)
        wy = np.ascontiguousarray(np.asarray(weightY, dtype=np.float64).T).reshape(-1)
        assert wx.size == (nx - 1) * ny and wy.size == nx * (ny - 1)
        check(lib().dotsocp_weights_set_planes(self._h, ptr(wx), ptr(wy)))
        return self

    def restrict(self, geometric=False):
        """levels 1 .. from level 0: downSample_q, or downSample_barrier (exp of the restricted log) when geometric"""
        check(lib().dotsocp_weights_restrict(self._h, 1 if geometric else 0))
        return self

    def get(self, level):
        out = np.empty(self.size(level))
        check(lib().dotsocp_weights_get(self._h, int(level), ptr(out)))
        return out

    def log10_mean(self, level):
        m = C.c_double()
        check(lib().dotsocp_weights_log10_mean(self._h, int(level), C.byref(m)))
        return m.value

    @property
    def gpu_launches(self):
        return float(lib().dotsocp_weights_launch_count(self._h))

    def level(self, level):
        return DeviceWeight(self, int(level))

    def close(self):
        if self._h:
            lib().dotsocp_weights_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DeviceWeight:
    """One level of a Weights pyramid, accepted wherever the drivers take model.weight (Session.upload / prolong_from)."""

    def __init__(self, pyramid, level):
        self.pyramid, self.level = pyramid, level

    def log10_mean(self):
        return self.pyramid.log10_mean(self.level)


class OutputBuffers:
    """Host arrays for Session.recover, allocated and touched page by page on background threads while the device is busy
    with the last level: a fresh 26 GB of output at 1024x1024x512 otherwise takes its first-touch page faults inside the
    device-to-host copies (1.6 s instead of 0.7 s).  get() joins the threads and returns the dict of arrays."""

    def __init__(self, session, fields=("rho", "Ex", "Ey", "q0", "bx", "by"), threads=4):
        import threading
        self._shapes = session.output_shapes(fields)
        self._out = {}
        self._thread = threading.Thread(target=self._fill, args=(threads,), daemon=True)
        self._thread.start()

    def _fill(self, threads):
        try:
            self._fill_impl(threads)
        except Exception:          # (e.g. MemoryError) recover() then allocates its own arrays and reports the real problem
            self._out = None

    def _fill_impl(self, threads):
        from concurrent.futures import ThreadPoolExecutor
        jobs = []
        for name, shp in self._shapes.items():
            a = np.empty(shp)
            self._out[name] = a
            v = a.reshape(-1)
            step = max(1 << 20, -(-v.size // 16))
            jobs += [(v, i, min(i + step, v.size)) for i in range(0, v.size, step)]
        touch = lambda job: job[0][job[1]:job[2]].fill(0.0)        # numpy releases the GIL inside fill
        if sum(j - i for _, i, j in jobs) < (1 << 22):              # small outputs: not worth a pool
            for job in jobs:
                touch(job)
        else:
            with ThreadPoolExecutor(threads) as ex:
                list(ex.map(touch, jobs))

    def get(self):
        self._thread.join()
        return self._out


class Session:
    """Device-resident state of one level (dotsocp_create / _upload / _run / _download / _destroy)."""

    def __init__(self, variant, nt, nx, ny=1, rank=0, world=1, nccl_id=None):
        """world > 1 with nccl_id=None: all `world` time slabs live in this process on the current device (emulation of
        the multi-GPU path, global host arrays); with a 128-byte nccl_id: this process owns slab `rank` (one process per
        GPU, NCCL between them) and upload/download take the slab-local parts (see slab.py)."""
        self._h = C.c_void_p()
        check(lib().dotsocp_create(C.byref(self._h), VARIANT[variant], int(nt), int(nx), int(ny), rank, world, nccl_id))
        self._describe(variant, nt, nx, ny, rank, world, nccl_id is not None, None)

    def _describe(self, variant, nt, nx, ny, rank, world, distributed, cuts):
        from .slab import default_cuts, local_sizes
        self.variant = variant
        self.rank, self.world, self.distributed = int(rank), int(world), bool(distributed)
        self.nt, self.nx, self.ny = int(nt), int(nx), int(ny)
        self.ncol = 6 if variant == "dot1d" else 10
        self.cuts = default_cuts(self.nt, self.world) if cuts is None else list(cuts)
        self.L = (self.nt - 1) * self.nx * self.ny
        self.Q = self.L + self.nt * (self.nx - 1) * self.ny + self.nt * self.nx * (self.ny - 1)
        self.N = self.nt * self.nx * self.ny
        self.nt_local, self.nc_local = self.nt, self.nt - 1
        if self.distributed:   # host arrays hold this slab's owned part only (slab.py)
            sz = local_sizes(self.rank, self.world, self.nt, self.nx, self.ny, self.cuts)
            self.L, self.Q, self.N = sz["L"], sz["Q"], sz["N"]
            P = self.nx * self.ny
            self.nt_local, self.nc_local = self.N // P, self.L // P

    @classmethod
    def refined(cls, coarse):
        """Fresh session of the next finer level (2n-1 nodes per refined axis) with the rank, world, communicator and --
        every cut doubled -- the time partition of `coarse` (dotsocp_create_refined): the target of prolong_from."""
        s = cls.__new__(cls)
        s._h = C.c_void_p()
        check(lib().dotsocp_create_refined(C.byref(s._h), coarse._h))
        ny = 2 * coarse.ny - 1 if coarse.ny > 1 else 1
        s._describe(coarse.variant, 2 * coarse.nt - 1, 2 * coarse.nx - 1, ny, coarse.rank, coarse.world, coarse.distributed,
                    [2 * v for v in coarse.cuts])
        return s

    def close(self):
        if self._h:
            lib().dotsocp_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload(self, phi, q, z, alpha, beta, c, weight=None):
        f = lambda a: np.ascontiguousarray(a, dtype=np.float64) if a.ndim == 1 else np.asfortranarray(a, dtype=np.float64)
        arrs = [None if a is None else f(np.asarray(a)) for a in (phi, q, z, alpha, beta, c)]
        assert arrs[0].size == self.N and arrs[1].size == self.Q and arrs[3].size == self.Q and arrs[5].size == self.N
        assert arrs[4].shape == (self.L, self.ncol)
        assert arrs[2] is None or arrs[2].shape == (self.L, self.ncol)   # z=None: inPALM never reads the incoming z
        assert (weight is None) == (self.variant != "wdot2d"), "weight is required for (and only for) the weighted variant"
        if isinstance(weight, DeviceWeight):      # resident pyramid: device-to-device, nothing crosses PCIe
            self.set_weight(weight)
            weight = None
        w = None if weight is None else np.ascontiguousarray(weight, dtype=np.float64).reshape(-1)
        assert w is None or w.size == self.Q, f"weight has {0 if w is None else w.size} entries, expected {self.Q}"
        check(lib().dotsocp_upload(self._h, *[ptr(a) for a in arrs], ptr(w)))

    def set_weight(self, dw):
        """weight of this (wdot2d) session from a level of a device pyramid (dotsocp_set_weight); the slab takes its own part"""
        check(lib().dotsocp_set_weight(self._h, dw.pyramid._h, dw.level))

    def download(self, out=None):
        """out = (phi, q, z, alpha, beta): fill these float64 arrays IN PLACE (the convention of the reference's MEX
        kernels and of dotsocp_solve_level); default: fresh arrays."""
        if out is not None:
            phi, q, z, alpha, beta = out
            for a, n in ((phi, self.N), (q, self.Q), (alpha, self.Q)):
                assert a.dtype == np.float64 and a.size == n and a.flags.c_contiguous and a.flags.writeable
            for a in (z, beta):
                assert a.dtype == np.float64 and a.shape == (self.L, self.ncol) and a.flags.f_contiguous and a.flags.writeable
        else:
            phi = np.empty(self.N)
            q = np.empty(self.Q)
            alpha = np.empty(self.Q)
            z = np.empty((self.L, self.ncol), order="F")
            beta = np.empty((self.L, self.ncol), order="F")
        check(lib().dotsocp_download(self._h, ptr(phi), ptr(q), ptr(z), ptr(alpha), ptr(beta)))
        return phi, q, z, alpha, beta

    def prolong_from(self, coarse, scal, c=None, weight=None, c_first=None, c_last=None):
        """Level transfer on the device (dotsocp_prolong): fill this fresh session of the refined grid (Session.refined for
        time slabs) from the finished `coarse` session.  scal: dict of the ProlongScal fields; the fine model.c (already
        divided by cScale) either whole (c) or as its two non-zero planes (c_first, c_last: nx*ny doubles each; a slab
        that owns neither the first nor the last time level may pass None); weight follows the upload convention."""
        P = self.nx * self.ny
        if c is not None:
            c = np.ascontiguousarray(c, dtype=np.float64)
            assert c.size == self.N and not self.distributed
            c_first, c_last = c[:P], c[-P:]
            assert not c[P:-P].any(), "model.c has a non-zero interior entry: unsupported"
        first = None if c_first is None else np.ascontiguousarray(c_first, dtype=np.float64).reshape(-1)
        last = None if c_last is None else np.ascontiguousarray(c_last, dtype=np.float64).reshape(-1)
        assert (first is None or first.size == P) and (last is None or last.size == P)
        if isinstance(weight, DeviceWeight):
            self.set_weight(weight)
            weight = None
        w = None if weight is None else np.ascontiguousarray(weight, dtype=np.float64).reshape(-1)
        assert w is None or w.size == self.Q, f"weight has {0 if w is None else w.size} entries, expected {self.Q}"
        ps = ProlongScal(**{k: float(v) for k, v in scal.items()})
        check(lib().dotsocp_prolong(coarse._h, self._h, C.byref(ps), ptr(first), ptr(last), ptr(w)))

    def output_shapes(self, fields=("rho", "Ex", "Ey", "q0", "bx", "by")):
        """shapes of the recovered fields (this slab's levels in a distributed session)"""
        shp = (self.nx, self.ny) if self.variant != "dot1d" else (self.nx,)
        return {name: ((self.nt_local if name in ("rho", "Ex", "Ey") else self.nc_local),) + shp
                for name in ("rho", "Ex", "Ey", "q0", "bx", "by")
                if name in fields and not (self.variant == "dot1d" and name in ("Ey", "by"))}

    def recover(self, alpha_recover, q_recover, rho0, rho1, fields=("rho", "Ex", "Ey", "q0", "bx", "by"), stats=True, out=None):
        """Output recovery on the device (dotsocp_recover) from the state of a finished run: recoverOrgVar + recover_RhoE +
        recover_q + check_massConservation + transport cost.  rho0 / rho1: MATLAB-shaped (ny, nx) densities [1-D: (nx,)].
        Returns (dict of C-order arrays (levels, nx, ny) -- this slab's levels in a distributed session --, sumRho,
        sumNegRho, w2cost)."""
        r0 = np.ascontiguousarray(np.asarray(rho0, dtype=np.float64).ravel(order="F"))
        r1 = np.ascontiguousarray(np.asarray(rho1, dtype=np.float64).ravel(order="F"))
        P = self.nx * self.ny
        assert r0.size == P and r1.size == P
        shapes = self.output_shapes(fields)
        if out is None:
            out = {name: np.empty(shp) for name, shp in shapes.items()}
        else:       # caller-provided arrays (OutputBuffers): filled in place
            assert set(out) == set(shapes)
            for name, shp in shapes.items():
                assert out[name].shape == shp and out[name].dtype == np.float64 and out[name].flags.c_contiguous, name
        sr = np.empty(self.nt) if stats else None
        sn = np.empty(self.nt) if stats else None
        w2 = C.c_double(float("nan"))
        rs = RecoverScal(float(alpha_recover), float(q_recover))
        check(lib().dotsocp_recover(self._h, C.byref(rs), ptr(r0), ptr(r1), *[ptr(out.get(k)) if out.get(k) is None else out[k].ctypes.data
                                                                               for k in ("rho", "Ex", "Ey", "q0", "bx", "by")],
                                    ptr(sr), ptr(sn), C.byref(w2) if stats else None))
        return out, sr, sn, (w2.value if stats else None)

    def run(self, level_opts):
        hb = HistBuffers(level_opts.maxit)
        res = LevelResult()
        check(lib().dotsocp_run(self._h, C.byref(level_opts), C.byref(hb.c), C.byref(res)))
        return hb, res

    # benchmark primitives -------------------------------------------------------------------------------------
    def iter_begin(self, level_opts):
        check(lib().dotsocp_iter_begin(self._h, C.byref(level_opts)))

    def iterate(self, n, per_kernel=False, kkt_every=0):
        """n iterations; kkt_every = k > 0: every k-th one is a check iteration (fused KKT sums + reduction + read-back)"""
        ms = C.c_float()
        k = (C.c_float * 4)()
        check(lib().dotsocp_iterate(self._h, int(n), int(kkt_every), C.byref(ms), k if per_kernel else None))
        return (ms.value, list(k)) if per_kernel else ms.value

    def iter_end(self):
        check(lib().dotsocp_iter_end(self._h))

    @property
    def launches(self):
        return lib().dotsocp_launch_count(self._h)


def _finish(var, method_id, hb, res):
    n = res.hist_len
    runHist = SimpleNamespace(kkt=hb.kkt[:n].copy(), time=hb.time[:n].copy(), iter=hb.iter[:n].copy(),
                              pdGap=hb.pdGap[:n].copy(), priVal=hb.priVal[:n].copy(), dualVal=hb.dualVal[:n].copy(), len=n)
    names = TIME_NAMES[method_id]
    var.time = {nm: res.times[i] for i, nm in enumerate(names)}
    var.time["Iters"] = res.iters
    var.name = METHOD_NAMES[method_id]
    var.cScale, var.dScale, var.D, var.E = res.cScale, res.dScale, res.D, res.E
    var.gpu_launches = res.gpu_launches
    return runHist, res.sigma


def _solve(variant, method, var, opts, model):
    o = make_level_opts(variant, method, var, opts, model)
    weight = getattr(model, "weight", None) if variant == "wdot2d" else None
    if variant == "wdot2d" and weight is None:
        raise ValueError("solver_wsocp_*: model.weight is required")
    with Session(variant, o.nt, o.nx, o.ny) as s:
        # inPALM overwrites z (solver_socp_inPALM.m:199) before reading it: the incoming z stays on the host
        z_dead = o.method in (METHOD["inPALM"], METHOD["sGS-inPALM"]) and o.maxit >= 1
        s.upload(var.phi, var.q, None if z_dead else var.z, var.alpha, var.beta, model.c, weight)
        hb, res = s.run(o)
        var.phi, var.q, var.z, var.alpha, var.beta = s.download()
    return _finish(var, o.method, hb, res)


def _variant_of(model):
    if getattr(model, "weight", None) is not None:
        return "wdot2d"
    return "dot1d" if (getattr(model, "ny", None) in (None, 1) or getattr(model, "dim", 2) == 1) else "dot2d"


def solver_socp_inPALM(var, opts, model):
    """socp/dot2d/algorithms/solver_socp_inPALM.m:1 and socp/dot1d/algorithms/solver_socp_inPALM.m:1
    (ALG2 is this loop with opts.tau = 1, solver_dotsocp2d.m:133-137)."""
    v = _variant_of(model)
    if v == "wdot2d":
        raise ValueError("use solver_wsocp_inPALM for weighted models")
    return _solve(v, "inPALM", var, opts, model)


def solver_wsocp_inPALM(var, opts, model):
    """socp/wdot2d/algorithms/solver_wsocp_inPALM.m:1"""
    return _solve("wdot2d", "inPALM", var, opts, model)


def solver_socp_PALM(var, opts, model):
    """socp/dot2d/algorithms/solver_socp_PALM.m:1"""
    return _solve("dot2d", "PALM", var, opts, model)


def solver_socp_accADMM(var, opts, model):
    """socp/dot2d/algorithms/solver_socp_accADMM.m:1"""
    return _solve("dot2d", "acc-ADMM", var, opts, model)


def solver_socp_accsGSADMM(var, opts, model):
    """socp/dot2d/algorithms/solver_socp_accsGSADMM.m:1"""
    return _solve("dot2d", "acc-sGS-ADMM", var, opts, model)


def solver_socp_sGSinPALM(var, opts, model):
    """socp/dot2d/algorithms/solver_socp_sGSinPALM.m:1 (the phi-step is one red-black symmetric Gauss-Seidel sweep, mexsGS)"""
    return _solve("dot2d", "sGS-inPALM", var, opts, model)


def solver_wsocp_accADMM(var, opts, model):
    """socp/wdot2d/algorithms/solver_wsocp_accADMM.m:1"""
    return _solve("wdot2d", "acc-ADMM", var, opts, model)
