"""TEST INFRASTRUCTURE ONLY (oracle/) -- CPU restatement of the reference's iteration hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs
may import this module; the product (``dotsocp_b200``) never does.

What it is: a literal numpy/scipy restatement of the reference's MATLAB glue around its native MEX
kernels, for the three variants ``socp/dot2d``, ``socp/wdot2d``, ``socp/dot1d``.  The native kernels
themselves are called through ``oracle.kernels`` (genuine reference binaries when ``oracle/_ref`` is
present, else the bit-identical C / numpy restatements).

Parity pin: the reference ships NO tests, golden vectors or logs and neither MATLAB
nor Octave exists offline, so
  * the KERNEL level is pinned bit-for-bit against the reference's own binaries (tests/test_oracle_kernels.py),
  * the SOLVER level (this file) is "parity unpinned by the reference": it is validated only through
    invariants (adjointness, A'A*poisson(rhs) == rhs, DCT == orthonormal DCT-II, closed-form W2 costs) and
    self-generated goldens (tests/golden/), not against a MATLAB run.

MATLAB -> numpy conventions: vectors are 1-D float64 arrays in MATLAB linear (column-major) order, so a
MATLAB ``reshape(v, ny, nx, nt)`` is the C-order view ``v.reshape(nt, nx, ny)``; ``z``/``beta`` are
``(L, ncol)`` Fortran-ordered.  ``movmean(x,2,dim,'Endpoints','discard')`` == adjacent-pair average.
All file:line citations are relative to /root/reference.
"""
from __future__ import annotations

import math
import time
from types import SimpleNamespace

import numpy as np
import scipy.fft as sfft
import scipy.sparse as sp

from . import kernels as K

INF = float("inf")


# ================================================================================================
# small helpers
# ================================================================================================
class Handle(SimpleNamespace):
    """Stands in for VarHandle / ModelHandle (socp/dot2d/utils/VarHandle.m:1-32, ModelHandle.m:1-32)."""

    def copy(self):
        return Handle(**self.__dict__)


def mmax(vals):
    """MATLAB max() ignores NaN."""
    v = [x for x in vals if not (isinstance(x, float) and math.isnan(x))]
    return max(v) if v else float("nan")


def normL2(x, h):
    """socp/dot2d/utils/normL2.m:4"""
    return math.sqrt(h) * float(np.linalg.norm(np.ravel(x, order="K")))


def FnormL2(x, h):
    """socp/dot2d/utils/FnormL2.m:4"""
    return math.sqrt(h) * float(np.linalg.norm(np.ravel(x, order="K")))


def pair_mean(x, axis):
    """movmean(x, 2, dim, 'Endpoints', 'discard')"""
    n = x.shape[axis]
    a = np.take(x, np.arange(0, n - 1), axis=axis)
    b = np.take(x, np.arange(1, n), axis=axis)
    return (a + b) / 2


# ================================================================================================
# operators, 2-D                                                   socp/dot2d/utils/*.m
# ================================================================================================
def _fwd_diff(n, hinv):
    """spdiags([-tmp, tmp], [0, 1], n-1, n)  (initialize.m:69,78,85)"""
    return sp.diags([np.full(n - 1, -hinv), np.full(n - 1, hinv)], [0, 1], shape=(n - 1, n), format="csr")


def gene_grad2d(nt, nx, ny):
    """initialize.m:35-39, 67-87 : A = [kron(Dt,Ixy); kron(kron(It,Dx),Iy); kron(Itx,Dy)]"""
    ht, hx, hy = 1 / (nt - 1), 1 / (nx - 1), 1 / (ny - 1) if ny > 1 else 1.0
    gt = sp.kron(_fwd_diff(nt, 1 / ht), sp.identity(nx * ny), format="csr")
    gx = sp.kron(sp.kron(sp.identity(nt), _fwd_diff(nx, 1 / hx)), sp.identity(ny), format="csr")
    gy = sp.kron(sp.identity(nt * nx), _fwd_diff(ny, 1 / hy), format="csr")
    return sp.vstack([gt, gx, gy], format="csr")


def initialize2d(rho0, rho1, nt):
    """socp/dot2d/utils/initialize.m:1-64.  rho0/rho1: MATLAB (ny, nx) arrays."""
    rho0 = np.asarray(rho0, dtype=np.float64)
    rho1 = np.asarray(rho1, dtype=np.float64)
    ny, nx = rho0.shape
    n = nx * ny * nt
    model = Handle(rho0=rho0, rho1=rho1, nx=nx, ny=ny, nt=nt, dim=2)
    qInd = Handle(bx=(nt - 1) * nx * ny + 1)
    qInd.by = nt * (nx - 1) * ny + qInd.bx
    ht, hx, hy = 1 / (nt - 1), 1 / (nx - 1), 1 / (ny - 1)
    model.grad = gene_grad2d(nt, nx, ny)
    model.gradT = model.grad.T.tocsr()
    c = np.zeros(n)
    c[: nx * ny] = -rho0.ravel(order="F") / ht
    c[n - nx * ny:] = rho1.ravel(order="F") / ht
    model.c = c
    xs = np.arange(nx) * hx  # 0:hx:1
    ys = np.arange(ny) * hy
    xx, yy = np.meshgrid(xs, ys)  # (ny, nx)
    phi2 = 0.5 * (xx ** 2 + yy ** 2)
    phi = np.tile(phi2.ravel(order="F"), nt)
    lenA = (nt - 1) * nx * ny
    Q = model.grad.shape[0]
    var = Handle(qInd=qInd, phi=phi, z=np.zeros((lenA, 10), order="F"), beta=np.zeros((lenA, 10), order="F"),
                 q=np.zeros(Q), alpha=np.zeros(Q))
    return var, model


def initialize_FFTkernel(nt, nx, ny=None):
    """socp/dot2d/utils/initialize_FFTkernel.m:6-15 (3 args) ; dot1d/utils/initialize_FFTkernel.m:6-13.
    Returned in the C-order view (nt, nx, ny) / (nt, nx)."""
    CT = (2 * (nt - 1) ** 2) * (1 - np.cos(np.pi * np.arange(nt) / nt))
    CX = (2 * (nx - 1) ** 2) * (1 - np.cos(np.pi * np.arange(nx) / nx))
    if ny is None:
        kernel = CX[None, :] + CT[:, None]
    else:
        CY = (2 * (ny - 1) ** 2) * (1 - np.cos(np.pi * np.arange(ny) / ny))
        kernel = (CY[None, None, :] + CX[None, :, None]) + CT[:, None, None]
    kernel[kernel == 0] = 1
    return kernel


def oper_poisson(kernel, rhs, workers=1):
    """oper_poisson3dim.m:4 / dot1d oper_poisson.m:4 : idctn(dctn(rhs) ./ kernel).
    mirt_dctn/mirt_idctn (mirt_dctn.m:64-96, mirt_idctn.m:59-95) == orthonormal DCT-II and its inverse
    (weights 2*exp(-i*pi*k/2n)/sqrt(2n), first one /sqrt(2))."""
    a = rhs.reshape(kernel.shape)
    a = sfft.dctn(a, type=2, norm="ortho", workers=workers)
    a = a / kernel
    a = sfft.idctn(a, type=2, norm="ortho", workers=workers)
    return a.ravel()


def oper_q2d(ny, nx, nt, D, E, weight=None):
    """socp/dot2d/utils/oper_q.m:13-26 ; socp/wdot2d/utils/oper_q.m:15-28 (with weight)."""
    tmp = (E / D) ** 2
    if weight is None:
        c1, c2 = 1 + 2 * tmp, 1 + tmp
    else:
        c1, c2 = 2 * tmp, tmp
    a = np.full((nt - 1, nx, ny), c1)
    b = np.full((nt, nx - 1, ny), c1)
    c = np.full((nt, nx, ny - 1), c1)
    b[[0, -1]] = c2
    c[[0, -1]] = c2
    d = np.concatenate([a.ravel(), b.ravel(), c.ravel()])
    if weight is not None:
        d = d + weight ** 2
    return d


def compute_kkt_dot_complement2d(q, alpha, z2, sigma, h, nt, nx, ny, qInd, cScale, dScale, D, E, weight=None):
    """socp/dot2d/utils/compute_kkt_dot_complement.m:1-19 ; wdot2d version :1-20 (Dalpha = weight.*alpha)."""
    Dalpha = alpha if weight is None else weight * alpha
    L = qInd.bx - 1
    nb = qInd.by - 1
    rhoT = (sigma * cScale * D) * Dalpha[:L]
    sq = ((dScale / E) * z2[:, 1:9]) ** 2
    ssum = sq[:, 0].copy()
    for j in range(1, 8):  # MATLAB sum(.,2): column by column
        ssum = ssum + sq[:, j]
    rhoFq = rhoT + (dScale / D) * q[:L] + ssum / 4
    rhoFq[rhoFq < 0] = 0
    dotcomplem = normL2(rhoT - rhoFq, h)
    normRho = normL2(rhoT, h)
    norm_rhoFq = normL2(rhoFq, h)
    pad = np.zeros((nt + 1, nx, ny))
    pad[1:nt] = rhoT.reshape(nt - 1, nx, ny)
    rho = pair_mean(pad, 0)  # (nt, nx, ny)
    rhoBx = (dScale / D) * (pair_mean(rho, 1).ravel() * q[L:nb])
    rhoBy = (dScale / D) * (pair_mean(rho, 2).ravel() * q[nb:])
    mx = (sigma * cScale * D) * Dalpha[L:nb]
    my = (sigma * cScale * D) * Dalpha[nb:]
    mRhoB = math.sqrt(normL2(mx - rhoBx, h) ** 2 + normL2(my - rhoBy, h) ** 2)
    normM = math.sqrt(normL2(mx, h) ** 2 + normL2(my, h) ** 2)
    normRhoB = math.sqrt(normL2(rhoBx, h) ** 2 + normL2(rhoBy, h) ** 2)
    return dotcomplem, normRho, norm_rhoFq, mRhoB, normM, normRhoB


# ================================================================================================
# operators, 1-D                                                   socp/dot1d/utils/*.m
# ================================================================================================
def initialize1d(rho0, rho1, nt):
    """socp/dot1d/utils/initialize.m:1-58"""
    rho0 = np.asarray(rho0, dtype=np.float64).ravel()
    rho1 = np.asarray(rho1, dtype=np.float64).ravel()
    nx = rho0.size
    n = nx * nt
    model = Handle(rho0=rho0, rho1=rho1, nx=nx, nt=nt, ny=1, dim=1)
    qInd = Handle(bx=(nt - 1) * nx + 1)
    ht, hx = 1 / (nt - 1), 1 / (nx - 1)
    gt = sp.kron(_fwd_diff(nt, 1 / ht), sp.identity(nx), format="csr")
    gx = sp.kron(sp.identity(nt), _fwd_diff(nx, 1 / hx), format="csr")
    model.grad = sp.vstack([gt, gx], format="csr")
    model.gradT = model.grad.T.tocsr()
    c = np.zeros(n)
    c[:nx] = -rho0 / ht
    c[n - nx:] = rho1 / ht
    model.c = c
    xx = np.arange(nx) * hx
    phi = np.tile(0.5 * xx ** 2, nt)
    lenA = (nt - 1) * nx
    Q = model.grad.shape[0]
    var = Handle(qInd=qInd, phi=phi, z=np.zeros((lenA, 6), order="F"), beta=np.zeros((lenA, 6), order="F"),
                 q=np.zeros(Q), alpha=np.zeros(Q))
    return var, model


def oper_q1d(nx, nt, D, E):
    """socp/dot1d/utils/oper_q.m:1-22"""
    tmp = (E / D) ** 2
    c1, c2 = 1 + 2 * tmp, 1 + tmp
    a = np.full((nt - 1, nx), c1)
    b = np.full((nt, nx - 1), c1)
    b[[0, -1]] = c2
    return np.concatenate([a.ravel(), b.ravel()])


def compute_kkt_dot_complement1d(q, alpha, z2, sigma, h, nt, nx, qInd, cScale, dScale, D, E):
    """socp/dot1d/utils/compute_kkt_dot_complement.m:1-17"""
    L = qInd.bx - 1
    rhoT = (sigma * cScale * D) * alpha[:L]
    sq = ((dScale / E) * z2[:, 1:5]) ** 2
    ssum = sq[:, 0].copy()
    for j in range(1, 4):
        ssum = ssum + sq[:, j]
    rhoFq = rhoT + (dScale / D) * q[:L] + ssum / 4
    rhoFq[rhoFq < 0] = 0
    dotcomplem = normL2(rhoT - rhoFq, h)
    normRho = normL2(rhoT, h)
    norm_rhoFq = normL2(rhoFq, h)
    pad = np.zeros((nt + 1, nx))
    pad[1:nt] = rhoT.reshape(nt - 1, nx)
    rho = pair_mean(pad, 0)
    rhoBx = (dScale / D) * (pair_mean(rho, 1).ravel() * q[L:])
    mx = (sigma * cScale * D) * alpha[L:]
    normM = normL2(mx, h)
    normRhoB = normL2(rhoBx, h)
    mRhoB = normL2(mx - rhoBx, h)
    return dotcomplem, normRho, norm_rhoFq, mRhoB, normM, normRhoB


# ================================================================================================
# dimension dispatch used by the iteration loops
# ================================================================================================
class _Ops:
    def __init__(self, model, var, workers=1):
        self.dim = model.dim
        self.nt, self.nx, self.ny = model.nt, model.nx, model.ny
        self.qInd = var.qInd
        self.workers = workers
        if self.dim == 2:
            self.h = 1 / (self.nx * self.ny * self.nt)
            self.ncol = 10
        else:
            self.h = 1 / (self.nx * self.nt)
            self.ncol = 6

    def BFd(self, z2, q, S, DF):
        if self.dim == 2:
            K.mexBFd(z2, q, self.nt, self.nx, self.ny, S, DF)
        else:
            K.mexBFd1d(z2, q, self.nt, self.nx, S, DF)

    def BFdConj(self, q2, z, S):
        if self.dim == 2:
            K.mexBFdConj(q2, z, self.nt, self.nx, self.ny, S)
        else:
            K.mexBFdConj1d(q2, z, self.nt, self.nx, S)

    def fftkernel(self):
        return initialize_FFTkernel(self.nt, self.nx, self.ny if self.dim == 2 else None)

    def poisson(self, kernel, rhs):
        return oper_poisson(kernel, rhs, self.workers)

    def oper_q(self, D, E, weight=None):
        if self.dim == 2:
            return oper_q2d(self.ny, self.nx, self.nt, D, E, weight)
        return oper_q1d(self.nx, self.nt, D, E)

    def kkt_dot(self, q, alpha, z2, sigma, cScale, dScale, D, E, weight=None):
        if self.dim == 2:
            return compute_kkt_dot_complement2d(q, alpha, z2, sigma, self.h, self.nt, self.nx, self.ny, self.qInd,
                                                cScale, dScale, D, E, weight)
        return compute_kkt_dot_complement1d(q, alpha, z2, sigma, self.h, self.nt, self.nx, self.qInd,
                                            cScale, dScale, D, E)


UPDATE_RULE = np.array([  # solver_socp_inPALM.m:39-51
    [1.1, 1.10], [1.2, 1.15], [1.5, 1.20], [2, 1.26], [2.5, 1.28], [3.33, 1.32],
    [5, 1.35], [10, 1.40], [20, 1.60], [40, 1.80], [50, 2.00]])


def _get_factor(xi, rule):
    """adjust_lagrangianParam.m:47-59"""
    factor = 1
    for i in range(rule.shape[0]):
        if xi >= rule[i, 0]:
            factor = rule[i, 1]
            continue
        break
    return factor


def adjust_lagrangianParam(sigma, xi, rule, bound=(1e-3, 1e3)):
    """socp/dot2d/utils/adjust_lagrangianParam.m:14-39"""
    lower, upper = bound
    if xi >= 1:
        factor = _get_factor(xi, rule)
    elif xi < 1:
        factor = 1 / _get_factor(1 / xi, rule)
    else:
        raise FloatingPointError("xi is NaN (MATLAB would error: 'factor' undefined)")
    if factor != 1:
        sigmaOld = sigma
        sigma = max(min(sigma * factor, upper), lower)
        factor = sigma / sigmaOld
    return sigma, factor


def IfAdjustSigma(it, last_it):
    """solver_socp_inPALM.m:361-379"""
    passed = it - last_it
    if it < 20 and passed >= 3:
        return True
    if it < 50 and passed >= 6:
        return True
    if it < 100 and passed >= 10:
        return True
    if it < 200 and passed >= 15:
        return True
    if it < 500 and passed >= 25:
        return True
    return passed >= 40


def _opt(opts, name, default):
    return opts[name] if name in opts else default


# ================================================================================================
# iteration loops
# ================================================================================================
def solver_socp_inPALM(var, opts, model, workers=1, palm=False, trace=None):
    """socp/dot2d/algorithms/solver_socp_inPALM.m:1-359   (palm=False, model.dim==2, no model.weight)
    socp/wdot2d/algorithms/solver_wsocp_inPALM.m:1-366    (model has .weight)
    socp/dot1d/algorithms/solver_socp_inPALM.m:1-358      (model.dim==1)
    socp/dot2d/algorithms/solver_socp_PALM.m:1-371        (palm=True)
    Returns (runHist, sigma); mutates var in place like the reference (handle semantics).
    Extra (not in the reference output struct): runHist.priVal / runHist.dualVal per check."""
    weight = getattr(model, "weight", None)
    weighted = weight is not None
    ops = _Ops(model, var, workers)
    checkPD = _opt(opts, "checkPrimDualFeas", False if weighted else True)     # :20-24 / wsocp :25-29
    time_limit = _opt(opts, "time_limit", 3600)
    tau, sigma, maxit, tol = opts["tau"], opts["sigma"], int(opts["maxit"]), opts["tol"]
    checkSByS = opts["ifCheckStepByStep"]
    lastSigmaIt = -INF
    cScale, dScale, D, E = var.cScale, var.dScale, var.D, var.E                # :54-59
    scaleBF = E / D
    scaleD = E / dScale
    use_feasOrg = 0
    tol_feasOrg = 5 * tol
    rescale = 1 if _opt(opts, "scaling", False) else 0                         # :64-68
    firstScaleIter, SecondScaleIter, checkRescaleIters, ratioThreshold = 10, 50, 100, 1.2
    maxFeas, relGap = INF, INF
    h = ops.h
    A, AT = model.grad, model.gradT
    c = model.c
    phi, q, z, alpha, beta = var.phi, var.q, var.z, var.alpha, var.beta        # :89-93
    var.phi = var.q = var.z = var.alpha = var.beta = None
    z = np.asfortranarray(z)
    kernel = D ** 2 * ops.fftkernel()                                          # :96
    diagQInv = 1 / ops.oper_q(D, E, weight)                                    # :97
    norm_c = model.normc
    norm_d = None if weighted else model.normd
    alpha = alpha / sigma                                                      # :102-104
    beta = np.asfortranarray(beta / sigma)
    c = c / sigma
    sigmaScale = 1
    kktConst = 1
    hist = Handle(kkt=[], time=[], iter=[], pdGap=[], priVal=[], dualVal=[])
    stopCondition = [0, 2, 5, 6] if checkPD else [0, 2, 5]                     # :117-121 (1-based 1,3,6,7)
    T = dict(lineq=0.0, proj=0.0, q=0.0, mult=0.0, kkt=0.0, q0=0.0)
    z2 = np.zeros(z.shape, order="F")                                          # :131-133
    q2 = np.zeros(q.shape)
    ops.BFd(z2, q, scaleBF, scaleD)
    if palm:                                                                   # PALM :137-138
        tmp_q = A @ phi
        ops.BFd(z, tmp_q, scaleBF, scaleD)
    wq = (lambda v: weight * v) if weighted else (lambda v: v)
    clock_total = time.perf_counter()
    it = 0
    for it in range(1, maxit + 1):
        # ---- rescaling :138-190 ----
        scaleYes = 0
        if rescale >= 3 and it % checkRescaleIters == 0:
            normPhi, normQ, normZ = normL2(phi, h), normL2(q, h), FnormL2(z, h)
            normAlpha, normBeta = sigma * normL2(alpha, h), sigma * FnormL2(beta, h)
            normPhis = max(normPhi, normQ, normZ)
            normAlps = max(normAlpha, normBeta)
            ratio = max(normAlps, normPhis) / min(normAlps, normPhis)
            if ratio > ratioThreshold:
                scaleYes = 1
        if ((rescale == 1 and maxFeas < 2e-2 and it >= firstScaleIter and relGap < 5e-2)
                or (rescale == 2 and maxFeas < 5e-3 and it >= SecondScaleIter and relGap < 1e-2)
                or scaleYes):
            if not scaleYes:
                normPhi, normQ, normZ = normL2(phi, h), normL2(q, h), FnormL2(z, h)
                normAlpha, normBeta = sigma * normL2(alpha, h), sigma * FnormL2(beta, h)
                normPhis = max(normPhi, normQ, normZ)
                normAlps = max(normAlpha, normBeta)
            dScale2, cScale2 = normPhis, normAlps
            sigma = sigma * (cScale2 / dScale2)
            c = c * dScale2 / cScale2 ** 2
            norm_c = norm_c / cScale2
            if not weighted:
                norm_d = norm_d / dScale2
            alpha = alpha * dScale2 / cScale2 ** 2
            beta = beta * dScale2 / cScale2 ** 2
            if not palm:
                q = q / dScale2                                                # :177 (absent in PALM :181)
            z = z / dScale2
            dScale = dScale2 * dScale
            cScale = cScale2 * cScale
            scaleD = E / dScale
            sigmaScale = sigmaScale * (cScale2 / dScale2)
            if palm:
                tmp_q = tmp_q / dScale2                                        # PALM :191
            else:
                ops.BFd(z2, q, scaleBF, scaleD)                                # :187
            rescale += 1
            if trace is not None:
                trace.append(("rescale", it, dScale2, cScale2))
        if palm:                                                               # PALM :196-200
            t0 = time.perf_counter()
            ops.BFdConj(q2, z + beta, scaleBF)
            q = (tmp_q + alpha + q2) * diagQInv
            T["q0"] += time.perf_counter() - t0
        # ---- step phi :192-195 ----
        t0 = time.perf_counter()
        phi = ops.poisson(kernel, AT @ (wq(q) - alpha) + c)
        T["lineq"] += time.perf_counter() - t0
        # ---- step z :197-200 ----
        t0 = time.perf_counter()
        if palm:
            ops.BFd(z2, q, scaleBF, scaleD)                                    # PALM :209
        K.mexProjSoc(z, np.asfortranarray(z2 - beta))
        T["proj"] += time.perf_counter() - t0
        # ---- step q :202-207 ----
        t0 = time.perf_counter()
        tmp_q = A @ phi
        ops.BFdConj(q2, np.asfortranarray(z + beta), scaleBF)
        if weighted:
            q = (weight * (tmp_q + alpha) + q2) * diagQInv                     # wsocp :212
        else:
            q = (tmp_q + alpha + q2) * diagQInv
        T["q"] += time.perf_counter() - t0
        # ---- step alpha, beta :209-216 ----
        t0 = time.perf_counter()
        resi_alpha = tmp_q - wq(q)
        ops.BFd(z2, q, scaleBF, scaleD)
        resi_beta = z - z2
        alpha = alpha + tau * resi_alpha
        beta = beta + tau * resi_beta
        T["mult"] += time.perf_counter() - t0
        # ---- kkt :218-324 ----
        t0 = time.perf_counter()
        adjustSigmaYes = IfAdjustSigma(it, lastSigmaIt)
        check = checkSByS or adjustSigmaYes or it == maxit or (time.perf_counter() - clock_total) > time_limit
        stop = False
        if check:
            ops.BFdConj(q2, np.asfortranarray(beta), scaleBF)
            Dalpha = wq(alpha)
            norm_q = normL2(q, h)
            norm_z = FnormL2(z, h)
            norm_Aphi = normL2(tmp_q, h)
            norm_alpha = sigma * normL2(alpha, h)
            norm_beta = sigma * FnormL2(beta, h)
            norm_FBbeta = sigma * normL2(q2, h)
            primFea1 = normL2(resi_alpha, h)
            primFea2 = FnormL2(resi_beta, h)
            dualFea1 = sigma * normL2(AT @ alpha - c, h)
            dualFea2 = sigma * normL2(q2 + Dalpha, h)
            K.mexProjSoc(z2, np.asfortranarray(z - sigma * beta))
            complem = FnormL2(z - z2, h)
            ops.BFd(z2, q, scaleBF, scaleD)
            dotcomplem, normRho, norm_rhoFq, mRhoB, normM, normRhoB = ops.kkt_dot(
                q, alpha, z2, sigma, cScale, dScale, D, E, weight)
            den2o = (kktConst * E / dScale + norm_q + norm_z) if weighted else (kktConst * E / dScale + norm_d)
            den2 = (kktConst + norm_q + norm_z) if weighted else (kktConst + norm_d)
            KKTResiOrg = [
                primFea1 / (kktConst * D / dScale + norm_Aphi + norm_q),
                primFea2 / den2o,
                dualFea1 / (kktConst / cScale + norm_c),
                complem / (kktConst * E / dScale + norm_z + norm_beta),
                dualFea2 / (kktConst / cScale / D + norm_FBbeta + norm_alpha),
                dotcomplem / (kktConst + normRho + norm_rhoFq),
                mRhoB / (kktConst + normM + normRhoB)]
            KKTResi = [
                primFea1 / (kktConst + norm_Aphi + norm_q),
                primFea2 / den2,
                dualFea1 / (kktConst + norm_c),
                complem / (kktConst + norm_z + norm_beta),
                dualFea2 / (kktConst + norm_FBbeta + norm_alpha)]
            priVal = (sigma * cScale * dScale * h) * float(np.dot(wq(q), alpha))
            dualVal = (sigma * cScale * dScale * h) * float(np.dot(c, phi))
            pdGap = abs(priVal - dualVal) / (1 + abs(priVal) + abs(dualVal))
            hist.kkt.append(KKTResiOrg)
            hist.time.append(time.perf_counter() - clock_total)
            hist.iter.append(it)
            hist.pdGap.append(pdGap)
            hist.priVal.append(priVal)
            hist.dualVal.append(dualVal)
            if trace is not None:
                trace.append(("check", it, sigma, list(KKTResiOrg), list(KKTResi), priVal, dualVal))
            if mmax([KKTResiOrg[i] for i in stopCondition]) < tol or (time.perf_counter() - clock_total) > time_limit:
                stop = True
            else:
                if mmax(KKTResi) < tol_feasOrg:
                    use_feasOrg = 1
                if adjustSigmaYes:
                    lastSigmaIt = it
                    if use_feasOrg:
                        resiPri, resiDual = mmax(KKTResiOrg[0:2]), mmax([KKTResiOrg[2], KKTResiOrg[4]])
                    else:
                        resiPri, resiDual = mmax(KKTResi[0:2]), mmax([KKTResi[2], KKTResi[4]])
                    sigma, factor = adjust_lagrangianParam(sigma, resiPri / resiDual, UPDATE_RULE)
                    if factor != 1:
                        alpha = alpha / factor
                        beta = beta / factor
                        c = c / factor
                if rescale > 0:
                    maxFeas = mmax(KKTResi)
                    relGap = pdGap
        T["kkt"] += time.perf_counter() - t0
        if stop:
            break
    time_total = time.perf_counter() - clock_total
    # ---- output :328-357 ----
    var.name = "Proximal ALM" if palm else "Inexact Proximal ALM"
    var.phi, var.q, var.z = phi, q, z
    var.alpha = sigma * alpha
    var.beta = sigma * beta
    if palm:
        var.time = {"Step_1_Q_Step": T["q0"], "Step_2_1_FFT": T["lineq"], "Step_2_2_ProjSOC": T["proj"],
                    "Step_3_Q_Step": T["q"], "Step_4_Multiplier": T["mult"], "KKT": T["kkt"],
                    "Total_Time": time_total, "Iters": it}
    else:
        var.time = {"Step_1_1_FFT": T["lineq"], "Step_1_2_ProjSOC": T["proj"], "Step_2_Q_Step": T["q"],
                    "Step_3_Multiplier": T["mult"], "KKT": T["kkt"], "Total_Time": time_total, "Iters": it}
    var.cScale, var.dScale, var.D, var.E = cScale, dScale, D, E
    runHist = _finish_hist(hist)
    return runHist, sigma / sigmaScale


SGS_UPDATE_RULE = np.array([[1.5, 1.20], [2, 1.26], [2.5, 1.28], [3.33, 1.32], [5, 1.35], [10, 1.40]])   # sGSinPALM.m:37-44


def IfAdjustSigma_sGS(it, last_it, scale=1):
    """solver_socp_sGSinPALM.m:431-456"""
    passed = it - last_it
    it = it / scale
    passed = passed / scale
    if it < 20 and passed >= 5:
        return True
    if it < 50 and passed >= 10:
        return True
    if it < 100 and passed >= 20:
        return True
    if it < 200 and passed >= 35:
        return True
    if it < 500 and passed >= 50:
        return True
    return passed >= 100


def solver_socp_sGSinPALM(var, opts, model, workers=1, trace=None):
    """socp/dot2d/algorithms/solver_socp_sGSinPALM.m:1-429 : inPALM whose phi-step is ONE symmetric red-black Gauss-Seidel
    sweep (mexsGS, :205) instead of the DCT Poisson solve, with the sGS-specific sigma voting (:322-360).  2-D, unweighted."""
    ops = _Ops(model, var, workers)
    checkPD = _opt(opts, "checkPrimDualFeas", True)                            # :20-24
    time_limit = _opt(opts, "time_limit", 3600)
    tau, sigma, maxit, tol = opts["tau"], opts["sigma"], int(opts["maxit"]), opts["tol"]
    checkSByS = opts["ifCheckStepByStep"]
    lastSigmaIt = -INF
    sGSits = 1
    cScale, dScale, D, E = var.cScale, var.dScale, var.D, var.E                # :47-55
    scaleBF = E / D
    scaleD = E / dScale
    scaleLap = D ** 2
    use_feasOrg = False
    tol_feasOrg = 5 * tol
    rescale = 1 if _opt(opts, "scaling", False) else 0
    firstScaleIter, SecondScaleIter, checkRescaleIters, ratioThreshold = 10, 50, 100, 1.2
    maxFeas, relGap = INF, INF
    hist_n, victory, initialSigmaScale, stablePhase = 19, 12, 1.10, False     # :76-80
    nx, ny, nt = model.nx, model.ny, model.nt
    h = 1 / (nx * ny * nt)
    A, AT = model.grad, model.gradT
    c = model.c
    phi, q, z, alpha, beta = var.phi, var.q, var.z, var.alpha, var.beta        # :92-96
    var.phi = var.q = var.z = var.alpha = var.beta = None
    z = np.asfortranarray(z)
    diagQInv = 1 / ops.oper_q(D, E, None)                                      # :99
    norm_c, norm_d = model.normc, model.normd
    alpha = alpha / sigma
    beta = np.asfortranarray(beta / sigma)
    c = c / sigma
    sigmaScale = 1
    sigma_adjust_it_gap = max(1, (nt * nx * ny) ** (1 / 3) / 33)               # :109
    sigma_adjust_val_gap = 0.95
    sgs_superior_yes = False
    tol_sgs_blocks = 5 * tol
    kktConst = 1
    hist = Handle(kkt=[], time=[], iter=[], pdGap=[], priVal=[], dualVal=[])
    FeasRatio = np.full(maxit + 1, INF)                                        # 1-based like the reference
    stopCondition = [0, 2, 5, 6] if checkPD else [0, 2, 5]
    T = dict(lineq=0.0, proj=0.0, q=0.0, mult=0.0, kkt=0.0)
    z2 = np.zeros(z.shape, order="F")
    q2 = np.zeros(q.shape)
    ops.BFd(z2, q, scaleBF, scaleD)
    phi = phi - h * phi.sum()                                                  # :142  phi - integralL2(phi, h)
    KKTResi = None
    norm_Aphi = norm_q = None
    clock_total = time.perf_counter()
    it = 0
    for it in range(1, maxit + 1):
        # ---- rescaling :147-201 (phi is rescaled too, :184) ----
        scaleYes = 0
        if rescale >= 3 and it % checkRescaleIters == 0:
            normPhi, normQ, normZ = normL2(phi, h), normL2(q, h), FnormL2(z, h)
            normAlpha, normBeta = sigma * normL2(alpha, h), sigma * FnormL2(beta, h)
            normPhis = max(normPhi, normQ, normZ)
            normAlps = max(normAlpha, normBeta)
            ratio = max(normAlps, normPhis) / min(normAlps, normPhis)
            if ratio > ratioThreshold:
                scaleYes = 1
        if ((rescale == 1 and maxFeas < 2e-2 and it >= firstScaleIter and relGap < 5e-2)
                or (rescale == 2 and maxFeas < 5e-3 and it >= SecondScaleIter and relGap < 1e-2)
                or scaleYes):
            if not scaleYes:
                normPhi, normQ, normZ = normL2(phi, h), normL2(q, h), FnormL2(z, h)
                normAlpha, normBeta = sigma * normL2(alpha, h), sigma * FnormL2(beta, h)
                normPhis = max(normPhi, normQ, normZ)
                normAlps = max(normAlpha, normBeta)
            dScale2, cScale2 = normPhis, normAlps
            sigma = sigma * (cScale2 / dScale2)
            c = c * dScale2 / cScale2 ** 2
            norm_c = norm_c / cScale2
            norm_d = norm_d / dScale2
            alpha = alpha * dScale2 / cScale2 ** 2
            beta = beta * dScale2 / cScale2 ** 2
            phi = phi / dScale2
            q = q / dScale2
            dScale = dScale2 * dScale
            cScale = cScale2 * cScale
            scaleD = E / dScale
            sigmaScale = sigmaScale * (cScale2 / dScale2)
            ops.BFd(z2, q, scaleBF, scaleD)
            rescale += 1
            if trace is not None:
                trace.append(("rescale", it, dScale2, cScale2))
        # ---- step phi :203-206 ----
        t0 = time.perf_counter()
        K.mexsGS(phi, AT @ (q - alpha) + c, 0, scaleLap, nt, nx, ny, sGSits)
        T["lineq"] += time.perf_counter() - t0
        t0 = time.perf_counter()
        adjustSigmaYes = IfAdjustSigma_sGS(it, lastSigmaIt, sigma_adjust_it_gap)
        check = checkSByS or adjustSigmaYes or it == maxit or (time.perf_counter() - clock_total) > time_limit
        if check:
            tmp_resi_sGS = AT @ (A @ phi - q + alpha) - c                      # :212-216
            resi_sGS_blocks = normL2(tmp_resi_sGS[0::2], h)
        T["kkt"] += time.perf_counter() - t0
        # ---- step z :219-222 ----
        t0 = time.perf_counter()
        K.mexProjSoc(z, np.asfortranarray(z2 - beta))
        T["proj"] += time.perf_counter() - t0
        # ---- step q :224-229 ----
        t0 = time.perf_counter()
        tmp_q = A @ phi
        ops.BFdConj(q2, np.asfortranarray(z + beta), scaleBF)
        q = (tmp_q + alpha + q2) * diagQInv
        T["q"] += time.perf_counter() - t0
        # ---- step alpha, beta :231-238 ----
        t0 = time.perf_counter()
        resi_alpha = tmp_q - q
        ops.BFd(z2, q, scaleBF, scaleD)
        resi_beta = z - z2
        alpha = alpha + tau * resi_alpha
        beta = beta + tau * resi_beta
        T["mult"] += time.perf_counter() - t0
        # ---- kkt :240-416 ----
        t0 = time.perf_counter()
        stop = False
        if check:
            ops.BFdConj(q2, np.asfortranarray(beta), scaleBF)
            norm_q = normL2(q, h)
            norm_z = FnormL2(z, h)
            norm_Aphi = normL2(tmp_q, h)
            norm_alpha = sigma * normL2(alpha, h)
            norm_beta = sigma * FnormL2(beta, h)
            norm_FBbeta = sigma * normL2(q2, h)
            primFea1 = normL2(resi_alpha, h)
            primFea2 = FnormL2(resi_beta, h)
            dualFea1 = sigma * normL2(AT @ alpha - c, h)
            dualFea2 = sigma * normL2(q2 + alpha, h)
            K.mexProjSoc(z2, np.asfortranarray(z - sigma * beta))
            complem = FnormL2(z - z2, h)
            ops.BFd(z2, q, scaleBF, scaleD)
            dotcomplem, normRho, norm_rhoFq, mRhoB, normM, normRhoB = ops.kkt_dot(q, alpha, z2, sigma, cScale, dScale, D, E, None)
            KKTResiOrg = [
                primFea1 / (kktConst * D / dScale + norm_Aphi + norm_q),
                primFea2 / (kktConst * E / dScale + norm_d),
                dualFea1 / (kktConst / cScale + norm_c),
                complem / (kktConst * E / dScale + norm_z + norm_beta),
                dualFea2 / (kktConst / cScale / D + norm_FBbeta + norm_alpha),
                dotcomplem / (kktConst + normRho + norm_rhoFq),
                mRhoB / (kktConst + normM + normRhoB)]
            KKTResi = [
                primFea1 / (kktConst + norm_Aphi + norm_q),
                primFea2 / (kktConst + norm_d),
                dualFea1 / (kktConst + norm_c),
                complem / (kktConst + norm_z + norm_beta),
                dualFea2 / (kktConst + norm_FBbeta + norm_alpha)]
            FeasRatio[it] = mmax(KKTResi[0:2]) / mmax([KKTResi[2], KKTResi[4]])
            priVal = (sigma * cScale * dScale * h) * float(np.dot(q, alpha))
            dualVal = (sigma * cScale * dScale * h) * float(np.dot(c, phi))
            pdGap = abs(priVal - dualVal) / (1 + abs(priVal) + abs(dualVal))
            hist.kkt.append(KKTResiOrg)
            hist.time.append(time.perf_counter() - clock_total)
            hist.iter.append(it)
            hist.pdGap.append(pdGap)
            hist.priVal.append(priVal)
            hist.dualVal.append(dualVal)
            error = mmax([KKTResiOrg[i] for i in stopCondition])
            if trace is not None:
                trace.append(("check", it, sigma, list(KKTResiOrg), list(KKTResi), priVal, dualVal))
            if error < tol or (time.perf_counter() - clock_total) > time_limit:
                stop = True
            else:
                if (not use_feasOrg) and mmax(KKTResi) < tol_feasOrg:
                    use_feasOrg = True
                kkt_sgs_blocks = math.sqrt(normL2(AT @ resi_alpha, h) ** 2 + (dualFea1 / sigma) ** 2)      # :322-323
                sgs_superior_yes = resi_sGS_blocks < sigma_adjust_val_gap * kkt_sgs_blocks
                if adjustSigmaYes:
                    lastSigmaIt = it
                    feasRatioHist = FeasRatio[max(1, it - hist_n): it + 1]
                    meanFeasRatio = float(np.mean(feasRatioHist))
                    primWinTimes = int(np.sum(feasRatioHist < 1))
                    dualWinTimes = int(np.sum(feasRatioHist > 1))
                    adjust_sigma_yes_2 = (sgs_superior_yes or (error < tol_sgs_blocks)
                                          or ((dualWinTimes >= victory) and (meanFeasRatio > 1)))
                    if adjust_sigma_yes_2:
                        if it > 2500:
                            stablePhase = True
                        if (((primWinTimes >= victory) and (meanFeasRatio < 1))
                                or ((dualWinTimes >= victory) and (meanFeasRatio > 1))):
                            factor = 1
                            if stablePhase:
                                sigma, factor = adjust_lagrangianParam(sigma, meanFeasRatio, SGS_UPDATE_RULE)
                            else:
                                if meanFeasRatio < 1:
                                    factor = 1 / initialSigmaScale
                                elif meanFeasRatio > 1:
                                    factor = initialSigmaScale
                                sigma = sigma * factor
                            if factor != 1:
                                alpha = alpha / factor
                                beta = beta / factor
                                c = c / factor
                if rescale > 0:
                    maxFeas = mmax(KKTResi)
                    relGap = pdGap
        elif sgs_superior_yes:                                                 # :385-402
            primFea1 = normL2(resi_alpha, h)
            dualFea1 = sigma * normL2(AT @ alpha - c, h)
            if use_feasOrg:
                relaPrimFeaDec = primFea1 / ((kktConst * D / dScale + norm_Aphi + norm_q) * KKTResi[0])
                KKTResi[0] = KKTResi[0] * relaPrimFeaDec
                KKTResi[1] = KKTResi[1] * relaPrimFeaDec
                KKTResi[2] = dualFea1 / (kktConst / cScale + norm_c)
            else:
                relaPrimFeaDec = primFea1 / ((kktConst + norm_Aphi + norm_q) * KKTResi[0])
                KKTResi[0] = KKTResi[0] * relaPrimFeaDec
                KKTResi[1] = KKTResi[1] * relaPrimFeaDec
                KKTResi[2] = dualFea1 / (kktConst + norm_c)
            FeasRatio[it] = mmax(KKTResi[0:2]) / mmax([KKTResi[2], KKTResi[4]])
        else:
            FeasRatio[it] = FeasRatio[it - 1]
        T["kkt"] += time.perf_counter() - t0
        if stop:
            break
    time_total = time.perf_counter() - clock_total
    var.name = "Symmetric Gauss-seidel based inPALM"
    var.phi, var.q, var.z = phi, q, z
    var.alpha = sigma * alpha
    var.beta = sigma * beta
    var.time = {"Step_1_1_sGS": T["lineq"], "Step_1_2_ProjSOC": T["proj"], "Step_2_Q_Step": T["q"], "Step_3_Multiplier": T["mult"],
                "KKT": T["kkt"], "Total_Time": time_total, "Iters": it}
    var.cScale, var.dScale, var.D, var.E = cScale, dScale, D, E
    return _finish_hist(hist), sigma / sigmaScale


def _finish_hist(hist):
    n = len(hist.iter)
    return Handle(kkt=np.array(hist.kkt, dtype=np.float64).reshape(n, 7), time=np.array(hist.time, dtype=np.float64),
                  iter=np.array(hist.iter, dtype=np.float64), pdGap=np.array(hist.pdGap, dtype=np.float64),
                  priVal=np.array(hist.priVal, dtype=np.float64), dualVal=np.array(hist.dualVal, dtype=np.float64), len=n)


def _copyvar(phi, z, q, alpha, beta):
    """solver_socp_accADMM.m:480-484 (deep copies; numpy arrays are never copy-on-write)."""
    return phi.copy(), z.copy(order="F"), q.copy(), alpha.copy(), beta.copy(order="F")


def solver_socp_accsGSADMM(var, opts, model, workers=1, trace=None):
    """socp/dot2d/algorithms/solver_socp_accsGSADMM.m:1-568 : acc-ADMM whose phi-step is one symmetric red-black Gauss-Seidel
    sweep (mexsGS, :256), with the sGS check schedule and sigma voting of solver_socp_sGSinPALM.m (stable phase after 1500)."""
    return solver_socp_accADMM(var, opts, model, workers, trace, sgs=True)


def solver_socp_accADMM(var, opts, model, workers=1, trace=None, sgs=False):
    """socp/dot2d/algorithms/solver_socp_accADMM.m:1-458 ; socp/wdot2d/algorithms/solver_wsocp_accADMM.m:1-463 ;
    sgs=True: socp/dot2d/algorithms/solver_socp_accsGSADMM.m (line numbers marked sGS:)."""
    weight = getattr(model, "weight", None)
    weighted = weight is not None
    ops = _Ops(model, var, workers)
    restart = _opt(opts, "restart", 100)
    stepRho = _opt(opts, "rho", 2)
    stepAlpha = _opt(opts, "theta", 2)
    HalpernYes = stepAlpha == 2
    checkPD = _opt(opts, "checkPrimDualFeas", False if weighted else True)
    time_limit = _opt(opts, "time_limit", 3600)
    sigma, maxit, tol = opts["sigma"], int(opts["maxit"]), opts["tol"]
    checkSByS = opts["ifCheckStepByStep"]
    lastSigmaIt = -INF
    cScale, dScale, D, E = var.cScale, var.dScale, var.D, var.E
    scaleBF, scaleD = E / D, E / dScale
    use_feasOrg, tol_feasOrg = 0, 5 * tol
    rescale = 1 if _opt(opts, "scaling", False) else 0
    firstScaleIter, SecondScaleIter, checkRescaleIters, ratioThreshold = 10, 50, 200, 1.2      # :94-97
    maxFeas, relGap = INF, INF
    h = ops.h
    A, AT, c = model.grad, model.gradT, model.c
    phi, q, z, alpha, beta = var.phi, var.q, np.asfortranarray(var.z), var.alpha, var.beta
    var.phi = var.q = var.z = var.alpha = var.beta = None
    kernel = None if sgs else D ** 2 * ops.fftkernel()
    diagQInv = 1 / ops.oper_q(D, E, weight)
    norm_c = model.normc
    norm_d = None if weighted else model.normd
    alpha = alpha / sigma
    beta = np.asfortranarray(beta / sigma)
    c = c / sigma
    sigmaScale, kktConst = 1, 1
    hist = Handle(kkt=[], time=[], iter=[], pdGap=[], priVal=[], dualVal=[])
    stopCondition = [0, 2, 5, 6] if checkPD else [0, 2, 5]
    T = dict(lineq=0.0, proj=0.0, q=0.0, mult=0.0, kkt=0.0, interp=0.0)
    z2 = np.zeros(z.shape, order="F")
    q2 = np.zeros(q.shape)
    if sgs:                                                                                  # sGS:76-120,147,164-166
        nx_, ny_, nt_ = model.nx, model.ny, model.nt
        scaleLap = D ** 2
        hist_n, victory, initialSigmaScale, stablePhase = 19, 12, 1.10, False
        sigma_adjust_it_gap = max(1, (nt_ * nx_ * ny_) ** (1 / 3) / 33)
        sigma_adjust_val_gap, sgs_superior_yes, tol_sgs_blocks = 0.95, False, 5 * tol
        FeasRatio = np.full(maxit + 1, INF)
        KKTResi = None
        norm_Aphi = norm_q = None
        phi = phi - h * phi.sum()
    phiOld, zOld, qOld, alphaOld, betaOld = _copyvar(phi, z, q, alpha, beta)                 # :157
    k = 0
    if HalpernYes:
        phi0, z0, q0, alpha0, beta0 = _copyvar(phi, z, q, alpha, beta)                        # :161-163
    wq = (lambda v: weight * v) if weighted else (lambda v: v)
    clock_total = time.perf_counter()
    it = 0
    for it in range(1, maxit + 1):
        scaleYes = 0
        if rescale >= 3 and it % checkRescaleIters == 0:
            normPhis = max(normL2(phi, h), normL2(q, h), FnormL2(z, h))
            normAlps = max(sigma * normL2(alpha, h), sigma * FnormL2(beta, h))
            if max(normAlps, normPhis) / min(normAlps, normPhis) > ratioThreshold:
                scaleYes = 1
        if ((rescale == 1 and maxFeas < 2e-2 and it >= firstScaleIter and relGap < 5e-2)
                or (rescale == 2 and maxFeas < 5e-3 and it >= SecondScaleIter and relGap < 1e-2)
                or scaleYes):
            if not scaleYes:
                normPhis = max(normL2(phi, h), normL2(q, h), FnormL2(z, h))
                normAlps = max(sigma * normL2(alpha, h), sigma * FnormL2(beta, h))
            dScale2, cScale2 = normPhis, normAlps
            sigma = sigma * (cScale2 / dScale2)
            c = c * dScale2 / cScale2 ** 2
            norm_c = norm_c / cScale2
            if not weighted:
                norm_d = norm_d / dScale2
            alpha = alpha * dScale2 / cScale2 ** 2
            beta = beta * dScale2 / cScale2 ** 2
            phi = phi / dScale2                                                              # :207
            q = q / dScale2
            z = z / dScale2
            dScale = dScale2 * dScale
            cScale = cScale2 * cScale
            scaleD = E / dScale
            sigmaScale = sigmaScale * (cScale2 / dScale2)
            k = 0                                                                            # :217-222
            phiOld, zOld, qOld, alphaOld, betaOld = _copyvar(phi, z, q, alpha, beta)
            if HalpernYes:
                phi0, z0, q0, alpha0, beta0 = _copyvar(phi, z, q, alpha, beta)
            rescale += 1
            if trace is not None:
                trace.append(("rescale", it, dScale2, cScale2))
        # step q :227-232
        t0 = time.perf_counter()
        ops.BFdConj(q2, np.asfortranarray(z + beta), scaleBF)
        tmp_q = A @ phi
        if weighted:
            q = (weight * (tmp_q + alpha) + q2) * diagQInv
        else:
            q = (tmp_q + alpha + q2) * diagQInv
        T["q"] += time.perf_counter() - t0
        # step alpha, beta :234-239
        t0 = time.perf_counter()
        ops.BFd(z2, q, scaleBF, scaleD)
        alpha = alpha + tmp_q - wq(q)
        beta = beta + z - z2
        T["mult"] += time.perf_counter() - t0
        # step phi :241-244
        t0 = time.perf_counter()
        if sgs:
            phi = phi.copy()                                                                 # in-place MEX write; Old/anchor copies are deep (sGS:563-566)
            K.mexsGS(phi, AT @ (q - alpha) + c, 0, scaleLap, nt_, nx_, ny_, 1)                # sGS:256
        else:
            phi = ops.poisson(kernel, AT @ (wq(q) - alpha) + c)
        T["lineq"] += time.perf_counter() - t0
        if sgs:                                                                              # sGS:259-269
            t0 = time.perf_counter()
            adjustSigmaYes = IfAdjustSigma_sGS(it, lastSigmaIt, sigma_adjust_it_gap)
            check = checkSByS or adjustSigmaYes or it == maxit or (time.perf_counter() - clock_total) > time_limit
            if check:
                tmp_resi_sGS = AT @ (A @ phi - q + alpha) - c
                resi_sGS_blocks = normL2(tmp_resi_sGS[0::2], h)
            T["kkt"] += time.perf_counter() - t0
        # step z :246-249   (in place into z; the Old/anchor copies are deep, :480-484)
        t0 = time.perf_counter()
        znew = np.empty(z.shape, order="F")
        K.mexProjSoc(znew, np.asfortranarray(z2 - beta))
        z = znew
        T["proj"] += time.perf_counter() - t0
        # kkt :251-367
        t0 = time.perf_counter()
        if not sgs:
            adjustSigmaYes = IfAdjustSigma(it, lastSigmaIt)
            check = checkSByS or adjustSigmaYes or it == maxit or (time.perf_counter() - clock_total) > time_limit
        stop = False
        if check:
            ops.BFdConj(q2, np.asfortranarray(beta), scaleBF)
            tmp_q = A @ phi
            norm_q, norm_z, norm_Aphi = normL2(q, h), FnormL2(z, h), normL2(tmp_q, h)
            norm_alpha, norm_beta = sigma * normL2(alpha, h), sigma * FnormL2(beta, h)
            norm_FBbeta = sigma * normL2(q2, h)
            K.mexProjSoc(z2, np.asfortranarray(z - sigma * beta))
            complem = FnormL2(z - z2, h)
            ops.BFd(z2, q, scaleBF, scaleD)
            primFea1 = normL2(tmp_q - wq(q), h)
            primFea2 = FnormL2(z - z2, h)
            dualFea1 = sigma * normL2(AT @ alpha - c, h)
            dualFea2 = sigma * normL2(q2 + wq(alpha), h)
            dotcomplem, normRho, norm_rhoFq, mRhoB, normM, normRhoB = ops.kkt_dot(
                q, alpha, z2, sigma, cScale, dScale, D, E, weight)
            den2o = (kktConst * E / dScale + norm_q + norm_z) if weighted else (kktConst * E / dScale + norm_d)
            den2 = (kktConst + norm_q + norm_z) if weighted else (kktConst + norm_d)
            KKTResiOrg = [
                primFea1 / (kktConst * D / dScale + norm_Aphi + norm_q), primFea2 / den2o,
                dualFea1 / (kktConst / cScale + norm_c), complem / (kktConst * E / dScale + norm_z + norm_beta),
                dualFea2 / (kktConst / cScale / D + norm_FBbeta + norm_alpha),
                dotcomplem / (kktConst + normRho + norm_rhoFq), mRhoB / (kktConst + normM + normRhoB)]
            KKTResi = [
                primFea1 / (kktConst + norm_Aphi + norm_q), primFea2 / den2, dualFea1 / (kktConst + norm_c),
                complem / (kktConst + norm_z + norm_beta), dualFea2 / (kktConst + norm_FBbeta + norm_alpha)]
            priVal = (sigma * cScale * dScale * h) * float(np.dot(wq(q), alpha))
            dualVal = (sigma * cScale * dScale * h) * float(np.dot(c, phi))
            pdGap = abs(priVal - dualVal) / (1 + abs(priVal) + abs(dualVal))
            hist.kkt.append(KKTResiOrg); hist.time.append(time.perf_counter() - clock_total)
            hist.iter.append(it); hist.pdGap.append(pdGap); hist.priVal.append(priVal); hist.dualVal.append(dualVal)
            if trace is not None:
                trace.append(("check", it, sigma, list(KKTResiOrg), list(KKTResi), priVal, dualVal))
            error = mmax([KKTResiOrg[i] for i in stopCondition])
            if sgs:
                FeasRatio[it] = mmax(KKTResi[0:2]) / mmax([KKTResi[2], KKTResi[4]])           # sGS:323
            if error < tol or (time.perf_counter() - clock_total) > time_limit:
                stop = True
            else:
                if mmax(KKTResi) < tol_feasOrg:
                    use_feasOrg = 1
                factor = 1
                if sgs:                                                                      # sGS:360-412
                    kkt_sgs_blocks = math.sqrt(normL2(AT @ (tmp_q - q), h) ** 2 + (dualFea1 / sigma) ** 2)
                    sgs_superior_yes = resi_sGS_blocks < sigma_adjust_val_gap * kkt_sgs_blocks
                    if adjustSigmaYes:
                        lastSigmaIt = it
                        feasRatioHist = FeasRatio[max(1, it - hist_n): it + 1]
                        meanFeasRatio = float(np.mean(feasRatioHist))
                        primWinTimes = int(np.sum(feasRatioHist < 1))
                        dualWinTimes = int(np.sum(feasRatioHist > 1))
                        if sgs_superior_yes or error < tol_sgs_blocks or (dualWinTimes >= victory and meanFeasRatio > 1):
                            if it > 1500:
                                stablePhase = True
                            if ((primWinTimes >= victory and meanFeasRatio < 1)
                                    or (dualWinTimes >= victory and meanFeasRatio > 1)):
                                if stablePhase:
                                    sigma, factor = adjust_lagrangianParam(sigma, meanFeasRatio, SGS_UPDATE_RULE)
                                else:
                                    if meanFeasRatio < 1:
                                        factor = 1 / initialSigmaScale
                                    elif meanFeasRatio > 1:
                                        factor = initialSigmaScale
                                    sigma = sigma * factor
                elif adjustSigmaYes:
                    lastSigmaIt = it
                    if use_feasOrg:
                        resiPri, resiDual = mmax(KKTResiOrg[0:2]), mmax([KKTResiOrg[2], KKTResiOrg[4]])
                    else:
                        resiPri, resiDual = mmax(KKTResi[0:2]), mmax([KKTResi[2], KKTResi[4]])
                    sigma, factor = adjust_lagrangianParam(sigma, resiPri / resiDual, UPDATE_RULE)
                if sgs or adjustSigmaYes:
                    if factor != 1:
                        alpha = alpha / factor
                        alphaOld = alphaOld / factor
                        beta = beta / factor
                        betaOld = betaOld / factor
                        c = c / factor
                        k = 0
                        if HalpernYes:
                            phi0, z0, q0, alpha0, beta0 = _copyvar(phi, z, q, alpha, beta)
                if rescale > 0:
                    maxFeas = mmax(KKTResi)
                    relGap = pdGap
        elif sgs and sgs_superior_yes:                                                       # sGS:420-437
            primFea1 = normL2(A @ phi - q, h)
            dualFea1 = sigma * normL2(AT @ alpha - c, h)
            if use_feasOrg:
                dec = primFea1 / ((kktConst * D / dScale + norm_Aphi + norm_q) * KKTResi[0])
                KKTResi[0], KKTResi[1] = KKTResi[0] * dec, KKTResi[1] * dec
                KKTResi[2] = dualFea1 / (kktConst / cScale + norm_c)
            else:
                dec = primFea1 / ((kktConst + norm_Aphi + norm_q) * KKTResi[0])
                KKTResi[0], KKTResi[1] = KKTResi[0] * dec, KKTResi[1] * dec
                KKTResi[2] = dualFea1 / (kktConst + norm_c)
            FeasRatio[it] = mmax(KKTResi[0:2]) / mmax([KKTResi[2], KKTResi[4]])
        elif sgs:
            FeasRatio[it] = FeasRatio[it - 1]
        T["kkt"] += time.perf_counter() - t0
        if stop:
            break
        # step interpolation :369-423
        t0 = time.perf_counter()
        if HalpernYes:
            c1 = 1 / (k + 2)
            c2 = (k + 1) / (k + 2)
            phi = c1 * phi0 + c2 * ((1 - stepRho) * phiOld + stepRho * phi)
            z = c1 * z0 + c2 * ((1 - stepRho) * zOld + stepRho * z)
            q = c1 * q0 + c2 * ((1 - stepRho) * qOld + stepRho * q)
            alpha = c1 * alpha0 + c2 * ((1 - stepRho) * alphaOld + stepRho * alpha)
            beta = c1 * beta0 + c2 * ((1 - stepRho) * betaOld + stepRho * beta)
            k += 1
            phiOld, zOld, qOld, alphaOld, betaOld = _copyvar(phi, z, q, alpha, beta)
            if k >= restart:
                k = 0
                phi0, z0, q0, alpha0, beta0 = _copyvar(phi, z, q, alpha, beta)
        else:
            phiHat = (1 - stepRho) * phiOld + stepRho * phi
            zHat = (1 - stepRho) * zOld + stepRho * z
            qHat = (1 - stepRho) * qOld + stepRho * q
            alphaHat = (1 - stepRho) * alphaOld + stepRho * alpha
            betaHat = (1 - stepRho) * betaOld + stepRho * beta
            c1 = stepAlpha / (2 * (k + stepAlpha))
            if k == 0:
                phi = (1 - c1) * phiOld + c1 * phiHat
                z = (1 - c1) * zOld + c1 * zHat
                q = (1 - c1) * qOld + c1 * qHat
                alpha = (1 - c1) * alphaOld + c1 * alphaHat
                beta = (1 - c1) * betaOld + c1 * betaHat
            else:
                c2 = k / (k + stepAlpha)
                phi = (1 - c1) * phiOld + (c1 + c2) * phiHat - c2 * phiHatOld
                z = (1 - c1) * zOld + (c1 + c2) * zHat - c2 * zHatOld
                q = (1 - c1) * qOld + (c1 + c2) * qHat - c2 * qHatOld
                alpha = (1 - c1) * alphaOld + (c1 + c2) * alphaHat - c2 * alphaHatOld
                beta = (1 - c1) * betaOld + (c1 + c2) * betaHat - c2 * betaHatOld
            k += 1
            phiOld, zOld, qOld, alphaOld, betaOld = _copyvar(phi, z, q, alpha, beta)
            if k >= restart:
                k = 0
            else:
                phiHatOld, zHatOld, qHatOld, alphaHatOld, betaHatOld = phiHat, zHat, qHat, alphaHat, betaHat
        z = np.asfortranarray(z)
        beta = np.asfortranarray(beta)
        T["interp"] += time.perf_counter() - t0
    time_total = time.perf_counter() - clock_total
    var.name = "Accelerated symmetric Gauss-Seidel based ADMM" if sgs else "Accelerated ADMM"
    var.phi, var.q, var.z = phi, q, z
    var.alpha = sigma * alpha
    var.beta = sigma * beta
    if sgs:                                                                                  # sGS:512-513
        var.time = {"Step_1_1_sGS": T["lineq"], "Step_1_2_ProjSOC": T["proj"], "Step_2_Multiplier": T["mult"],
                    "Step_3_Q_Step": T["q"], "Step_4_Interp": T["interp"], "KKT": T["kkt"], "Total_Time": time_total, "Iters": it}
    else:
        var.time = {"Step_1_Q_Step": T["q"], "Step_2_Multiplier": T["mult"], "Step_3_1_FFT": T["lineq"],
                    "Step_3_2_ProjSOC": T["proj"], "KKT": T["kkt"], "Interp": T["interp"], "Total_Time": time_total,
                    "Iters": it}
    var.cScale, var.dScale, var.D, var.E = cScale, dScale, D, E
    return _finish_hist(hist), sigma / sigmaScale


# ================================================================================================
# level transfer                                   socp/*/utils/{downSample_phi,interpolate,jump_nextLevel,...}.m
# ================================================================================================
def downSample_phi2d(v):
    """socp/dot2d/utils/downSample_phi.m:5-34 (literal, incl. the first-corner formula that uses v(2,1) twice)."""
    v = np.asarray(v, dtype=np.float64)
    Mx, My = v.shape[0] - 1, v.shape[1] - 1
    Mxc, Myc = Mx // 2, My // 2
    vc = np.zeros((Mxc + 1, Myc + 1))
    ind = np.arange(3, Mx, 2) - 1  # MATLAB 3:2:(Mx-1) -> 0-based  (same index set used for both dims, as in the file)
    I = ind[:, None]
    J = ind[None, :]
    vc[1:Mxc, 1:Myc] = (4 * v[I, J] + 2 * (v[I - 1, J] + v[I + 1, J] + v[I, J - 1] + v[I, J + 1])
                        + (v[I - 1, J - 1] + v[I - 1, J + 1] + v[I + 1, J - 1] + v[I + 1, J + 1])) / 16
    vc[0, 1:Myc] = (4 * v[0, ind] + 2 * (v[1, ind] + v[0, ind - 1] + v[0, ind + 1]) + (v[1, ind - 1] + v[1, ind + 1])) / 12
    vc[Mxc, 1:Myc] = (4 * v[Mx, ind] + 2 * (v[Mx - 1, ind] + v[Mx, ind - 1] + v[Mx, ind + 1])
                      + (v[Mx - 1, ind - 1] + v[Mx - 1, ind + 1])) / 12
    vc[1:Mxc, 0] = (4 * v[ind, 0] + 2 * (v[ind - 1, 0] + v[ind + 1, 0] + v[ind, 1]) + (v[ind - 1, 1] + v[ind + 1, 1])) / 12
    vc[1:Mxc, Myc] = (4 * v[ind, My] + 2 * (v[ind - 1, My] + v[ind + 1, My] + v[ind, My - 1])
                      + (v[ind - 1, My - 1] + v[ind + 1, My - 1])) / 12
    vc[0, 0] = (4 * v[0, 0] + 2 * (v[1, 0] + v[0, 1]) + v[1, 0]) / 9
    vc[0, Myc] = (4 * v[0, My] + 2 * (v[1, My] + v[0, My - 1]) + v[1, My - 1]) / 9
    vc[Mxc, 0] = (4 * v[Mx, 0] + 2 * (v[Mx - 1, 0] + v[Mx, 1]) + v[Mx - 1, 1]) / 9
    vc[Mxc, Myc] = (4 * v[Mx, My] + 2 * (v[Mx - 1, My] + v[Mx, My - 1]) + v[Mx - 1, My - 1]) / 9
    return vc


def downSample_phi1d(phi):
    """socp/dot1d/utils/downSample_phi.m:4-14"""
    phi = np.asarray(phi, dtype=np.float64).ravel()
    ln = phi.size - 1
    lenc = ln // 2
    phic = np.zeros(lenc + 1)
    ind = np.arange(3, ln, 2) - 1
    phic[1:lenc] = 0.5 * phi[ind] + 0.25 * (phi[ind - 1] + phi[ind + 1])
    phic[0] = (2 / 3) * phi[0] + (1 / 3) * phi[1]
    phic[-1] = (1 / 3) * phi[-2] + (2 / 3) * phi[-1]
    return phic


def _interp_phi(phi, shape):
    """interpolate.m:46-71 (2-D) / dot1d interpolate.m:38-58 : nodal linear interpolation, axis by axis
    (y, then x, then t).  `shape` is the coarse C-order shape (nt, nx[, ny])."""
    a = phi.reshape(shape)
    for ax in range(a.ndim - 1, -1, -1):  # y first (last C axis), t last
        n = a.shape[ax]
        sh = list(a.shape)
        sh[ax] = 2 * (n - 1) + 1
        r = np.zeros(sh)
        odd = [slice(None)] * a.ndim
        even = [slice(None)] * a.ndim
        odd[ax] = slice(0, None, 2)
        even[ax] = slice(1, None, 2)
        r[tuple(odd)] = a
        r[tuple(even)] = pair_mean(a, ax)
        a = r
    return a.ravel()


def _interp_tstagger(f):
    """interpolate.m:20-43 : nearest in t (each coarse cell layer -> two fine layers), linear in y then x."""
    a = np.repeat(f, 2, axis=0)
    for ax in range(a.ndim - 1, 0, -1):
        n = a.shape[ax]
        sh = list(a.shape)
        sh[ax] = 2 * (n - 1) + 1
        r = np.zeros(sh)
        odd = [slice(None)] * a.ndim
        even = [slice(None)] * a.ndim
        odd[ax] = slice(0, None, 2)
        even[ax] = slice(1, None, 2)
        r[tuple(odd)] = a
        r[tuple(even)] = pair_mean(a, ax)
        a = r
    return a


def interpolate(var, model):
    """socp/dot2d/utils/interpolate.m:1-15 ; dot1d/utils/interpolate.m:1-13 (mutates and returns var)."""
    nt, nx, ny = model.nt, model.nx, model.ny
    shape = (nt, nx, ny) if model.dim == 2 else (nt, nx)
    cshape = (nt - 1,) + shape[1:]
    var.phi = _interp_phi(var.phi, shape)
    ncol = var.beta.shape[1]
    cols = [_interp_tstagger(var.beta[:, j].reshape(cshape)).ravel() for j in range(ncol)]
    var.beta = np.asfortranarray(np.stack(cols, axis=1))
    return var


def jump_nextLevel(var, model, rho0, rho1, nt, weight=None):
    """socp/dot2d/utils/jump_nextLevel.m:1-18 ; wdot2d :1-17 ; dot1d :1-18"""
    varR = interpolate(var, model).copy()
    if model.dim == 2:
        var_init, modelR = initialize2d(rho0, rho1, nt)
    else:
        var_init, modelR = initialize1d(rho0, rho1, nt)
    varR.qInd = var_init.qInd
    varR.z = var_init.z
    if weight is not None:
        modelR.weight = weight
        varR.q = (modelR.grad @ varR.phi) / weight
    else:
        varR.q = modelR.grad @ varR.phi
    varR.alpha = var_init.alpha
    nb = np.asfortranarray(-varR.beta)
    if model.dim == 2:
        K.mexBFdConj(varR.alpha, nb, modelR.nt, modelR.nx, modelR.ny, 1.0)
    else:
        K.mexBFdConj1d(varR.alpha, nb, modelR.nt, modelR.nx, 1.0)
    if weight is not None:
        varR.alpha = varR.alpha / weight
    return varR, modelR


def _prolong_linear(nC):
    """downSample_q.m:25-31"""
    nR = 2 * (nC - 1) + 1
    i = np.concatenate([np.arange(0, nR, 2), np.tile(np.arange(1, nR - 1, 2), 2)])
    j = np.concatenate([np.arange(nC), np.arange(nC - 1), np.arange(1, nC)])
    v = np.concatenate([np.ones(nC), np.full(2 * (nC - 1), 0.5)])
    return sp.csr_matrix((v, (i, j)), shape=(nR, nC))


def _prolong_nearest(nC):
    """downSample_q.m:33-39"""
    nR = 2 * nC
    i = np.concatenate([np.arange(0, nR, 2), np.arange(1, nR, 2)])
    j = np.concatenate([np.arange(nC), np.arange(nC)])
    return sp.csr_matrix((np.ones(2 * nC), (i, j)), shape=(nR, nC))


def _restrictions(nt, nx, ny):
    nt2, nx2, ny2 = (nt + 1) // 2, (nx + 1) // 2, (ny + 1) // 2
    PT = sp.kron(sp.kron(_prolong_nearest(nt2 - 1), _prolong_linear(nx2)), _prolong_linear(ny2), format="csc")
    PX = sp.kron(sp.kron(_prolong_linear(nt2), _prolong_nearest(nx2 - 1)), _prolong_linear(ny2), format="csc")
    PY = sp.kron(sp.kron(_prolong_linear(nt2), _prolong_linear(nx2)), _prolong_nearest(ny2 - 1), format="csc")
    out = []
    for P in (PT, PX, PY):
        colsum = np.asarray(P.sum(axis=0)).ravel()
        out.append((P @ sp.diags(1 / colsum)).T.tocsr())
    return out


def downSample_q(nt, nx, ny, q):
    """socp/wdot2d/utils/downSample_q.m:4-19"""
    RT, RX, RY = _restrictions(nt, nx, ny)
    bx = (nt - 1) * nx * ny
    by = bx + nt * (nx - 1) * ny
    return np.concatenate([RT @ q[:bx], RX @ q[bx:by], RY @ q[by:]])


def downSample_barrier(nt, nx, ny, weight):
    """socp/wdot2d/utils/downSample_barrier.m:4-24 (geometric = log-average restriction)"""
    return np.exp(downSample_q(nt, nx, ny, np.log(weight)))


# ================================================================================================
# output recovery                                                 socp/*/utils/{recover_RhoE,recover_q}.m
# ================================================================================================
def recover_RhoE(var, model):
    """socp/dot2d/utils/recover_RhoE.m:13-25 (wdot2d: alpha = weight.*alpha, :11) ; dot1d :12-20.
    Returns C-order arrays (nt, nx, ny) / (nt, nx)."""
    nt, nx, ny = model.nt, model.nx, model.ny
    alpha = var.alpha
    if getattr(model, "weight", None) is not None:
        alpha = model.weight * alpha
    L = var.qInd.bx - 1
    if model.dim == 2:
        nb = var.qInd.by - 1
        rho = alpha[:L].reshape(nt - 1, nx, ny)
        rho = np.concatenate([model.rho0.T[None], (rho[:-1] + rho[1:]) / 2, model.rho1.T[None]], axis=0)
        Ex = alpha[L:nb].reshape(nt, nx - 1, ny).copy()
        Ex[[0, -1]] = 2 * Ex[[0, -1]]
        zx = np.zeros((nt, 1, ny))
        Ex = np.concatenate([zx, (Ex[:, :-1] + Ex[:, 1:]) / 2, zx], axis=1)
        Ey = alpha[nb:].reshape(nt, nx, ny - 1).copy()
        Ey[[0, -1]] = 2 * Ey[[0, -1]]
        zy = np.zeros((nt, nx, 1))
        Ey = np.concatenate([zy, (Ey[:, :, :-1] + Ey[:, :, 1:]) / 2, zy], axis=2)
        return rho, Ex, Ey
    rho = alpha[:L].reshape(nt - 1, nx)
    rho = np.concatenate([model.rho0[None], (rho[:-1] + rho[1:]) / 2, model.rho1[None]], axis=0)
    Ex = alpha[L:].reshape(nt, nx - 1).copy()
    Ex[[0, -1]] = 2 * Ex[[0, -1]]
    zx = np.zeros((nt, 1))
    Ex = np.concatenate([zx, (Ex[:, :-1] + Ex[:, 1:]) / 2, zx], axis=1)
    return rho, Ex


def recover_q(var, model):
    """socp/dot2d/utils/recover_q.m:12-22 ; dot1d :11-18"""
    nt, nx, ny = model.nt, model.nx, model.ny
    q = var.q
    L = var.qInd.bx - 1
    if model.dim == 2:
        nb = var.qInd.by - 1
        q0 = q[:L].reshape(nt - 1, nx, ny)
        bx = q[L:nb].reshape(nt, nx - 1, ny)
        zx = np.zeros((nt, 1, ny))
        bx = np.concatenate([zx, (bx[:, :-1] + bx[:, 1:]) / 2, zx], axis=1)
        bx = (bx[:-1] + bx[1:]) / 2
        by = q[nb:].reshape(nt, nx, ny - 1)
        zy = np.zeros((nt, nx, 1))
        by = np.concatenate([zy, (by[:, :, :-1] + by[:, :, 1:]) / 2, zy], axis=2)
        by = (by[:-1] + by[1:]) / 2
        return q0, bx, by
    q0 = q[:L].reshape(nt - 1, nx)
    bx = q[L:].reshape(nt, nx - 1)
    zx = np.zeros((nt, 1))
    bx = np.concatenate([zx, (bx[:, :-1] + bx[:, 1:]) / 2, zx], axis=1)
    bx = (bx[:-1] + bx[1:]) / 2
    return q0, bx


def check_massConservation(rho, tol=1e-2):
    """socp/dot2d/utils/check_massConservation.m:16-34 : per-time-layer mass and negative mass."""
    nt = rho.shape[0]
    r2 = rho.reshape(nt, -1)
    sumRho = r2.mean(axis=1)
    sumNeg = np.where(r2 < 0, r2, 0.0).mean(axis=1)
    err = max(np.abs(sumRho - 1).max(), np.abs(sumNeg).max())
    return err <= tol, sumRho, sumNeg


# ================================================================================================
# multilevel drivers                                              socp/*/solver_*dotsocp*.m
# ================================================================================================
def InitialScaling(var, model, scalingYes, lastLevelKKT, variant):
    """solver_dotsocp2d.m:304-365 ; solver_wdotsocp2d.m:297-342 ; solver_dotsocp1d.m:263-313.
    variant in {"dot2d","wdot2d","dot1d"}."""
    h = 1 / var.phi.size
    hMean = h ** (1 / 2) if variant == "dot1d" else h ** (1 / 3)
    if lastLevelKKT is None or not hasattr(var, "E2"):
        Escale2 = math.sqrt(2)
    elif variant == "wdot2d":
        safeguard = 4
        Escale2 = var.E2 * min(safeguard, max(1 / safeguard, math.sqrt(lastLevelKKT[0] / lastLevelKKT[1])))
    else:
        ratio = math.sqrt(lastLevelKKT[0] / lastLevelKKT[1])
        lowerRatio = 0.8333
        if ratio < lowerRatio:
            Escale2 = var.E2 * max(1 / math.sqrt(2), ratio / lowerRatio)
        else:
            Escale2 = var.E2 * min(math.sqrt(2), max(1, ratio))
    if scalingYes:
        norm_c = normL2(model.c, h) * math.sqrt(model.nt)
        norm_d = math.sqrt(2)
        if variant == "wdot2d":
            adjust = 10 ** float(np.mean(np.log10(model.weight + 1e-10)))
            D = math.sqrt(2) * math.sqrt(hMean) * adjust
            E = D / Escale2
            cScale = max(1, norm_c * math.sqrt(hMean) / adjust)
            dScale = E * norm_d * math.sqrt(adjust)
        else:
            D = math.sqrt(2) * math.sqrt(hMean)
            E = D / Escale2
            cScale = max(1, norm_c * math.sqrt(hMean))
            dScale = E * norm_d
        model.normc = norm_c / cScale
        model.normd = norm_d * E / dScale
        if variant == "dot2d":
            model.c = (1.0 / cScale) * model.c
        else:
            model.c = model.c / cScale
        model.grad = (D * model.grad).tocsr()
        model.gradT = model.grad.T.tocsr()
        if variant == "dot2d":
            var.phi = (1.0 / dScale) * var.phi
            var.q = (D / dScale) * var.q
            var.z = (E / dScale) * var.z
            var.alpha = (1.0 / cScale / D) * var.alpha
            var.beta = (1.0 / cScale / E) * var.beta
        else:
            var.phi = (1 / dScale) * var.phi
            var.q = (D / dScale) * var.q
            var.z = (E / dScale) * var.z
            var.alpha = (1 / cScale / D) * var.alpha
            var.beta = (1 / cScale / E) * var.beta
    else:
        cScale = dScale = D = E = 1
        model.normc = normL2(model.c, h)
        model.normd = math.sqrt(2)
    var.cScale, var.dScale, var.D, var.E, var.E2 = cScale, dScale, D, E, Escale2


def recoverOrgVar(var):
    """solver_dotsocp2d.m:368-386"""
    cScale, dScale, D, E = var.cScale, var.dScale, var.D, var.E
    var.phi = dScale * var.phi
    var.z = (dScale / E) * var.z
    var.q = (dScale / D) * var.q
    var.alpha = (cScale * D) * var.alpha
    var.beta = (cScale * E) * var.beta


def catRunHist(ML, rh):
    """solver_dotsocp2d.m:389-407"""
    ML.kkt = rh.kkt.copy() if ML.kkt is None else np.concatenate([ML.kkt, rh.kkt], axis=0)
    ML.pdGap = rh.pdGap.copy() if ML.pdGap is None else np.concatenate([ML.pdGap, rh.pdGap])
    if ML.time is None or ML.time.size == 0:
        ML.time = rh.time.copy()
    else:
        rh.time = ML.time[-1] + rh.time
        ML.time = np.concatenate([ML.time, rh.time])
    if ML.iter is None or ML.iter.size == 0:
        ML.iter = rh.iter.copy()
    else:
        ML.iter = np.concatenate([ML.iter, ML.iter[-1] + rh.iter])
    ML.len += rh.len
    return ML, rh


def _solve_multilevel(variant, rho0, rho1, nt, levelN, opts, method, barrier=None, workers=1, trace=None):
    opts = dict(opts)
    if not (isinstance(levelN, (int, np.integer)) and levelN >= 1):
        raise ValueError("Invalid input at position 4 (Number of levels in multilevel strategy)")
    valid = {"dot2d": ["PALM", "inPALM", "ALG2", "acc-ADMM", "sGS-inPALM", "acc-sGS-ADMM"],
             "wdot2d": ["inPALM", "ALG2", "acc-ADMM"], "dot1d": ["inPALM", "ALG2"]}[variant]
    if method not in valid:
        raise ValueError("Invalid input at position 6 (Solving method)")
    sgsMethod = method in ("sGS-inPALM", "acc-sGS-ADMM")                         # solver_dotsocp2d.m:93
    admmMaxIt, sgsMaxIt = 3000, 6000                                             # :96-97
    opts.setdefault("ifCheckStepByStep", False)
    scalingYes = opts.setdefault("scaling", True)
    optsML = dict(opts)
    if "maxit" not in opts:
        optsML["maxit"] = 1e4 if variant == "wdot2d" else (sgsMaxIt if sgsMethod else admmMaxIt)     # :114-121
    optsML["tolFactor"] = -1 if optsML["tol"] > 0.99e-3 else -0.5
    tolLowerBound = 1e-5 if variant == "dot1d" else 1e-4
    if variant == "dot2d":
        if method in ("PALM", "inPALM", "sGS-inPALM"):
            optsML["tau"] = 1.9
        elif method == "ALG2":
            optsML["tau"] = 1.0
    else:
        if method == "inPALM":
            optsML["tau"] = 1.9
        elif method == "ALG2":
            optsML["tau"] = 1.0
    if "sigma" not in optsML:
        optsML["sigma"] = 1 if ("scaling" in opts and not sgsMethod) else 0.1   # :139-146
    optsML.setdefault("time_limit", 3600)
    weight = opts.get("weight") if variant == "wdot2d" else None
    # ---- preparation :159-178 ----
    rho0s, rho1s, nts, tols, weights = [None] * levelN, [None] * levelN, [None] * levelN, [None] * levelN, [None] * levelN
    rho0s[-1], rho1s[-1], nts[-1], tols[-1], weights[-1] = np.asarray(rho0, float), np.asarray(rho1, float), nt, optsML["tol"], weight
    nxs, nys = [None] * levelN, [None] * levelN
    if variant != "dot1d":
        nys[-1], nxs[-1] = rho0s[-1].shape
    for lv in range(levelN - 2, -1, -1):
        nts[lv] = (nts[lv + 1] - 1) // 2 + 1
        tols[lv] = max(tols[lv + 1] * 2 ** optsML["tolFactor"], tolLowerBound)
        if variant == "dot1d":
            rho0s[lv] = downSample_phi1d(rho0s[lv + 1])
            rho1s[lv] = downSample_phi1d(rho1s[lv + 1])
        else:
            nxs[lv], nys[lv] = (nxs[lv + 1] + 1) // 2, (nys[lv + 1] + 1) // 2
            rho0s[lv] = downSample_phi2d(rho0s[lv + 1])
            rho1s[lv] = downSample_phi2d(rho1s[lv + 1])
        if variant == "wdot2d" and barrier is not None:
            rho0s[lv], rho1s[lv], _ = ensure_barrier_validity(rho0s[lv], rho1s[lv], barrier)
            weights[lv] = downSample_barrier(nts[lv + 1], nxs[lv + 1], nys[lv + 1], weights[lv + 1])
        else:
            if variant == "wdot2d":
                weights[lv] = downSample_q(nts[lv + 1], nxs[lv + 1], nys[lv + 1], weights[lv + 1])
            N = rho0s[lv].size
            rho0s[lv] = rho0s[lv] / (rho0s[lv].sum() / N)
            rho1s[lv] = rho1s[lv] / (rho1s[lv].sum() / N)
    # ---- multilevel :181-250 ----
    timeML = [None] * (levelN + 1)
    lastLevelKKT = None
    clockML = time.perf_counter()
    init = initialize1d if variant == "dot1d" else initialize2d
    var, model = init(rho0s[0], rho1s[0], nts[0])
    if variant == "wdot2d":
        model.weight = weights[0]
    ML = Handle(kkt=None, time=None, iter=None, pdGap=None, len=0)
    runHist = None
    level_iters = []
    for level in range(levelN):
        InitialScaling(var, model, scalingYes, lastLevelKKT, variant)
        o2 = dict(optsML)
        o2["tol"] = tols[level]
        if trace is not None:
            trace.append(("level", level, nts[level]))
        if method == "PALM":
            runHist, sigma = solver_socp_inPALM(var, o2, model, workers, palm=True, trace=trace)
        elif method in ("inPALM", "ALG2"):
            runHist, sigma = solver_socp_inPALM(var, o2, model, workers, trace=trace)
        elif method == "sGS-inPALM":                                             # :210-216: sGS on the last level only
            if level == levelN - 1:
                runHist, sigma = solver_socp_sGSinPALM(var, o2, model, workers, trace=trace)
            else:
                o2["maxit"] = admmMaxIt
                runHist, sigma = solver_socp_inPALM(var, o2, model, workers, trace=trace)
        elif method == "acc-sGS-ADMM":                                           # :217-223
            if level == levelN - 1:
                runHist, sigma = solver_socp_accsGSADMM(var, o2, model, workers, trace=trace)
            else:
                o2["maxit"] = admmMaxIt
                o2["tau"] = 1.9
                runHist, sigma = solver_socp_inPALM(var, o2, model, workers, trace=trace)
        else:
            runHist, sigma = solver_socp_accADMM(var, o2, model, workers, trace=trace)
        recoverOrgVar(var)
        timeML[level] = var.time
        level_iters.append(var.time["Iters"])
        ML, runHist = catRunHist(ML, runHist)
        if level < levelN - 1:
            optsML["time_limit"] = optsML["time_limit"] - var.time["Total_Time"]
            optsML["sigma"] = 10 ** (math.log10(optsML["sigma"] * sigma) / 2)
            var, model = jump_nextLevel(var, model, rho0s[level + 1], rho1s[level + 1], nts[level + 1],
                                        weights[level + 1] if variant == "wdot2d" else None)
            lastLevelKKT = runHist.kkt[-1, :]
    multilevelTime = time.perf_counter() - clockML
    # ---- output :253-299 ----
    output = Handle()
    if variant == "dot1d":
        output.rho, output.Ex = recover_RhoE(var, model)
        output.q0, output.bx = recover_q(var, model)
    else:
        output.rho, output.Ex, output.Ey = recover_RhoE(var, model)
        output.q0, output.bx, output.by = recover_q(var, model)
    output.massOK, output.sumRho, output.sumNegRho = check_massConservation(output.rho, 1e-2)
    output.var, output.model, output.level_iters, output.sigma = var, model, level_iters, sigma
    timeML[levelN] = {"ML_Time": multilevelTime}
    return output, timeML, ML, runHist


def solver_dotsocp2d(rho0, rho1, nt, levelN, opts, method="inPALM", workers=1, trace=None):
    """socp/dot2d/solver_dotsocp2d.m:1-300"""
    return _solve_multilevel("dot2d", rho0, rho1, nt, levelN, opts, method, None, workers, trace)


def solver_wdotsocp2d(rho0, rho1, nt, levelN, opts, method="inPALM", barrier=None, workers=1, trace=None):
    """socp/wdot2d/solver_wdotsocp2d.m:1-294"""
    return _solve_multilevel("wdot2d", rho0, rho1, nt, levelN, opts, method, barrier, workers, trace)


def solver_dotsocp1d(rho0, rho1, nt, levelN, opts, method="inPALM", workers=1, trace=None):
    """socp/dot1d/solver_dotsocp1d.m:1-260"""
    return _solve_multilevel("dot1d", rho0, rho1, nt, levelN, opts, method, None, workers, trace)


# ================================================================================================
# input generators (the benchmark/test inputs)                    examples/*/*.m
# ================================================================================================
def _normalize2d(rho0, rho1, nx, ny, lowerBound=0):
    """examples/dot2d/get_example.m:45-46"""
    rho0 = ((nx * ny / rho0.sum()) * rho0 + lowerBound) / (1 + lowerBound)
    rho1 = ((nx * ny / rho1.sum()) * rho1 + lowerBound) / (1 + lowerBound)
    return rho0, rho1


def gene_example1(nx, ny):
    """examples/dot2d/gene_example1.m:4-25 : Gaussian (0.25,0.75) -> (0.75,0.25), covariance 0.05*I.
    NB the file builds (nx, ny)-shaped arrays whose FIRST index follows `x`."""
    mu1, mu2, sigma = 0.25, 0.75, 0.05
    sinv = np.linalg.inv(np.array([[sigma, 0], [0, sigma]]))

    def Normal(x, y, mean):
        return math.sqrt(np.linalg.det(sinv)) / (2 * math.pi) * np.exp(
            -0.5 * (sinv[0, 0] * (x - mean[0]) ** 2 + sinv[0, 1] * (x - mean[0]) * (y - mean[1])
                    + sinv[1, 1] * (y - mean[1]) ** 2))
    x = np.tile(np.linspace(0, 1, nx).reshape(nx, 1), (1, ny))
    y = np.tile(np.linspace(0, 1, ny).reshape(1, ny), (nx, 1))
    return Normal(x, y, (mu1, mu2)), Normal(x, y, (mu2, mu1))


def gene_example2(nx, ny):
    """examples/dot2d/gene_example2.m:4-18 : one Gaussian (sigma .1) -> four Gaussians (sigma .05)."""
    mu1 = 0.25
    mu2 = 1 - mu1
    s1, s2 = 0.1, 0.05
    Y, X = np.meshgrid(np.linspace(0, 1, nx), np.linspace(0, 1, ny))  # (ny, nx)

    def g(a, b, s):
        return np.exp(-((X - a) ** 2 + (Y - b) ** 2) / (2 * s ** 2))
    rho0 = g(mu1, mu1, s1)
    rho1 = g(mu1, mu1, s2) + g(mu1, mu2, s2) + g(mu2, mu1, s2) + g(mu2, mu2, s2)
    return rho0, rho1


def gene_exampleCircle(nx, ny):
    """examples/dot2d/gene_exampleCircle.m:4-25 : indicator discs, pure translation by (0.5,-0.5)."""
    hx, hy = 1 / (nx - 1), 1 / (ny - 1)
    xx, yy = np.meshgrid(np.arange(nx) * hx, np.arange(ny) * hy)
    rho0 = ((xx - 0.25) ** 2 + (yy - 0.75) ** 2 < 0.25 ** 2).astype(float)
    rho1 = ((xx - 0.75) ** 2 + (yy - 0.25) ** 2 < 0.25 ** 2).astype(float)
    return rho0, rho1


def get_example2d(problem, nx, ny, lowerBound=0):
    """examples/dot2d/get_example.m:24-46 (analytic problems only)"""
    gen = {"example1": gene_example1, "example2": gene_example2, "circle": gene_exampleCircle}[problem]
    rho0, rho1 = gen(nx, ny)
    return _normalize2d(rho0, rho1, nx, ny, lowerBound)


def get_example1d(problem, nx, lowerBound=0):
    """examples/dot1d/get_example.m:9-19 ; gene_example_gaussian.m:5-21 ; gene_example_box.m:4-12"""
    x = np.linspace(0, 1, nx)
    if problem == "gaussian":
        def Normal(x, mean, sinv):
            return math.sqrt(sinv) / (2 * math.pi) * np.exp(-0.5 * (sinv * (x - mean) ** 2))
        s1 = 0.01
        s2 = s1 / 4
        rho0, rho1 = Normal(x, 0.3, 1 / s1), Normal(x, 0.7, 1 / s2)
    elif problem == "box":
        rho0 = ((x >= 0.1) & (x <= 0.5)).astype(float)
        rho1 = ((x >= 0.85) & (x <= 0.95)).astype(float)
    else:
        raise ValueError("Novalid input: 'Problem'")
    rho0 = ((nx / rho0.sum()) * rho0 + lowerBound) / (1 + lowerBound)
    rho1 = ((nx / rho1.sum()) * rho1 + lowerBound) / (1 + lowerBound)
    return rho0, rho1


def gene_weight_circle(nt, nx, ny):
    """examples/wdot2d/gene_weight_circle.m:4-27"""
    hx, hy = 1 / (nx - 1), 1 / (ny - 1)
    xS = np.linspace(.5 * hx, 1 - .5 * hx, nx - 1)
    xC = np.linspace(0, 1, nx)
    yS = np.linspace(.5 * hy, 1 - .5 * hy, ny - 1)
    yC = np.linspace(0, 1, ny)

    def circ(xx, yy):
        return np.sqrt((xx - .5) ** 2 + (yy - .5) ** 2)
    xx, yy = np.meshgrid(xS, yC)            # (ny, nx-1)
    wX = circ(xx, yy)
    wX = wX * (ny * (nx - 1) / wX.sum())
    xx, yy = np.meshgrid(xC, yS)            # (ny-1, nx)
    wY = circ(xx, yy)
    wY = wY * (ny * (nx - 1) / wY.sum())
    wT = np.ones((nt - 1) * nx * ny)
    return np.concatenate([wT, np.tile(wX.ravel(order="F"), nt), np.tile(wY.ravel(order="F"), nt)])


def gene_barrier_of_love_heart():
    """examples/wdot2d/gene_barrier_of_love_heart.m:4-14"""
    def heart(x, y, s):
        return ((s * (x - 0.5)) ** 2 + (s * (y - 0.5)) ** 2 - 1) ** 3 - (s * (x - 0.5)) ** 2 * (s * (y - 0.5)) ** 3
    return lambda x, y: (heart(x, y + 0.05, 2.5) > 0) | (heart(x, y, 15) <= 0)


def gene_exampleLoveHeart(nx, ny):
    """examples/wdot2d/gene_exampleLoveHeart.m:4-27"""
    c1, c2, r1, r2 = (0.7, 0.3), (0.345, 0.625), 0.09, 0.09
    s1, s2 = r1 / 3, r2 / 3
    Y, X = np.meshgrid(np.linspace(0, 1, nx), np.linspace(0, 1, ny))

    def g(a, b, s):
        return np.exp(-((X - a) ** 2 + (Y - b) ** 2) / (2 * s ** 2))
    rho0 = g(c1[0], c1[1], s1)
    rho0[(X - c1[0]) ** 2 + (Y - c1[1]) ** 2 > r1 ** 2] = 0
    rho1 = g(c2[0], c2[1], s2)
    rho1[(X - c2[0]) ** 2 + (Y - c2[1]) ** 2 > r2 ** 2] = 0
    return rho0, rho1


def get_weight_by_barrier(nx, ny, nt, barrier, barrierWeight=1e6):
    """examples/wdot2d/get_weight_by_barrier.m:12-33"""
    hx, hy = 1 / (nx - 1), 1 / (ny - 1)
    xS = np.linspace(.5 * hx, 1 - .5 * hx, nx - 1)
    xC = np.linspace(0, 1, nx)
    yS = np.linspace(.5 * hy, 1 - .5 * hy, ny - 1)
    yC = np.linspace(0, 1, ny)
    xx, yy = np.meshgrid(xS, yC)
    mask = barrier(xx.T, yy.T) > 0
    wX = np.ones((ny, nx - 1))
    wX[mask.T] = barrierWeight
    xx, yy = np.meshgrid(xC, yS)
    mask = barrier(xx.T, yy.T) > 0
    wY = np.ones((ny - 1, nx))
    wY[mask.T] = barrierWeight
    wT = np.ones((nt - 1) * nx * ny)
    return np.concatenate([wT, np.tile(wX.ravel(order="F"), nt), np.tile(wY.ravel(order="F"), nt)])


def ensure_barrier_validity(rho0, rho1, barrier):
    """examples/wdot2d/ensure_barrier_validity.m:4-16"""
    ny, nx = rho0.shape
    xx, yy = np.meshgrid(np.linspace(0, 1, nx), np.linspace(0, 1, ny))
    b = barrier(xx.T, yy.T).astype(float)
    filterVal = b.mean()
    bc = b.T > filterVal
    rho0 = rho0.copy()
    rho1 = rho1.copy()
    rho0[bc] = 0
    rho1[bc] = 0
    rho0 = (nx * ny / rho0.sum()) * rho0
    rho1 = (nx * ny / rho1.sum()) * rho1
    return rho0, rho1, bc


def w2_cost(output, dim=2):
    """Implementation-independent transport cost  sum_cells |m|^2 / rho * h  (SURVEY.md §8d), from output.{rho,Ex[,Ey]}."""
    rho = output.rho
    m2 = output.Ex ** 2 + (output.Ey ** 2 if dim == 2 else 0.0)
    mask = rho > 1e-12
    val = np.zeros_like(rho)
    val[mask] = m2[mask] / rho[mask]
    return float(val.mean())
