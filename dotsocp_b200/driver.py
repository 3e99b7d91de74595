"""Host-side mirror of the reference's multilevel drivers, calling the GPU level solver.

    [output, timeML, runHistML, runHist] = solver_dotsocp2d(rho0, rho1, nt, levelN, opts, method)   socp/dot2d/solver_dotsocp2d.m:1
                                           solver_wdotsocp2d(..., method, barrier)                  socp/wdot2d/solver_wdotsocp2d.m:1
                                           solver_dotsocp1d(...)                                    socp/dot1d/solver_dotsocp1d.m:1

In a MATLAB deployment these drivers stay byte-identical and only ``algorithms/solver_*socp_*.m`` is swapped for the
MEX gateway; this module exists so that the same end-to-end flow (level set-up, InitialScaling,
level solve on the GPU, recoverOrgVar, jump_nextLevel, output recovery) can be driven and tested from Python.
It is written against the staggered-grid stencils directly (no sparse matrices): model.grad is the triple
(D/ht, D/hx, D/hy).

Array conventions: densities rho0/rho1 are MATLAB-shaped (ny, nx) [1-D: (nx,)]; all solver vectors are 1-D arrays in
MATLAB linear order == C-order views (nt, nx, ny).
"""
from __future__ import annotations

import math
import time
from types import SimpleNamespace

import numpy as np

from . import ops
from . import solver as S


# ------------------------------------------------------------------------------------------------ set-up
def initialize(rho0, rho1, nt):
    """socp/dot2d/utils/initialize.m:1-64 ; socp/dot1d/utils/initialize.m:1-58 (dimension from rho0.ndim)."""
    rho0 = np.asarray(rho0, dtype=np.float64)
    rho1 = np.asarray(rho1, dtype=np.float64)
    if rho0.ndim == 1:
        nx, ny, dim = rho0.size, 1, 1
        r0, r1 = rho0, rho1
    else:
        ny, nx = rho0.shape
        dim = 2
        r0, r1 = rho0.ravel(order="F"), rho1.ravel(order="F")
    n = nt * nx * ny
    ht, hx = 1 / (nt - 1), 1 / (nx - 1)
    hy = 1 / (ny - 1) if ny > 1 else 1.0
    model = SimpleNamespace(rho0=rho0, rho1=rho1, nt=nt, nx=nx, ny=ny, dim=dim,
                            grad=(1 / ht, 1 / hx, (1 / hy) if ny > 1 else 0.0))
    c = np.zeros(n)
    c[: nx * ny] = -r0 / ht
    c[n - nx * ny:] = r1 / ht
    model.c = c
    model.c_first, model.c_last = c[: nx * ny], c[n - nx * ny:]     # views: model.c is zero in between (initialize.m:41-44)
    phi = np.tile(initial_phi_plane(nx, ny), nt)
    ncol = 10 if dim == 2 else 6
    L = (nt - 1) * nx * ny
    Q = L + nt * (nx - 1) * ny + nt * nx * (ny - 1)
    qInd = SimpleNamespace(bx=L + 1, by=L + nt * (nx - 1) * ny + 1)
    var = SimpleNamespace(qInd=qInd, phi=phi, z=np.zeros((L, ncol), order="F"), beta=np.zeros((L, ncol), order="F"),
                          q=np.zeros(Q), alpha=np.zeros(Q))
    return var, model


def initial_phi_plane(nx, ny):
    """one time level of the initial phi = (x^2 + y^2) / 2 (initialize.m:46-52), C order (nx, ny); 1-D: x^2 / 2"""
    xs = np.arange(nx) * (1 / (nx - 1))
    if ny > 1:
        ys = np.arange(ny) * (1 / (ny - 1))
        return (0.5 * (xs[:, None] ** 2 + ys[None, :] ** 2)).ravel()
    return 0.5 * xs ** 2


def initial_state_local(model, dScale, tc0, tc1, tn0, tn1, with_z=True):
    """The reference's initial state (initialize.m:46-64: phi = |x|^2/2 on every level, everything else zero) after
    InitialScaling (solver_dotsocp2d.m:338-342; dScale = None: unscaled), for node levels [tn0, tn1) / cell layers [tc0, tc1)
    only -- a time slab's part of split_state(initialize(...)) without the full-grid arrays (the zeros stay zeros under the
    scaling).  model.c_first / c_last are the already scaled planes (level_model + scaling_scalars).
    Returns (phi, q, z or None, alpha, beta, c)."""
    nt, nx, ny = model.nt, model.nx, model.ny
    P = nx * ny
    plane = initial_phi_plane(nx, ny)
    if dScale is not None:
        plane = (1 / dScale) * plane
    phi = np.tile(plane, tn1 - tn0)
    Lloc = (tc1 - tc0) * P
    Qloc = Lloc + (tn1 - tn0) * ((nx - 1) * ny + nx * (ny - 1))
    ncol = 10 if model.dim == 2 else 6
    c = np.zeros((tn1 - tn0) * P)
    if tn0 == 0:
        c[:P] = np.ravel(model.c_first)
    if tn1 == nt:
        c[c.size - P:] = np.ravel(model.c_last)
    z = np.zeros((Lloc, ncol), order="F") if with_z else None
    return phi, np.zeros(Qloc), z, np.zeros(Qloc), np.zeros((Lloc, ncol), order="F"), c


def normL2(x, h):
    return math.sqrt(h) * float(np.linalg.norm(np.ravel(x, order="K")))


def level_model(rho0, rho1, nt):
    """The part of initialize.m:1-44 that a level with device-side transitions needs: the grid, model.grad as the triple of
    forward-difference weights and the two non-zero planes of model.c -- no N-sized host array is built."""
    rho0 = np.asarray(rho0, dtype=np.float64)
    rho1 = np.asarray(rho1, dtype=np.float64)
    if rho0.ndim == 1:
        nx, ny, dim = rho0.size, 1, 1
        r0, r1 = rho0, rho1
    else:
        ny, nx = rho0.shape
        dim = 2
        r0, r1 = rho0.ravel(order="F"), rho1.ravel(order="F")
    ht, hx = 1 / (nt - 1), 1 / (nx - 1)
    hy = 1 / (ny - 1) if ny > 1 else 1.0
    model = SimpleNamespace(rho0=rho0, rho1=rho1, nt=nt, nx=nx, ny=ny, dim=dim, c=None, c_first=-r0 / ht, c_last=r1 / ht,
                            grad=(1 / ht, 1 / hx, (1 / hy) if ny > 1 else 0.0))
    L = (nt - 1) * nx * ny
    qInd = SimpleNamespace(bx=L + 1, by=L + nt * (nx - 1) * ny + 1)
    return SimpleNamespace(qInd=qInd), model


def norm_c_planes(model):
    """||model.c||_2 from its two non-zero planes (the same value for a full model.c and for level_model's planes)"""
    a, b = np.ravel(model.c_first), np.ravel(model.c_last)
    return math.sqrt(float(np.dot(a, a)) + float(np.dot(b, b)))


def _scale_c(model, fn):
    """apply fn to model.c -- the full vector (whose planes are views of it) or, without one, the two planes"""
    if model.c is not None:
        n = model.c_first.size
        model.c = fn(model.c)
        model.c_first, model.c_last = model.c[:n], model.c[model.c.size - n:]
    else:
        model.c_first, model.c_last = fn(model.c_first), fn(model.c_last)


def scaling_scalars(N, model, scalingYes, lastLevelKKT, E2_prev, variant):
    """The scalar half of InitialScaling (solver_dotsocp2d.m:304-336 ; solver_wdotsocp2d.m:297-320 ; solver_dotsocp1d.m
    :263-290): returns (cScale, dScale, D, E, Escale2) and rescales model.c / model.grad / model.normc / model.normd.
    Needs only host scalars, model.c and the weight, so the resident multilevel path can run it before the state exists."""
    h = 1 / N
    hMean = h ** (1 / 2) if variant == "dot1d" else h ** (1 / 3)
    if lastLevelKKT is None or E2_prev is None:
        Escale2 = math.sqrt(2)
    elif variant == "wdot2d":
        Escale2 = E2_prev * min(4, max(1 / 4, math.sqrt(lastLevelKKT[0] / lastLevelKKT[1])))
    else:
        ratio = math.sqrt(lastLevelKKT[0] / lastLevelKKT[1])
        if ratio < 0.8333:
            Escale2 = E2_prev * max(1 / math.sqrt(2), ratio / 0.8333)
        else:
            Escale2 = E2_prev * min(math.sqrt(2), max(1, ratio))
    if scalingYes:
        norm_c = (math.sqrt(h) * norm_c_planes(model)) * math.sqrt(model.nt)
        norm_d = math.sqrt(2)
        if variant == "wdot2d":
            if isinstance(model.weight, S.DeviceWeight):     # level of a device pyramid: fixed-order reduction in HBM
                adjust = 10 ** model.weight.log10_mean()
            else:
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
        _scale_c(model, (lambda c: (1.0 / cScale) * c) if variant == "dot2d" else (lambda c: c / cScale))
        model.grad = tuple(D * g for g in model.grad)
    else:
        cScale = dScale = D = E = 1
        model.normc = math.sqrt(h) * norm_c_planes(model)
        model.normd = math.sqrt(2)
    return cScale, dScale, D, E, Escale2


def InitialScaling(var, model, scalingYes, lastLevelKKT, variant):
    """solver_dotsocp2d.m:304-365 ; solver_wdotsocp2d.m:297-342 ; solver_dotsocp1d.m:263-313"""
    cScale, dScale, D, E, Escale2 = scaling_scalars(var.phi.size, model, scalingYes, lastLevelKKT, getattr(var, "E2", None),
                                                    variant)
    if scalingYes:
        var.phi = (1 / dScale) * var.phi
        var.q = (D / dScale) * var.q
        var.z = (E / dScale) * var.z
        var.alpha = (1 / cScale / D) * var.alpha
        var.beta = (1 / cScale / E) * var.beta
    var.cScale, var.dScale, var.D, var.E, var.E2 = cScale, dScale, D, E, Escale2


def _scaled(a, s, inplace):
    """s * a ; inplace: overwrite a (chunks on a few threads -- numpy releases the GIL inside the multiply), same bits"""
    if not inplace or a.size < (1 << 20) or not a.flags.writeable:
        return s * a
    from concurrent.futures import ThreadPoolExecutor
    v = a.reshape(-1, order="A")
    step = -(-v.size // 16)
    with ThreadPoolExecutor(8) as ex:
        list(ex.map(lambda i: np.multiply(v[i:i + step], s, out=v[i:i + step]), range(0, v.size, step)))
    return a


def recoverOrgVar(var, inplace=False):
    """solver_dotsocp2d.m:368-386.  inplace=True (used by the multilevel drivers on the arrays they have just downloaded and
    own) overwrites the arrays instead of allocating new ones."""
    cScale, dScale, D, E = var.cScale, var.dScale, var.D, var.E
    var.phi = _scaled(var.phi, dScale, inplace)
    var.z = _scaled(var.z, dScale / E, inplace)
    var.q = _scaled(var.q, dScale / D, inplace)
    var.alpha = _scaled(var.alpha, cScale * D, inplace)
    var.beta = _scaled(var.beta, cScale * E, inplace)


# ------------------------------------------------------------------------------------------------ level transfer
def _refine_axis(a, ax):
    """insert the pair averages between neighbours along `ax` (n -> 2n-1): linear nodal interpolation"""
    n = a.shape[ax]
    sh = list(a.shape)
    sh[ax] = 2 * n - 1
    r = np.empty(sh)
    lo = [slice(None)] * a.ndim
    hi = [slice(None)] * a.ndim
    ev = [slice(None)] * a.ndim
    od = [slice(None)] * a.ndim
    lo[ax], hi[ax] = slice(0, n - 1), slice(1, n)
    ev[ax], od[ax] = slice(0, None, 2), slice(1, None, 2)
    r[tuple(ev)] = a
    r[tuple(od)] = (a[tuple(lo)] + a[tuple(hi)]) / 2
    return r


def interpolate(var, model):
    """socp/dot2d/utils/interpolate.m:1-85 ; dot1d/utils/interpolate.m:1-72 : phi linear in y, x, t; beta nearest in t,
    linear in y then x, column by column."""
    shape = (model.nt, model.nx, model.ny) if model.dim == 2 else (model.nt, model.nx)
    a = var.phi.reshape(shape)
    for ax in range(a.ndim - 1, -1, -1):
        a = _refine_axis(a, ax)
    var.phi = a.ravel()
    cshape = (model.nt - 1,) + shape[1:]
    cols = []
    for j in range(var.beta.shape[1]):
        b = np.repeat(var.beta[:, j].reshape(cshape), 2, axis=0)
        for ax in range(b.ndim - 1, 0, -1):
            b = _refine_axis(b, ax)
        cols.append(b.ravel())
    var.beta = np.asfortranarray(np.stack(cols, axis=1))
    return var


def grad_apply(phi, nt, nx, ny, grad):
    """q = model.grad * phi with the forward-difference stencils (row: (-g) phi_i + g phi_{i+1}), initialize.m:67-87"""
    gt, gx, gy = grad
    p = phi.reshape(nt, nx, ny)
    q0 = (-gt) * p[:-1] + gt * p[1:]
    bx = (-gx) * p[:, :-1] + gx * p[:, 1:]
    parts = [q0.ravel(), bx.ravel()]
    if ny > 1:
        by = (-gy) * p[:, :, :-1] + gy * p[:, :, 1:]
        parts.append(by.ravel())
    return np.concatenate(parts)


def jump_nextLevel(var, model, rho0, rho1, nt, weight=None):
    """socp/dot2d/utils/jump_nextLevel.m:1-18 ; wdot2d :1-17 ; dot1d :1-18.  alpha = (BF)^*(-beta) runs on the GPU
    through the same C-ABI kernel that replaces mexBFdConj (jump_nextLevel.m:16)."""
    varR = SimpleNamespace(**interpolate(var, model).__dict__)
    var_init, modelR = initialize(rho0, rho1, nt)
    varR.qInd = var_init.qInd
    varR.z = var_init.z
    varR.q = grad_apply(varR.phi, modelR.nt, modelR.nx, modelR.ny, modelR.grad)
    if weight is not None:
        modelR.weight = weight
        varR.q = varR.q / weight
    varR.alpha = var_init.alpha
    nb = np.asfortranarray(-varR.beta)
    if modelR.dim == 2:
        ops.mexBFdConj(varR.alpha, nb, modelR.nt, modelR.nx, modelR.ny, 1.0)
    else:
        ops.mexBFdConj1d(varR.alpha, nb, modelR.nt, modelR.nx, 1.0)
    if weight is not None:
        varR.alpha = varR.alpha / weight
    return varR, modelR


def downSample_phi(v):
    """socp/dot2d/utils/downSample_phi.m:5-34 (full weighting; literal, incl. the first corner that uses v(2,1) twice)
    and socp/dot1d/utils/downSample_phi.m:4-14."""
    v = np.asarray(v, dtype=np.float64)
    if v.ndim == 1:
        ln = v.size - 1
        lenc = ln // 2
        out = np.zeros(lenc + 1)
        ind = np.arange(2, ln - 1, 2)
        out[1:lenc] = 0.5 * v[ind] + 0.25 * (v[ind - 1] + v[ind + 1])
        out[0] = (2 / 3) * v[0] + (1 / 3) * v[1]
        out[-1] = (1 / 3) * v[-2] + (2 / 3) * v[-1]
        return out
    Mx, My = v.shape[0] - 1, v.shape[1] - 1
    Mxc, Myc = Mx // 2, My // 2
    vc = np.zeros((Mxc + 1, Myc + 1))
    ind = np.arange(2, Mx - 1, 2)
    I, J = ind[:, None], ind[None, :]
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


def _restrict_linear(a, ax):
    """transpose of the column-normalised linear prolongation (downSample_q.m:25-31, :10-12): weights (.5,1,.5)/2
    inside, (1,.5)/1.5 at the two ends."""
    a = np.moveaxis(a, ax, 0)
    nR = a.shape[0]
    nC = (nR + 1) // 2
    out = np.empty((nC,) + a.shape[1:])
    out[1:-1] = (a[2:-2:2] + 0.5 * (a[1:-3:2] + a[3:-1:2])) / 2.0
    out[0] = (a[0] + 0.5 * a[1]) / 1.5
    out[-1] = (a[-1] + 0.5 * a[-2]) / 1.5
    return np.moveaxis(out, 0, ax)


def _restrict_nearest(a, ax):
    """transpose of the column-normalised nearest prolongation (downSample_q.m:33-39): pair average"""
    a = np.moveaxis(a, ax, 0)
    out = (a[0::2] + a[1::2]) / 2.0
    return np.moveaxis(out, 0, ax)


def downSample_q(nt, nx, ny, q):
    """socp/wdot2d/utils/downSample_q.m:4-19, as separable stencils instead of sparse kron products"""
    L = (nt - 1) * nx * ny
    nb = L + nt * (nx - 1) * ny
    a = q[:L].reshape(nt - 1, nx, ny)
    a = _restrict_linear(_restrict_linear(_restrict_nearest(a, 0), 1), 2)
    b = q[L:nb].reshape(nt, nx - 1, ny)
    b = _restrict_linear(_restrict_nearest(_restrict_linear(b, 0), 1), 2)
    c = q[nb:].reshape(nt, nx, ny - 1)
    c = _restrict_nearest(_restrict_linear(_restrict_linear(c, 0), 1), 2)
    return np.concatenate([a.ravel(), b.ravel(), c.ravel()])


def downSample_barrier(nt, nx, ny, weight):
    """socp/wdot2d/utils/downSample_barrier.m:4-24"""
    return np.exp(downSample_q(nt, nx, ny, np.log(weight)))


def _stag_grids(nx, ny):
    hx, hy = 1 / (nx - 1), 1 / (ny - 1)
    return (np.linspace(.5 * hx, 1 - .5 * hx, nx - 1), np.linspace(0, 1, nx),
            np.linspace(.5 * hy, 1 - .5 * hy, ny - 1), np.linspace(0, 1, ny))


def weight_planes_circle(nx, ny):
    """The two (x,y) planes of examples/wdot2d/gene_weight_circle.m:6-22: distance to (.5,.5) on the bx / by edge grids, each
    normalised to the sum ny*(nx-1) (the reference uses that count for BOTH planes, :18,:22).  MATLAB-shaped (ny, nx-1), (ny-1, nx)."""
    xS, xC, yS, yC = _stag_grids(nx, ny)
    dist = lambda xx, yy: np.sqrt((xx - .5) ** 2 + (yy - .5) ** 2)
    wX = dist(*np.meshgrid(xS, yC))
    wX = wX * (ny * (nx - 1) / wX.sum())
    wY = dist(*np.meshgrid(xC, yS))
    wY = wY * (ny * (nx - 1) / wY.sum())
    return wX, wY


def weight_planes_barrier(nx, ny, barrier, barrierWeight=1e6):
    """The two planes of examples/wdot2d/get_weight_by_barrier.m:12-28: barrierWeight where barrier(x, y) > 0 on the bx / by
    edge grids, 1 elsewhere.  barrier takes (nx', ny')-shaped coordinate arrays like the reference's function handles."""
    xS, xC, yS, yC = _stag_grids(nx, ny)
    xx, yy = np.meshgrid(xS, yC)
    wX = np.where((barrier(xx.T, yy.T) > 0).T, float(barrierWeight), 1.0)
    xx, yy = np.meshgrid(xC, yS)
    wY = np.where((barrier(xx.T, yy.T) > 0).T, float(barrierWeight), 1.0)
    return wX, wY


def weight_from_planes(nt, wX, wY):
    """[ones ; repmat(weightX, nt) ; repmat(weightY, nt)] (gene_weight_circle.m:24-27, get_weight_by_barrier.m:30-33) on the host"""
    ny, nxm1 = wX.shape
    return np.concatenate([np.ones((nt - 1) * (nxm1 + 1) * ny), np.tile(wX.ravel(order="F"), nt), np.tile(wY.ravel(order="F"), nt)])


def ensure_barrier_validity(rho0, rho1, barrier):
    """examples/wdot2d/ensure_barrier_validity.m:4-16"""
    ny, nx = rho0.shape
    xx, yy = np.meshgrid(np.linspace(0, 1, nx), np.linspace(0, 1, ny))
    b = np.asarray(barrier(xx.T, yy.T), dtype=float)
    bc = b.T > b.mean()
    rho0 = rho0.copy()
    rho1 = rho1.copy()
    rho0[bc] = 0
    rho1[bc] = 0
    return (nx * ny / rho0.sum()) * rho0, (nx * ny / rho1.sum()) * rho1, bc


# ------------------------------------------------------------------------------------------------ output recovery
def _pair_mean_padded(e, ax):
    """zeros at both ends of `ax`, the means of adjacent pairs in between: (n) -> (n+1), written into one fresh array"""
    shape = list(e.shape)
    shape[ax] += 1
    out = np.zeros(shape)
    lo = [slice(None)] * e.ndim
    hi = [slice(None)] * e.ndim
    mid = [slice(None)] * e.ndim
    lo[ax], hi[ax], mid[ax] = slice(0, -1), slice(1, None), slice(1, -1)
    dst = out[tuple(mid)]
    np.add(e[tuple(lo)], e[tuple(hi)], out=dst)
    dst /= 2
    return out


def recover_RhoE(var, model):
    """socp/dot2d/utils/recover_RhoE.m:13-25 (wdot2d :11 multiplies alpha by the weight) ; dot1d :12-20.
    Returns C-order arrays (nt, nx, ny) [1-D: (nt, nx)]."""
    nt, nx, ny = model.nt, model.nx, model.ny
    alpha = var.alpha
    if getattr(model, "weight", None) is not None:
        alpha = model.weight * alpha
    L = (nt - 1) * nx * ny
    nb = L + nt * (nx - 1) * ny
    shp = (nx, ny) if model.dim == 2 else (nx,)
    r0 = model.rho0.T if model.dim == 2 else model.rho0
    r1 = model.rho1.T if model.dim == 2 else model.rho1
    a0 = alpha[:L].reshape((nt - 1,) + shp)
    rho = np.empty((nt,) + shp)
    rho[0], rho[-1] = r0, r1
    np.add(a0[:-1], a0[1:], out=rho[1:-1])
    rho[1:-1] /= 2

    def centre(e, ax):
        e = e.copy()                      # first and last time level count twice (recover_RhoE.m:17-18)
        e[0] *= 2
        e[-1] *= 2
        return _pair_mean_padded(e, ax)
    if model.dim == 2:
        Ex = centre(alpha[L:nb].reshape(nt, nx - 1, ny), 1)
        Ey = centre(alpha[nb:].reshape(nt, nx, ny - 1), 2)
        return rho, Ex, Ey
    return rho, centre(alpha[L:].reshape(nt, nx - 1), 1)


def recover_q(var, model):
    """socp/dot2d/utils/recover_q.m:12-22 ; dot1d :11-18"""
    nt, nx, ny = model.nt, model.nx, model.ny
    q = var.q
    L = (nt - 1) * nx * ny
    nb = L + nt * (nx - 1) * ny

    def centre(e, ax):
        e = _pair_mean_padded(e, ax)
        out = e[:-1] + e[1:]
        out /= 2
        return out
    if model.dim == 2:
        return (q[:L].reshape(nt - 1, nx, ny), centre(q[L:nb].reshape(nt, nx - 1, ny), 1),
                centre(q[nb:].reshape(nt, nx, ny - 1), 2))
    return q[:L].reshape(nt - 1, nx), centre(q[L:].reshape(nt, nx - 1), 1)


def check_massConservation(rho, tol=1e-2):
    """socp/dot2d/utils/check_massConservation.m:16-34"""
    r2 = rho.reshape(rho.shape[0], -1)
    sumRho = r2.mean(axis=1)
    sumNeg = np.where(r2 < 0, r2, 0.0).mean(axis=1)
    err = max(np.abs(sumRho - 1).max(), np.abs(sumNeg).max())
    return bool(err <= tol), sumRho, sumNeg


def w2_cost(output, dim=2):
    """transport cost  mean over the space-time grid of |m|^2/rho  (implementation-independent check)"""
    rho = output.rho
    m2 = output.Ex ** 2 + (output.Ey ** 2 if dim == 2 else 0.0)
    val = np.zeros_like(rho)
    mask = rho > 1e-12
    val[mask] = m2[mask] / rho[mask]
    return float(val.mean())


# ------------------------------------------------------------------------------------------------ multilevel drivers
def _cat_hist(ML, rh):
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


_VALID = {"dot2d": ["PALM", "inPALM", "ALG2", "acc-ADMM", "sGS-inPALM", "acc-sGS-ADMM"],
          "wdot2d": ["inPALM", "ALG2", "acc-ADMM"], "dot1d": ["inPALM", "ALG2"]}


def _multilevel(variant, rho0, rho1, nt, levelN, opts, method, barrier=None):
    opts = dict(opts)
    if not (isinstance(levelN, (int, np.integer)) and levelN >= 1):
        raise ValueError("Invalid input at position 4 (Number of levels in multilevel strategy)")
    if method not in _VALID[variant]:
        raise ValueError("Invalid input at position 6 (Solving method)")
    sgsMethod = method in ("sGS-inPALM", "acc-sGS-ADMM")                 # solver_dotsocp2d.m:93-97
    admmMaxIt, sgsMaxIt = 3000, 6000
    opts.setdefault("ifCheckStepByStep", False)
    scalingYes = opts.setdefault("scaling", True)
    optsML = dict(opts)
    if "maxit" not in opts:
        optsML["maxit"] = int(1e4) if variant == "wdot2d" else (sgsMaxIt if sgsMethod else admmMaxIt)
    optsML["tolFactor"] = -1 if optsML["tol"] > 0.99e-3 else -0.5
    tolLowerBound = 1e-5 if variant == "dot1d" else 1e-4
    if method in ("PALM", "inPALM", "sGS-inPALM"):
        optsML["tau"] = 1.9
    elif method == "ALG2":
        optsML["tau"] = 1.0
    optsML.setdefault("sigma", 0.1 if sgsMethod else 1)                  # :139-146
    optsML.setdefault("time_limit", 3600)
    weight = opts.get("weight") if variant == "wdot2d" else None
    planes = opts.get("weight_planes") if variant == "wdot2d" else None     # (weightX, weightY) of the generators
    if variant == "wdot2d" and weight is None and planes is None:
        raise ValueError("opts.weight is required")
    # opts["weights_on_device"] (implied by weight_planes): the finest weight, its restriction chain and the log-means live in a
    # device pyramid (solver.Weights): no Q-sized host array per level.  Resident path only.
    on_device = variant == "wdot2d" and (planes is not None or bool(opts.get("weights_on_device", False)))
    for k in ("weight", "weight_planes", "weights_on_device"):
        optsML.pop(k, None)
    rho0s, rho1s, nts, tols, weights = ([None] * levelN for _ in range(5))
    rho0s[-1], rho1s[-1] = np.asarray(rho0, float), np.asarray(rho1, float)
    nts[-1], tols[-1], weights[-1] = int(nt), optsML["tol"], weight
    nxs, nys = [None] * levelN, [None] * levelN
    if variant != "dot1d":
        nys[-1], nxs[-1] = rho0s[-1].shape
    pyramid = None
    if on_device:
        if not optsML.get("resident", True):
            raise ValueError("weights_on_device needs the resident multilevel path")
        pyramid = S.Weights(nts[-1], nxs[-1], nys[-1], levelN)
        if planes is not None:
            pyramid.set_planes(*planes)
        else:
            pyramid.set(weight)
        pyramid.restrict(geometric=barrier is not None)        # downSample_barrier / downSample_q chain (:179-187)
        weights = [pyramid.level(levelN - 1 - lv) for lv in range(levelN)]
        weight = None
    for lv in range(levelN - 2, -1, -1):
        nts[lv] = (nts[lv + 1] - 1) // 2 + 1
        tols[lv] = max(tols[lv + 1] * 2 ** optsML["tolFactor"], tolLowerBound)
        rho0s[lv] = downSample_phi(rho0s[lv + 1])
        rho1s[lv] = downSample_phi(rho1s[lv + 1])
        if variant != "dot1d":
            nxs[lv], nys[lv] = (nxs[lv + 1] + 1) // 2, (nys[lv + 1] + 1) // 2
        if variant == "wdot2d" and barrier is not None:
            rho0s[lv], rho1s[lv], _ = ensure_barrier_validity(rho0s[lv], rho1s[lv], barrier)
            if pyramid is None:
                weights[lv] = downSample_barrier(nts[lv + 1], nxs[lv + 1], nys[lv + 1], weights[lv + 1])
        else:
            if variant == "wdot2d" and pyramid is None:
                weights[lv] = downSample_q(nts[lv + 1], nxs[lv + 1], nys[lv + 1], weights[lv + 1])
            N = rho0s[lv].size
            rho0s[lv] = rho0s[lv] / (rho0s[lv].sum() / N)
            rho1s[lv] = rho1s[lv] / (rho1s[lv].sum() / N)
    timeML = [None] * (levelN + 1)
    lastLevelKKT = None
    clk = time.perf_counter()
    # state resident in HBM between levels (device-side transitions) unless opts["resident"] = False asks for the
    # reference-shaped download -> host transfer -> upload loop; both give the same bits
    resident = bool(optsML.pop("resident", True))
    opts.pop("resident", None)
    if resident:      # the coarsest state is built slab by slab (initial_state_local): no full-grid host array
        var, model = level_model(rho0s[0], rho1s[0], nts[0])
    else:
        var, model = initialize(rho0s[0], rho1s[0], nts[0])
    if variant == "wdot2d":
        model.weight = weights[0]
    ML = SimpleNamespace(kkt=None, time=None, iter=None, pdGap=None, len=0)
    runHist = None
    level_iters, launches = [], 0.0
    sigma = optsML["sigma"]
    if resident:
        try:
            return _multilevel_resident(variant, method, levelN, optsML, scalingYes, var, model, rho0s, rho1s, nts, tols, weights,
                                        timeML, ML, clk)
        finally:
            if pyramid is not None:
                pyramid.close()
    for level in range(levelN):
        InitialScaling(var, model, scalingYes, lastLevelKKT, variant)
        o2 = dict(optsML)
        o2["tol"] = tols[level]
        if method == "PALM":
            runHist, sigma = S.solver_socp_PALM(var, o2, model)
        elif method in ("inPALM", "ALG2"):
            runHist, sigma = (S.solver_wsocp_inPALM if variant == "wdot2d" else S.solver_socp_inPALM)(var, o2, model)
        elif sgsMethod:                                                  # :210-223: the sGS loop on the last level only
            if level == levelN - 1:
                runHist, sigma = (S.solver_socp_sGSinPALM if method == "sGS-inPALM" else S.solver_socp_accsGSADMM)(var, o2, model)
            else:
                o2["maxit"] = admmMaxIt
                o2["tau"] = 1.9
                runHist, sigma = S.solver_socp_inPALM(var, o2, model)
        else:
            runHist, sigma = (S.solver_wsocp_accADMM if variant == "wdot2d" else S.solver_socp_accADMM)(var, o2, model)
        recoverOrgVar(var, inplace=True)          # the level solver returned freshly downloaded arrays
        timeML[level] = var.time
        level_iters.append(var.time["Iters"])
        launches += getattr(var, "gpu_launches", 0.0)
        _cat_hist(ML, runHist)
        if level < levelN - 1:
            optsML["time_limit"] = optsML["time_limit"] - var.time["Total_Time"]
            optsML["sigma"] = 10 ** (math.log10(optsML["sigma"] * sigma) / 2)
            var, model = jump_nextLevel(var, model, rho0s[level + 1], rho1s[level + 1], nts[level + 1],
                                        weights[level + 1] if variant == "wdot2d" else None)
            lastLevelKKT = runHist.kkt[-1, :]
    output = SimpleNamespace()
    if variant == "dot1d":
        output.rho, output.Ex = recover_RhoE(var, model)
        output.q0, output.bx = recover_q(var, model)
    else:
        output.rho, output.Ex, output.Ey = recover_RhoE(var, model)
        output.q0, output.bx, output.by = recover_q(var, model)
    output.massOK, output.sumRho, output.sumNegRho = check_massConservation(output.rho, 1e-2)
    output.var, output.model, output.level_iters, output.sigma, output.gpu_launches = var, model, level_iters, sigma, launches
    timeML[levelN] = {"ML_Time": time.perf_counter() - clk}
    return output, timeML, ML, runHist


def _multilevel_resident(variant, method, levelN, optsML, scalingYes, var, model, rho0s, rho1s, nts, tols, weights, timeML, ML,
                         clk):
    """The same multilevel loop with the state resident in HBM between levels (the default; opts["resident"] = False selects
    the host-transition loop): the transitions of solver_dotsocp2d.m:230-250 run on the device (dotsocp_prolong) instead of
    download -> host -> upload, and after the last level the outputs (rho, Ex, Ey, q0, bx, by, mass check) are recovered on
    the device as well (dotsocp_recover), so only 6N doubles ever cross PCIe.  Bit-identical to the host path.

    opts["slabs"]: None = one GPU; an int k = k time slabs emulated on this GPU; {"rank", "world", "nccl_id"} = this process
    owns slab `rank` of a one-process-per-GPU run (every rank calls the driver with the same arguments; the output fields
    then hold the rank's own time levels, output.slab = (first level, one past the last), the scalars are global).
    opts["return_state"] = True also downloads the final iterates into output.var (recoverOrgVar applied)."""
    from . import slab as SL
    weighted = variant == "wdot2d"
    slabs = optsML.pop("slabs", None)
    return_state = bool(optsML.pop("return_state", False))
    if isinstance(slabs, dict):
        rank, world, ident, distributed = int(slabs["rank"]), int(slabs["world"]), slabs["nccl_id"], True
    else:
        rank, world, ident, distributed = 0, int(slabs or 1), None, False
    # coarsest level: the scalar half of InitialScaling on the host, the arrays (phi = |x|^2/2, zeros) slab by slab
    cS0, dS0, D0, E0, E20 = scaling_scalars(model.nt * model.nx * model.ny, model, scalingYes, None, None, variant)
    var.cScale, var.dScale, var.D, var.E, var.E2 = cS0, dS0, D0, E0, E20
    sess = S.Session(variant, model.nt, model.nx, model.ny, rank=rank, world=world, nccl_id=ident)
    sgs_last = method if method in ("sGS-inPALM", "acc-sGS-ADMM") else None   # coarse levels: inPALM (tau 1.9, maxit 3000), :210-223
    mname = "inPALM" if method in ("inPALM", "ALG2", "sGS-inPALM", "acc-sGS-ADMM") else method
    first_is_inpalm = mname == "inPALM" and not (sgs_last == "acc-sGS-ADMM" and levelN == 1)   # acc loops read the incoming z
    z_dead = first_is_inpalm and int(optsML["maxit"]) >= 1
    tr0 = SL.partition(model.nt, world, sess.cuts)[rank] if distributed else (0, model.nt - 1, 0, model.nt)
    w0 = model.weight if weighted else None
    if distributed and w0 is not None and not isinstance(w0, S.DeviceWeight):   # (a device level is taken slab by slab by the session)
        w0 = SL.split_state(rank, world, model.nt, model.nx, model.ny, None, None, None, None, None, None, w0, cuts=sess.cuts)[6]
    state0 = initial_state_local(model, dS0 if scalingYes else None, *tr0, with_z=not z_dead) + (w0,)
    sess.upload(*state0)
    del state0
    level_iters, launches, runHist, sigma = [], 0.0, None, optsML["sigma"]
    outbuf = None
    try:
        for level in range(levelN):
            o2 = dict(optsML)
            o2["tol"] = tols[level]
            lname = mname
            if sgs_last:
                if level == levelN - 1:
                    lname = sgs_last
                else:
                    o2["maxit"] = 3000
                    o2["tau"] = 1.9
            lo = S.make_level_opts(variant, lname, var, o2, model)
            if level == levelN - 1:     # the output arrays are allocated and their pages faulted in while the last level runs
                outbuf = S.OutputBuffers(sess)
            hb, res = sess.run(lo)
            runHist, sigma = S._finish(var, lo.method, hb, res)      # var.cScale/dScale/D/E after in-loop rescaling, var.time
            timeML[level] = var.time
            level_iters.append(var.time["Iters"])
            launches += var.gpu_launches
            _cat_hist(ML, runHist)
            if level == levelN - 1:
                break
            optsML["time_limit"] = optsML["time_limit"] - var.time["Total_Time"]
            optsML["sigma"] = 10 ** (math.log10(optsML["sigma"] * sigma) / 2)
            var_f, model_f = level_model(rho0s[level + 1], rho1s[level + 1], nts[level + 1])
            if weighted:
                model_f.weight = weights[level + 1]
            gt, gx, gy = model_f.grad                               # unscaled 1/ht, 1/hx, 1/hy
            Nf = model_f.nt * model_f.nx * model_f.ny
            cS, dS, D, E, E2 = scaling_scalars(Nf, model_f, scalingYes, runHist.kkt[-1, :], var.E2, variant)
            scal = dict(phi_recover=var.dScale, beta_recover=var.cScale * var.E, grad_t=gt, grad_x=gx, grad_y=gy,
                        phi_scale=1 / dS, q_scale=D / dS, alpha_scale=1 / cS / D, beta_scale=1 / cS / E)
            fine = S.Session.refined(sess)
            try:
                w_f = model_f.weight if weighted else None
                if weighted and distributed and not isinstance(w_f, S.DeviceWeight):
                    w_f = SL.split_state(rank, world, model_f.nt, model_f.nx, model_f.ny, None, None, None, None, None, None, w_f,
                                         cuts=fine.cuts)[6]
                fine.prolong_from(sess, scal, weight=w_f, c_first=model_f.c_first, c_last=model_f.c_last)
            except Exception:
                fine.close()
                raise
            sess.close()
            sess = fine
            var = SimpleNamespace(qInd=var_f.qInd, cScale=cS, dScale=dS, D=D, E=E, E2=E2)
            model = model_f
        # ---- output (solver_dotsocp2d.m:268-287) on the device
        fields, sumRho, sumNeg, w2 = sess.recover(var.cScale * var.D, var.dScale / var.D, model.rho0, model.rho1, out=outbuf.get())
        if return_state:
            var.phi, var.q, var.z, var.alpha, var.beta = sess.download()
            recoverOrgVar(var, inplace=True)
        tc0, tc1, tn0, tn1 = SL.partition(model.nt, world, sess.cuts)[rank] if distributed else (0, model.nt - 1, 0, model.nt)
    finally:
        sess.close()
    output = SimpleNamespace(**fields)
    output.sumRho, output.sumNegRho = sumRho, sumNeg
    output.massOK = bool(max(np.abs(sumRho - 1).max(), np.abs(sumNeg).max()) <= 1e-2)     # check_massConservation.m:26-34
    output.w2 = w2
    output.slab = (tn0, tn1)
    output.var = var if return_state else None
    output.model, output.level_iters, output.sigma, output.gpu_launches = model, level_iters, sigma, launches
    timeML[levelN] = {"ML_Time": time.perf_counter() - clk}
    return output, timeML, ML, runHist


def solver_dotsocp2d(rho0, rho1, nt, levelN, opts, method="inPALM"):
    """socp/dot2d/solver_dotsocp2d.m:1"""
    return _multilevel("dot2d", rho0, rho1, nt, levelN, opts, method)


def solver_wdotsocp2d(rho0, rho1, nt, levelN, opts, method="inPALM", barrier=None):
    """socp/wdot2d/solver_wdotsocp2d.m:1"""
    return _multilevel("wdot2d", rho0, rho1, nt, levelN, opts, method, barrier)


def solver_dotsocp1d(rho0, rho1, nt, levelN, opts, method="inPALM"):
    """socp/dot1d/solver_dotsocp1d.m:1"""
    return _multilevel("dot1d", rho0, rho1, nt, levelN, opts, method)
