"""TEST INFRASTRUCTURE ONLY (oracle/): the reference's four native kernels, three ways.

backend "ref"   -> the genuine reference MEX binaries through oracle/refmex.py (only where oracle/_ref exists)
backend "c"     -> oracle/ckernels.c (plain-C restatement, bit-identical to the binaries; see its header)
backend "numpy" -> vectorised numpy restatement below (same arithmetic, also bit-identical: every value
                   is produced by the same sequence of IEEE double operations)

All functions keep the reference's in-place convention: the FIRST argument is overwritten
(SURVEY.md §8b "Ownership").  Arrays are float64; 2-D arrays are column-major (order='F').
"""
from __future__ import annotations

import ctypes
import os
import struct

import numpy as np

from . import refmex

_HERE = os.path.dirname(os.path.abspath(__file__))
_CLIB_PATH = os.path.join(_HERE, "liboracle_kernels.so")
_clib = None

# mexBFd.mexa64 .rodata @0x2000 (0x3fe6a09e667f3bd1): the literal 0.707106781186548, not sqrt(1/2)
INV_SQRT2_LITERAL = struct.unpack("<d", bytes.fromhex("d13b7f669ea0e63f"))[0]


def c_available() -> bool:
    return os.path.exists(_CLIB_PATH)


def _c():
    global _clib
    if _clib is None:
        lib = ctypes.CDLL(_CLIB_PATH)
        P, I, D, Z = ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_ssize_t
        lib.oracle_BFd.argtypes = [P, P, I, I, I, D, D]
        lib.oracle_BFdConj.argtypes = [P, P, I, I, I, D]
        lib.oracle_ProjSoc.argtypes = [P, P, Z, I]
        lib.oracle_BFd1d.argtypes = [P, P, I, I, D, D]
        lib.oracle_BFdConj1d.argtypes = [P, P, I, I, D]
        for f in (lib.oracle_BFd, lib.oracle_BFdConj, lib.oracle_ProjSoc, lib.oracle_BFd1d, lib.oracle_BFdConj1d):
            f.restype = None
        _clib = lib
    return _clib


def default_backend() -> str:
    forced = os.environ.get("DOTSOCP_ORACLE_BACKEND")
    if forced:
        return forced
    if refmex.available():
        return "ref"
    if c_available():
        return "c"
    return "numpy"


def _chk(a, shape=None):
    assert isinstance(a, np.ndarray) and a.dtype == np.float64
    assert a.ndim == 1 or a.flags.f_contiguous
    if shape is not None:
        assert a.shape == shape, (a.shape, shape)


def sizes2d(nt, nx, ny):
    L = (nt - 1) * nx * ny
    nbx = nt * (nx - 1) * ny
    nby = nt * nx * (ny - 1)
    return L, nbx, nby


# ------------------------------------------------------------------------------------------------
# numpy restatements (mexBFd.mexa64 @0x1120/@0x11a0/@0x1310; mexBFdConj.mexa64 @0x1120/@0x1160/@0x1310;
# mexProjSoc.mexa64 @0x1170/@0x1440).  C-order views (t,x,y) of the MATLAB column-major data.
# ------------------------------------------------------------------------------------------------
def _np_BFd(z, q, nt, nx, ny, S, DF):
    L, nbx, nby = sizes2d(nt, nx, ny)
    SF = INV_SQRT2_LITERAL * S
    q0 = q[:L]
    bx = q[L:L + nbx].reshape(nt, nx - 1, ny)
    by = q[L + nbx:].reshape(nt, nx, ny - 1)
    zc = [z[:, j].reshape(nt - 1, nx, ny) for j in range(10)]
    p = q0 * S
    z[:, 0] = DF - p
    z[:, 9] = p + DF
    bxs = bx * SF
    bys = by * SF
    zc[1][:, 1:, :] = bxs[:-1]
    zc[2][:, :-1, :] = bxs[:-1]
    zc[3][:, 1:, :] = bxs[1:]
    zc[4][:, :-1, :] = bxs[1:]
    zc[5][:, :, 1:] = bys[:-1]
    zc[6][:, :, :-1] = bys[:-1]
    zc[7][:, :, 1:] = bys[1:]
    zc[8][:, :, :-1] = bys[1:]


def _np_BFdConj(q, z, nt, nx, ny, S):
    L, nbx, nby = sizes2d(nt, nx, ny)
    SF = INV_SQRT2_LITERAL * S
    zc = [z[:, j].reshape(nt - 1, nx, ny) for j in range(10)]
    q[:L] = (z[:, 9] - z[:, 0]) * S
    bx = q[L:L + nbx].reshape(nt, nx - 1, ny)
    by = q[L + nbx:].reshape(nt, nx, ny - 1)
    up = zc[1][:, 1:, :] + zc[2][:, :-1, :]          # cells (t, x+1) col1 + (t, x) col2   -> node level t
    bx[0] = up[0] * SF
    bx[nt - 1] = (zc[3][nt - 2, 1:, :] + zc[4][nt - 2, :-1, :]) * SF
    if nt > 2:
        bx[1:nt - 1] = ((up[1:] + zc[3][:-1, 1:, :]) + zc[4][:-1, :-1, :]) * SF
    up = zc[5][:, :, 1:] + zc[6][:, :, :-1]
    by[0] = up[0] * SF
    by[nt - 1] = (zc[7][nt - 2, :, 1:] + zc[8][nt - 2, :, :-1]) * SF
    if nt > 2:
        by[1:nt - 1] = ((up[1:] + zc[7][:-1, :, 1:]) + zc[8][:-1, :, :-1]) * SF


def _np_rownorm(inp):
    """Eigen rowwise().norm() of columns 1..N-1 in the binary's association order (see ckernels.c)."""
    M, N = inp.shape
    n = N - 1
    s = [inp[:, k + 1] * inp[:, k + 1] for k in range(n)]
    acc = s[0].copy()
    kend = (n - 1) & ~3
    k = 1
    if kend > 1:
        while k < kend:
            acc = acc + ((s[k + 3] + s[k + 2]) + (s[k + 1] + s[k]))
            k += 4
    while k < n:
        acc = acc + s[k]
        k += 1
    if M % 2 == 1:  # odd tail row: plain sequential sum (@0x1668)
        a = s[0][M - 1]
        for kk in range(1, n):
            a = a + s[kk][M - 1]
        acc[M - 1] = a
    return np.sqrt(acc)


def _np_ProjSoc(out, inp):
    M, N = inp.shape
    nrm = _np_rownorm(inp)
    v0 = inp[:, 0].copy()
    with np.errstate(divide="ignore", invalid="ignore"):
        r = (v0 / nrm + 1.0) * 0.5
    gt = r > 1.0
    lt = 0.0 > r
    coef = np.where(gt, 1.0, np.where(lt, 0.0, r))
    keep = gt | (~gt & ~lt & (r == 1.0))
    with np.errstate(invalid="ignore"):
        res = inp[:, 1:] * coef[:, None]
        c0 = np.where(keep, v0, coef * nrm)
    out[:, 1:] = res
    out[:, 0] = c0


def _embed1d(z6, L):
    z10 = np.zeros((L, 10), order="F")
    z10[:, 0:5] = z6[:, 0:5]
    z10[:, 9] = z6[:, 5]
    return z10


# ------------------------------------------------------------------------------------------------
# public API (reference signatures)
# ------------------------------------------------------------------------------------------------
def mexBFd(z2, q, nt, nx, ny, scaleBF, scaleD, backend=None):
    nt, nx, ny = int(nt), int(nx), int(ny)
    L, nbx, nby = sizes2d(nt, nx, ny)
    _chk(z2, (L, 10)); _chk(q, (L + nbx + nby,))
    b = backend or default_backend()
    if b == "ref":
        refmex.mexBFd(z2, q, nt, nx, ny, scaleBF, scaleD)
    elif b == "c":
        _c().oracle_BFd(z2.ctypes.data, q.ctypes.data, nt, nx, ny, float(scaleBF), float(scaleD))
    else:
        _np_BFd(z2, q, nt, nx, ny, float(scaleBF), float(scaleD))


def mexBFdConj(q2, z, nt, nx, ny, scaleBF, backend=None):
    nt, nx, ny = int(nt), int(nx), int(ny)
    L, nbx, nby = sizes2d(nt, nx, ny)
    _chk(q2, (L + nbx + nby,)); _chk(z, (L, 10))
    b = backend or default_backend()
    if b == "ref":
        refmex.mexBFdConj(q2, z, nt, nx, ny, scaleBF)
    elif b == "c":
        _c().oracle_BFdConj(q2.ctypes.data, z.ctypes.data, nt, nx, ny, float(scaleBF))
    else:
        _np_BFdConj(q2, z, nt, nx, ny, float(scaleBF))


def mexProjSoc(out, inp, backend=None):
    _chk(out); _chk(inp)
    assert out.shape == inp.shape and inp.ndim == 2
    b = backend or default_backend()
    if b == "ref":
        refmex.mexProjSoc(out, inp)
    elif b == "c":
        _c().oracle_ProjSoc(out.ctypes.data, inp.ctypes.data, inp.shape[0], inp.shape[1])
    else:
        _np_ProjSoc(out, inp)


def mexBFd1d(z, q, nt, nx, scale, dfactor, backend=None):
    nt, nx = int(nt), int(nx)
    L = (nt - 1) * nx
    _chk(z, (L, 6)); _chk(q, (L + nt * (nx - 1),))
    b = backend or default_backend()
    if b == "ref":
        refmex.mexBFd1d(z, q, nt, nx, scale, dfactor)
    elif b == "c":
        _c().oracle_BFd1d(z.ctypes.data, q.ctypes.data, nt, nx, float(scale), float(dfactor))
    else:  # ny = 1 degenerate of the 2-D kernel
        z10 = _embed1d(z, L)
        _np_BFd(z10, q, nt, nx, 1, float(scale), float(dfactor))
        z[:, 0:5] = z10[:, 0:5]
        z[:, 5] = z10[:, 9]


def mexBFdConj1d(q, z, nt, nx, scale, backend=None):
    nt, nx = int(nt), int(nx)
    L = (nt - 1) * nx
    _chk(q, (L + nt * (nx - 1),)); _chk(z, (L, 6))
    b = backend or default_backend()
    if b == "ref":
        refmex.mexBFdConj1d(q, z, nt, nx, scale)
    elif b == "c":
        _c().oracle_BFdConj1d(q.ctypes.data, z.ctypes.data, nt, nx, float(scale))
    else:
        _np_BFdConj(q, _embed1d(z, L), nt, nx, 1, float(scale))


# ------------------------------------------------------------------------------------------------ mexsGS
def _np_sGS(phi, rhs, ep, scale, nt, nx, ny, its):
    """mexsGS.mexa64 (internal name mexRBsGSscaling) restated from the disassembly: `its` symmetric red-black Gauss-Seidel
    sweeps for  scale*(A'A + ep*I) phi = rhs  on the (ny, nx, nt) node grid, in place.

      mexFunction @0x23a0 : H = (1/(nx-1))^2 / scale ; C = ((nt-1)/(nx-1))^2 ; eH = H * (scale*ep) ;
                            COE = 1 / ((n_t*C + n_s) + eH)  with n_t in {1,2} time neighbours (2C = C + C) and n_s in
                            {2,3,4} space neighbours (Inside 2C+4, Face1 C+4, Face2 2C+3, Edge1 C+3, Edge2 2+2C, Corner C+2)
                            order of the half sweeps: odd ; its x [ even (corners last) ; odd ]      (parity of t+x+y)
      RBGS_inside @0x1380, RBGS_face @0x1c80 (+ transforms @0x1120/@0x11b0), RBGS_edge @0x17a0 (+ @0x1250/@0x12e0),
      RBGS_corner @0x1530 : every node of the half sweep's parity becomes
                            ((S + T) + H*rhs) * COE,   S = left-to-right sum of the existing neighbours in the order
                            x-1, x+1, y-1, y+1 ;  T = C*(phi[t-1] + phi[t+1])  or  C*phi[the one time neighbour].
    The binary walks rows with a start that toggles 1 <-> 2 and uses NY-based counts and 2*NX strides on faces and edges, so
    it is only meaningful for nx == ny, both odd; nodes of one parity have all neighbours in the other, so
    the order of the updates inside a half sweep does not matter."""
    assert nx == ny and nx % 2 == 1 and nt % 2 == 1, "mexsGS needs nx == ny and odd node counts"
    p = phi.reshape(nt, nx, ny)
    r = rhs.reshape(nt, nx, ny)
    hx = 1.0 / (nx - 1.0)
    H = (hx * hx) / scale
    c1 = (nt - 1.0) / (nx - 1.0)
    C = c1 * c1
    eH = H * (scale * ep)
    t, x, y = np.meshgrid(np.arange(nt), np.arange(nx), np.arange(ny), indexing="ij")
    par = (t + x + y) & 1
    nts = (t > 0).astype(int) + (t < nt - 1)
    nss = (x > 0).astype(int) + (x < nx - 1) + (y > 0) + (y < ny - 1)
    twoC = C + C
    coe = 1.0 / ((np.where(nts == 2, twoC, C) + nss.astype(float)) + eH)

    def half(parity):
        S = np.zeros_like(p)
        started = np.zeros(p.shape, dtype=bool)
        for ax, sh in ((1, -1), (1, 1), (2, -1), (2, 1)):
            nb = np.roll(p, -sh, axis=ax)
            idx = x if ax == 1 else y
            n = nx if ax == 1 else ny
            ok = (idx + sh >= 0) & (idx + sh <= n - 1)
            S = np.where(ok, np.where(started, S + nb, nb), S)
            started |= ok
        tm, tp = np.roll(p, 1, axis=0), np.roll(p, -1, axis=0)
        T = np.where(nts == 2, (tm + tp) * C, np.where(t == 0, tp, tm) * C)
        new = ((S + T) + r * H) * coe
        m = par == parity
        p[m] = new[m]
    half(1)
    for _ in range(int(its)):
        half(0)
        half(1)


def mexsGS(phi, rhs, ep, scale, nt, nx, ny, its, backend=None):
    """mexsGS(phi, rhs, ep, scale, nt, nx, ny, its): in place into phi   (solver_socp_sGSinPALM.m:205)"""
    b = backend or default_backend()
    if b == "ref":
        refmex.mexsGS(phi, rhs, float(ep), float(scale), float(nt), float(nx), float(ny), float(its))
    else:
        _np_sGS(phi, rhs, float(ep), float(scale), int(nt), int(nx), int(ny), int(its))
