"""Kernel-level entry points with the reference's MEX signatures, running on the GPU through the C ABI.

Same names, argument order, in-place convention (the FIRST argument receives the result) and array layout as the
reference's pre-built MEX kernels:

    mexBFd(z2, q, nt, nx, ny, scaleBF, scaleD)      socp/dot2d/algorithms/solver_socp_inPALM.m:133
    mexBFdConj(q2, z, nt, nx, ny, scaleBF)          socp/dot2d/algorithms/solver_socp_inPALM.m:205
    mexProjSoc(out, in)                             socp/dot2d/algorithms/solver_socp_inPALM.m:199
    mexsGS(phi, rhs, ep, scale, nt, nx, ny, its)    socp/dot2d/algorithms/solver_socp_sGSinPALM.m:205
    mexBFd1d(z, q, nt, nx, scale, dFactor)          socp/dot1d/algorithms/solver_socp_inPALM.m:132
    mexBFdConj1d(q, z, nt, nx, scale)               socp/dot1d/algorithms/solver_socp_inPALM.m:204
    oper_poisson3dim / oper_poisson                  socp/dot2d/utils/oper_poisson3dim.m:4, dot1d/utils/oper_poisson.m:4

They copy host buffers to the device and back on every call, so they are parity/utility entry points, not the
performance path (that is the device-resident session in ``solver.py``).
"""
from __future__ import annotations

import numpy as np

from ._lib import check, lib, ptr


def _vec(a, n, name):
    if not (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.size == n and (a.flags.c_contiguous or a.flags.f_contiguous)):
        raise ValueError(f"{name}: expected a contiguous float64 array with {n} elements")


def _mat(a, shape, name):
    if not (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.shape == shape and a.flags.f_contiguous):
        raise ValueError(f"{name}: expected a column-major float64 array of shape {shape}")


def _sizes(nt, nx, ny):
    L = (nt - 1) * nx * ny
    return L, L + nt * (nx - 1) * ny + nt * nx * (ny - 1)


def mexBFd(z2, q, nt, nx, ny, scaleBF, scaleD):
    nt, nx, ny = int(nt), int(nx), int(ny)
    L, Q = _sizes(nt, nx, ny)
    _mat(z2, (L, 10), "z2"); _vec(q, Q, "q")
    check(lib().dotsocp_mexBFd(ptr(z2), ptr(q), nt, nx, ny, float(scaleBF), float(scaleD)))


def mexBFdConj(q2, z, nt, nx, ny, scaleBF):
    nt, nx, ny = int(nt), int(nx), int(ny)
    L, Q = _sizes(nt, nx, ny)
    _vec(q2, Q, "q2"); _mat(z, (L, 10), "z")
    check(lib().dotsocp_mexBFdConj(ptr(q2), ptr(z), nt, nx, ny, float(scaleBF)))


def mexProjSoc(out, inp):
    if not (isinstance(inp, np.ndarray) and inp.ndim == 2):
        raise ValueError("in: expected a 2-D array")
    _mat(inp, inp.shape, "in"); _mat(out, inp.shape, "out")
    check(lib().dotsocp_mexProjSoc(ptr(out), ptr(inp), inp.shape[0], inp.shape[1]))


def mexsGS(phi, rhs, ep, scale, nt, nx, ny, its):
    """`its` symmetric red-black Gauss-Seidel sweeps for scale*(A'A + ep*I) phi = rhs, in place into phi (nx == ny, odd sizes)"""
    nt, nx, ny = int(nt), int(nx), int(ny)
    _vec(phi, nt * nx * ny, "phi"); _vec(rhs, nt * nx * ny, "rhs")
    check(lib().dotsocp_mexsGS(ptr(phi), ptr(rhs), float(ep), float(scale), nt, nx, ny, int(its)))


def mexBFd1d(z, q, nt, nx, scale, dFactor):
    nt, nx = int(nt), int(nx)
    L = (nt - 1) * nx
    _mat(z, (L, 6), "z"); _vec(q, L + nt * (nx - 1), "q")
    check(lib().dotsocp_mexBFd1d(ptr(z), ptr(q), nt, nx, float(scale), float(dFactor)))


def mexBFdConj1d(q, z, nt, nx, scale):
    nt, nx = int(nt), int(nx)
    L = (nt - 1) * nx
    _vec(q, L + nt * (nx - 1), "q"); _mat(z, (L, 6), "z")
    check(lib().dotsocp_mexBFdConj1d(ptr(q), ptr(z), nt, nx, float(scale)))


def oper_poisson3dim(rhs, nt, nx, ny, D=1.0):
    """phi = idctn(dctn(rhs) ./ (D^2 * initialize_FFTkernel(nt,nx,ny))); rhs in MATLAB linear order (N doubles)."""
    rhs = np.ascontiguousarray(rhs, dtype=np.float64).ravel()
    _vec(rhs, nt * nx * ny, "rhs")
    phi = np.empty_like(rhs)
    check(lib().dotsocp_poisson(ptr(phi), ptr(rhs), int(nt), int(nx), int(ny), float(D)))
    return phi


def oper_poisson(rhs, nt, nx, D=1.0):
    """1-D variant (dot1d/utils/oper_poisson.m:4): a 2-D DCT solve over (nx, nt)."""
    return oper_poisson3dim(rhs, nt, nx, 1, D)


def dctn(a, nt, nx, ny, inverse=False):
    """Orthonormal DCT-II (or its inverse) along every axis of the C-order (nt,nx,ny) view: mirt_dctn / mirt_idctn."""
    out = np.array(a, dtype=np.float64, order="C").ravel()
    _vec(out, nt * nx * ny, "a")
    check(lib().dotsocp_dctn(ptr(out), int(nt), int(nx), int(ny), 1 if inverse else 0))
    return out
