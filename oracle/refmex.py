"""TEST INFRASTRUCTURE ONLY (oracle/): loader for the reference's genuine MEX binaries.

The reference ships its native kernels only as pre-built MEX files
(``socp/<variant>/utils/mex*.mexa64``, no source).  ``oracle/Makefile`` copies them verbatim
into ``oracle/_ref/`` and builds a tiny ``libmx.so``/``libmex.so`` stand-in (``oracle/mxshim.c``);
this module ``dlopen``s them and calls their own ``mexFunction(nlhs, plhs, nrhs, prhs)``.

Call convention mirrored exactly:
  mexBFd(z2, q, nt, nx, ny, scaleBF, scaleD)       in place into z2   (solver_socp_inPALM.m:133)
  mexBFdConj(q2, z, nt, nx, ny, scaleBF)           in place into q2   (solver_socp_inPALM.m:205)
  mexProjSoc(out, in)                              in place into out  (solver_socp_inPALM.m:199)
  mexBFd1d(z, q, nt, nx, scale, dFactor)           dot1d/algorithms/solver_socp_inPALM.m:132
  mexBFdConj1d(q, z, nt, nx, scale)                dot1d/algorithms/solver_socp_inPALM.m:204
Arrays are float64, column-major (Fortran order) exactly as MATLAB hands them over.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")


class _MX(ctypes.Structure):
    _fields_ = [("pr", ctypes.c_void_p), ("m", ctypes.c_size_t), ("n", ctypes.c_size_t)]


_libs: dict[str, ctypes.CDLL] = {}
_shim_loaded = False


def available() -> bool:
    need = ["libmx.so", "libmex.so", "dot2d/mexBFd.mexa64", "dot2d/mexBFdConj.mexa64",
            "dot2d/mexProjSoc.mexa64", "dot1d/mexBFd1d.mexa64", "dot1d/mexBFdConj1d.mexa64"]
    return all(os.path.exists(os.path.join(REF_DIR, f)) for f in need)


def _load(rel: str) -> ctypes.CDLL:
    global _shim_loaded
    if not _shim_loaded:
        ctypes.CDLL(os.path.join(REF_DIR, "libmx.so"), mode=ctypes.RTLD_GLOBAL)
        ctypes.CDLL(os.path.join(REF_DIR, "libmex.so"), mode=ctypes.RTLD_GLOBAL)
        _shim_loaded = True
    if rel not in _libs:
        lib = ctypes.CDLL(os.path.join(REF_DIR, rel))
        lib.mexFunction.restype = None
        lib.mexFunction.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        _libs[rel] = lib
    return _libs[rel]


def _as_mx(a, keep):
    """Wrap a numpy array / python scalar as a fake mxArray (see mxshim.c)."""
    if np.isscalar(a):
        arr = np.array([[float(a)]], dtype=np.float64, order="F")
        keep.append(arr)
        return _MX(arr.ctypes.data, 1, 1)
    assert isinstance(a, np.ndarray) and a.dtype == np.float64
    assert a.flags.f_contiguous or a.ndim == 1, "MATLAB arrays are column-major"
    if a.ndim == 1:
        m, n = a.shape[0], 1
    else:
        m, n = a.shape[0], int(np.prod(a.shape[1:]))
    return _MX(a.ctypes.data, m, n)


def _call(rel: str, *args) -> None:
    lib = _load(rel)
    keep: list = []
    mxs = [_as_mx(a, keep) for a in args]
    arr = (ctypes.POINTER(_MX) * len(mxs))(*[ctypes.pointer(m) for m in mxs])
    lib.mexFunction(0, None, len(mxs), ctypes.cast(arr, ctypes.c_void_p))


def mexBFd(z2, q, nt, nx, ny, scaleBF, scaleD):
    _call("dot2d/mexBFd.mexa64", z2, q, nt, nx, ny, scaleBF, scaleD)


def mexBFdConj(q2, z, nt, nx, ny, scaleBF):
    _call("dot2d/mexBFdConj.mexa64", q2, z, nt, nx, ny, scaleBF)


def mexProjSoc(out, inp):
    assert out.shape == inp.shape
    _call("dot2d/mexProjSoc.mexa64", out, inp)


def mexBFd1d(z, q, nt, nx, scale, dfactor):
    _call("dot1d/mexBFd1d.mexa64", z, q, nt, nx, scale, dfactor)


def mexBFdConj1d(q, z, nt, nx, scale):
    _call("dot1d/mexBFdConj1d.mexa64", q, z, nt, nx, scale)


def mexsGS(phi, rhs, ep, scale, nt, nx, ny, its):
    """mexsGS(phi, rhs, ep, scale, nt, nx, ny, its) -- red-black symmetric Gauss-Seidel."""
    _call("dot2d/mexsGS.mexa64", phi, rhs, ep, scale, nt, nx, ny, its)
