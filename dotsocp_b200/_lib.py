"""ctypes binding of libdotsocp.so (the C ABI declared in include/dotsocp.h).

The product path is the CUDA library and nothing else: if the shared object is missing, or no CUDA device is
usable, every compute call raises -- there is no CPU fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DOTSOCP_LIB") or os.path.join(_HERE, "libdotsocp.so")   # override only for A/B experiments

VARIANT = {"dot2d": 0, "wdot2d": 1, "dot1d": 2}
METHOD = {"inPALM": 0, "ALG2": 0, "PALM": 1, "acc-ADMM": 2, "sGS-inPALM": 3, "acc-sGS-ADMM": 4}
NTIMES = 8


class DotsocpError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libdotsocp error {code}: {msg}")
        self.code = code


class LevelOpts(C.Structure):
    _fields_ = [
        ("variant", C.c_int32), ("method", C.c_int32),
        ("nt", C.c_int32), ("nx", C.c_int32), ("ny", C.c_int32),
        ("maxit", C.c_int32), ("ifCheckStepByStep", C.c_int32), ("scaling", C.c_int32),
        ("checkPrimDualFeas", C.c_int32), ("restart", C.c_int32),
        ("tau", C.c_double), ("sigma", C.c_double), ("tol", C.c_double), ("time_limit", C.c_double),
        ("rho", C.c_double), ("theta", C.c_double),
        ("cScale", C.c_double), ("dScale", C.c_double), ("D", C.c_double), ("E", C.c_double),
        ("normc", C.c_double), ("normd", C.c_double),
        ("grad_t", C.c_double), ("grad_x", C.c_double), ("grad_y", C.c_double),
    ]


class LevelResult(C.Structure):
    _fields_ = [
        ("iters", C.c_int32), ("hist_len", C.c_int32), ("sigma", C.c_double),
        ("cScale", C.c_double), ("dScale", C.c_double), ("D", C.c_double), ("E", C.c_double),
        ("times", C.c_double * NTIMES), ("gpu_launches", C.c_double),
    ]


class ProlongScal(C.Structure):
    _fields_ = [("phi_recover", C.c_double), ("beta_recover", C.c_double),
                ("grad_t", C.c_double), ("grad_x", C.c_double), ("grad_y", C.c_double),
                ("phi_scale", C.c_double), ("q_scale", C.c_double), ("alpha_scale", C.c_double), ("beta_scale", C.c_double)]


class RecoverScal(C.Structure):
    _fields_ = [("alpha_recover", C.c_double), ("q_recover", C.c_double)]


class Hist(C.Structure):
    _fields_ = [
        ("cap", C.c_int32), ("kkt", C.c_void_p), ("time", C.c_void_p), ("iter", C.c_void_p),
        ("pdGap", C.c_void_p), ("priVal", C.c_void_p), ("dualVal", C.c_void_p),
    ]


EXPORTS = [
    "dotsocp_last_error", "dotsocp_version", "dotsocp_device_count", "dotsocp_set_device",
    "dotsocp_mexBFd", "dotsocp_mexBFdConj", "dotsocp_mexProjSoc", "dotsocp_mexsGS", "dotsocp_mexBFd1d", "dotsocp_mexBFdConj1d",
    "dotsocp_poisson", "dotsocp_dctn", "dotsocp_solve_level", "dotsocp_release_cached",
    "dotsocp_nccl_unique_id", "dotsocp_create", "dotsocp_destroy", "dotsocp_upload", "dotsocp_download",
    "dotsocp_create_refined", "dotsocp_prolong", "dotsocp_recover", "dotsocp_run", "dotsocp_iter_begin", "dotsocp_iterate", "dotsocp_iter_end", "dotsocp_launch_count",
    "dotsocp_weights_create", "dotsocp_weights_destroy", "dotsocp_weights_set", "dotsocp_weights_set_planes", "dotsocp_weights_restrict",
    "dotsocp_weights_get", "dotsocp_weights_dims", "dotsocp_weights_log10_mean", "dotsocp_weights_launch_count", "dotsocp_set_weight",
]

_lib = None


def lib():
    """Load libdotsocp.so (once).  Fails loudly when the CUDA extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DotsocpError(-2, f"{LIB_PATH} is missing: build it with `make -C dotsocp_b200/csrc` "
                               "(or __graft_entry__.build()); there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    P, I, D, I64 = C.c_void_p, C.c_int, C.c_double, C.c_int64
    L.dotsocp_last_error.restype = C.c_char_p
    L.dotsocp_launch_count.restype = C.c_double
    L.dotsocp_launch_count.argtypes = [P]
    L.dotsocp_set_device.argtypes = [I]
    L.dotsocp_mexBFd.argtypes = [P, P, I, I, I, D, D]
    L.dotsocp_mexBFdConj.argtypes = [P, P, I, I, I, D]
    L.dotsocp_mexProjSoc.argtypes = [P, P, I64, I]
    L.dotsocp_mexsGS.argtypes = [P, P, D, D, I, I, I, I]
    L.dotsocp_mexBFd1d.argtypes = [P, P, I, I, D, D]
    L.dotsocp_mexBFdConj1d.argtypes = [P, P, I, I, D]
    L.dotsocp_poisson.argtypes = [P, P, I, I, I, D]
    L.dotsocp_dctn.argtypes = [P, I, I, I, I]
    L.dotsocp_solve_level.argtypes = [C.POINTER(LevelOpts), P, P, P, P, P, P, P, C.POINTER(Hist), C.POINTER(LevelResult)]
    L.dotsocp_release_cached.argtypes = []
    L.dotsocp_release_cached.restype = None
    L.dotsocp_nccl_unique_id.argtypes = [P]
    L.dotsocp_create.argtypes = [C.POINTER(P), I, I, I, I, I, I, P]
    L.dotsocp_destroy.argtypes = [P]
    L.dotsocp_destroy.restype = None
    L.dotsocp_upload.argtypes = [P, P, P, P, P, P, P, P]
    L.dotsocp_download.argtypes = [P, P, P, P, P, P]
    L.dotsocp_prolong.argtypes = [P, P, C.POINTER(ProlongScal), P, P, P]
    L.dotsocp_create_refined.argtypes = [C.POINTER(P), P]
    L.dotsocp_recover.argtypes = [P, C.POINTER(RecoverScal), P, P, P, P, P, P, P, P, P, P, P]
    L.dotsocp_run.argtypes = [P, C.POINTER(LevelOpts), C.POINTER(Hist), C.POINTER(LevelResult)]
    L.dotsocp_iter_begin.argtypes = [P, C.POINTER(LevelOpts)]
    L.dotsocp_iterate.argtypes = [P, I, I, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.dotsocp_iter_end.argtypes = [P]
    L.dotsocp_weights_create.argtypes = [C.POINTER(P), I, I, I, I]
    L.dotsocp_weights_destroy.argtypes = [P]
    L.dotsocp_weights_destroy.restype = None
    L.dotsocp_weights_set.argtypes = [P, P]
    L.dotsocp_weights_set_planes.argtypes = [P, P, P]
    L.dotsocp_weights_restrict.argtypes = [P, I]
    L.dotsocp_weights_get.argtypes = [P, I, P]
    L.dotsocp_weights_dims.argtypes = [P, I, C.POINTER(I), C.POINTER(I), C.POINTER(I)]
    L.dotsocp_weights_log10_mean.argtypes = [P, I, C.POINTER(D)]
    L.dotsocp_weights_launch_count.argtypes = [P]
    L.dotsocp_weights_launch_count.restype = D
    L.dotsocp_set_weight.argtypes = [P, P, I]
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise DotsocpError(rc, lib().dotsocp_last_error().decode(errors="replace"))


def ptr(a):
    """Pointer of a float64 array that is contiguous in memory order (vectors, or column-major matrices)."""
    if a is None:
        return None
    assert isinstance(a, np.ndarray) and a.dtype == np.float64, "float64 ndarray expected"
    if a.ndim >= 2 and min(a.shape) > 1:
        assert a.flags.f_contiguous, "matrices cross the C ABI in column-major (MATLAB) order"
    else:
        assert a.flags.c_contiguous or a.flags.f_contiguous, "array must be contiguous"
    return a.ctypes.data


class HistBuffers:
    """Caller-allocated runHist storage (dotsocp_hist)."""

    def __init__(self, cap):
        cap = max(int(cap), 1)
        self.kkt = np.full((cap, 7), np.inf)
        self.time = np.full(cap, np.inf)
        self.iter = np.full(cap, np.inf)
        self.pdGap = np.full(cap, np.inf)
        self.priVal = np.full(cap, np.inf)
        self.dualVal = np.full(cap, np.inf)
        # kkt is ROW-major on purpose (dotsocp_hist: kkt[i*7 + j]): hand over the raw address, not ptr()
        self.c = Hist(cap, self.kkt.ctypes.data, ptr(self.time), ptr(self.iter), ptr(self.pdGap), ptr(self.priVal),
                      ptr(self.dualVal))
