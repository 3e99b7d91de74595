"""dotsocp_b200 -- B200 (sm_100a) implementation of DOTSOCP's ADMM / inPALM iteration hot path.

Layout
  csrc/        hand-written CUDA kernels + the C ABI (include/dotsocp.h) -> libdotsocp.so (built in-tree)
  _lib.py      ctypes binding of the C ABI (fails loudly when the library or a GPU is missing; no CPU fallback)
  ops.py       kernel-level entry points with the reference's MEX names (mexBFd, mexBFdConj, mexProjSoc, ...)
  solver.py    solver-level mirror: solver_socp_inPALM(var, opts, model), ... and the device-resident Session
  driver.py    host-side mirror of the multilevel drivers solver_dotsocp2d / solver_wdotsocp2d / solver_dotsocp1d
"""
from . import _lib, driver, ops, slab, solver  # noqa: F401
from ._lib import DotsocpError  # noqa: F401
from .driver import solver_dotsocp1d, solver_dotsocp2d, solver_wdotsocp2d  # noqa: F401
from .solver import (Session, Weights, solver_socp_accADMM, solver_socp_inPALM, solver_socp_PALM,  # noqa: F401
                     solver_socp_accsGSADMM, solver_socp_sGSinPALM, solver_wsocp_accADMM, solver_wsocp_inPALM)

__all__ = ["DotsocpError", "ops", "solver", "driver", "Session", "Weights", "solver_socp_inPALM", "solver_socp_PALM", "solver_socp_accADMM", "solver_socp_sGSinPALM", "solver_socp_accsGSADMM",
           "solver_wsocp_inPALM", "solver_wsocp_accADMM", "solver_dotsocp2d", "solver_wdotsocp2d", "solver_dotsocp1d"]
