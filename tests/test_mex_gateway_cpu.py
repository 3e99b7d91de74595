"""The MEX gateway must compile against the (stub) MEX C API and only use the C-ABI functions include/dotsocp.h declares."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_mex_gateway_compiles_against_stub_mex_h():
    src = os.path.join(ROOT, "matlab", "mex", "mexDotSocpGPU.cpp")
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-I" + os.path.join(ROOT, "matlab", "mex", "stub"),
                        "-I" + os.path.join(ROOT, "include"), src], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    used = set(re.findall(r"\b(dotsocp_[a-zA-Z0-9_]+)\s*\(", open(src).read()))
    hdr = open(os.path.join(ROOT, "include", "dotsocp.h")).read()
    for name in used:
        assert re.search(r"\b%s\s*\(" % name, hdr), name


def test_matlab_wrappers_keep_reference_names():
    want = {"dot2d": ["solver_socp_inPALM.m", "solver_socp_PALM.m", "solver_socp_accADMM.m", "solver_socp_sGSinPALM.m", "solver_socp_accsGSADMM.m"],
            "wdot2d": ["solver_wsocp_inPALM.m", "solver_wsocp_accADMM.m"], "dot1d": ["solver_socp_inPALM.m"]}
    for v, files in want.items():
        for f in files:
            p = os.path.join(ROOT, "matlab", "socp", v, "algorithms", f)
            txt = open(p).read()
            assert txt.startswith("function [runHist, sigma] = %s(var, opts, model)" % f[:-2]), p
            assert "dotsocp_gpu_level" in txt
