"""Parity at a BASELINE grid (sorted last on purpose: the largest case of the suite).  The CPU oracle needs five minutes for
this solve, so its history is a committed golden (tests/golden/solver_c3.json, made by make_golden_c3.py)."""
import json
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_three_level_solve_at_256x256x128_matches_the_oracle_golden(gpu):
    """example1 at 256x256x128 cells, 3 levels, inPALM, tol 1e-4 (BASELINE config 2 grid): same iterations per level and
    check schedule, KKT history within 1e-8, objective within 1e-6 relative."""
    import dotsocp_b200 as dp
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_c3", os.path.join(ROOT, "tests", "golden", "make_golden_c3.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    with open(os.path.join(ROOT, "tests", "golden", "solver_c3.json")) as f:
        gold = json.load(f)
    nt, nx, ny = 129, 257, 257
    r0, r1 = mg.densities(nx, ny)
    out, _, ML, rh = dp.solver_dotsocp2d(r0, r1, nt, 3, {"tol": 1e-4, "maxit": 3000}, "inPALM")
    assert [int(v) for v in out.level_iters] == gold["level_iters"]
    assert [int(v) for v in ML.iter] == gold["hist_iter"]
    assert np.abs(ML.kkt - np.array(gold["kkt"])).max() < 1e-8
    assert abs(rh.priVal[-1] - gold["priVal"]) <= 1e-6 * abs(gold["priVal"])
