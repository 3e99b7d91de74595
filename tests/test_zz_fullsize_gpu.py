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


def _load_cfg():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_baseline_configs",
                                                  os.path.join(ROOT, "tests", "golden", "make_golden_baseline_configs.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    with open(os.path.join(ROOT, "tests", "golden", "solver_baseline_configs.json")) as f:
        return m, json.load(f)


def _check_against_golden(out, ML, rh, gold, dim):
    """north-star tolerances: iterations to tolerance within 2 %, objective and W2 cost within 1e-6 relative, residual
    history within 1e-8 (compared when the check schedules coincide)"""
    from oracle import dotsocp_oracle as O
    for got, ref in zip(out.level_iters, gold["level_iters"]):
        assert abs(int(got) - ref) <= max(1, round(0.02 * ref)), (list(out.level_iters), gold["level_iters"])
    assert abs(rh.priVal[-1] - gold["priVal"]) <= 1e-6 * abs(gold["priVal"])
    assert abs(O.w2_cost(out, dim) - gold["w2"]) <= 1e-6 * abs(gold["w2"])
    if [int(v) for v in ML.iter] == gold["hist_iter"]:
        assert np.abs(ML.kkt - np.array(gold["kkt"])).max() < 1e-8


@pytest.mark.gpu
def test_baseline_config0_dot1d_demo_defaults(gpu):
    """BASELINE.json configs[0]: demo_dot1d.m defaults (nt = 33, nx = 1025, 3 levels, tol 1e-5, Gaussian instance)."""
    import dotsocp_b200 as dp
    m, gold = _load_cfg()
    rho0, rho1, nt, levelN, opts = m.config_dot1d()
    out, _, ML, rh = dp.solver_dotsocp1d(rho0, rho1, nt, levelN, opts, "inPALM")
    _check_against_golden(out, ML, rh, gold["dot1d_demo_default"], 1)


@pytest.mark.gpu
def test_baseline_config2_weighted_256x256x128(gpu):
    """BASELINE.json configs[2]: weighted 2-D DOT, circle weight, 256x256x128 cells, 3 levels, tol 1e-3."""
    import dotsocp_b200 as dp
    m, gold = _load_cfg()
    rho0, rho1, nt, levelN, opts = m.config_wdot2d()
    out, _, ML, rh = dp.solver_wdotsocp2d(rho0, rho1, nt, levelN, opts, "inPALM")
    _check_against_golden(out, ML, rh, gold["wdot2d_circle_256x256x128"], 2)
