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


def _c4_module():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_c4", os.path.join(ROOT, "tests", "golden", "make_golden_c4.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def check_c4_golden(hb, res, state, gold, tol_kkt=1e-8):
    """shared by this test and tests/dist_parity_big.py (the same run on 2/4/8 GPUs)"""
    m = _c4_module()
    n = res.hist_len
    assert res.iters == gold["iters"] and [int(v) for v in hb.iter[:n]] == gold["hist_iter"]
    assert np.abs(hb.kkt[:n] - np.array(gold["kkt"])).max() < tol_kkt, np.abs(hb.kkt[:n] - np.array(gold["kkt"])).max()
    assert np.abs(hb.pdGap[:n] - np.array(gold["pdGap"])).max() < 1e-8
    for name in ("priVal", "dualVal"):
        a, b = getattr(hb, name)[:n], np.array(gold[name])
        assert np.abs(a - b).max() <= 1e-6 * np.abs(b).max(), name
    assert abs(res.sigma - gold["sigma"]) <= 1e-12 * gold["sigma"]
    assert [res.cScale, res.dScale, res.D, res.E] == pytest.approx(gold["scal"], rel=1e-12)     # after the in-loop rescalings
    for a, name in zip(state, ("phi", "q", "z", "alpha", "beta")):
        fp = gold["final"][name]
        v = a.reshape(-1, order="F" if a.ndim == 2 else "C")
        nrm = float(np.sqrt(np.dot(v, v)))
        assert abs(nrm - fp["norm2"]) <= 1e-9 * fp["norm2"], (name, nrm, fp["norm2"])
        got = np.array([v[i] for i in m.sample_index(v.size)])
        scale = max(1.0, float(np.abs(np.array(fp["samples"])).max()))
        assert np.abs(got - np.array(fp["samples"])).max() <= 1e-9 * scale, name


@pytest.mark.gpu
def test_baseline_config3_mixture_512x512x256_checked_iterations(gpu):
    """BASELINE.json configs[3]: example2 (Gaussian -> mixture) at 512x512x256 cells, one level from the reference's initial
    state, 12 inPALM iterations with ifCheckStepByStep: every KKT row, the objective values and fingerprints of the final
    iterates against the CPU oracle's golden (tests/golden/solver_c4.json, 14 minutes and 50 GB on the CPU)."""
    import dotsocp_b200 as dp
    from dotsocp_b200 import driver, solver
    with open(os.path.join(ROOT, "tests", "golden", "solver_c4.json")) as f:
        gold = json.load(f)
    nt, nx, ny = gold["grid_nodes"]
    m = _c4_module()
    rho0, rho1 = m.problem(nx, ny)
    var, model = driver.initialize(rho0, rho1, nt)
    driver.InitialScaling(var, model, True, None, "dot2d")
    opts = {"tol": 1e-4, "maxit": gold["iters"], "tau": 1.9, "sigma": 1.0, "ifCheckStepByStep": True, "scaling": True}
    o = solver.make_level_opts("dot2d", "inPALM", var, opts, model)
    with dp.Session("dot2d", nt, nx, ny) as s:
        s.upload(var.phi, var.q, None, var.alpha, var.beta, model.c)
        del var
        hb, res = s.run(o)
        state = s.download()
    check_c4_golden(hb, res, state, gold)
