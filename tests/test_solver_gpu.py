"""GPU parity of the solver-level path (C ABI -> CUDA) against the CPU oracle on the same inputs.

Tolerances are the north star's: objective within 1e-6 relative, KKT residual history within 1e-8 absolute, and the
same iteration count (the +-2 % allowance is not used at these sizes: counts must match exactly)."""
import numpy as np
import pytest

from oracle import dotsocp_oracle as O

pytestmark = pytest.mark.gpu


def _compare(out_g, rh_g, ML_g, out_o, rh_o, ML_o, dim=2):
    assert out_g.level_iters == out_o.level_iters, (out_g.level_iters, out_o.level_iters)
    assert ML_g.len == ML_o.len
    assert np.array_equal(ML_g.iter, ML_o.iter)
    assert np.abs(ML_g.kkt - ML_o.kkt).max() < 1e-8
    assert np.abs(ML_g.pdGap - ML_o.pdGap).max() < 1e-8
    assert abs(rh_g.priVal[-1] - rh_o.priVal[-1]) <= 1e-6 * abs(rh_o.priVal[-1])
    assert abs(rh_g.dualVal[-1] - rh_o.dualVal[-1]) <= 1e-6 * abs(rh_o.dualVal[-1])
    assert abs(out_g.sigma - out_o.sigma) <= 1e-9 * abs(out_o.sigma)
    assert np.abs(out_g.rho - out_o.rho).max() < 1e-6
    assert np.abs(out_g.Ex - out_o.Ex).max() < 1e-6
    from dotsocp_b200 import driver
    wg, wo = driver.w2_cost(out_g, dim), O.w2_cost(out_o, dim)
    assert abs(wg - wo) <= 1e-6 * abs(wo)
    assert out_g.massOK


@pytest.mark.parametrize("problem,n,nt,levelN,method", [
    ("example1", 17, 9, 1, "inPALM"),
    ("example1", 33, 17, 2, "inPALM"),
    ("example2", 33, 17, 2, "ALG2"),
    ("circle", 33, 17, 1, "inPALM"),
    ("example1", 65, 33, 3, "inPALM"),      # BASELINE config 2
])
def test_dot2d_multilevel_parity(gpu, problem, n, nt, levelN, method):
    import dotsocp_b200 as dp
    rho0, rho1 = O.get_example2d(problem, n, n)
    opts = {"tol": 1e-4, "maxit": 3000}
    out_g, _, ML_g, rh_g = dp.solver_dotsocp2d(rho0, rho1, nt, levelN, opts, method)
    out_o, _, ML_o, rh_o = O.solver_dotsocp2d(rho0, rho1, nt, levelN, opts, method, workers=4)
    _compare(out_g, rh_g, ML_g, out_o, rh_o, ML_o)
    assert out_g.gpu_launches > 0


def test_dot2d_rectangular_grid(gpu):
    """nx != ny, single level (the reference's 2-D down-sampling assumes square grids)."""
    import dotsocp_b200 as dp
    rho0, rho1 = O.get_example2d("example2", 41, 29)      # MATLAB (ny, nx) = (29, 41)
    opts = {"tol": 1e-4, "maxit": 600}
    out_g, _, ML_g, rh_g = dp.solver_dotsocp2d(rho0, rho1, 13, 1, opts, "inPALM")
    out_o, _, ML_o, rh_o = O.solver_dotsocp2d(rho0, rho1, 13, 1, opts, "inPALM")
    _compare(out_g, rh_g, ML_g, out_o, rh_o, ML_o)


def test_wdot2d_weighted_parity(gpu):
    import dotsocp_b200 as dp
    n, nt = 33, 17
    rho0, rho1 = O.get_example2d("example1", n, n)
    w = O.gene_weight_circle(nt, n, n)
    opts = {"tol": 1e-3, "maxit": 10000, "weight": w}
    out_g, _, ML_g, rh_g = dp.solver_wdotsocp2d(rho0, rho1, nt, 2, opts, "inPALM")
    out_o, _, ML_o, rh_o = O.solver_wdotsocp2d(rho0, rho1, nt, 2, opts, "inPALM")
    _compare(out_g, rh_g, ML_g, out_o, rh_o, ML_o)


def test_wdot2d_barrier_parity(gpu):
    import dotsocp_b200 as dp
    n, nt = 33, 17
    barrier = O.gene_barrier_of_love_heart()
    rho0, rho1 = O.gene_exampleLoveHeart(n, n)
    rho0, rho1 = O._normalize2d(rho0, rho1, n, n)
    w = O.get_weight_by_barrier(n, n, nt, barrier)
    rho0, rho1, _ = O.ensure_barrier_validity(rho0, rho1, barrier)
    opts = {"tol": 1e-3, "maxit": 3000, "weight": w}
    out_g, _, ML_g, rh_g = dp.solver_wdotsocp2d(rho0, rho1, nt, 2, opts, "inPALM", barrier)
    out_o, _, ML_o, rh_o = O.solver_wdotsocp2d(rho0, rho1, nt, 2, opts, "inPALM", barrier)
    _compare(out_g, rh_g, ML_g, out_o, rh_o, ML_o)


@pytest.mark.parametrize("nx,nt,levelN,tol", [(129, 9, 1, 1e-4), (257, 17, 2, 1e-5)])
def test_dot1d_parity(gpu, nx, nt, levelN, tol):
    import dotsocp_b200 as dp
    rho0, rho1 = O.get_example1d("gaussian", nx)
    opts = {"tol": tol, "maxit": 3000}
    out_g, _, ML_g, rh_g = dp.solver_dotsocp1d(rho0, rho1, nt, levelN, opts, "inPALM")
    out_o, _, ML_o, rh_o = O.solver_dotsocp1d(rho0, rho1, nt, levelN, opts, "inPALM")
    _compare(out_g, rh_g, ML_g, out_o, rh_o, ML_o, dim=1)


def test_level_solver_mirrors_reference_side_effects(gpu):
    """solver_socp_inPALM(var, opts, model): var mutated in place, alpha/beta returned times sigma, time table filled."""
    import dotsocp_b200 as dp
    from dotsocp_b200 import driver
    rho0, rho1 = O.get_example2d("example1", 17, 17)
    var, model = driver.initialize(rho0, rho1, 9)
    driver.InitialScaling(var, model, True, None, "dot2d")
    vo, mo = O.initialize2d(rho0, rho1, 9)
    O.InitialScaling(vo, mo, True, None, "dot2d")
    opts = {"tol": 1e-4, "maxit": 25, "tau": 1.9, "sigma": 1.0, "ifCheckStepByStep": True, "scaling": True}
    rh_g, sg = dp.solver_socp_inPALM(var, opts, model)
    rh_o, so = O.solver_socp_inPALM(vo, opts, mo)
    assert rh_g.len == rh_o.len == 25            # step-by-step checks: one history row per iteration
    assert np.abs(rh_g.kkt - rh_o.kkt).max() < 1e-9
    assert abs(sg - so) <= 1e-12 * abs(so)
    for name in ("phi", "q", "alpha"):
        assert np.abs(getattr(var, name) - getattr(vo, name)).max() < 1e-9, name
    assert np.abs(var.z - vo.z).max() < 1e-9 and np.abs(var.beta - vo.beta).max() < 1e-9
    assert set(["Step_1_1_FFT", "Step_1_2_ProjSOC", "Step_2_Q_Step", "Step_3_Multiplier", "KKT", "Total_Time", "Iters"]) <= set(var.time)
    assert (var.cScale, var.dScale) == pytest.approx((vo.cScale, vo.dScale), rel=1e-12)


def test_error_behaviour(gpu):
    import dotsocp_b200 as dp
    from dotsocp_b200 import _lib
    rho0, rho1 = O.get_example2d("example1", 9, 9)
    with pytest.raises(ValueError):
        dp.solver_dotsocp2d(rho0, rho1, 5, 1, {"tol": 1e-3}, "nonsense")
    with pytest.raises(ValueError):
        dp.solver_dotsocp2d(rho0, rho1, 5, 0, {"tol": 1e-3}, "inPALM")
    with pytest.raises(_lib.DotsocpError):
        dp.Session("dot1d", 5, 9, 3)             # 1-D variant with ny != 1
    with pytest.raises(_lib.DotsocpError):
        dp.Session("dot2d", 1, 9, 9)


@pytest.mark.parametrize("method,n,nt,levelN", [("PALM", 17, 9, 1), ("acc-ADMM", 17, 9, 1), ("acc-ADMM", 33, 17, 2),
                                                 ("PALM", 33, 17, 2)])
def test_dot2d_palm_and_accadmm_parity(gpu, method, n, nt, levelN):
    import dotsocp_b200 as dp
    rho0, rho1 = O.get_example2d("example1", n, n)
    opts = {"tol": 1e-4, "maxit": 3000}
    out_g, _, ML_g, rh_g = dp.solver_dotsocp2d(rho0, rho1, nt, levelN, opts, method)
    out_o, _, ML_o, rh_o = O.solver_dotsocp2d(rho0, rho1, nt, levelN, opts, method, workers=4)
    _compare(out_g, rh_g, ML_g, out_o, rh_o, ML_o)


def test_wdot2d_accadmm_parity(gpu):
    import dotsocp_b200 as dp
    n, nt = 17, 9
    rho0, rho1 = O.get_example2d("example1", n, n)
    w = O.gene_weight_circle(nt, n, n)
    opts = {"tol": 1e-3, "maxit": 10000, "weight": w}
    out_g, _, ML_g, rh_g = dp.solver_wdotsocp2d(rho0, rho1, nt, 2, opts, "acc-ADMM")
    out_o, _, ML_o, rh_o = O.solver_wdotsocp2d(rho0, rho1, nt, 2, opts, "acc-ADMM")
    _compare(out_g, rh_g, ML_g, out_o, rh_o, ML_o)


def test_golden_solver_histories(gpu):
    """committed goldens (tests/golden/solver.json): iteration counts exact, KKT history within 1e-8, objective 1e-6 rel"""
    import json, os
    import dotsocp_b200 as dp
    from dotsocp_b200 import driver
    with open(os.path.join(os.path.dirname(__file__), "golden", "solver.json")) as f:
        gold = {c["name"]: c for c in json.load(f)}
    rho0, rho1 = O.get_example2d("example1", 17, 17)
    runs = {
        "dot2d_example1_17x17x9_L2_inPALM": lambda: dp.solver_dotsocp2d(rho0, rho1, 9, 2, {"tol": 1e-4, "maxit": 3000}, "inPALM"),
        "dot2d_example1_17x17x9_L1_accADMM": lambda: dp.solver_dotsocp2d(rho0, rho1, 9, 1, {"tol": 1e-4, "maxit": 3000}, "acc-ADMM"),
        "dot2d_example1_17x17x9_L1_PALM": lambda: dp.solver_dotsocp2d(rho0, rho1, 9, 1, {"tol": 1e-4, "maxit": 3000}, "PALM"),
        "wdot2d_example1_circle_17x17x9_L2_inPALM": lambda: dp.solver_wdotsocp2d(
            rho0, rho1, 9, 2, {"tol": 1e-3, "maxit": 10000, "weight": O.gene_weight_circle(9, 17, 17)}, "inPALM"),
        "dot1d_gaussian_129x9_L2_inPALM": lambda: dp.solver_dotsocp1d(*O.get_example1d("gaussian", 129), 9, 2,
                                                                       {"tol": 1e-5, "maxit": 3000}, "inPALM"),
        "dot2d_example2_33x33x17_L2_ALG2": lambda: dp.solver_dotsocp2d(*O.get_example2d("example2", 33, 33), 17, 2,
                                                                        {"tol": 1e-4, "maxit": 3000}, "ALG2"),
    }
    for name, fn in runs.items():
        out, _, ML, rh = fn()
        g = gold[name]
        assert [int(v) for v in out.level_iters] == g["level_iters"], name
        assert ML.iter.tolist() == g["hist_iter"], name
        assert np.abs(ML.kkt - np.array(g["kkt"])).max() < 1e-8, name
        assert abs(rh.priVal[-1] - g["priVal"]) <= 1e-6 * abs(g["priVal"]), name
        assert abs(driver.w2_cost(out, 1 if name.startswith("dot1d") else 2) - g["w2"]) <= 1e-6 * abs(g["w2"]), name


@pytest.mark.parametrize("tsolve", ["pipelined", "transpose"])
@pytest.mark.parametrize("world", [2, 3, 4])
@pytest.mark.parametrize("variant", ["dot2d", "wdot2d"])
def test_time_slab_partition_emulated_on_one_gpu(gpu, world, variant, tsolve, monkeypatch):
    """The multi-GPU code path (time slabs, ghost planes, transposed t-pass, reduced KKT sums) with all slabs on ONE
    device (device-to-device copies instead of NCCL): must reproduce the single-slab solve."""
    import dotsocp_b200 as dp
    from dotsocp_b200 import driver, solver
    # t-direction of the Poisson solve across slabs: pipelined Thomas sweeps with one carry plane per boundary (default), with
    # the modes cut into 3 chunks here, or the two all-to-all transposes (DOTSOCP_TSOLVE=transpose); both read when the session
    # is created
    if tsolve == "transpose":
        monkeypatch.setenv("DOTSOCP_TSOLVE", "transpose")
    else:
        monkeypatch.setenv("DOTSOCP_TCHUNKS", "3")
    n, nt = 33, 17
    rho0, rho1 = O.get_example2d("example2", n, n)
    weight = O.gene_weight_circle(nt, n, n) if variant == "wdot2d" else None
    results = []
    for w in (1, world):
        var, model = driver.initialize(rho0, rho1, nt)
        if weight is not None:
            model.weight = weight
        driver.InitialScaling(var, model, True, None, variant)
        opts = {"tol": 1e-4 if variant == "dot2d" else 1e-3, "maxit": 400, "tau": 1.9, "sigma": 1.0,
                "ifCheckStepByStep": False, "scaling": True}
        o = solver.make_level_opts(variant, "inPALM", var, opts, model)
        with dp.Session(variant, nt, n, n, world=w) as s:
            s.upload(var.phi, var.q, var.z, var.alpha, var.beta, model.c, weight)
            hb, res = s.run(o)
            state = s.download()
        results.append((hb, res, state))
    (hb1, r1, s1), (hbw, rw, sw) = results
    # every sum is reduced per time level and then over the levels in a fixed order, and every kernel's per-element
    # arithmetic is independent of the partition: the slab count must not change a single bit
    assert r1.iters == rw.iters and r1.hist_len == rw.hist_len
    assert np.array_equal(hb1.kkt[:r1.hist_len], hbw.kkt[:rw.hist_len])
    assert np.array_equal(hb1.priVal[:r1.hist_len], hbw.priVal[:rw.hist_len])
    assert r1.sigma == rw.sigma
    for a, b, name in zip(s1, sw, ("phi", "q", "z", "alpha", "beta")):
        assert np.array_equal(a, b), name


@pytest.mark.parametrize("xchg", ["copy", "direct"])
@pytest.mark.parametrize("world", [2, 4])
def test_time_slabs_with_fused_transpose_pack(gpu, world, xchg, monkeypatch):
    """nx = 129 uses the register-FFT DCT kernel, whose x passes write/read the packed all-to-all buffer directly.
    xchg = direct: the x pass and the t-solve store straight into the destination slab's buffers (DOTSOCP_XCHG, read when
    the session is created)."""
    import dotsocp_b200 as dp
    monkeypatch.setenv("DOTSOCP_XCHG", xchg)
    monkeypatch.setenv("DOTSOCP_TSOLVE", "transpose")      # the default t-solve of the slabs is the pipelined Thomas sweep
    from dotsocp_b200 import driver, solver
    nt, nx, ny = (33 if world == 2 else 9), 129, 12    # 16 levels per slab: all 8 pipeline groups of the transposes are used
    rng = np.random.default_rng(5)
    rho0 = np.abs(rng.standard_normal((ny, nx))) + 0.1
    rho1 = np.abs(rng.standard_normal((ny, nx))) + 0.1
    rho0 *= rho0.size / rho0.sum()
    rho1 *= rho1.size / rho1.sum()
    out = []
    for w in (1, world):
        var, model = driver.initialize(rho0, rho1, nt)
        driver.InitialScaling(var, model, True, None, "dot2d")
        opts = {"tol": 1e-12, "maxit": 60, "tau": 1.9, "sigma": 1.0, "ifCheckStepByStep": False, "scaling": True}
        o = solver.make_level_opts("dot2d", "inPALM", var, opts, model)
        with dp.Session("dot2d", nt, nx, ny, world=w) as s:
            s.upload(var.phi, var.q, var.z, var.alpha, var.beta, model.c)
            hb, res = s.run(o)
            out.append((hb, res, s.download()))
    (hb1, r1, s1), (hbw, rw, sw) = out
    assert r1.iters == rw.iters == 60 and r1.hist_len == rw.hist_len
    assert np.array_equal(hb1.kkt[:r1.hist_len], hbw.kkt[:rw.hist_len])
    for a, b, name in zip(s1, sw, ("phi", "q", "z", "alpha", "beta")):
        assert np.array_equal(a, b), name


def _level_run(variant, nt, nx, ny, rho0, rho1, weight, opts, world=1):
    import dotsocp_b200 as dp
    from dotsocp_b200 import driver, solver
    var, model = driver.initialize(rho0, rho1, nt)
    if weight is not None:
        model.weight = weight
    driver.InitialScaling(var, model, True, None, variant)
    o = solver.make_level_opts(variant, "inPALM", var, opts, model)
    with dp.Session(variant, nt, nx, ny, world=world) as s:
        s.upload(var.phi, var.q, var.z, var.alpha, var.beta, model.c, weight)
        hb, res = s.run(o)
        return hb, res, s.download()


@pytest.mark.parametrize("variant", ["dot2d", "wdot2d", "dot1d"])
def test_fused_kkt_check_matches_the_standalone_kkt_kernels(gpu, variant, monkeypatch):
    """inPALM check iterations accumulate the KKT sums inside k_qstep / k_mult (no extra pass over the 10-column arrays);
    DOTSOCP_KKT=separate (read when the session is created) selects the stand-alone KKT kernels that PALM / acc-ADMM use.
    Same check schedule and sigma decisions, residuals equal up to the rounding of a different summation order, and the
    iterates -- which the sums only steer -- identical bit for bit.  Also on emulated slabs."""
    if variant == "dot1d":
        nt, nx, ny = 17, 257, 1
        rho0, rho1 = O.get_example1d("gaussian", nx)
        weight, opts = None, {"tol": 1e-5, "maxit": 300}
    else:
        nt, nx, ny = 17, 37, 33
        rho0, rho1 = O.get_example2d("example2", nx, ny)
        weight = O.gene_weight_circle(nt, nx, ny) if variant == "wdot2d" else None
        opts = {"tol": 1e-4 if weight is None else 1e-3, "maxit": 300}
    opts.update(tau=1.9, sigma=1.0, ifCheckStepByStep=False, scaling=True)
    for world in (1, 3):
        monkeypatch.setenv("DOTSOCP_KKT", "separate")
        hb_s, r_s, st_s = _level_run(variant, nt, nx, ny, rho0, rho1, weight, opts, world)
        monkeypatch.delenv("DOTSOCP_KKT")
        hb_f, r_f, st_f = _level_run(variant, nt, nx, ny, rho0, rho1, weight, opts, world)
        assert r_f.iters == r_s.iters and r_f.hist_len == r_s.hist_len > 5
        assert np.array_equal(hb_f.iter[:r_f.hist_len], hb_s.iter[:r_s.hist_len])
        assert np.abs(hb_f.kkt[:r_f.hist_len] - hb_s.kkt[:r_s.hist_len]).max() < 1e-13
        assert np.abs(hb_f.priVal[:r_f.hist_len] - hb_s.priVal[:r_s.hist_len]).max() < 1e-13
        assert np.abs(hb_f.pdGap[:r_f.hist_len] - hb_s.pdGap[:r_s.hist_len]).max() < 1e-13
        assert r_f.sigma == r_s.sigma
        for a, b, name in zip(st_f, st_s, ("phi", "q", "z", "alpha", "beta")):
            assert np.array_equal(a, b), (name, world)


@pytest.mark.parametrize("variant", ["dot2d", "wdot2d"])
def test_register_prefetch_march_is_bit_identical_to_the_plain_march(gpu, variant, monkeypatch):
    """k_mult prefetches the values of the next time step -- into a second register set (software pipeline, unrolled by two) or
    with cp.async into a shared-memory ring; DOTSOCP_KM_PF=0 selects the plain load-then-compute march.  Odd and even numbers
    of levels, slabs and pieces."""
    for nt, world in ((18, 1), (17, 1), (23, 3)):
        nx, ny = 41, 35
        rho0, rho1 = O.get_example2d("example2", nx, ny)
        weight = O.gene_weight_circle(nt, nx, ny) if variant == "wdot2d" else None
        opts = {"tol": 1e-12, "maxit": 30, "tau": 1.9, "sigma": 1.0, "ifCheckStepByStep": False, "scaling": True}
        monkeypatch.setenv("DOTSOCP_KM_PF", "0")
        hb1, r1, st1 = _level_run(variant, nt, nx, ny, rho0, rho1, weight, opts, world)
        for mode in ("1", "2"):       # 1: second register set, 2: cp.async ring in shared memory
            monkeypatch.setenv("DOTSOCP_KM_PF", mode)
            hb2, r2, st2 = _level_run(variant, nt, nx, ny, rho0, rho1, weight, opts, world)
            assert r1.iters == r2.iters == 30
            assert np.array_equal(hb1.kkt[:r1.hist_len], hb2.kkt[:r2.hist_len])
            for a, b, name in zip(st1, st2, ("phi", "q", "z", "alpha", "beta")):
                assert np.array_equal(a, b), (name, nt, world, mode)
        monkeypatch.delenv("DOTSOCP_KM_PF")


CHUNK_SCRIPT = """
import sys, hashlib
import numpy as np
sys.path.insert(0, %r)
from oracle import dotsocp_oracle as O
import dotsocp_b200 as dp
from dotsocp_b200 import driver, solver
n, nt = 65, 33
rho0, rho1 = O.get_example2d("example2", n, n)
var, model = driver.initialize(rho0, rho1, nt)
driver.InitialScaling(var, model, True, None, "dot2d")
opts = {"tol": 1e-12, "maxit": 40, "tau": 1.9, "sigma": 1.0, "ifCheckStepByStep": False, "scaling": True}
o = solver.make_level_opts("dot2d", "inPALM", var, opts, model)
h = hashlib.sha256()
for world in (1, 2):
    with dp.Session("dot2d", nt, n, n, world=world) as s:
        s.upload(var.phi, var.q, var.z, var.alpha, var.beta, model.c)
        hb, res = s.run(o)
        for a in s.download():
            h.update(np.ascontiguousarray(a).tobytes())
        h.update(hb.kkt[:res.hist_len].tobytes())
print("HASH", h.hexdigest())
"""


def test_time_chunked_mult_kernel_is_bit_identical(gpu, tmp_path):
    """k_mult cuts the time range into pieces (grid.z) to fill the last round of CTAs; every piece replays the cell layer
    below it like a slab does, so any number of pieces must give bit-identical iterates (single slab and emulated slabs)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "chunks.py"
    script.write_text(CHUNK_SCRIPT % root)
    hashes = {}
    for n in ("1", "3", "8"):
        e = dict(os.environ, DOTSOCP_KM_CHUNKS=n)
        r = subprocess.run([sys.executable, str(script)], env=e, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        hashes[n] = [l for l in r.stdout.splitlines() if l.startswith("HASH")][0]
    assert hashes["1"] == hashes["3"] == hashes["8"], hashes


@pytest.mark.parametrize("variant,theta,rho,restart", [("dot2d", 3.0, 1.6, 25), ("dot2d", 5.0, 2.0, 7), ("wdot2d", 3.0, 1.8, 40)])
def test_accadmm_general_extrapolation_parity(gpu, variant, theta, rho, restart):
    """opts.theta != 2 leaves the Halpern iteration for the three-term extrapolation of solver_socp_accADMM.m:389-417
    (reachable only through the level solver: the drivers never set theta)."""
    import dotsocp_b200 as dp
    from dotsocp_b200 import driver
    n, nt = 17, 9
    rho0, rho1 = O.get_example2d("example1", n, n)
    weight = O.gene_weight_circle(nt, n, n) if variant == "wdot2d" else None
    var, model = driver.initialize(rho0, rho1, nt)
    vo, mo = O.initialize2d(rho0, rho1, nt)
    if weight is not None:
        model.weight = weight
        mo.weight = weight
    driver.InitialScaling(var, model, True, None, variant)
    O.InitialScaling(vo, mo, True, None, variant)
    opts = {"tol": 1e-4 if weight is None else 1e-3, "maxit": 400, "sigma": 1.0, "ifCheckStepByStep": False, "scaling": True,
            "theta": theta, "rho": rho, "restart": restart}
    fn_g = dp.solver_socp_accADMM if weight is None else dp.solver_wsocp_accADMM
    rh_g, sg = fn_g(var, opts, model)
    rh_o, so = O.solver_socp_accADMM(vo, opts, mo)
    assert rh_g.len == rh_o.len and int(var.time["Iters"]) == int(vo.time["Iters"])
    assert np.abs(rh_g.kkt - rh_o.kkt).max() < 1e-8
    assert abs(sg - so) <= 1e-12 * abs(so)
    for name in ("phi", "q", "alpha", "z", "beta"):
        a, b = np.asarray(getattr(var, name)), np.asarray(getattr(vo, name))
        assert np.abs(a - b).max() <= 1e-8 * max(1.0, np.abs(b).max()), name


def test_inpalm_never_reads_the_incoming_z(gpu):
    """solver_socp_inPALM.m:199 overwrites z before any read, so the upload may omit it (dotsocp_solve_level does): same
    iterates bit for bit with a garbage z, and the calls that do need z refuse to run without it."""
    import dotsocp_b200 as dp
    from dotsocp_b200 import _lib, driver, solver
    n, nt = 33, 17
    rho0, rho1 = O.get_example2d("example2", n, n)
    var, model = driver.initialize(rho0, rho1, nt)
    driver.InitialScaling(var, model, True, None, "dot2d")
    opts = {"tol": 1e-4, "maxit": 60, "tau": 1.9, "sigma": 1.0, "ifCheckStepByStep": False, "scaling": True}
    o = solver.make_level_opts("dot2d", "inPALM", var, opts, model)
    states = []
    for z in (var.z, None, np.full_like(var.z, 1e300)):
        with dp.Session("dot2d", nt, n, n) as s:
            s.upload(var.phi, var.q, z, var.alpha, var.beta, model.c)
            hb, res = s.run(o)
            states.append((hb.kkt[:res.hist_len].copy(), s.download()))
    for kkt, st in states[1:]:
        assert np.array_equal(kkt, states[0][0])
        for a, b in zip(st, states[0][1]):
            assert np.array_equal(a, b)
    with dp.Session("dot2d", nt, n, n) as s:
        s.upload(var.phi, var.q, None, var.alpha, var.beta, model.c)
        with pytest.raises(_lib.DotsocpError):
            s.download()                                  # z neither uploaded nor computed
        with pytest.raises(_lib.DotsocpError):
            s.run(solver.make_level_opts("dot2d", "PALM", var, opts, model))
        with pytest.raises(_lib.DotsocpError):
            s.run(solver.make_level_opts("dot2d", "acc-ADMM", var, opts, model))
        hb, res = s.run(o)                                # inPALM is fine
        assert np.array_equal(hb.kkt[:res.hist_len], states[0][0])


@pytest.mark.parametrize("case", ["dot2d-inPALM", "dot2d-accADMM", "wdot2d-inPALM", "dot1d-inPALM", "dot2d-noscaling"])
def test_resident_multilevel_transitions_match_host_transitions(gpu, case):
    """opts["resident"] (the default): recoverOrgVar + interpolate + jump_nextLevel + InitialScaling run on the device
    (dotsocp_prolong) and only the last level is downloaded; resident=False is the download -> host -> upload loop.  Every value is rounded where the host path rounds it, so the two solves agree
    bit for bit: same iteration counts, same KKT history, same iterates."""
    import dotsocp_b200 as dp
    if case.startswith("dot1d"):
        rho0, rho1 = O.get_example1d("gaussian", 257)
        run = lambda o: dp.solver_dotsocp1d(rho0, rho1, 17, 3, o, "inPALM")
        opts = {"tol": 1e-5, "maxit": 3000}
    elif case.startswith("wdot2d"):
        n, nt = 33, 17
        rho0, rho1 = O.get_example2d("example1", n, n)
        opts = {"tol": 1e-3, "maxit": 10000, "weight": O.gene_weight_circle(nt, n, n)}
        run = lambda o: dp.solver_wdotsocp2d(rho0, rho1, nt, 2, o, "inPALM")
    else:
        n, nt = 65, 33
        rho0, rho1 = O.get_example2d("example2", n, n)
        opts = {"tol": 1e-4, "maxit": 3000}
        if case.endswith("noscaling"):
            opts["scaling"] = False
        method = "acc-ADMM" if case.endswith("accADMM") else "inPALM"
        levels = 2 if case.endswith("accADMM") else 3
        run = lambda o: dp.solver_dotsocp2d(rho0, rho1, nt, levels, o, method)
    out_h, _, ML_h, rh_h = run(dict(opts, resident=False))
    out_r, _, ML_r, rh_r = run(dict(opts, resident=True, return_state=True))
    assert [int(v) for v in out_r.level_iters] == [int(v) for v in out_h.level_iters]
    assert ML_r.len == ML_h.len and np.array_equal(ML_r.iter, ML_h.iter)
    assert np.array_equal(ML_r.kkt, ML_h.kkt)
    for name in ("phi", "q", "z", "alpha", "beta"):
        a, b = np.asarray(getattr(out_r.var, name)), np.asarray(getattr(out_h.var, name))
        assert np.array_equal(a, b), name
    assert out_r.sigma == out_h.sigma
    # the outputs of the resident path come from the device (dotsocp_recover): recover_RhoE / recover_q / the mass check bit for bit
    from dotsocp_b200 import driver
    dim = 1 if case.startswith("dot1d") else 2
    for name in ("rho", "Ex", "Ey", "q0", "bx", "by")[:: 1]:
        if dim == 1 and name in ("Ey", "by"):
            continue
        assert np.array_equal(getattr(out_r, name), getattr(out_h, name)), name
    assert np.abs(out_r.sumRho - out_h.sumRho).max() < 1e-13 and np.abs(out_r.sumNegRho - out_h.sumNegRho).max() < 1e-13
    assert out_r.massOK == out_h.massOK
    assert abs(out_r.w2 - driver.w2_cost(out_h, dim)) <= 1e-12 * abs(out_r.w2)


@pytest.mark.parametrize("case", ["dot2d", "wdot2d", "dot1d", "dot2d-ALG2"])
def test_multilevel_solve_on_time_slabs_is_bit_identical_to_one_slab(gpu, case):
    """Whole multilevel solves with the time axis cut into 2, 3 and 4 slabs (emulated on one GPU): refined sessions keep the
    coarse partition (every cut doubled), the level transfer fills every slab from the coarse slab with the same index, the
    outputs are recovered slab by slab -- and nothing may depend on the partition: iteration counts, KKT history and all
    six output fields are compared bit for bit with the single-slab solve."""
    import dotsocp_b200 as dp
    if case == "dot1d":
        rho0, rho1 = O.get_example1d("gaussian", 257)
        run = lambda o: dp.solver_dotsocp1d(rho0, rho1, 33, 3, o, "inPALM")
        opts = {"tol": 1e-5, "maxit": 3000}
    elif case == "wdot2d":
        n, nt = 33, 33
        rho0, rho1 = O.get_example2d("example1", n, n)
        opts = {"tol": 1e-3, "maxit": 10000, "weight": O.gene_weight_circle(nt, n, n)}
        run = lambda o: dp.solver_wdotsocp2d(rho0, rho1, nt, 3, o, "inPALM")
    else:
        n, nt = 65, 33
        rho0, rho1 = O.get_example2d("example2", n, n)
        opts = {"tol": 1e-4, "maxit": 3000}
        run = lambda o: dp.solver_dotsocp2d(rho0, rho1, nt, 3, o, "ALG2" if case.endswith("ALG2") else "inPALM")
    ref = run(dict(opts))
    # (1-D: the x transform pairs consecutive TIME levels into one complex sequence, so bit-identity needs slabs that start on
    # even levels -- 2 and 4 slabs of the 8 coarsest cell layers do, 3 do not; the 2-D transforms pair lines inside a level)
    for k in ((2, 4) if case == "dot1d" else (2, 3, 4)):
        got = run(dict(opts, slabs=k))
        assert [int(v) for v in got[0].level_iters] == [int(v) for v in ref[0].level_iters], k
        assert np.array_equal(got[2].kkt, ref[2].kkt) and np.array_equal(got[2].iter, ref[2].iter), k
        for name in ("rho", "Ex", "Ey", "q0", "bx", "by"):
            if case == "dot1d" and name in ("Ey", "by"):
                continue
            assert np.array_equal(getattr(got[0], name), getattr(ref[0], name)), (name, k)
        assert np.array_equal(got[0].sumRho, ref[0].sumRho) and got[0].w2 == ref[0].w2
        assert got[0].sigma == ref[0].sigma


@pytest.mark.parametrize("n,nt,levelN", [(17, 9, 1), (33, 17, 2)])
def test_dot2d_sgs_inpalm_parity(gpu, n, nt, levelN):
    """sGS-inPALM (solver_socp_sGSinPALM.m): the phi-step is one symmetric red-black Gauss-Seidel sweep instead of the DCT solve,
    with its own check schedule and sigma voting; coarse levels run inPALM (solver_dotsocp2d.m:210-216).  Resident and host
    transitions, and emulated time slabs (the sGS phi-step exchanges ghost planes only)."""
    import dotsocp_b200 as dp
    rho0, rho1 = O.get_example2d("example1", n, n)
    opts = {"tol": 1e-4}
    out_o, _, ML_o, rh_o = O.solver_dotsocp2d(rho0, rho1, nt, levelN, opts, "sGS-inPALM", workers=4)
    for extra in ({}, {"resident": False}, {"slabs": 2}):
        out_g, _, ML_g, rh_g = dp.solver_dotsocp2d(rho0, rho1, nt, levelN, dict(opts, **extra), "sGS-inPALM")
        _compare(out_g, rh_g, ML_g, out_o, rh_o, ML_o)


@pytest.mark.parametrize("n,nt,levelN", [(17, 9, 1), (17, 9, 2)])
def test_dot2d_acc_sgs_admm_parity(gpu, n, nt, levelN):
    """acc-sGS-ADMM (solver_socp_accsGSADMM.m): Halpern-accelerated ADMM with the Gauss-Seidel phi-step and the sGS sigma voting"""
    import dotsocp_b200 as dp
    rho0, rho1 = O.get_example2d("example1", n, n)
    opts = {"tol": 1e-4}
    out_o, _, ML_o, rh_o = O.solver_dotsocp2d(rho0, rho1, nt, levelN, opts, "acc-sGS-ADMM", workers=4)
    for extra in ({}, {"resident": False}):
        out_g, _, ML_g, rh_g = dp.solver_dotsocp2d(rho0, rho1, nt, levelN, dict(opts, **extra), "acc-sGS-ADMM")
        _compare(out_g, rh_g, ML_g, out_o, rh_o, ML_o)


@pytest.mark.parametrize("nt,nx,ny,levels", [(9, 17, 17, 3), (17, 9, 33, 2), (5, 5, 9, 2), (3, 3, 3, 2)])
def test_device_weight_pyramid_matches_host_restriction(gpu, nt, nx, ny, levels):
    """the restriction chain downSample_q.m / downSample_barrier.m, the generators' time replication and
    mean(log10(w + 1e-10)) on the device.  Arithmetic chain: bit-identical to the host mirror (same separable order), which
    agrees with the oracle's sparse kron products to 1e-14; geometric chain: to the rounding of log / exp."""
    import dotsocp_b200 as dp
    from dotsocp_b200 import driver, solver
    rng = np.random.default_rng(7)
    Q = (nt - 1) * nx * ny + nt * (nx - 1) * ny + nt * nx * (ny - 1)
    w = np.abs(rng.standard_normal(Q)) + 0.1
    with solver.Weights(nt, nx, ny, levels) as pyr:
        pyr.set(w).restrict(False)
        assert np.array_equal(pyr.get(0), w)
        host, dims = w, (nt, nx, ny)
        for l in range(1, levels):
            host = driver.downSample_q(*dims, host)
            dims = tuple((d + 1) // 2 for d in dims)
            assert pyr.dims(l) == dims
            got = pyr.get(l)
            assert np.array_equal(got, host), (l, np.abs(got - host).max())
        assert np.allclose(host, O.downSample_q(*[2 * d - 1 for d in dims], pyr.get(levels - 2)), rtol=1e-13, atol=0)
        for l in range(levels):
            ref = float(np.mean(np.log10(pyr.get(l) + 1e-10)))
            assert abs(pyr.log10_mean(l) - ref) <= 1e-13 * max(1.0, abs(ref))
        pyr.restrict(True)
        host, dims = w, (nt, nx, ny)
        for l in range(1, levels):
            host = driver.downSample_barrier(*dims, host)
            dims = tuple((d + 1) // 2 for d in dims)
            assert np.allclose(pyr.get(l), host, rtol=1e-13, atol=0)
        wX, wY = driver.weight_planes_circle(nx, ny)
        pyr.set_planes(wX, wY)
        assert np.array_equal(pyr.get(0), O.gene_weight_circle(nt, nx, ny))
        assert pyr.gpu_launches > 0
    with pytest.raises(dp.DotsocpError):
        solver.Weights(8, 17, 17, 2)       # an even node count cannot be halved
    with solver.Weights(5, 5, 5, 2) as pyr:
        with pytest.raises(dp.DotsocpError):
            pyr.get(1)                      # nothing computed yet


@pytest.mark.parametrize("case", ["circle", "barrier", "circle-slabs"])
def test_wdot2d_with_device_resident_weights(gpu, case):
    """the weighted multilevel solve with the weights taken from the device pyramid (planes in, nothing Q-sized on the host)
    against the same solve with host weights: same iterations; KKT history and outputs agree to rounding (the log-mean of
    `adjust` is reduced in a different order, the barrier chain uses the device's log / exp)"""
    import dotsocp_b200 as dp
    from dotsocp_b200 import driver
    n, nt = 33, 17
    barrier = None
    if case == "barrier":
        barrier = O.gene_barrier_of_love_heart()
        rho0, rho1 = O.gene_exampleLoveHeart(n, n)
        rho0, rho1 = O._normalize2d(rho0, rho1, n, n)
        rho0, rho1, _ = O.ensure_barrier_validity(rho0, rho1, barrier)
        planes = driver.weight_planes_barrier(n, n, barrier)
        opts = {"tol": 1e-3, "maxit": 3000}
    else:
        rho0, rho1 = O.get_example2d("example1", n, n)
        planes = driver.weight_planes_circle(n, n)
        opts = {"tol": 1e-3, "maxit": 10000}
    if case.endswith("slabs"):
        opts["slabs"] = 3
    host = dp.solver_wdotsocp2d(rho0, rho1, nt, 2, dict(opts, weight=driver.weight_from_planes(nt, *planes)), "inPALM", barrier)
    dev = dp.solver_wdotsocp2d(rho0, rho1, nt, 2, dict(opts, weight_planes=planes), "inPALM", barrier)
    dev2 = dp.solver_wdotsocp2d(rho0, rho1, nt, 2, dict(opts, weight=driver.weight_from_planes(nt, *planes), weights_on_device=True),
                                "inPALM", barrier)
    for got in (dev, dev2):
        assert [int(v) for v in got[0].level_iters] == [int(v) for v in host[0].level_iters]
        loose = case == "barrier"
        assert np.abs(got[2].kkt - host[2].kkt).max() < (1e-8 if loose else 1e-9)
        for name in ("rho", "Ex", "Ey", "q0", "bx", "by"):
            # (barrier: weights of 1e6 multiply alpha in recover_RhoE and amplify the rounding differences; 1e-6 is _compare's bound)
            assert np.abs(getattr(got[0], name) - getattr(host[0], name)).max() < (1e-6 if case == "barrier" else 1e-8), name
        assert abs(got[0].w2 - host[0].w2) <= (1e-6 if loose else 1e-9) * abs(host[0].w2)
    assert np.array_equal(dev[2].kkt, dev2[2].kkt)     # planes or an uploaded finest level: the same pyramid
