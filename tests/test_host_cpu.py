"""CPU tests of the product's host side: the C-ABI library loads and exports everything include/dotsocp.h declares,
fails loudly without a GPU, and the Python mirror of the reference's driver logic agrees with the oracle."""
import os
import re

import numpy as np
import pytest

from oracle import dotsocp_oracle as O
from oracle import kernels as K

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(built):
    from dotsocp_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "dotsocp.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(dotsocp_[A-Za-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    lib = _lib.lib()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/dotsocp.h but not exported by libdotsocp.so"
    assert declared == set(_lib.EXPORTS)
    assert lib.dotsocp_version() >= 100


def test_no_cpu_fallback(built):
    """Without a usable CUDA device every compute entry point must fail with DOTSOCP_ENODEV (never compute on the CPU)."""
    from dotsocp_b200 import _lib, ops
    if _lib.lib().dotsocp_device_count() > 0:
        pytest.skip("a GPU is present")
    z = np.zeros((4, 10), order="F")
    with pytest.raises(_lib.DotsocpError) as e:
        ops.mexBFd(z, np.zeros(K.sizes2d(2, 2, 2)[0] + 8), 2, 2, 2, 1.0, 1.0)
    assert e.value.code == -2
    with pytest.raises(_lib.DotsocpError):
        import dotsocp_b200 as dp
        dp.Session("dot2d", 5, 5, 5)


def test_struct_layout_matches_header(built):
    """ctypes mirrors of dotsocp_level_opts / _result: field order and sizes as declared (no compute call)."""
    import ctypes as C
    from dotsocp_b200 import _lib
    assert C.sizeof(_lib.LevelOpts) == 10 * 4 + 15 * 8
    assert C.sizeof(_lib.LevelResult) == 2 * 4 + 5 * 8 + 8 * 8 + 8
    assert [f[0] for f in _lib.LevelOpts._fields_][:5] == ["variant", "method", "nt", "nx", "ny"]


def test_driver_setup_matches_oracle():
    from dotsocp_b200 import driver
    rho0, rho1 = O.get_example2d("example2", 9, 9)
    var, model = driver.initialize(rho0, rho1, 5)
    vo, mo = O.initialize2d(rho0, rho1, 5)
    assert np.array_equal(var.phi, vo.phi) and np.array_equal(model.c, mo.c)
    driver.InitialScaling(var, model, True, None, "dot2d")
    O.InitialScaling(vo, mo, True, None, "dot2d")
    for k in ("cScale", "dScale", "D", "E", "E2"):
        assert getattr(var, k) == getattr(vo, k)
    assert model.normc == mo.normc and model.normd == mo.normd
    assert np.array_equal(var.phi, vo.phi) and np.array_equal(model.c, mo.c)
    # grad triple == the three magnitudes of the oracle's sparse gradient
    from dotsocp_b200.solver import grad_scalars
    assert grad_scalars(model) == pytest.approx(grad_scalars(mo), rel=1e-15)
    rng = np.random.default_rng(0)
    phi = rng.standard_normal(5 * 9 * 9)
    assert np.allclose(driver.grad_apply(phi, 5, 9, 9, model.grad), mo.grad @ phi, rtol=0, atol=1e-12)


def test_driver_level_transfer_matches_oracle():
    from dotsocp_b200 import driver
    rng = np.random.default_rng(1)
    rho0, rho1 = O.get_example2d("example1", 9, 9)
    var, model = driver.initialize(rho0, rho1, 5)
    vo, mo = O.initialize2d(rho0, rho1, 5)
    var.phi = rng.standard_normal(var.phi.size); vo.phi = var.phi.copy()
    var.beta = np.asfortranarray(rng.standard_normal(var.beta.shape)); vo.beta = var.beta.copy(order="F")
    a = driver.interpolate(var, model)
    b = O.interpolate(vo, mo)
    assert np.array_equal(a.phi, b.phi) and np.array_equal(a.beta, b.beta)
    v = rng.standard_normal((17, 17))
    assert np.array_equal(driver.downSample_phi(v), O.downSample_phi2d(v))
    v1 = rng.standard_normal(33)
    assert np.array_equal(driver.downSample_phi(v1), O.downSample_phi1d(v1))
    w = np.abs(rng.standard_normal(K.sizes2d(9, 17, 17)[0] + sum(K.sizes2d(9, 17, 17)[1:]))) + 0.1
    assert np.allclose(driver.downSample_q(9, 17, 17, w), O.downSample_q(9, 17, 17, w), rtol=1e-14, atol=0)
    assert np.allclose(driver.downSample_barrier(9, 17, 17, w), O.downSample_barrier(9, 17, 17, w), rtol=1e-13, atol=0)


def test_driver_output_recovery_matches_oracle():
    from dotsocp_b200 import driver
    rng = np.random.default_rng(2)
    rho0, rho1 = O.get_example2d("example1", 9, 9)
    var, model = driver.initialize(rho0, rho1, 5)
    vo, mo = O.initialize2d(rho0, rho1, 5)
    var.alpha = rng.standard_normal(var.alpha.size); vo.alpha = var.alpha.copy()
    var.q = rng.standard_normal(var.q.size); vo.q = var.q.copy()
    for a, b in zip(driver.recover_RhoE(var, model), O.recover_RhoE(vo, mo)):
        assert np.array_equal(a, b)
    for a, b in zip(driver.recover_q(var, model), O.recover_q(vo, mo)):
        assert np.array_equal(a, b)
    r0, r1 = O.get_example1d("gaussian", 17)
    v1, m1 = driver.initialize(r0, r1, 5)
    o1, n1 = O.initialize1d(r0, r1, 5)
    assert np.array_equal(v1.phi, o1.phi) and np.array_equal(m1.c, n1.c)
    v1.alpha = rng.standard_normal(v1.alpha.size); o1.alpha = v1.alpha.copy()
    for a, b in zip(driver.recover_RhoE(v1, m1), O.recover_RhoE(o1, n1)):
        assert np.array_equal(a, b)


def test_weight_plane_generators_match_oracle():
    """the product's own plane generators (gene_weight_circle.m / get_weight_by_barrier.m) against the oracle's full weights"""
    from dotsocp_b200 import driver
    for nt, nx, ny in ((5, 9, 9), (9, 17, 33)):
        wX, wY = driver.weight_planes_circle(nx, ny)
        assert wX.shape == (ny, nx - 1) and wY.shape == (ny - 1, nx)
        assert np.array_equal(driver.weight_from_planes(nt, wX, wY), O.gene_weight_circle(nt, nx, ny))
    barrier = O.gene_barrier_of_love_heart()
    wX, wY = driver.weight_planes_barrier(17, 17, barrier)
    assert np.array_equal(driver.weight_from_planes(9, wX, wY), O.get_weight_by_barrier(17, 17, 9, barrier))
    assert (wX == 1e6).any() and (wX == 1.0).any()


def test_slab_local_initial_state_equals_split_of_the_full_initial_state():
    """the resident multilevel driver builds the coarsest state slab by slab (driver.initial_state_local): bit-identical to
    split_state(initialize + InitialScaling), for 2-D and 1-D, scaled and unscaled, 1 and 3 slabs"""
    from dotsocp_b200 import driver
    from dotsocp_b200 import slab as SL
    for dim in (2, 1):
        r0, r1 = O.get_example2d("example1", 9, 17) if dim == 2 else O.get_example1d("gaussian", 33)
        nt, variant = 9, ("dot2d" if dim == 2 else "dot1d")
        for scalingYes in (True, False):
            var, model = driver.initialize(r0, r1, nt)
            driver.InitialScaling(var, model, scalingYes, None, variant)
            _, m2 = driver.level_model(r0, r1, nt)
            cS, dS, D, E, _ = driver.scaling_scalars(m2.nt * m2.nx * m2.ny, m2, scalingYes, None, None, variant)
            assert (cS, dS, D, E) == (var.cScale, var.dScale, var.D, var.E)
            assert m2.normc == model.normc and m2.normd == model.normd and m2.grad == model.grad
            for world in (1, 3):
                for rank in range(world):
                    tr = SL.partition(nt, world)[rank]
                    ref = SL.split_state(rank, world, model.nt, model.nx, model.ny, var.phi, var.q, var.z, var.alpha, var.beta, model.c)
                    got = driver.initial_state_local(m2, dS if scalingYes else None, *tr)
                    for a, b in zip(ref[:6], got):
                        assert a.shape == b.shape and np.array_equal(a, b)
            assert driver.initial_state_local(m2, None, 0, nt - 1, 0, nt, with_z=False)[2] is None


def test_output_buffers_prefault_in_the_background():
    from types import SimpleNamespace
    from dotsocp_b200 import solver
    fake = SimpleNamespace(output_shapes=lambda fields: {"rho": (5, 9, 9), "q0": (4, 9, 9), "big": (3, 1200, 1200)})
    out = solver.OutputBuffers(fake).get()
    assert set(out) == {"rho", "q0", "big"} and out["big"].shape == (3, 1200, 1200) and out["rho"].flags.c_contiguous
    assert not out["big"].any()
