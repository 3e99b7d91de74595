"""GPU parity of the kernel-level C-ABI entry points against the oracle (bit-exact for the cell-local kernels)."""
import numpy as np
import pytest
import scipy.fft as sfft

from oracle import kernels as K
from oracle import dotsocp_oracle as O

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a).view(np.uint64) if a.flags.c_contiguous else np.asfortranarray(a).T.copy().view(np.uint64)


def same_bits(a, b):
    return np.array_equal(np.asarray(a).ravel(order="K").view(np.uint64), np.asarray(b).ravel(order="K").view(np.uint64))


GRIDS = [(3, 3, 4), (5, 4, 3), (9, 17, 17), (4, 5, 7), (2, 2, 2), (17, 33, 31), (33, 65, 65)]


@pytest.mark.parametrize("nt,nx,ny", GRIDS)
def test_mexBFd_bit_exact(gpu, nt, nx, ny):
    from dotsocp_b200 import ops
    rng = np.random.default_rng(1)
    L, nbx, nby = K.sizes2d(nt, nx, ny)
    q = rng.standard_normal(L + nbx + nby)
    S, DF = 1.2345, 0.777
    z_o = np.full((L, 10), -7.5, order="F")
    z_g = z_o.copy(order="F")
    K.mexBFd(z_o, q, nt, nx, ny, S, DF)
    ops.mexBFd(z_g, q, nt, nx, ny, S, DF)
    assert same_bits(z_o, z_g)          # includes the untouched boundary entries (-7.5)


@pytest.mark.parametrize("nt,nx,ny", GRIDS)
def test_mexBFdConj_bit_exact_and_adjoint(gpu, nt, nx, ny):
    from dotsocp_b200 import ops
    rng = np.random.default_rng(2)
    L, nbx, nby = K.sizes2d(nt, nx, ny)
    z = np.asfortranarray(rng.standard_normal((L, 10)))
    q_o = np.zeros(L + nbx + nby)
    q_g = np.full(L + nbx + nby, 3.0)
    K.mexBFdConj(q_o, z, nt, nx, ny, 0.9)
    ops.mexBFdConj(q_g, z, nt, nx, ny, 0.9)
    assert same_bits(q_o, q_g)
    # <BF q, z> == <q, (BF)^* z>   (mexBFd with DF = 0 on a zeroed z2)
    q = rng.standard_normal(L + nbx + nby)
    bfq = np.zeros((L, 10), order="F")
    ops.mexBFd(bfq, q, nt, nx, ny, 0.9, 0.0)
    assert abs(np.vdot(bfq, z) - np.dot(q, q_g)) <= 1e-12 * (1 + abs(np.dot(q, q_g)))


@pytest.mark.parametrize("M,N", [(1, 10), (2, 10), (5, 10), (1001, 10), (1000, 6), (33, 6), (64, 3), (17, 2), (40, 12), (9, 7)])
def test_mexProjSoc_bit_exact(gpu, M, N):
    from dotsocp_b200 import ops
    rng = np.random.default_rng(3)
    v = np.asfortranarray(rng.standard_normal((M, N)))
    v[0, 0] = 25.0                       # strictly inside the cone
    if M > 4:
        v[1, 0] = -50.0                  # inside the polar cone -> 0
        v[2, 1:] = 0.0                   # |x| = 0, t != 0
        v[3, :] = 0.0                    # 0/0 -> NaN row, like the binary
        v[4, 0] = np.linalg.norm(v[4, 1:])   # on the boundary
    o_o = np.zeros((M, N), order="F")
    o_g = np.zeros((M, N), order="F")
    K.mexProjSoc(o_o, v)
    ops.mexProjSoc(o_g, v)
    assert same_bits(o_o, o_g)


def test_mexProjSoc_properties(gpu):
    from dotsocp_b200 import ops
    rng = np.random.default_rng(4)
    v = np.asfortranarray(rng.standard_normal((4096, 10)) * 3)
    p = np.zeros_like(v, order="F")
    ops.mexProjSoc(p, v)
    nrm = np.linalg.norm(p[:, 1:], axis=1)
    assert np.all(nrm <= p[:, 0] * (1 + 1e-12) + 1e-14)          # in the cone
    nz = np.abs(p).sum(axis=1) > 0                               # rows projected onto the apex give 0/0 = NaN when
    p2 = np.zeros_like(v, order="F")                             # projected again, exactly like the reference binary
    ops.mexProjSoc(p2, p)
    assert np.allclose(p2[nz], p[nz], rtol=0, atol=1e-13)        # idempotent
    m = np.zeros_like(v, order="F")
    ops.mexProjSoc(m, np.asfortranarray(-v))
    assert np.allclose(v, p - m, rtol=0, atol=1e-12)             # Moreau: v = P_K(v) - P_K(-v) (self-dual cone)


@pytest.mark.parametrize("nt,nx", [(3, 4), (9, 17), (5, 2), (2, 3), (33, 129)])
def test_1d_kernels_bit_exact(gpu, nt, nx):
    from dotsocp_b200 import ops
    rng = np.random.default_rng(5)
    L = (nt - 1) * nx
    Q = L + nt * (nx - 1)
    q = rng.standard_normal(Q)
    z_o = np.full((L, 6), 2.5, order="F")
    z_g = z_o.copy(order="F")
    K.mexBFd1d(z_o, q, nt, nx, 0.8, 1.1)
    ops.mexBFd1d(z_g, q, nt, nx, 0.8, 1.1)
    assert same_bits(z_o, z_g)
    z = np.asfortranarray(rng.standard_normal((L, 6)))
    q_o, q_g = np.zeros(Q), np.zeros(Q)
    K.mexBFdConj1d(q_o, z, nt, nx, 0.8)
    ops.mexBFdConj1d(q_g, z, nt, nx, 0.8)
    assert same_bits(q_o, q_g)


DCT_GRIDS = [(5, 9, 9), (9, 17, 17), (17, 33, 33), (33, 65, 65), (7, 12, 20), (40, 50, 70), (3, 129, 5), (65, 7, 257),
             (2, 300, 3), (33, 1, 1),
             # the line lengths of the BASELINE grids (257, 513, 1025 and the next one), along x and along y: every
             # instantiation of the register-FFT kernel (M = 512 ... 4096) in both of its line layouts
             (4, 257, 6), (3, 513, 4), (2, 1025, 3), (3, 5, 513), (2, 3, 1025), (2, 2049, 2)]


@pytest.mark.parametrize("nt,nx,ny", DCT_GRIDS)
def test_dctn_matches_orthonormal_dct2(gpu, nt, nx, ny):
    from dotsocp_b200 import ops
    rng = np.random.default_rng(6)
    a = rng.standard_normal((nt, nx, ny))
    f = ops.dctn(a.ravel(), nt, nx, ny).reshape(nt, nx, ny)
    ref = sfft.dctn(a, type=2, norm="ortho")
    assert np.abs(f - ref).max() <= 5e-13 * max(1.0, np.abs(ref).max())
    b = ops.dctn(f.ravel(), nt, nx, ny, inverse=True).reshape(nt, nx, ny)
    assert np.abs(b - a).max() <= 5e-13


@pytest.mark.parametrize("nt,nx,ny", [(9, 17, 17), (17, 33, 33), (33, 65, 65), (12, 40, 33), (65, 129, 129)])
def test_poisson_matches_reference_formula(gpu, nt, nx, ny):
    from dotsocp_b200 import ops
    rng = np.random.default_rng(7)
    rhs = rng.standard_normal(nt * nx * ny)
    D = 0.37
    phi = ops.oper_poisson3dim(rhs, nt, nx, ny, D)
    ker = D ** 2 * O.initialize_FFTkernel(nt, nx, ny)
    ref = O.oper_poisson(ker, rhs)
    assert np.abs(phi - ref).max() <= 1e-11 * np.abs(ref).max()
    # A'A phi == rhs - mean(rhs) up to the DC convention (kernel(0)=1): check through the oracle's sparse gradient
    A = (D * O.gene_grad2d(nt, nx, ny)).tocsr()
    r = A.T @ (A @ phi)
    assert np.abs(r - (rhs - rhs.mean())).max() <= 1e-9 * np.abs(rhs).max()


def test_poisson_1d_variant(gpu):
    from dotsocp_b200 import ops
    rng = np.random.default_rng(8)
    nt, nx = 33, 257
    rhs = rng.standard_normal(nt * nx)
    phi = ops.oper_poisson(rhs, nt, nx, 0.5)
    ref = O.oper_poisson(0.25 * O.initialize_FFTkernel(nt, nx), rhs)
    assert np.abs(phi - ref).max() <= 1e-11 * np.abs(ref).max()


@pytest.mark.parametrize("nt,nx,ny", [(17, 33, 20), (33, 129, 65)])
def test_poisson_t_direction_thomas_equals_transform(gpu, nt, nx, ny, monkeypatch):
    """The t direction is a tridiagonal solve by default; DOTSOCP_TSOLVE=dct (read when the plan is created) selects the
    fused DCT_t -> ./kernel -> IDCT_t pass.  Both must match the reference formula."""
    from dotsocp_b200 import ops
    rng = np.random.default_rng(21)
    rhs = rng.standard_normal(nt * nx * ny)
    D = 0.61
    ref = O.oper_poisson(D ** 2 * O.initialize_FFTkernel(nt, nx, ny), rhs)
    monkeypatch.delenv("DOTSOCP_TSOLVE", raising=False)
    thomas = ops.oper_poisson3dim(rhs, nt, nx, ny, D)
    monkeypatch.setenv("DOTSOCP_TSOLVE", "dct")
    viadct = ops.oper_poisson3dim(rhs, nt, nx, ny, D)
    for got in (thomas, viadct):
        assert np.abs(got - ref).max() <= 1e-11 * np.abs(ref).max()
    assert np.abs(thomas - viadct).max() > 0.0 or nt < 3   # they really are two different code paths


@pytest.mark.parametrize("nt,n,its,ep,scale", [(5, 5, 1, 0.0, 1.0), (9, 7, 1, 0.0, 0.37), (3, 9, 2, 0.5, 2.5), (17, 33, 1, 0.0, 1.3e-2),
                                                (33, 17, 3, 1e-3, 0.7), (3, 3, 1, 0.0, 1.0), (65, 129, 1, 0.0, 2.1e-3)])
def test_mexsGS_bit_exact(gpu, nt, n, its, ep, scale):
    """red-black symmetric Gauss-Seidel sweeps (mexsGS.mexa64): one thread per node of the half sweep's parity, bit-identical to
    the reference binary (through the oracle, whose numpy restatement is itself pinned to the binary)"""
    from dotsocp_b200 import ops
    from oracle import kernels as K
    rng = np.random.default_rng(nt * 1000 + n)
    phi = rng.standard_normal(nt * n * n)
    rhs = rng.standard_normal(nt * n * n)
    want, got = phi.copy(), phi.copy()
    K.mexsGS(want, rhs.copy(), ep, scale, nt, n, n, its)
    ops.mexsGS(got, rhs, ep, scale, nt, n, n, its)
    assert np.array_equal(got, want)
    from dotsocp_b200 import _lib
    with pytest.raises(_lib.DotsocpError):
        ops.mexsGS(np.zeros(5 * 5 * 7), np.zeros(5 * 5 * 7), 0.0, 1.0, 5, 5, 7, 1)      # nx != ny: the binary does not handle it
