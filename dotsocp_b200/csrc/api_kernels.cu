// api_kernels.cu -- kernel-level C-ABI entry points (host buffers in, host buffers out): the GPU counterparts of the
// reference's MEX kernels and of oper_poisson3dim.  Parity/utility path; the performance path is the session in solver.cu.
#include "../../include/dotsocp.h"
#include "errs.h"
#include "kernels.h"

using namespace dsocp;

// ------------------------------------------------------------------------------------------------ kernel-level entry points
struct DevBuf {
    double* p = nullptr;
    ~DevBuf() { cudaFree(p); }
    int alloc(size_t n)
    {
        cudaError_t e = cudaMalloc(&p, (n ? n : 1) * sizeof(double));
        if (e != cudaSuccess) { cudaGetLastError(); return set_err(DOTSOCP_ENOMEM, "cudaMalloc(%zu doubles): %s", n, cudaGetErrorString(e)); }
        return 0;
    }
};

extern "C" int dotsocp_mexBFd(double* z2, const double* q, int nt, int nx, int ny, double scaleBF, double scaleD)
{
    if (!z2 || !q || nt < 2 || nx < 1 || ny < 1) return set_err(DOTSOCP_EINVAL, "mexBFd: bad arguments");
    int rc = require_device();
    if (rc) return rc;
    const Geo g = make_geo(nt, nx, ny);
    DevBuf dz, dq;
    if ((rc = dz.alloc(10 * g.L)) || (rc = dq.alloc(g.Q))) return rc;
    // in-place semantics: entries the kernel does not write keep the caller's values
    CU(cudaMemcpy(dz.p, z2, (size_t)10 * g.L * sizeof(double), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(dq.p, q, (size_t)g.Q * sizeof(double), cudaMemcpyHostToDevice));
    launch_bfd(g, scaleBF, scaleD, dq.p, dz.p, 0);
    CU(cudaGetLastError());
    CU(cudaMemcpy(z2, dz.p, (size_t)10 * g.L * sizeof(double), cudaMemcpyDeviceToHost));
    return DOTSOCP_OK;
}

extern "C" int dotsocp_mexBFdConj(double* q2, const double* z, int nt, int nx, int ny, double scaleBF)
{
    if (!q2 || !z || nt < 2 || nx < 1 || ny < 1) return set_err(DOTSOCP_EINVAL, "mexBFdConj: bad arguments");
    int rc = require_device();
    if (rc) return rc;
    const Geo g = make_geo(nt, nx, ny);
    DevBuf dz, dq;
    if ((rc = dz.alloc(10 * g.L)) || (rc = dq.alloc(g.Q))) return rc;
    CU(cudaMemcpy(dz.p, z, (size_t)10 * g.L * sizeof(double), cudaMemcpyHostToDevice));
    launch_bfdconj(g, scaleBF, dz.p, dq.p, 0);
    CU(cudaGetLastError());
    CU(cudaMemcpy(q2, dq.p, (size_t)g.Q * sizeof(double), cudaMemcpyDeviceToHost));
    return DOTSOCP_OK;
}

extern "C" int dotsocp_mexProjSoc(double* out, const double* in, int64_t M, int N)
{
    if (!out || !in || M < 0 || N < 1) return set_err(DOTSOCP_EINVAL, "mexProjSoc: bad arguments");
    int rc = require_device();
    if (rc) return rc;
    if (M == 0) return DOTSOCP_OK;
    DevBuf di, dout;
    if ((rc = di.alloc((size_t)M * N)) || (rc = dout.alloc((size_t)M * N))) return rc;
    CU(cudaMemcpy(di.p, in, (size_t)M * N * sizeof(double), cudaMemcpyHostToDevice));
    launch_projsoc(M, N, di.p, dout.p, 0);
    CU(cudaGetLastError());
    CU(cudaMemcpy(out, dout.p, (size_t)M * N * sizeof(double), cudaMemcpyDeviceToHost));
    return DOTSOCP_OK;
}

extern "C" int dotsocp_mexBFd1d(double* z, const double* q, int nt, int nx, double scale, double dFactor)
{
    if (!z || !q || nt < 2 || nx < 1) return set_err(DOTSOCP_EINVAL, "mexBFd:invalidInput");
    int rc = require_device();
    if (rc) return rc;
    const Geo g = make_geo(nt, nx, 1);
    DevBuf d6, d10, dq;
    if ((rc = d6.alloc(6 * g.L)) || (rc = d10.alloc(10 * g.L)) || (rc = dq.alloc(g.Q))) return rc;
    CU(cudaMemcpy(d6.p, z, (size_t)6 * g.L * sizeof(double), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(dq.p, q, (size_t)g.Q * sizeof(double), cudaMemcpyHostToDevice));
    launch_cols6to10(d6.p, d10.p, g.L, 0);
    launch_bfd(g, scale, dFactor, dq.p, d10.p, 0);
    launch_cols10to6(d10.p, d6.p, g.L, 0);
    CU(cudaGetLastError());
    CU(cudaMemcpy(z, d6.p, (size_t)6 * g.L * sizeof(double), cudaMemcpyDeviceToHost));
    return DOTSOCP_OK;
}

extern "C" int dotsocp_mexBFdConj1d(double* q, const double* z, int nt, int nx, double scale)
{
    if (!z || !q || nt < 2 || nx < 1) return set_err(DOTSOCP_EINVAL, "mexBFd:invalidInput");
    int rc = require_device();
    if (rc) return rc;
    const Geo g = make_geo(nt, nx, 1);
    DevBuf d6, d10, dq;
    if ((rc = d6.alloc(6 * g.L)) || (rc = d10.alloc(10 * g.L)) || (rc = dq.alloc(g.Q))) return rc;
    CU(cudaMemcpy(d6.p, z, (size_t)6 * g.L * sizeof(double), cudaMemcpyHostToDevice));
    launch_cols6to10(d6.p, d10.p, g.L, 0);
    launch_bfdconj(g, scale, d10.p, dq.p, 0);
    CU(cudaGetLastError());
    CU(cudaMemcpy(q, dq.p, (size_t)g.Q * sizeof(double), cudaMemcpyDeviceToHost));
    return DOTSOCP_OK;
}

extern "C" int dotsocp_mexsGS(double* phi, const double* rhs, double ep, double scale, int nt, int nx, int ny, int its)
{
    if (!phi || !rhs || nt < 3 || nx < 3 || ny < 3 || its < 0) return set_err(DOTSOCP_EINVAL, "bad arguments");
    const Geo g = make_geo(nt, nx, ny);
    if (!sgs_supported(g)) return set_err(DOTSOCP_EINVAL, "mexsGS handles only nx == ny with odd node counts (got %d x %d x %d)", nt, nx, ny);
    int rc = require_device();
    if (rc) return rc;
    DevBuf a, b;
    if ((rc = a.alloc(g.N)) || (rc = b.alloc(g.N))) return rc;
    CU(cudaMemcpy(a.p, phi, (size_t)g.N * sizeof(double), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(b.p, rhs, (size_t)g.N * sizeof(double), cudaMemcpyHostToDevice));
    launch_sgs_half(g, ep, scale, 1, b.p, a.p, 0, nt, 0);            // odd ; its x [even ; odd]   (mexFunction @0x2598-0x2629)
    for (int i = 0; i < its; i++) {
        launch_sgs_half(g, ep, scale, 0, b.p, a.p, 0, nt, 0);
        launch_sgs_half(g, ep, scale, 1, b.p, a.p, 0, nt, 0);
    }
    CU(cudaGetLastError());
    CU(cudaMemcpy(phi, a.p, (size_t)g.N * sizeof(double), cudaMemcpyDeviceToHost));
    return DOTSOCP_OK;
}

static int dct_common(double* out, const double* in, int nt, int nx, int ny, int what, double D)
{
    if (!out || !in || nt < 1 || nx < 1 || ny < 1) return set_err(DOTSOCP_EINVAL, "bad arguments");
    int rc = require_device();
    if (rc) return rc;
    const i64 N = (i64)nt * nx * ny;
    DevBuf a, b;
    if ((rc = a.alloc(N)) || (rc = b.alloc(N))) return rc;
    CU(cudaMemcpy(a.p, in, (size_t)N * sizeof(double), cudaMemcpyHostToDevice));
    PoissonPlan* pp = poisson_plan_create(nt, nx, ny);
    int prc = 0;
    if (what == 2) prc = poisson_solve(pp, a.p, b.p, D * D, 0, nullptr);
    else { CU(cudaMemcpy(b.p, a.p, (size_t)N * sizeof(double), cudaMemcpyDeviceToDevice)); prc = poisson_dctn(pp, b.p, what == 1, 0, nullptr); }
    cudaError_t e = cudaDeviceSynchronize();
    poisson_plan_destroy(pp);
    if (prc) return prc;   // unsupported geometry: dotsocp_last_error() says which
    if (e != cudaSuccess) return set_err(DOTSOCP_ECUDA, "transform kernels: %s", cudaGetErrorString(e));
    CU(cudaGetLastError());
    CU(cudaMemcpy(out, b.p, (size_t)N * sizeof(double), cudaMemcpyDeviceToHost));
    return DOTSOCP_OK;
}

extern "C" int dotsocp_poisson(double* phi, const double* rhs, int nt, int nx, int ny, double D)
{
    return dct_common(phi, rhs, nt, nx, ny, 2, D);
}
extern "C" int dotsocp_dctn(double* a, int nt, int nx, int ny, int inverse)
{
    return dct_common(a, a, nt, nx, ny, inverse ? 1 : 0, 1.0);
}
