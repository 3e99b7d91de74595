// prolong.cu -- level transfer of the multilevel drivers on the device.
//
// Reference: socp/dot2d/utils/interpolate.m:46-84 (phi: linear nodal interpolation along y, then x, then t; beta: nearest
// in t, linear along y then x, column by column), jump_nextLevel.m:5-16 (q = A phi, alpha = (BF)^*(-beta)),
// solver_dotsocp2d.m:368-386 (recoverOrgVar) and :304-365 (InitialScaling).  Every value goes through the same sequence
// of individually rounded operations as the host path (dotsocp_b200/driver.py), so a solve with resident transitions is
// bit-identical to one that downloads, transfers on the host and uploads again.  Compiled with -fmad=false.
#include "kernels.h"

namespace dsocp {

__device__ __forceinline__ double avg2(double a, double b) { return dmul(dadd(a, b), 0.5); }

// fine node (tf,xf,yf) <- coarse array (values pre-multiplied by `rec`, the recoverOrgVar factor)
__global__ void __launch_bounds__(256) k_prolong_phi(Geo gf, Geo gc, double rec, const double* __restrict__ pc, double* __restrict__ pf)
{
    const i64 p = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    const int tf = blockIdx.y;
    if (p >= gf.P) return;
    const int xf = (int)(p / gf.ny), yf = (int)(p - (i64)xf * gf.ny);
    const bool ry = gc.ny > 1;                     // the 1-D variant has no y direction to refine
    const int tc = tf >> 1, xc = xf >> 1, yc = ry ? (yf >> 1) : 0;
    const bool ot = tf & 1, ox = xf & 1, oy = ry && (yf & 1);
    auto V = [&](int t, int x, int y) { return dmul(rec, pc[(i64)t * gc.P + (i64)x * gc.ny + y]); };
    auto Y = [&](int t, int x) { return oy ? avg2(V(t, x, yc), V(t, x, yc + 1)) : V(t, x, yc); };
    auto X = [&](int t) { return ox ? avg2(Y(t, xc), Y(t, xc + 1)) : Y(t, xc); };
    pf[(i64)tf * gf.P + p] = ot ? avg2(X(tc), X(tc + 1)) : X(tc);
}

// fine cell (tf,xf,yf), all 10 planes <- coarse cell layer tf/2
__global__ void __launch_bounds__(256) k_prolong_beta(Geo gf, Geo gc, double rec, const double* __restrict__ bc, double* __restrict__ bf)
{
    const i64 p = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    const int tf = blockIdx.y;
    if (p >= gf.P) return;
    const int xf = (int)(p / gf.ny), yf = (int)(p - (i64)xf * gf.ny);
    const bool ry = gc.ny > 1;
    const int tc = tf >> 1, xc = xf >> 1, yc = ry ? (yf >> 1) : 0;
    const bool ox = xf & 1, oy = ry && (yf & 1);
#pragma unroll 1
    for (int j = 0; j < 10; j++) {
        const double* src = bc + (i64)j * gc.L + (i64)tc * gc.P;
        auto V = [&](int x, int y) { return dmul(rec, src[(i64)x * gc.ny + y]); };
        auto Y = [&](int x) { return oy ? avg2(V(x, yc), V(x, yc + 1)) : V(x, yc); };
        bf[(i64)j * gf.L + (i64)tf * gf.P + p] = ox ? avg2(Y(xc), Y(xc + 1)) : Y(xc);
    }
}

// q = scale * ((A phi) [./ weight])  with the UNSCALED forward differences of the fine grid (initialize.m:67-87)
__global__ void __launch_bounds__(256) k_prolong_q(Geo g, double gt, double gx, double gy, double scale, const double* __restrict__ phi,
                                                   const double* __restrict__ weight, double* __restrict__ q)
{
    const i64 p = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    const int t = blockIdx.y;
    if (p >= g.P) return;
    const int x = (int)(p / g.ny), y = (int)(p - (i64)x * g.ny);
    const i64 n = (i64)t * g.P + p;
    const double ph = phi[n];
    auto put = [&](i64 e, double gr, double next) {
        double v = dadd(dmul(-gr, ph), dmul(gr, next));
        if (weight) v = v / weight[e];
        q[e] = dmul(scale, v);
    };
    if (t < g.nt - 1) put(n, gt, phi[n + g.P]);
    if (x < g.nx - 1) put(g.L + (i64)t * g.PBX + (i64)x * g.ny + y, gx, phi[n + g.ny]);
    if (y < g.ny - 1) put(g.L + g.NBX + (i64)t * g.PBY + (i64)x * (g.ny - 1) + y, gy, phi[n + 1]);
}

// alpha = scale * ((-a) [./ weight]) : a = (BF)^* beta was computed on +beta, (BF)^*(-beta) = -(BF)^* beta exactly
__global__ void __launch_bounds__(256) k_prolong_alpha(i64 n, double scale, const double* __restrict__ weight, double* __restrict__ a)
{
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        double v = -a[i];
        if (weight) v = v / weight[i];
        a[i] = dmul(scale, v);
    }
}

__global__ void __launch_bounds__(256) k_mul_inplace(i64 n, double s, double* __restrict__ x)
{
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) x[i] = dmul(s, x[i]);
}

static unsigned stream_blocks(i64 n)
{
    i64 b = (n + 255) / 256;
    return (unsigned)(b > 148 * 16 ? 148 * 16 : (b < 1 ? 1 : b));
}

int launch_prolong(const Geo& gc, const Geo& gf, const ProlongScal& s, const double* phi_c, const double* beta_c, double* phi_f,
                   double* q_f, double* alpha_f, double* beta_f, const double* weight_f, cudaStream_t st)
{
    const dim3 gn((unsigned)((gf.P + 255) / 256), (unsigned)gf.nt), gcell((unsigned)((gf.P + 255) / 256), (unsigned)(gf.nt - 1));
    k_prolong_phi<<<gn, 256, 0, st>>>(gf, gc, s.phi_recover, phi_c, phi_f);                       // unscaled fine phi
    k_prolong_q<<<gn, 256, 0, st>>>(gf, s.grad_t, s.grad_x, s.grad_y, s.q_scale, phi_f, weight_f, q_f);
    k_mul_inplace<<<stream_blocks(gf.N), 256, 0, st>>>(gf.N, s.phi_scale, phi_f);
    k_prolong_beta<<<gcell, 256, 0, st>>>(gf, gc, s.beta_recover, beta_c, beta_f);                // unscaled fine beta
    launch_bfdconj(gf, 1.0, beta_f, alpha_f, st);                                                // mexBFdConj(alpha, ., 1)
    k_prolong_alpha<<<stream_blocks(gf.Q), 256, 0, st>>>(gf.Q, s.alpha_scale, weight_f, alpha_f);
    k_mul_inplace<<<stream_blocks(10 * gf.L), 256, 0, st>>>(10 * gf.L, s.beta_scale, beta_f);
    return 7;   // launches (launch_bfdconj counts as one)
}

}  // namespace dsocp
