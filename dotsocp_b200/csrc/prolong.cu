// prolong.cu -- level transfer of the multilevel drivers on the device.
//
// Reference: socp/dot2d/utils/interpolate.m:46-84 (phi: linear nodal interpolation along y, then x, then t; beta: nearest
// in t, linear along y then x, column by column), jump_nextLevel.m:5-16 (q = A phi, alpha = (BF)^*(-beta)),
// solver_dotsocp2d.m:368-386 (recoverOrgVar) and :304-365 (InitialScaling).  Every value goes through the same sequence
// of individually rounded operations as the host path (dotsocp_b200/driver.py), so a solve with resident transitions is
// bit-identical to one that downloads, transfers on the host and uploads again.  Compiled with -fmad=false.
//
// Every stage works on a range of fine time levels, so that a time slab fills exactly its own part (plus the ghost layer
// of beta below it, which the slab keeps redundantly) from the coarse slab with the same index: fine level tf reads the
// coarse levels tf/2 and tf/2 + 1, which the coarse slab backs (its own levels plus one ghost level per side) when the
// fine partition is the coarse one with every cut doubled (dotsocp_create_refined).
#include "kernels.h"

namespace dsocp {

__device__ __forceinline__ double avg2(double a, double b) { return dmul(dadd(a, b), 0.5); }

// fine node (tf,xf,yf) <- coarse array (values pre-multiplied by `rec`, the recoverOrgVar factor)
__global__ void __launch_bounds__(256) k_prolong_phi(Geo gf, Geo gc, int t0, double rec, const double* __restrict__ pc,
                                                     double* __restrict__ pf)
{
    const i64 p = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    const int tf = t0 + blockIdx.y;
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
__global__ void __launch_bounds__(256) k_prolong_beta(Geo gf, Geo gc, int t0, double rec, const double* __restrict__ bc,
                                                      double* __restrict__ bf)
{
    const i64 p = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    const int tf = t0 + blockIdx.y;
    const PlanePos pp = plane_pos(gf, p);
    if (!pp.ok) return;
    const int xf = pp.x, yf = pp.y;
    const bool ry = gc.ny > 1;
    const int tc = tf >> 1, xc = xf >> 1, yc = ry ? (yf >> 1) : 0;
    const bool ox = xf & 1, oy = ry && (yf & 1);
#pragma unroll 1
    for (int j = 0; j < 10; j++) {
        const double* src = bc + (i64)j * gc.L + (i64)tc * gc.PC;
        auto V = [&](int x, int y) { return dmul(rec, src[(i64)x * gc.py + y]); };
        auto Y = [&](int x) { return oy ? avg2(V(x, yc), V(x, yc + 1)) : V(x, yc); };
        bf[(i64)j * gf.L + (i64)tf * gf.PC + p] = ox ? avg2(Y(xc), Y(xc + 1)) : Y(xc);
    }
}

// q = scale * ((A phi) [./ weight])  with the UNSCALED forward differences of the fine grid (initialize.m:67-87)
__global__ void __launch_bounds__(256) k_prolong_q(Geo g, int t0, double gt, double gx, double gy, double scale,
                                                   const double* __restrict__ phi, const double* __restrict__ weight,
                                                   double* __restrict__ q)
{
    const i64 p = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    const int t = t0 + blockIdx.y;
    const PlanePos pp = plane_pos(g, p);
    if (!pp.ok) return;
    const int x = pp.x, y = pp.y;
    const i64 n = (i64)t * g.P + pp.pn;
    const double ph = phi[n];
    auto put = [&](i64 e, double gr, double next) {
        double v = dadd(dmul(-gr, ph), dmul(gr, next));
        if (weight) v = v / weight[e];
        q[e] = dmul(scale, v);
    };
    if (t < g.nt - 1) put((i64)t * g.PC + p, gt, phi[n + g.P]);
    if (x < g.nx - 1) put(g.L + (i64)t * g.PBX + p, gx, phi[n + g.ny]);
    if (y < g.ny - 1) put(g.L + g.NBX + (i64)t * g.PBY + (i64)x * g.pyb + y, gy, phi[n + 1]);
}

// alpha = scale * ((-a) [./ weight]) : a = (BF)^* beta was computed on +beta, (BF)^*(-beta) = -(BF)^* beta exactly
__global__ void __launch_bounds__(256) k_prolong_alpha(i64 n, double scale, const double* __restrict__ weight, double* __restrict__ a)
{
    // (pad entries of a pitched layout: a = 0 and weight = 1, see dotsocp_upload -- they stay zero)
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
static dim3 level_grid(const Geo& g, int nlev) { return dim3((unsigned)((g.PC + 255) / 256), (unsigned)nlev); }   // PC >= P

// unscaled fine phi on node levels [t0, t1)
void launch_prolong_phi(const Geo& gc, const Geo& gf, double rec, const double* phi_c, double* phi_f, int t0, int t1, cudaStream_t st)
{
    if (t1 > t0) k_prolong_phi<<<level_grid(gf, t1 - t0), 256, 0, st>>>(gf, gc, t0, rec, phi_c, phi_f);
}
// q on the edges owned by node levels [t0, t1) from the unscaled fine phi (needs phi on level t1 as well)
void launch_prolong_q(const Geo& gf, const ProlongScal& s, const double* phi_f, const double* weight_f, double* q_f, int t0, int t1,
                      cudaStream_t st)
{
    if (t1 > t0) k_prolong_q<<<level_grid(gf, t1 - t0), 256, 0, st>>>(gf, t0, s.grad_t, s.grad_x, s.grad_y, s.q_scale, phi_f, weight_f, q_f);
}
// unscaled fine beta on cell layers [c0, c1)
void launch_prolong_beta(const Geo& gc, const Geo& gf, double rec, const double* beta_c, double* beta_f, int c0, int c1, cudaStream_t st)
{
    if (c1 > c0) k_prolong_beta<<<level_grid(gf, c1 - c0), 256, 0, st>>>(gf, gc, c0, rec, beta_c, beta_f);
}
// x[0..n) *= s  /  alpha[0..n) = scale * ((-alpha) [./ weight])
void launch_mul_inplace(double* x, i64 n, double s, cudaStream_t st)
{
    if (n > 0) k_mul_inplace<<<stream_blocks(n), 256, 0, st>>>(n, s, x);
}
void launch_prolong_alpha(double* alpha, const double* weight, i64 n, double scale, cudaStream_t st)
{
    if (n > 0) k_prolong_alpha<<<stream_blocks(n), 256, 0, st>>>(n, scale, weight, alpha);
}

}  // namespace dsocp
