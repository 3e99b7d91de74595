// recover.cu -- output recovery and checks on the device.
//
// Reference: socp/dot2d/utils/recover_RhoE.m:13-25 (rho = pair average in t of alpha(q0 part), first / last level = rho0 /
// rho1; Ex, Ey = pair averages in x / y of alpha(bx / by part) with zero rows at the two ends, the first and last time
// level counted twice; wdot2d :11 multiplies alpha by the weight first), recover_q.m:12-22 (q0 as is; bx, by centred in
// x / y and then in t), check_massConservation.m:16-34 (mean and negative mean of rho per time level) and the transport
// cost mean(|m|^2 / rho) the parity tests use.  The inputs are the scaled device iterates of a
// finished level; recoverOrgVar (solver_dotsocp2d.m:368-386) is folded in as one product per value, exactly where the
// host path rounds, so the fields are bit-identical to download -> recoverOrgVar -> recover_RhoE on the host.
// One thread per node; every field is produced into one node-indexed scratch array (the Poisson rhs buffer, free after a
// run) and copied out, so the recovery needs no device memory of its own and only 3N (not 2Q + 20L) doubles cross PCIe.
#include "kernels.h"
#include "reduce.cuh"

namespace dsocp {

struct RecIdx {
    int t, x, y;
    i64 p;
};

template <bool WEIGHTED>
struct RecVal {
    const RecoverArgs& a;
    __device__ __forceinline__ double al(i64 e) const   // weight .* ((cScale*D) * alpha)
    {
        const double v = dmul(a.arec, a.alpha[e]);
        return WEIGHTED ? dmul(a.weight[e], v) : v;
    }
    // p = x*ny + y (node plane), pc = x*py + y (cell plane of the pitched staggered arrays)
    __device__ __forceinline__ double rho(int t, i64 p, i64 pc) const
    {
        const Geo& g = a.g;
        if (t == 0) return a.rho0[p];
        if (t == g.nt - 1) return a.rho1[p];
        return dadd(al((i64)(t - 1) * g.PC + pc), al((i64)t * g.PC + pc)) / 2.0;
    }
    // alpha on the bx edge (t, x, y), doubled on the first and last time level (recover_RhoE.m:17-18)
    __device__ __forceinline__ double ex_edge(int t, int x, int y) const
    {
        const Geo& g = a.g;
        const double v = al(g.L + (i64)t * g.PBX + (i64)x * g.py + y);
        return (t == 0 || t == g.nt - 1) ? dmul(2.0, v) : v;
    }
    __device__ __forceinline__ double ey_edge(int t, int x, int y) const
    {
        const Geo& g = a.g;
        const double v = al(g.L + g.NBX + (i64)t * g.PBY + (i64)x * g.pyb + y);
        return (t == 0 || t == g.nt - 1) ? dmul(2.0, v) : v;
    }
    __device__ __forceinline__ double Ex(int t, int x, int y) const
    {
        if (x == 0 || x == a.g.nx - 1) return 0.0;
        return dadd(ex_edge(t, x - 1, y), ex_edge(t, x, y)) / 2.0;
    }
    __device__ __forceinline__ double Ey(int t, int x, int y) const
    {
        if (y == 0 || y == a.g.ny - 1) return 0.0;
        return dadd(ey_edge(t, x, y - 1), ey_edge(t, x, y)) / 2.0;
    }
    // q centred in space on node level t (recover_q.m:15-16, 19-20), q = (dScale/D) * q
    __device__ __forceinline__ double bx_c(int t, int x, int y) const
    {
        const Geo& g = a.g;
        if (x == 0 || x == g.nx - 1) return 0.0;
        const i64 e = g.L + (i64)t * g.PBX + (i64)x * g.py + y;
        return dadd(dmul(a.qrec, a.q[e - g.py]), dmul(a.qrec, a.q[e])) / 2.0;
    }
    __device__ __forceinline__ double by_c(int t, int x, int y) const
    {
        const Geo& g = a.g;
        if (y == 0 || y == g.ny - 1) return 0.0;
        const i64 e = g.L + g.NBX + (i64)t * g.PBY + (i64)x * g.pyb + y;
        return dadd(dmul(a.qrec, a.q[e - 1]), dmul(a.qrec, a.q[e])) / 2.0;
    }
};

template <bool WEIGHTED, int WHICH>
__global__ void __launch_bounds__(256) k_recover(RecoverArgs a)
{
    const Geo& g = a.g;
    const i64 p = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    const int t = a.tr.tn0 + blockIdx.y;
    if (p >= g.P) return;
    const int x = (int)(p / g.ny), y = (int)(p - (i64)x * g.ny);
    const i64 pc = (i64)x * g.py + y;
    const RecVal<WEIGHTED> v{a};
    double r;
    if (WHICH == RC_RHO) r = v.rho(t, p, pc);
    else if (WHICH == RC_EX) r = v.Ex(t, x, y);
    else if (WHICH == RC_EY) r = v.Ey(t, x, y);
    else {
        if (t >= g.nt - 1) return;                       // cell-indexed fields: nt-1 layers
        if (WHICH == RC_Q0) r = dmul(a.qrec, a.q[(i64)t * g.PC + pc]);
        else if (WHICH == RC_BX) r = dadd(v.bx_c(t, x, y), v.bx_c(t + 1, x, y)) / 2.0;
        else r = dadd(v.by_c(t, x, y), v.by_c(t + 1, x, y)) / 2.0;
    }
    a.out[(i64)t * g.P + p] = r;
}

template <bool WEIGHTED, bool ONE_D>
__global__ void __launch_bounds__(256) k_recover_stats(RecoverArgs a)
{
    const Geo& g = a.g;
    const i64 p = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    const int t = a.tr.tn0 + blockIdx.y;
    double s[RS_COUNT];
#pragma unroll
    for (int k = 0; k < RS_COUNT; k++) s[k] = 0.0;
    if (p < g.P) {
        const int x = (int)(p / g.ny), y = (int)(p - (i64)x * g.ny);
        const RecVal<WEIGHTED> v{a};
        const double rho = v.rho(t, p, (i64)x * g.py + y);
        const double ex = v.Ex(t, x, y);
        double m2 = dmul(ex, ex);
        if (!ONE_D) {
            const double ey = v.Ey(t, x, y);
            m2 = dadd(m2, dmul(ey, ey));
        }
        s[RS_SUMRHO] = rho;
        s[RS_SUMNEG] = rho < 0.0 ? rho : 0.0;
        s[RS_W2] = rho > 1e-12 ? m2 / rho : 0.0;
    }
    block_reduce_store<RS_COUNT, 256>(s, a.partial);
}

void launch_recover(const RecoverArgs& a, int which, bool weighted, cudaStream_t st)
{
    const int nl = a.tr.tn1 - a.tr.tn0;
    if (nl <= 0) return;
    dim3 grid((unsigned)((a.g.P + 255) / 256), (unsigned)nl);
#define RC(W)                                                            \
    switch (which) {                                                     \
        case RC_RHO: k_recover<W, RC_RHO><<<grid, 256, 0, st>>>(a); break; \
        case RC_EX: k_recover<W, RC_EX><<<grid, 256, 0, st>>>(a); break;   \
        case RC_EY: k_recover<W, RC_EY><<<grid, 256, 0, st>>>(a); break;   \
        case RC_Q0: k_recover<W, RC_Q0><<<grid, 256, 0, st>>>(a); break;   \
        case RC_BX: k_recover<W, RC_BX><<<grid, 256, 0, st>>>(a); break;   \
        default: k_recover<W, RC_BY><<<grid, 256, 0, st>>>(a); break;      \
    }
    if (weighted) RC(true) else RC(false)
#undef RC
}

void launch_recover_stats(const RecoverArgs& a, bool weighted, bool one_d, cudaStream_t st)
{
    const int nl = a.tr.tn1 - a.tr.tn0;
    if (nl <= 0) return;
    dim3 grid((unsigned)((a.g.P + 255) / 256), (unsigned)nl);
    if (one_d) k_recover_stats<false, true><<<grid, 256, 0, st>>>(a);
    else if (weighted) k_recover_stats<true, false><<<grid, 256, 0, st>>>(a);
    else k_recover_stats<false, false><<<grid, 256, 0, st>>>(a);
    int slots[RS_COUNT];
    for (int k = 0; k < RS_COUNT; k++) slots[k] = k;
    level_reduce(a.partial, (int)grid.x, RS_COUNT, slots, a.tr.tn0, nl, a.lvl, st);
}

}  // namespace dsocp
