// kernels_update.cu -- the per-iteration stencil / projection / multiplier kernels of the DOT-SOCP loop (sm_100a).
//
// Compiled with -fmad=false: every product and sum below is a separately rounded IEEE double operation in the
// same order as the reference's scalar SSE2 MEX kernels and MATLAB expressions, so the cell-local arithmetic is
// bit-identical to the CPU path (only the DCT-based Poisson solve differs in rounding).
//
// Data layout (MATLAB column-major == C order (t,x,y), y fastest):
//   q/alpha/weight/q2 : [q0 (nt-1,nx,ny) | bx (nt,nx-1,ny) | by (nt,nx,ny-1)]
//   beta/z            : 10 planes of L = (nt-1)*nx*ny doubles (structure of arrays)
// All kernels are HBM-bound streaming kernels: warps run along y (coalesced), k_mult marches along t.
#include "kernels.h"
#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace dsocp {

__device__ __forceinline__ double inv_sqrt2_literal() { return __longlong_as_double((long long)DSOCP_INV_SQRT2_BITS); }

// ---------------------------------------------------------------------------------------------------------------
// Projection onto the second-order cone {(t,x): |x| <= t}, column 0 = t.          mexProjSoc.mexa64 @0x1170, @0x1440
// Row-norm association order = Eigen's 2-row packet loop (4-way unrolled):
//    acc = s1 + ((s5+s4)+(s3+s2)) ; acc += ((s9+s8)+(s7+s6))         (s_j = v_j^2)
// ONE_D: the row has 6 columns (c0..c4 and c5 stored in slot 9; slots 5..8 are structural zeros):
//    acc = s1 + ((s9+s4)+(s3+s2))
// r = (v0/nrm + 1)*0.5 ; r>1 -> identity ; 0>r -> 0 ; else scale by r (t = r*nrm unless r == 1).
// ---------------------------------------------------------------------------------------------------------------
template <bool ONE_D>
__device__ __forceinline__ void proj_soc(double (&v)[10])
{
    const double s1 = dmul(v[1], v[1]), s2 = dmul(v[2], v[2]), s3 = dmul(v[3], v[3]), s4 = dmul(v[4], v[4]);
    const double s9 = dmul(v[9], v[9]);
    double acc;
    if (ONE_D) {
        acc = dadd(s1, dadd(dadd(s9, s4), dadd(s3, s2)));
    } else {
        const double s5 = dmul(v[5], v[5]), s6 = dmul(v[6], v[6]), s7 = dmul(v[7], v[7]), s8 = dmul(v[8], v[8]);
        acc = dadd(s1, dadd(dadd(s5, s4), dadd(s3, s2)));
        acc = dadd(acc, dadd(dadd(s9, s8), dadd(s7, s6)));
    }
    const double nrm = sqrt(acc);
    const double r = dmul(dadd(v[0] / nrm, 1.0), 0.5);
    double coef;
    bool keep;
    if (r > 1.0) {
        coef = 1.0;
        keep = true;
    } else if (0.0 > r) {
        coef = 0.0;
        keep = false;
    } else {
        coef = r;
        keep = (r == 1.0);
    }
#pragma unroll
    for (int j = 1; j < 10; j++) v[j] = dmul(v[j], coef);
    v[0] = keep ? v[0] : dmul(coef, nrm);
}

// z2 = d + s BF q of one cell from the 9 staggered values that touch it.        mexBFd.mexa64 @0x1120/@0x11a0/@0x1310
// Out-of-domain neighbours give the structural zeros the reference never writes.
struct CellQ {
    double q0;
    double bxm, bx, bxm1, bx1;   // bx[t,x-1], bx[t,x], bx[t+1,x-1], bx[t+1,x]
    double bym, by, bym1, by1;   // by[t,y-1], by[t,y], by[t+1,y-1], by[t+1,y]
};
__device__ __forceinline__ void cell_z2(const CellQ& c, const IterScal& sc, bool hxm, bool hxp, bool hym, bool hyp,
                                        double (&z2)[10])
{
    const double p = dmul(c.q0, sc.S);
    z2[0] = dsub(sc.DF, p);
    z2[9] = dadd(p, sc.DF);
    z2[1] = hxm ? dmul(c.bxm, sc.SF) : 0.0;
    z2[2] = hxp ? dmul(c.bx, sc.SF) : 0.0;
    z2[3] = hxm ? dmul(c.bxm1, sc.SF) : 0.0;
    z2[4] = hxp ? dmul(c.bx1, sc.SF) : 0.0;
    z2[5] = hym ? dmul(c.bym, sc.SF) : 0.0;
    z2[6] = hyp ? dmul(c.by, sc.SF) : 0.0;
    z2[7] = hym ? dmul(c.bym1, sc.SF) : 0.0;
    z2[8] = hyp ? dmul(c.by1, sc.SF) : 0.0;
}

// z of a cell: either materialised (zmat != NULL: PALM / acc-ADMM keep z as state) or recomputed from the inputs
// of the z-step that produced it, z = Pi_Q(d + BF q_old - beta_old) (inPALM never stores z).
template <bool ONE_D>
__device__ __forceinline__ void load_z(const Geo& g, const IterScal& sc, const double* __restrict__ zmat,
                                       const double* __restrict__ q_old, const double* __restrict__ beta_old, int t, int x,
                                       int y, double (&z)[10])
{
    const i64 L = g.L, c = (i64)t * g.P + (i64)x * g.ny + y;
    if (zmat != nullptr) {
#pragma unroll
        for (int j = 0; j < 10; j++) z[j] = (ONE_D && j >= 5 && j <= 8) ? 0.0 : zmat[(i64)j * L + c];
        return;
    }
    const bool hxm = x > 0, hxp = x < g.nx - 1, hym = y > 0, hyp = y < g.ny - 1;
    const double* bx = q_old + L;
    const double* by = bx + g.NBX;
    const i64 ox = (i64)t * g.PBX + (i64)x * g.ny + y, oy = (i64)t * g.PBY + (i64)x * (g.ny - 1) + y;
    CellQ cq;
    cq.q0 = q_old[c];
    cq.bxm = hxm ? bx[ox - g.ny] : 0.0;
    cq.bx = hxp ? bx[ox] : 0.0;
    cq.bxm1 = hxm ? bx[ox + g.PBX - g.ny] : 0.0;
    cq.bx1 = hxp ? bx[ox + g.PBX] : 0.0;
    cq.bym = hym ? by[oy - 1] : 0.0;
    cq.by = hyp ? by[oy] : 0.0;
    cq.bym1 = hym ? by[oy + g.PBY - 1] : 0.0;
    cq.by1 = hyp ? by[oy + g.PBY] : 0.0;
    cell_z2(cq, sc, hxm, hxp, hym, hyp, z);
#pragma unroll
    for (int j = 0; j < 10; j++) z[j] = dsub(z[j], (ONE_D && j >= 5 && j <= 8) ? 0.0 : beta_old[(i64)j * L + c]);
    proj_soc<ONE_D>(z);
}

// ---------------------------------------------------------------------------------------------------------------
// Standalone mexBFd: one thread per cell, boundary entries NOT written (reference semantics).
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_bfd(Geo g, double S, double SF, double DF, const double* __restrict__ q,
                                             double* __restrict__ z)
{
    const i64 p = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    const int t = blockIdx.y;
    if (p >= g.P) return;
    const int x = (int)(p / g.ny), y = (int)(p - (i64)x * g.ny);
    const i64 c = (i64)t * g.P + p;
    const double* bx = q + g.L;
    const double* by = bx + g.NBX;
    const double pr = dmul(q[c], S);
    z[c] = dsub(DF, pr);
    z[9 * g.L + c] = dadd(pr, DF);
    if (x >= 1) {
        z[1 * g.L + c] = dmul(bx[(i64)t * g.PBX + (i64)(x - 1) * g.ny + y], SF);
        z[3 * g.L + c] = dmul(bx[(i64)(t + 1) * g.PBX + (i64)(x - 1) * g.ny + y], SF);
    }
    if (x <= g.nx - 2) {
        z[2 * g.L + c] = dmul(bx[(i64)t * g.PBX + (i64)x * g.ny + y], SF);
        z[4 * g.L + c] = dmul(bx[(i64)(t + 1) * g.PBX + (i64)x * g.ny + y], SF);
    }
    if (y >= 1) {
        z[5 * g.L + c] = dmul(by[(i64)t * g.PBY + (i64)x * (g.ny - 1) + (y - 1)], SF);
        z[7 * g.L + c] = dmul(by[(i64)(t + 1) * g.PBY + (i64)x * (g.ny - 1) + (y - 1)], SF);
    }
    if (y <= g.ny - 2) {
        z[6 * g.L + c] = dmul(by[(i64)t * g.PBY + (i64)x * (g.ny - 1) + y], SF);
        z[8 * g.L + c] = dmul(by[(i64)(t + 1) * g.PBY + (i64)x * (g.ny - 1) + y], SF);
    }
}

static double host_sf(double S)
{
    uint64_t b = DSOCP_INV_SQRT2_BITS;
    double d;
    memcpy(&d, &b, 8);
    return d * S;
}

void launch_bfd(const Geo& g, double S, double DF, const double* q, double* z, cudaStream_t st)
{
    dim3 grid((unsigned)((g.P + 255) / 256), (unsigned)(g.nt - 1));
    k_bfd<<<grid, 256, 0, st>>>(g, S, host_sf(S), DF, q, z);
}

// ---------------------------------------------------------------------------------------------------------------
// Standalone mexBFdConj: one thread per node; q0, bx, by entries owned by the node.   mexBFdConj.mexa64 @0x1120/1160/1310
// add order ((z1[t,x+1] + z2[t,x]) + z3[t-1,x+1]) + z4[t-1,x]
// ---------------------------------------------------------------------------------------------------------------
template <bool ADD2>
__global__ void __launch_bounds__(256) k_bfdconj(Geo g, int tn0, double S, double SF, const double* __restrict__ za,
                                                 const double* __restrict__ zb, double* __restrict__ q2)
{
    // ADD2: the argument is the elementwise sum za + zb (mexBFdConj(q2, z + beta, ...), solver_socp_accADMM.m:229)
    auto Z = [&](i64 idx) -> double { return ADD2 ? dadd(za[idx], zb[idx]) : za[idx]; };
    const i64 p = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    const int t = tn0 + blockIdx.y;
    if (p >= g.P) return;
    const int x = (int)(p / g.ny), y = (int)(p - (i64)x * g.ny);
    const i64 L = g.L;
    const i64 cu = (i64)t * g.P + p;         // cell (t,x,y)
    const i64 cd = cu - g.P;                 // cell (t-1,x,y)
    const bool up = t < g.nt - 1, dn = t > 0;
    if (up) q2[cu] = dmul(dsub(Z(9 * L + cu), Z(cu)), S);
    if (x < g.nx - 1) {
        double s;
        if (!dn)
            s = dadd(Z(1 * L + cu + g.ny), Z(2 * L + cu));
        else if (!up)
            s = dadd(Z(3 * L + cd + g.ny), Z(4 * L + cd));
        else {
            s = dadd(Z(1 * L + cu + g.ny), Z(2 * L + cu));
            s = dadd(s, Z(3 * L + cd + g.ny));
            s = dadd(s, Z(4 * L + cd));
        }
        q2[L + (i64)t * g.PBX + (i64)x * g.ny + y] = dmul(s, SF);
    }
    if (y < g.ny - 1) {
        double s;
        if (!dn)
            s = dadd(Z(5 * L + cu + 1), Z(6 * L + cu));
        else if (!up)
            s = dadd(Z(7 * L + cd + 1), Z(8 * L + cd));
        else {
            s = dadd(Z(5 * L + cu + 1), Z(6 * L + cu));
            s = dadd(s, Z(7 * L + cd + 1));
            s = dadd(s, Z(8 * L + cd));
        }
        q2[L + g.NBX + (i64)t * g.PBY + (i64)x * (g.ny - 1) + y] = dmul(s, SF);
    }
}

void launch_bfdconj(const Geo& g, double S, const double* z, double* q2, cudaStream_t st, const TRange* tr)
{
    const int tn0 = tr ? tr->tn0 : 0, tn1 = tr ? tr->tn1 : g.nt;
    dim3 grid((unsigned)((g.P + 255) / 256), (unsigned)(tn1 - tn0));
    k_bfdconj<false><<<grid, 256, 0, st>>>(g, tn0, S, host_sf(S), z, nullptr, q2);
}
void launch_bfdconj_sum(const Geo& g, double S, const double* za, const double* zb, double* q2, cudaStream_t st)
{
    dim3 grid((unsigned)((g.P + 255) / 256), (unsigned)g.nt);
    k_bfdconj<true><<<grid, 256, 0, st>>>(g, 0, S, host_sf(S), za, zb, q2);
}

// ---------------------------------------------------------------------------------------------------------------
// Standalone mexProjSoc for an arbitrary M x N column-major matrix (generic column count; the fused kernels use the
// unrolled proj_soc<> above).  Reproduces Eigen's packet order for paired rows and the sequential order of the
// odd tail row (mexProjSoc.mexa64 @0x1530 / @0x1668).
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_projsoc(i64 M, int N, const double* __restrict__ in, double* __restrict__ out)
{
    const i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    if (i >= M) return;
    const int n = N - 1;
    double nrm = 0.0;
    if (n > 0) {
        const bool packet = i < (M & ~(i64)1);
        double v = in[1 * M + i];
        double acc = dmul(v, v);
        int k = 1;
        if (packet) {
            const int kend = (n - 1) & ~3;
            if (kend > 1) {
                for (; k < kend; k += 4) {
                    const double a0 = in[(i64)(k + 1) * M + i], a1 = in[(i64)(k + 2) * M + i];
                    const double a2 = in[(i64)(k + 3) * M + i], a3 = in[(i64)(k + 4) * M + i];
                    const double hi = dadd(dmul(a3, a3), dmul(a2, a2));
                    const double lo = dadd(dmul(a1, a1), dmul(a0, a0));
                    acc = dadd(acc, dadd(hi, lo));
                }
            }
        }
        for (; k < n; k++) {
            const double a = in[(i64)(k + 1) * M + i];
            acc = dadd(acc, dmul(a, a));
        }
        nrm = sqrt(acc);
    }
    const double v0 = in[i];
    const double r = dmul(dadd(v0 / nrm, 1.0), 0.5);
    double coef;
    bool keep;
    if (r > 1.0) {
        coef = 1.0;
        keep = true;
    } else if (0.0 > r) {
        coef = 0.0;
        keep = false;
    } else {
        coef = r;
        keep = (r == 1.0);
    }
    for (int j = 1; j < N; j++) out[(i64)j * M + i] = dmul(in[(i64)j * M + i], coef);
    out[i] = keep ? v0 : dmul(coef, nrm);
}

void launch_projsoc(i64 M, int N, const double* in, double* out, cudaStream_t st)
{
    if (M <= 0) return;
    k_projsoc<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(M, N, in, out);
}

// ---------------------------------------------------------------------------------------------------------------
// q-step + alpha-step, one thread per node (t,x,y); the node owns q0[t,x,y], bx[t,x,y], by[t,x,y].
//   tmp_q = A*phi (CSR row: (-g)*phi_i + g*phi_{i+1})                                   solver_socp_inPALM.m:204
//   q     = ((tmp_q + alpha) + q2) .* diagQInv          | weighted: (w.*(tmp_q+alpha) + q2) .* diagQInv   :206 / wsocp :212
//   alpha = alpha + tau*(tmp_q - w.*q)                  | acc-ADMM: (alpha + tmp_q) - w.*q                :211,214 / accADMM :237
// ---------------------------------------------------------------------------------------------------------------
template <bool WEIGHTED, bool ACC>
__device__ __forceinline__ void q_update(i64 e, double aphi, double dinv_plain, double s2term, const IterScal& sc,
                                         const double* __restrict__ q2, const double* __restrict__ weight,
                                         double* __restrict__ alpha, double* __restrict__ qout,
                                         const double* __restrict__ tmpq_in, double* __restrict__ tmpq_out, bool upd_alpha)
{
    if (tmpq_in != nullptr) aphi = tmpq_in[e];      // PALM's first q-step re-uses the stored A*phi (solver_socp_PALM.m:198-199)
    if (tmpq_out != nullptr) tmpq_out[e] = aphi;
    const double a = alpha[e];
    const double q2v = q2[e];
    double qn, wq;
    if (WEIGHTED) {
        const double w = weight[e];
        const double dinv = 1.0 / dadd(s2term, dmul(w, w));
        qn = dmul(dadd(dmul(w, dadd(aphi, a)), q2v), dinv);
        wq = dmul(w, qn);
    } else {
        qn = dmul(dadd(dadd(aphi, a), q2v), dinv_plain);
        wq = qn;
    }
    qout[e] = qn;
    if (!upd_alpha) return;
    if (ACC)
        alpha[e] = dsub(dadd(a, aphi), wq);
    else
        alpha[e] = dadd(a, dmul(sc.tau, dsub(aphi, wq)));
}

template <bool WEIGHTED, bool ACC>
__global__ void __launch_bounds__(256) k_qstep(Geo g, int tn0, IterScal sc, const double* __restrict__ phi,
                                               const double* __restrict__ q2, const double* __restrict__ weight,
                                               double* __restrict__ alpha, double* __restrict__ qout,
                                               const double* __restrict__ tmpq_in, double* __restrict__ tmpq_out,
                                               bool upd_alpha)
{
    const i64 p = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    const int t = tn0 + blockIdx.y;
    if (p >= g.P) return;
    const int x = (int)(p / g.ny), y = (int)(p - (i64)x * g.ny);
    const i64 n = (i64)t * g.P + p;
    const double ph = phi[n];
    const bool edge_t = (t == 0) || (t == g.nt - 1);
    if (t < g.nt - 1) {
        const double aphi = dadd(dmul(-sc.gt, ph), dmul(sc.gt, phi[n + g.P]));
        q_update<WEIGHTED, ACC>(n, aphi, sc.dinv1, sc.s2x2, sc, q2, weight, alpha, qout, tmpq_in, tmpq_out, upd_alpha);
    }
    if (x < g.nx - 1) {
        const double aphi = dadd(dmul(-sc.gx, ph), dmul(sc.gx, phi[n + g.ny]));
        q_update<WEIGHTED, ACC>(g.L + (i64)t * g.PBX + (i64)x * g.ny + y, aphi, edge_t ? sc.dinv2 : sc.dinv1,
                                edge_t ? sc.s2x1 : sc.s2x2, sc, q2, weight, alpha, qout, tmpq_in, tmpq_out, upd_alpha);
    }
    if (y < g.ny - 1) {
        const double aphi = dadd(dmul(-sc.gy, ph), dmul(sc.gy, phi[n + 1]));
        q_update<WEIGHTED, ACC>(g.L + g.NBX + (i64)t * g.PBY + (i64)x * (g.ny - 1) + y, aphi,
                                edge_t ? sc.dinv2 : sc.dinv1, edge_t ? sc.s2x1 : sc.s2x2, sc, q2, weight, alpha, qout, tmpq_in,
                                tmpq_out, upd_alpha);
    }
}

void launch_qstep(const UpdateArgs& a, bool weighted, bool acc, cudaStream_t st, const double* tmpq_in, double* tmpq_out,
                  bool upd_alpha)
{
    dim3 grid((unsigned)((a.g.P + 255) / 256), (unsigned)(a.tr.tn1 - a.tr.tn0));
#define QS(W, A) \
    k_qstep<W, A><<<grid, 256, 0, st>>>(a.g, a.tr.tn0, a.sc, a.phi, a.q2, a.weight, a.alpha, a.q_new, tmpq_in, tmpq_out, upd_alpha)
    if (weighted) {
        if (acc) QS(true, true); else QS(true, false);
    } else {
        if (acc) QS(false, true); else QS(false, false);
    }
#undef QS
}

// ---------------------------------------------------------------------------------------------------------------
// k_mult: fused z-step + beta-step of iteration i, then the z-step inputs of iteration i+1.
//
// Each CTA owns a (TX-1) x (TY-1) tile of (x,y) columns (+1 halo row/column on the high side, recomputed) and
// marches along t.  Per cell and step:
//     z2o = d + BF q_old ; z = Pi_Q(z2o - beta)                   (mexBFd + mexProjSoc,   solver_socp_inPALM.m:199)
//     z2n = d + BF q_new ; beta += tau (z - z2n)                  (mexBFd,                :212-215)
//     w   = Pi_Q(z2n - beta) + beta                                (next iteration's z + beta, :199,:205)
//     q2  = s (BF)^* w                                             (mexBFdConj,            :205)  -> consumed by k_qstep
//     rhs = A'(w.*q_new - alpha) + c                               (next Poisson rhs,      :194)
// so beta (the 10L array that dominates the traffic) is read once and written once per iteration and z, z2, q2's
// 10-column temporaries never touch HBM.  The x+1 / y+1 neighbours' w (columns 1,3 / 5,7) come through shared
// memory; the t-1 layer's columns 3,4,7,8 are carried in registers.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}

template <bool WEIGHTED>
__device__ __forceinline__ double uval(double q, double a, double w)
{
    return WEIGHTED ? dsub(dmul(w, q), a) : dsub(q, a);
}

#ifndef KM_PAIRSYNC
#define KM_PAIRSYNC 0     // 1: no CTA-wide barrier in the time loop: a warp (= one x row of the tile) only waits for the row
                          //    above it (full/empty mbarrier pair per row), y neighbours are exchanged by warp shuffles.
                          //    Bit-exact; measured 5.14 ms against 5.01 ms with __syncthreads (512x512x256): the barrier is
                          //    not what limits the kernel.
#endif
#ifndef KM_TX
#define KM_TX 8          // tile rows (x) per CTA; 16 (one 512-thread CTA per SM) measured in profiles/README.md
#endif
#ifndef KM_TY
#define KM_TY 32         // tile columns (y, contiguous) per CTA
#endif
#ifndef KM_MIN_BLOCKS
#define KM_MIN_BLOCKS (KM_TX * KM_TY > 256 ? 1 : 2)
#endif
#ifndef KM_BULK
#define KM_BULK 0         // 1: interior CTAs stage every step's input rows with cp.async.bulk (TMA engine) two steps ahead.
                          // Bit-exact and fewer instructions per step, but measured SLOWER than the plain loads on B200
                          // (512x512x256: 5.67 vs 5.01 ms, 1024x1024x512: 48.0 vs 37.9 ms); so is KM_L2PF (5.52 / 49.2 ms).
                          // Deeper look-ahead loses more than the hidden latency wins; kept for experiments.
#endif
constexpr int KM_ROWD = 36;   // doubles per staged row: 32 (+1 for the y-1 neighbour) needed, + alignment slack, 288 B = 18 x 16 B
__host__ __device__ constexpr int km_nrows(int TX) { return 16 * TX + 3 * (TX + 1); }

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "KM_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@!p bra KM_WAIT_%=;\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
// 16-byte aligned global -> shared bulk copy, completion counted in bytes on the mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ unsigned ptr_par(const void* p) { return (unsigned)((unsigned long long)p >> 3) & 1u; }

#ifndef KM_L2PF
#define KM_L2PF 0         // 1: prefetch.global.L2 of the next step's lines
#endif
#ifndef KM_PREFETCH
#define KM_PREFETCH 0     // 1: cp.async ring for the next step's loads (measured: no gain, the kernel is issue/latency bound)
#endif
template <int TX, int TY, bool WEIGHTED, bool ONE_D, bool UPDATE, bool EDGE>
__device__ __forceinline__ void k_mult_body(const Geo& g, const TRange& tr, const IterScal& sc, const double* __restrict__ qo,
                                            const double* __restrict__ qn, const double* __restrict__ alpha,
                                            const double* __restrict__ weight, const double* __restrict__ beta,
                                            double* __restrict__ beta_out, double* __restrict__ q2,
                                            double* __restrict__ rhs, const double* __restrict__ c0,
                                            const double* __restrict__ c1)
{
    // dynamic shared memory: [2][4][NT] exchange of the w columns 1,3,5,7 + a two-stage ring [2][NV][NT] into which every
    // thread prefetches (cp.async, 8 B) the 21 values of its NEXT time step while it works on the current one, so the
    // HBM latency of the beta / q streams is overlapped with the two projections instead of being exposed once per step
    constexpr int NT = TX * TY, NV = 21;
    extern __shared__ __align__(16) double dyn_smem[];
    double (*sh)[4][TX][TY] = reinterpret_cast<double (*)[4][TX][TY]>(dyn_smem);
    double* ring = dyn_smem + 2 * 4 * NT;
    const int ly = threadIdx.x, lx = threadIdx.y;
    const int tid = lx * TY + ly;
    // y tiles vary fastest over the grid so that CTAs running side by side stream adjacent pieces of the same rows
    const int x = blockIdx.y * (TX - 1) + lx, y = blockIdx.x * (TY - 1) + ly;
    // EDGE = false: the whole tile (halo included) lies strictly inside the domain, every neighbour exists and all the
    // boundary predicates fold away at compile time (most CTAs of a large grid)
    const bool valid = EDGE ? ((x < g.nx) && (y < g.ny)) : true;
    const bool owner = valid && (lx < TX - 1) && (ly < TY - 1);
    const bool hxm = EDGE ? (valid && x > 0) : true, hxp = EDGE ? (valid && x < g.nx - 1) : true;
    const bool hym = EDGE ? (valid && y > 0) : true, hyp = EDGE ? (valid && y < g.ny - 1) : true;
    const i64 L = g.L;
    const i64 node = (i64)x * g.ny + y;
    const i64 ibx = (i64)x * g.ny + y, ibxm = ibx - g.ny;
    const i64 iby = (i64)x * (g.ny - 1) + y, ibym = iby - 1;
    const double* __restrict__ qo_bx = qo + L;
    const double* __restrict__ qo_by = qo_bx + g.NBX;
    const double* __restrict__ qn_bx = qn + L;
    const double* __restrict__ qn_by = qn_bx + g.NBX;
    const double* __restrict__ al_bx = alpha + L;
    const double* __restrict__ al_by = al_bx + g.NBX;
    const double* __restrict__ w_bx = WEIGHTED ? weight + L : nullptr;
    const double* __restrict__ w_by = WEIGHTED ? w_bx + g.NBX : nullptr;
    double* __restrict__ q2_bx = q2 + L;
    double* __restrict__ q2_by = q2_bx + g.NBX;

    // staggered q at node level t (carried) -- old and new iterate
    CellQ co, cn;
    co.q0 = cn.q0 = 0.0;
    co.bxm = co.bx = co.bym = co.by = co.bxm1 = co.bx1 = co.bym1 = co.by1 = 0.0;
    cn = co;
    // time slab: march over the owned node levels [tn0, tn1); a slab that does not start at t = 0 first replays the
    // cell layer below it (ghost layer, kept redundantly by both neighbours) to obtain the carried t-1 quantities
    const int t_start = tr.tc0 > 0 ? tr.tc0 - 1 : 0;
    {
        const i64 s0x = (i64)t_start * g.PBX, s0y = (i64)t_start * g.PBY;
        if (UPDATE) {
            if (hxm) co.bxm = qo_bx[s0x + ibxm];
            if (hxp) co.bx = qo_bx[s0x + ibx];
            if (hym) co.bym = qo_by[s0y + ibym];
            if (hyp) co.by = qo_by[s0y + iby];
        }
        if (hxm) cn.bxm = qn_bx[s0x + ibxm];
        if (hxp) cn.bx = qn_bx[s0x + ibx];
        if (hym) cn.bym = qn_by[s0y + ibym];
        if (hyp) cn.by = qn_by[s0y + iby];
    }

    double wp3n = 0.0, wp4 = 0.0, wp7n = 0.0, wp8 = 0.0, u0p = 0.0;

    // ---- bulk-copy ring (interior CTAs) ---------------------------------------------------------------------------------
    // Every input of a time step is a set of rows of 32 (33) consecutive doubles.  Thread r < NROWS owns row r of the
    // stage: slots 0..9 beta planes, 10 q0 new, 11 q0 old, 12 alpha0 (TX rows each, cell level t), then bx new / bx old at
    // level t+1 (TX+1 rows from x0-1), by new / by old at level t+1 (TX rows, from y0-1), alpha_bx (TX+1 rows) and alpha_by
    // (TX rows) at level t.  A row is fetched as 288 bytes from the 16-byte aligned address at or below its first element
    // (the grids have odd lengths, so rows start on odd multiples of 8 bytes half of the time); readers add the parity of
    // the row's first address to their index.  Step t+2 is issued right after the barrier of step t: two stages.
    constexpr bool BULK = KM_BULK && !EDGE && !KM_PREFETCH;
    constexpr int NROWS = km_nrows(TX), STAGE_D = NROWS * KM_ROWD;
    constexpr int R_QN0 = 10 * TX, R_QO0 = 11 * TX, R_A0 = 12 * TX, R_QNBX = 13 * TX, R_QOBX = R_QNBX + TX + 1,
                  R_QNBY = R_QOBX + TX + 1, R_QOBY = R_QNBY + TX, R_ALBX = R_QOBY + TX, R_ALBY = R_ALBX + TX + 1;
    double* bring = dyn_smem + 2 * 4 * NT;
    unsigned long long* mbar = reinterpret_cast<unsigned long long*>(bring + 2 * STAGE_D);
    const char* p_src = nullptr;      // producer: first byte of my row at the first step
    i64 p_stride = 0;                 //           bytes per time step
    bool p_cell = true, p_on = false; //           row exists only while t is a cell level / row is used at all
    if (BULK) {
        const int x0 = blockIdx.y * (TX - 1), y0 = blockIdx.x * (TY - 1);
        if (tid < NROWS) {
            const int r = tid;
            const double* base;
            i64 idx, str;
            p_on = true;
            if (r < R_QNBX) {          // cell-indexed planes
                const int slot = r / TX, row = r - slot * TX;
                base = slot < 10 ? beta + (i64)slot * L : slot == 10 ? qn : slot == 11 ? qo : alpha;
                if (slot == 11 && !UPDATE) p_on = false;
                idx = (i64)t_start * g.P + (i64)(x0 + row) * g.ny + y0;
                str = g.P;
            } else if (r < R_QNBY) {   // bx at level t+1, rows x0-1 ..
                const bool old = r >= R_QOBX;
                const int row = r - (old ? R_QOBX : R_QNBX);
                base = old ? qo_bx : qn_bx;
                if (old && !UPDATE) p_on = false;
                idx = (i64)(t_start + 1) * g.PBX + (i64)(x0 - 1 + row) * g.ny + y0;
                str = g.PBX;
            } else if (r < R_ALBX) {   // by at level t+1, from y0-1
                const bool old = r >= R_QOBY;
                const int row = r - (old ? R_QOBY : R_QNBY);
                base = old ? qo_by : qn_by;
                if (old && !UPDATE) p_on = false;
                idx = (i64)(t_start + 1) * g.PBY + (i64)(x0 + row) * (g.ny - 1) + y0 - 1;
                str = g.PBY;
            } else if (r < R_ALBY) {   // alpha_bx at level t
                base = al_bx;
                idx = (i64)t_start * g.PBX + (i64)(x0 - 1 + (r - R_ALBX)) * g.ny + y0;
                str = g.PBX;
                p_cell = false;
            } else {                   // alpha_by at level t
                base = al_by;
                idx = (i64)t_start * g.PBY + (i64)(x0 + (r - R_ALBY)) * (g.ny - 1) + y0 - 1;
                str = g.PBY;
                p_cell = false;
            }
            p_src = reinterpret_cast<const char*>(base + idx);
            p_stride = str * (i64)sizeof(double);
        }
        if (tid == 0) {
            mbar_init(&mbar[0], 1);
            mbar_init(&mbar[1], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
    }
    constexpr bool PAIR = KM_PAIRSYNC && !BULK && !KM_PREFETCH;
    static_assert(!PAIR || TY == 32, "one warp per tile row");
    __shared__ unsigned long long xbar[2][2][TX];   // [full | empty][buffer][row]
    if (PAIR) {
        if (tid < 4 * TX) mbar_init(&xbar[0][0][0] + tid, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        __syncthreads();
    }
    auto fill = [&](int tt) {   // issue the rows of step tt into stage (tt - t_start) & 1
        if (!BULK) return;
        const int k = tt - t_start;
        const bool cellt = tt < g.nt - 1;
        unsigned long long* bar = &mbar[k & 1];
        if (tid == 0) {
            const int rows = cellt ? (NROWS - (UPDATE ? 0 : 3 * TX + 1)) : (2 * TX + 1);
            mbar_expect_tx(bar, (unsigned)(rows * KM_ROWD * sizeof(double)));
        }
        if (p_on && (cellt || !p_cell)) {
            const unsigned long long a = (unsigned long long)(p_src + (i64)k * p_stride) & ~15ull;
            bulk_g2s(bring + (size_t)(k & 1) * STAGE_D + (size_t)tid * KM_ROWD, reinterpret_cast<const void*>(a),
                     (unsigned)(KM_ROWD * sizeof(double)), bar);
        }
    };
    // parity of the first address of my rows (low word arithmetic is enough), advanced every step
    unsigned cw = 0, bw = 0, yw = 0;
    unsigned par_c = 0, par_bxn = 0, par_bxo = 0, par_byn = 0, par_byo = 0, par_albx = 0, par_alby = 0;
    if (BULK) {
        const int y0 = blockIdx.x * (TY - 1);
        cw = (unsigned)((i64)t_start * g.P + (i64)x * g.ny + y0);                       // cell planes, row x
        bw = (unsigned)((i64)(t_start + 1) * g.PBX + (i64)(x - 1) * g.ny + y0);         // bx planes at t+1, row x-1
        yw = (unsigned)((i64)(t_start + 1) * g.PBY + (i64)x * (g.ny - 1) + y0 - 1);     // by planes at t+1, row x (from y0-1)
        par_c = ptr_par(beta) | (ptr_par(qn) << 1) | (ptr_par(qo) << 2) | (ptr_par(alpha) << 3) | ((unsigned)(L & 1) << 4);
        par_bxn = ptr_par(qn_bx); par_bxo = ptr_par(qo_bx); par_byn = ptr_par(qn_by); par_byo = ptr_par(qo_by);
        par_albx = ptr_par(al_bx) ^ (unsigned)(g.PBX & 1); par_alby = ptr_par(al_by) ^ (unsigned)(g.PBY & 1);
        fill(t_start);
        if (t_start + 1 < tr.tn1) fill(t_start + 1);
    }

    // ring slots: 0..9 beta, 10 q0 new, 11 q0 old, 12 alpha0, 13..16 new bx/by at level t+1, 17..20 old bx/by at t+1
    auto prefetch = [&](int tt) {
        if (!KM_PREFETCH) return;
        if (tt < g.nt - 1 && valid) {
            double* dst = ring + (size_t)(tt & 1) * NV * NT + tid;
            const i64 cc = (i64)tt * g.P + node;
#pragma unroll
            for (int j = 0; j < 10; j++)
                if (!(ONE_D && j >= 5 && j <= 8)) cp_async8(dst + j * NT, beta + (i64)j * L + cc);
            cp_async8(dst + 10 * NT, qn + cc);
            if (UPDATE) cp_async8(dst + 11 * NT, qo + cc);
            cp_async8(dst + 12 * NT, alpha + cc);
            const i64 o1x = (i64)(tt + 1) * g.PBX, o1y = (i64)(tt + 1) * g.PBY;
            if (hxm) cp_async8(dst + 13 * NT, qn_bx + o1x + ibxm);
            if (hxp) cp_async8(dst + 14 * NT, qn_bx + o1x + ibx);
            if (hym) cp_async8(dst + 15 * NT, qn_by + o1y + ibym);
            if (hyp) cp_async8(dst + 16 * NT, qn_by + o1y + iby);
            if (UPDATE) {
                if (hxm) cp_async8(dst + 17 * NT, qo_bx + o1x + ibxm);
                if (hxp) cp_async8(dst + 18 * NT, qo_bx + o1x + ibx);
                if (hym) cp_async8(dst + 19 * NT, qo_by + o1y + ibym);
                if (hyp) cp_async8(dst + 20 * NT, qo_by + o1y + iby);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    prefetch(t_start);
    // KM_L2PF: no ring, but ask the L2 for the next step's lines one step ahead (prefetch.global.L2 needs no registers and
    // no shared memory), so that the loads of the next step find their data on chip
    auto l2pf = [&](const double* p) {
#if KM_L2PF
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#endif
    };

    for (int t = t_start; t < tr.tn1; t++) {
        const int buf = t & 1;
        const bool cell = t < g.nt - 1;
        const bool emit = t >= tr.tn0;      // false only on the replayed ghost layer
        const i64 cidx = (i64)t * g.P + node;
        double w[10];
        double a0 = 0.0, wt0 = 1.0;
#pragma unroll
        for (int j = 0; j < 10; j++) w[j] = 0.0;
        if (t + 1 < tr.tn1) prefetch(t + 1);
        // bulk ring: wait for this step's stage, then all reads below are shared-memory loads at fixed offsets
        const int kst = t - t_start;
        const double* S = bring + (size_t)(kst & 1) * STAGE_D + ly;
        const unsigned nyp = (unsigned)(g.ny & 1);
        const unsigned pc = cw & 1u, pbx = bw & 1u, pby = yw & 1u;
        if (BULK) mbar_wait(&mbar[kst & 1], (unsigned)(kst >> 1) & 1u);
#define RB(rowbase, row, off) S[((rowbase) + (row)) * KM_ROWD + (int)(off)]
        if (KM_L2PF && valid && t + 1 < g.nt - 1 && t + 1 < tr.tn1) {
            const i64 cn1 = cidx + g.P;
#pragma unroll
            for (int j = 0; j < 10; j++)
                if (!(ONE_D && j >= 5 && j <= 8)) l2pf(beta + (i64)j * L + cn1);
            l2pf(qn + cn1);
            l2pf(alpha + cn1);
            if (UPDATE) l2pf(qo + cn1);
            const i64 o2x = (i64)(t + 2) * g.PBX + ibx, o2y = (i64)(t + 2) * g.PBY + iby;
            if (hxp) { l2pf(qn_bx + o2x); l2pf(al_bx + o2x - g.PBX); if (UPDATE) l2pf(qo_bx + o2x); }
            if (hyp) { l2pf(qn_by + o2y); l2pf(al_by + o2y - g.PBY); if (UPDATE) l2pf(qo_by + o2y); }
        }
        // the level-t alpha (and weight) values of the rhs stencil are only needed after the barrier: issue them now so
        // that their latency hides behind the two projections
        double al_xm = 0.0, al_x = 0.0, al_ym = 0.0, al_y = 0.0, wt_xm = 1.0, wt_x = 1.0, wt_ym = 1.0, wt_y = 1.0;
        if (owner) {
            const i64 ox = (i64)t * g.PBX, oy = (i64)t * g.PBY;
            if (BULK) {
                al_xm = RB(R_ALBX, lx, pbx ^ par_albx);
                al_x = RB(R_ALBX, lx + 1, pbx ^ nyp ^ par_albx);
                al_ym = RB(R_ALBY, lx, pby ^ par_alby);
                al_y = RB(R_ALBY, lx, 1 + (pby ^ par_alby));
            } else {
                if (hxm) al_xm = al_bx[ox + ibxm];
                if (hxp) al_x = al_bx[ox + ibx];
                if (hym) al_ym = al_by[oy + ibym];
                if (hyp) al_y = al_by[oy + iby];
            }
            if (WEIGHTED) {
                if (hxm) wt_xm = w_bx[ox + ibxm];
                if (hxp) wt_x = w_bx[ox + ibx];
                if (hym) wt_ym = w_by[oy + ibym];
                if (hyp) wt_y = w_by[oy + iby];
                if (cell) wt0 = weight[cidx];
            }
        }
        if (KM_PREFETCH) {
            if (t + 1 < tr.tn1) asm volatile("cp.async.wait_group 1;" ::: "memory");
            else asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        const int sb = PAIR ? (kst & 1) : buf;
        // the row below me has consumed what I wrote two steps ago into this buffer (passes at once on the first two steps)
        if (PAIR && lx >= 1) mbar_wait(&xbar[1][sb][lx], ((unsigned)(kst >> 1) & 1u) ^ 1u);
        if (cell && valid) {
            const double* src = ring + (size_t)buf * NV * NT + tid;
            const i64 o1x = (i64)(t + 1) * g.PBX, o1y = (i64)(t + 1) * g.PBY;
#define LDV(slot, gexpr, bexpr) (BULK ? (bexpr) : KM_PREFETCH ? src[(slot) * NT] : (gexpr))
            double b[10];
#pragma unroll
            for (int j = 0; j < 10; j++)
                b[j] = (ONE_D && j >= 5 && j <= 8) ? 0.0
                                                   : LDV(j, beta[(i64)j * L + cidx], RB(j * TX, lx, (pc ^ par_c ^ ((j & 1) & (par_c >> 4))) & 1u));
            cn.q0 = LDV(10, qn[cidx], RB(R_QN0, lx, (pc ^ (par_c >> 1)) & 1u));
            a0 = LDV(12, alpha[cidx], RB(R_A0, lx, (pc ^ (par_c >> 3)) & 1u));
            cn.bxm1 = hxm ? LDV(13, qn_bx[o1x + ibxm], RB(R_QNBX, lx, pbx ^ par_bxn)) : 0.0;
            cn.bx1 = hxp ? LDV(14, qn_bx[o1x + ibx], RB(R_QNBX, lx + 1, pbx ^ nyp ^ par_bxn)) : 0.0;
            cn.bym1 = hym ? LDV(15, qn_by[o1y + ibym], RB(R_QNBY, lx, pby ^ par_byn)) : 0.0;
            cn.by1 = hyp ? LDV(16, qn_by[o1y + iby], RB(R_QNBY, lx, 1 + (pby ^ par_byn))) : 0.0;
            double z2n[10];
            cell_z2(cn, sc, hxm, hxp, hym, hyp, z2n);
            if (UPDATE) {
                co.q0 = LDV(11, qo[cidx], RB(R_QO0, lx, (pc ^ (par_c >> 2)) & 1u));
                co.bxm1 = hxm ? LDV(17, qo_bx[o1x + ibxm], RB(R_QOBX, lx, pbx ^ par_bxo)) : 0.0;
                co.bx1 = hxp ? LDV(18, qo_bx[o1x + ibx], RB(R_QOBX, lx + 1, pbx ^ nyp ^ par_bxo)) : 0.0;
                co.bym1 = hym ? LDV(19, qo_by[o1y + ibym], RB(R_QOBY, lx, pby ^ par_byo)) : 0.0;
                co.by1 = hyp ? LDV(20, qo_by[o1y + iby], RB(R_QOBY, lx, 1 + (pby ^ par_byo))) : 0.0;
#undef LDV
                double v[10];
                cell_z2(co, sc, hxm, hxp, hym, hyp, v);
#pragma unroll
                for (int j = 0; j < 10; j++) v[j] = dsub(v[j], b[j]);
                proj_soc<ONE_D>(v);  // v = z
#pragma unroll
                for (int j = 0; j < 10; j++) {
                    b[j] = dadd(b[j], dmul(sc.tau, dsub(v[j], z2n[j])));
                    if (owner && !(ONE_D && j >= 5 && j <= 8)) beta_out[(i64)j * L + cidx] = b[j];
                }
            }
            // z-step input of the next iteration
#pragma unroll
            for (int j = 0; j < 10; j++) w[j] = dsub(z2n[j], b[j]);
            proj_soc<ONE_D>(w);
#pragma unroll
            for (int j = 0; j < 10; j++) w[j] = dadd(w[j], b[j]);
            sh[sb][0][lx][ly] = w[1];
            sh[sb][1][lx][ly] = w[3];
            if (!PAIR) {
                sh[sb][2][lx][ly] = w[5];
                sh[sb][3][lx][ly] = w[7];
            }
        }
        double w1n_ = 0.0, w3n_ = 0.0, w5n_ = 0.0, w7n_ = 0.0;   // w columns 1,3 of (x+1,y) and 5,7 of (x,y+1)
        if (PAIR) {
            __syncwarp();
            if (lx >= 1 && ly == 0) mbar_arrive(&xbar[0][sb][lx]);          // my row is published
            w5n_ = __shfl_down_sync(0xffffffffu, w[5], 1);
            w7n_ = __shfl_down_sync(0xffffffffu, w[7], 1);
            if (lx < TX - 1) {
                mbar_wait(&xbar[0][sb][lx + 1], (unsigned)(kst >> 1) & 1u);     // the row above is published
                w1n_ = sh[sb][0][lx + 1][ly];
                w3n_ = sh[sb][1][lx + 1][ly];
                __syncwarp();
                if (ly == 0) mbar_arrive(&xbar[1][sb][lx + 1]);              // ... and consumed
            }
        } else {
            __syncthreads();
            if (lx < TX - 1) { w1n_ = sh[sb][0][lx + 1][ly]; w3n_ = sh[sb][1][lx + 1][ly]; }
            if (ly < TY - 1) { w5n_ = sh[sb][2][lx][ly + 1]; w7n_ = sh[sb][3][lx][ly + 1]; }
        }
        // every thread has read stage kst: refill it with the rows of step t+2
        if (BULK && t + 2 < tr.tn1) fill(t + 2);
        cw += (unsigned)g.P;
        bw += (unsigned)g.PBX;
        yw += (unsigned)g.PBY;
        if (owner) {
            if (cell) {
                if (emit) q2[cidx] = dmul(dsub(w[9], w[0]), sc.S);
            }
            const double u0 = cell ? uval<WEIGHTED>(cn.q0, a0, wt0) : 0.0;
            // rhs = A' u + c : CSR-transpose row order (t-1 edge, t edge, x-1, x, y-1, y)
            double acc = 0.0;
            bool first = true;
#define ADDTERM(val)                         \
    {                                        \
        const double tv_ = (val);            \
        acc = first ? tv_ : dadd(acc, tv_);  \
        first = false;                       \
    }
            if (t > 0) ADDTERM(dmul(sc.gt, u0p));
            if (cell) ADDTERM(dmul(-sc.gt, u0));
            if (hxp || hxm) {
                const i64 ox = (i64)t * g.PBX;
                if (hxm) ADDTERM(dmul(sc.gx, uval<WEIGHTED>(cn.bxm, al_xm, wt_xm)));
                if (hxp) {
                    ADDTERM(dmul(-sc.gx, uval<WEIGHTED>(cn.bx, al_x, wt_x)));
                    const double w1n = cell ? w1n_ : 0.0;
                    const double w3n = cell ? w3n_ : 0.0;
                    double s;
                    if (t == 0)
                        s = dadd(w1n, w[2]);
                    else if (!cell)
                        s = dadd(wp3n, wp4);
                    else
                        s = dadd(dadd(dadd(w1n, w[2]), wp3n), wp4);
                    if (emit) q2_bx[ox + ibx] = dmul(s, sc.SF);
                    wp3n = w3n;
                }
            }
            if (hyp || hym) {
                const i64 oy = (i64)t * g.PBY;
                if (hym) ADDTERM(dmul(sc.gy, uval<WEIGHTED>(cn.bym, al_ym, wt_ym)));
                if (hyp) {
                    ADDTERM(dmul(-sc.gy, uval<WEIGHTED>(cn.by, al_y, wt_y)));
                    const double w5n = cell ? w5n_ : 0.0;
                    const double w7n = cell ? w7n_ : 0.0;
                    double s;
                    if (t == 0)
                        s = dadd(w5n, w[6]);
                    else if (!cell)
                        s = dadd(wp7n, wp8);
                    else
                        s = dadd(dadd(dadd(w5n, w[6]), wp7n), wp8);
                    if (emit) q2_by[oy + iby] = dmul(s, sc.SF);
                    wp7n = w7n;
                }
            }
#undef ADDTERM
            double cv = 0.0;
            if (t == 0) cv = c0[node];
            else if (!cell) cv = c1[node];
            if (emit) rhs[cidx] = dadd(first ? 0.0 : acc, cv);
            u0p = u0;
            wp4 = w[4];
            wp8 = w[8];
        }
#undef RB
        // advance the carried node-level values
        cn.bxm = cn.bxm1; cn.bx = cn.bx1; cn.bym = cn.bym1; cn.by = cn.by1;
        if (UPDATE) { co.bxm = co.bxm1; co.bx = co.bx1; co.bym = co.bym1; co.by = co.by1; }
    }
}

template <int TX, int TY, bool WEIGHTED, bool ONE_D, bool UPDATE>
__global__ void __launch_bounds__(TX* TY, KM_MIN_BLOCKS) k_mult(Geo g, TRange tr, int nchunk, IterScal sc, const double* __restrict__ qo,
                                                 const double* __restrict__ qn, const double* __restrict__ alpha,
                                                 const double* __restrict__ weight, const double* __restrict__ beta,
                                                 double* __restrict__ beta_out, double* __restrict__ q2,
                                                 double* __restrict__ rhs, const double* __restrict__ c0,
                                                 const double* __restrict__ c1)
{
    if (nchunk > 1) {
        // the time range is cut into nchunk pieces (blockIdx.z), each marched by its own CTA exactly like the slab of a
        // multi-GPU run: a piece that does not start at the range's first cell layer replays the layer below it
        const int i = blockIdx.z, nc = tr.tc1 - tr.tc0;
        const int a = tr.tc0 + (int)((i64)i * nc / nchunk), b = tr.tc0 + (int)((i64)(i + 1) * nc / nchunk);
        tr = TRange{a, b, i == 0 ? tr.tn0 : a, i == nchunk - 1 ? tr.tn1 : b};
    }
    const int x0 = blockIdx.y * (TX - 1), y0 = blockIdx.x * (TY - 1);
    const bool interior = !ONE_D && x0 >= 1 && x0 + TX - 1 <= g.nx - 2 && y0 >= 1 && y0 + TY - 1 <= g.ny - 2;
    if (interior)
        k_mult_body<TX, TY, WEIGHTED, ONE_D, UPDATE, false>(g, tr, sc, qo, qn, alpha, weight, beta, beta_out, q2, rhs, c0, c1);
    else
        k_mult_body<TX, TY, WEIGHTED, ONE_D, UPDATE, true>(g, tr, sc, qo, qn, alpha, weight, beta, beta_out, q2, rhs, c0, c1);
}

// number of time pieces that maximises (fraction of the last round that is filled) x (useful layers / marched layers)
static int km_pick_chunks(long long base_ctas, int ncells, int slots)
{
    int best = 1;
    double best_eff = 0.0;
    const int nmax = std::max(1, std::min(32, ncells / 4));
    // marches longer than 128 steps let the CTAs of a row drift apart (measured: 37.2 ms in one piece, 36.4 in 4 and 36.0 in 16
    // at 1024x1024x512), so long ranges always get a few pieces
    const int nmin = std::min(nmax, (ncells + 127) / 128);
    for (int n = nmin; n <= nmax; n++) {
        const long long total = base_ctas * n;
        const long long rounds = (total + slots - 1) / slots;
        const double eff = (double)total / (double)(rounds * slots) * (double)ncells / (double)(ncells + n - 1);
        if (eff > best_eff * 1.01) { best_eff = eff; best = n; }
    }
    return best;
}

void launch_mult(const UpdateArgs& a, bool weighted, bool one_d, bool update, cudaStream_t st)
{
    constexpr int TX = KM_TX, TY = KM_TY;
    dim3 block(TY, TX);
    dim3 grid((unsigned)((a.g.ny + TY - 2) / (TY - 1)), (unsigned)((a.g.nx + TX - 2) / (TX - 1)));
    const size_t ring = KM_PREFETCH ? (size_t)2 * 21 * TX * TY * sizeof(double)
                        : KM_BULK   ? (size_t)2 * km_nrows(TX) * KM_ROWD * sizeof(double) + 16
                                    : 0;
    const size_t smem = (size_t)2 * 4 * TX * TY * sizeof(double) + ring;
    // Every CTA marches the same number of time steps, so a grid that fills the machine 4.25 times runs for 5 full rounds
    // (512x512x256: 1258 CTAs on 296 slots).  Cutting the time range into pieces (grid.z) lets the CTA count land just below
    // a whole number of rounds; each extra piece costs one replayed cell layer.  DOTSOCP_KM_CHUNKS=n forces n pieces.
    static const int forced = [] { const char* e = getenv("DOTSOCP_KM_CHUNKS"); return e ? atoi(e) : 0; }();
#define KM(W, O, U)                                                                                                   \
    {                                                                                                                 \
        static int slots = 0;                                                                                         \
        if (!slots) {                                                                                                 \
            cudaFuncSetAttribute(k_mult<TX, TY, W, O, U>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);    \
            int per_sm = 0, dev = 0, sms = 0;                                                                         \
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_mult<TX, TY, W, O, U>, TX * TY, smem);           \
            cudaGetDevice(&dev);                                                                                      \
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);                                        \
            slots = per_sm > 0 && sms > 0 ? per_sm * sms : 1;                                                         \
        }                                                                                                             \
        const int nchunk = forced > 0 ? std::min(forced, std::max(1, a.tr.tc1 - a.tr.tc0))                            \
                                      : km_pick_chunks((long long)grid.x * grid.y, a.tr.tc1 - a.tr.tc0, slots);       \
        grid.z = (unsigned)nchunk;                                                                                    \
        k_mult<TX, TY, W, O, U><<<grid, block, smem, st>>>(a.g, a.tr, nchunk, a.sc, a.q_old, a.q_new, a.alpha,        \
                                                            a.weight, a.beta_in, a.beta_out, a.q2, a.rhs, a.c0, a.c1); \
    }
    if (one_d) {
        if (update) KM(false, true, true) else KM(false, true, false)
    } else if (weighted) {
        if (update) KM(true, false, true) else KM(true, false, false)
    } else {
        if (update) KM(false, false, true) else KM(false, false, false)
    }
#undef KM
}

// ---------------------------------------------------------------------------------------------------------------
// Unfused building blocks for the loops that keep z as state (PALM, acc-ADMM).
// rhs = A'(w.*q - alpha) + c                                                      solver_socp_accADMM.m:243
// ---------------------------------------------------------------------------------------------------------------
template <bool WEIGHTED>
__global__ void __launch_bounds__(256) k_rhs(Geo g, IterScal sc, const double* __restrict__ q, const double* __restrict__ alpha,
                                             const double* __restrict__ weight, const double* __restrict__ c0,
                                             const double* __restrict__ c1, double* __restrict__ rhs)
{
    const i64 p = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    const int t = blockIdx.y;
    if (p >= g.P) return;
    const int x = (int)(p / g.ny), y = (int)(p - (i64)x * g.ny);
    const i64 L = g.L, n = (i64)t * g.P + p;
    const bool up = t < g.nt - 1, dn = t > 0;
    auto u = [&](i64 e) -> double { return uval<WEIGHTED>(q[e], alpha[e], WEIGHTED ? weight[e] : 1.0); };
    double acc = 0.0;
    bool first = true;
#define ADDTERM(val)                         \
    {                                        \
        const double tv_ = (val);            \
        acc = first ? tv_ : dadd(acc, tv_);  \
        first = false;                       \
    }
    if (dn) ADDTERM(dmul(sc.gt, u(n - g.P)));
    if (up) ADDTERM(dmul(-sc.gt, u(n)));
    const i64 ox = L + (i64)t * g.PBX + (i64)x * g.ny + y, oy = L + g.NBX + (i64)t * g.PBY + (i64)x * (g.ny - 1) + y;
    if (x > 0) ADDTERM(dmul(sc.gx, u(ox - g.ny)));
    if (x < g.nx - 1) ADDTERM(dmul(-sc.gx, u(ox)));
    if (y > 0) ADDTERM(dmul(sc.gy, u(oy - 1)));
    if (y < g.ny - 1) ADDTERM(dmul(-sc.gy, u(oy)));
#undef ADDTERM
    double cv = 0.0;
    if (t == 0) cv = c0[p];
    else if (!up) cv = c1[p];
    rhs[n] = dadd(first ? 0.0 : acc, cv);
}
void launch_rhs(const Geo& g, const IterScal& sc, bool weighted, const double* q, const double* alpha, const double* weight,
                const double* c0, const double* c1, double* rhs, cudaStream_t st)
{
    dim3 grid((unsigned)((g.P + 255) / 256), (unsigned)g.nt);
    if (weighted) k_rhs<true><<<grid, 256, 0, st>>>(g, sc, q, alpha, weight, c0, c1, rhs);
    else k_rhs<false><<<grid, 256, 0, st>>>(g, sc, q, alpha, weight, c0, c1, rhs);
}

// cell-local multiplier updates with z materialised (in place; no neighbour cell is touched)
//   MODE 0 (PALM :219-224)     : beta = beta + tau*(z - z2(q))
//   MODE 1 (acc-ADMM :236-248) : beta = (beta + z) - z2(q) ; z = Pi_Q(z2(q) - beta)
//   MODE 2 (PALM :137-138)     : z = d + BF q on the non-boundary entries only (mexBFd(z, tmp_q, ...))
template <bool ONE_D, int MODE>
__global__ void __launch_bounds__(256) k_cells_update(Geo g, IterScal sc, const double* __restrict__ q, double* z, double* beta)
{
    const i64 p = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    const int t = blockIdx.y;
    if (p >= g.P) return;
    const int x = (int)(p / g.ny), y = (int)(p - (i64)x * g.ny);
    const bool hxm = x > 0, hxp = x < g.nx - 1, hym = y > 0, hyp = y < g.ny - 1;
    const i64 L = g.L, c = (i64)t * g.P + p;
    const double* bx = q + L;
    const double* by = bx + g.NBX;
    const i64 ox = (i64)t * g.PBX + (i64)x * g.ny + y, oy = (i64)t * g.PBY + (i64)x * (g.ny - 1) + y;
    CellQ cq;
    cq.q0 = q[c];
    cq.bxm = hxm ? bx[ox - g.ny] : 0.0;
    cq.bx = hxp ? bx[ox] : 0.0;
    cq.bxm1 = hxm ? bx[ox + g.PBX - g.ny] : 0.0;
    cq.bx1 = hxp ? bx[ox + g.PBX] : 0.0;
    cq.bym = hym ? by[oy - 1] : 0.0;
    cq.by = hyp ? by[oy] : 0.0;
    cq.bym1 = hym ? by[oy + g.PBY - 1] : 0.0;
    cq.by1 = hyp ? by[oy + g.PBY] : 0.0;
    double z2[10];
    cell_z2(cq, sc, hxm, hxp, hym, hyp, z2);
    if (MODE == 2) {
        const bool wr[10] = {true, hxm, hxp, hxm, hxp, hym, hyp, hym, hyp, true};
#pragma unroll
        for (int j = 0; j < 10; j++)
            if (wr[j] && !(ONE_D && j >= 5 && j <= 8)) z[(i64)j * L + c] = z2[j];
        return;
    }
    double zz[10], b[10];
#pragma unroll
    for (int j = 0; j < 10; j++) {
        const bool dead = ONE_D && j >= 5 && j <= 8;
        zz[j] = dead ? 0.0 : z[(i64)j * L + c];
        b[j] = dead ? 0.0 : beta[(i64)j * L + c];
        if (MODE == 0) b[j] = dadd(b[j], dmul(sc.tau, dsub(zz[j], z2[j])));
        else b[j] = dsub(dadd(b[j], zz[j]), z2[j]);
        if (!dead) beta[(i64)j * L + c] = b[j];
    }
    if (MODE == 1) {
#pragma unroll
        for (int j = 0; j < 10; j++) zz[j] = dsub(z2[j], b[j]);
        proj_soc<ONE_D>(zz);
#pragma unroll
        for (int j = 0; j < 10; j++)
            if (!(ONE_D && j >= 5 && j <= 8)) z[(i64)j * L + c] = zz[j];
    }
}
void launch_cells_update(const Geo& g, const IterScal& sc, bool one_d, int mode, const double* q, double* z, double* beta,
                         cudaStream_t st)
{
    dim3 grid((unsigned)((g.P + 255) / 256), (unsigned)(g.nt - 1));
#define CUK(O, M) k_cells_update<O, M><<<grid, 256, 0, st>>>(g, sc, q, z, beta)
    if (one_d) { if (mode == 0) CUK(true, 0); else if (mode == 1) CUK(true, 1); else CUK(true, 2); }
    else { if (mode == 0) CUK(false, 0); else if (mode == 1) CUK(false, 1); else CUK(false, 2); }
#undef CUK
}

// ---------------------------------------------------------------------------------------------------------------
// Deterministic reductions: every CTA writes its K partial sums, a single CTA adds them in a fixed tree order.
// ---------------------------------------------------------------------------------------------------------------
template <int K, int NT>
__device__ __forceinline__ void block_reduce_store(double (&s)[K], double* __restrict__ partial)
{
    __shared__ double red[K][NT / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; k++) {
        double v = s[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) red[k][wid] = v;
    }
    __syncthreads();
    if (threadIdx.x < K) {
        double v = 0.0;
        for (int i = 0; i < NT / 32; i++) v += red[threadIdx.x][i];
        partial[(i64)(blockIdx.y * gridDim.x + blockIdx.x) * K + threadIdx.x] = v;
    }
}

template <int K>
__global__ void __launch_bounds__(256) k_final_reduce(const double* __restrict__ partial, int nblocks, double* __restrict__ out)
{
    __shared__ double red[256];
    for (int k = 0; k < K; k++) {
        double v = 0.0;
        for (int i = threadIdx.x; i < nblocks; i += 256) v += partial[(i64)i * K + k];
        red[threadIdx.x] = v;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) {
            if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
            __syncthreads();
        }
        if (threadIdx.x == 0) out[k] = red[0];
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// KKT sums over cells (solver_socp_inPALM.m:228,231,236,240-241 and compute_kkt_dot_complement.m:2-8).
// ---------------------------------------------------------------------------------------------------------------
template <bool WEIGHTED, bool ONE_D>
__global__ void __launch_bounds__(256) k_kkt_cells(KktArgs a)
{
    const Geo& g = a.g;
    const i64 p = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    const int t = a.tr.tc0 + blockIdx.y;
    double s[KC_COUNT];
#pragma unroll
    for (int k = 0; k < KC_COUNT; k++) s[k] = 0.0;
    if (p < g.P) {
        const int x = (int)(p / g.ny), y = (int)(p - (i64)x * g.ny);
        const bool hxm = x > 0, hxp = x < g.nx - 1, hym = y > 0, hyp = y < g.ny - 1;
        const i64 L = g.L, c = (i64)t * g.P + p;
        const double* bx = a.q + L;
        const double* by = bx + g.NBX;
        CellQ cq;
        cq.q0 = a.q[c];
        const i64 ox = (i64)t * g.PBX + (i64)x * g.ny + y, oy = (i64)t * g.PBY + (i64)x * (g.ny - 1) + y;
        cq.bxm = hxm ? bx[ox - g.ny] : 0.0;
        cq.bx = hxp ? bx[ox] : 0.0;
        cq.bxm1 = hxm ? bx[ox + g.PBX - g.ny] : 0.0;
        cq.bx1 = hxp ? bx[ox + g.PBX] : 0.0;
        cq.bym = hym ? by[oy - 1] : 0.0;
        cq.by = hyp ? by[oy] : 0.0;
        cq.bym1 = hym ? by[oy + g.PBY - 1] : 0.0;
        cq.by1 = hyp ? by[oy + g.PBY] : 0.0;
        double z2[10], z[10], b[10], v[10];
        cell_z2(cq, a.sc, hxm, hxp, hym, hyp, z2);
        load_z<ONE_D>(a.g, a.sc, a.z, a.q_old, a.beta_old, t, x, y, z);
#pragma unroll
        for (int j = 0; j < 10; j++) {
            const bool dead = ONE_D && j >= 5 && j <= 8;
            b[j] = dead ? 0.0 : a.beta[(i64)j * L + c];
            s[KC_Z2] += z[j] * z[j];
            s[KC_BETA2] += b[j] * b[j];
            const double r = dsub(z[j], z2[j]);
            s[KC_PRIM2] += r * r;
            v[j] = dsub(z[j], dmul(a.sigma, b[j]));
        }
        proj_soc<ONE_D>(v);
#pragma unroll
        for (int j = 0; j < 10; j++) {
            const double r = dsub(z[j], v[j]);
            s[KC_COMPL] += r * r;
        }
        // DOT-level complementarity
        const double al = WEIGHTED ? dmul(a.weight[c], a.alpha[c]) : a.alpha[c];
        const double rhoT = dmul(dmul(dmul(a.sigma, a.cScale), a.D), al);
        const double sE = a.dScale / a.E;
        double ss = 0.0;
#pragma unroll
        for (int j = 1; j <= 8; j++) {
            const double e = dmul(sE, z2[j]);
            ss = (j == 1) ? dmul(e, e) : dadd(ss, dmul(e, e));
        }
        double rhoFq = dadd(dadd(rhoT, dmul(a.dScale / a.D, cq.q0)), ss / 4.0);
        if (rhoFq < 0.0) rhoFq = 0.0;
        const double dr = dsub(rhoT, rhoFq);
        s[KC_DOTC] = dr * dr;
        s[KC_RHOT] = rhoT * rhoT;
        s[KC_RHOFQ] = rhoFq * rhoFq;
    }
    block_reduce_store<KC_COUNT, 256>(s, a.partial);
}

int kkt_cells_blocks(const Geo& g) { return (int)((g.P + 255) / 256) * (g.nt - 1); }
int kkt_nodes_blocks(const Geo& g) { return (int)((g.P + 255) / 256) * g.nt; }
static int blocks_x(const Geo& g) { return (int)((g.P + 255) / 256); }

void launch_kkt_cells(const KktArgs& a, bool weighted, bool one_d, cudaStream_t st)
{
    const int nl = a.tr.tc1 - a.tr.tc0;
    dim3 grid((unsigned)blocks_x(a.g), (unsigned)nl);
    if (one_d)
        k_kkt_cells<false, true><<<grid, 256, 0, st>>>(a);
    else if (weighted)
        k_kkt_cells<true, false><<<grid, 256, 0, st>>>(a);
    else
        k_kkt_cells<false, false><<<grid, 256, 0, st>>>(a);
    k_final_reduce<KC_COUNT><<<1, 256, 0, st>>>(a.partial, blocks_x(a.g) * nl, a.out);
}

// ---------------------------------------------------------------------------------------------------------------
// z = Pi_Q(d + BF q_old - beta_old) materialised (final output, :334) and/or its squared Frobenius norm (rescale
// block, :141).  zout may alias beta_old (cell-local read-then-write).
// ---------------------------------------------------------------------------------------------------------------
template <bool ONE_D>
__global__ void __launch_bounds__(256) k_zstep(Geo g, int tc0, IterScal sc, const double* q_old, const double* beta_old,
                                               double* zout, double* __restrict__ partial)
{
    const i64 p = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    const int t = tc0 + blockIdx.y;
    double s[1] = {0.0};
    if (p < g.P) {
        const int x = (int)(p / g.ny), y = (int)(p - (i64)x * g.ny);
        double z[10];
        load_z<ONE_D>(g, sc, nullptr, q_old, beta_old, t, x, y, z);
        const i64 c = (i64)t * g.P + p;
#pragma unroll
        for (int j = 0; j < 10; j++) {
            s[0] += z[j] * z[j];
            if (zout != nullptr && !(ONE_D && j >= 5 && j <= 8)) zout[(i64)j * g.L + c] = z[j];
        }
    }
    block_reduce_store<1, 256>(s, partial);
}

void launch_zstep(const Geo& g, const IterScal& sc, bool one_d, const double* q_old, const double* beta_old, double* zout,
                  double* partial, double* out, cudaStream_t st, const TRange* tr)
{
    const int tc0 = tr ? tr->tc0 : 0, tc1 = tr ? tr->tc1 : g.nt - 1;
    dim3 grid((unsigned)blocks_x(g), (unsigned)(tc1 - tc0));
    if (one_d)
        k_zstep<true><<<grid, 256, 0, st>>>(g, tc0, sc, q_old, beta_old, zout, partial);
    else
        k_zstep<false><<<grid, 256, 0, st>>>(g, tc0, sc, q_old, beta_old, zout, partial);
    k_final_reduce<1><<<1, 256, 0, st>>>(partial, blocks_x(g) * (tc1 - tc0), out);
}

// ---------------------------------------------------------------------------------------------------------------
// KKT sums over nodes / staggered edges (solver_socp_inPALM.m:227-238,265-266; compute_kkt_dot_complement.m:10-18).
// ---------------------------------------------------------------------------------------------------------------
template <bool WEIGHTED>
__global__ void __launch_bounds__(256) k_kkt_nodes(KktArgs a)
{
    const Geo& g = a.g;
    const i64 p = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    const int t = a.tr.tn0 + blockIdx.y;
    double s[KN_COUNT];
#pragma unroll
    for (int k = 0; k < KN_COUNT; k++) s[k] = 0.0;
    if (p < g.P) {
        const int x = (int)(p / g.ny), y = (int)(p - (i64)x * g.ny);
        const i64 L = g.L, n = (i64)t * g.P + p;
        const bool up = t < g.nt - 1, dn = t > 0;
        const bool hxm = x > 0, hxp = x < g.nx - 1, hym = y > 0, hyp = y < g.ny - 1;
        const IterScal& sc = a.sc;
        const double ph = a.phi[n];
        const double scD = dmul(dmul(a.sigma, a.cScale), a.D);   // sigma*cScale*D
        const double dD = a.dScale / a.D;
        // Dalpha on the t-edges around this node and its x+1 / y+1 neighbours (for rho = pair-average in t)
        auto dal0 = [&](i64 c) -> double { return WEIGHTED ? dmul(a.weight[c], a.alpha[c]) : a.alpha[c]; };
        auto rho_at = [&](i64 pp) -> double {   // rho(t, node pp) = (rhoT[t-1] + rhoT[t]) / 2 with zero padding
            const double lo = dn ? dmul(scD, dal0((i64)(t - 1) * g.P + pp)) : 0.0;
            const double hi = up ? dmul(scD, dal0((i64)t * g.P + pp)) : 0.0;
            return dadd(lo, hi) / 2.0;
        };
        auto edge = [&](i64 e, double aphi, bool momentum, double rho_avg) {
            const double qv = a.q[e], av = a.alpha[e], fb = a.q2b[e];
            const double w = WEIGHTED ? a.weight[e] : 1.0;
            const double wq = WEIGHTED ? dmul(w, qv) : qv;
            const double wa = WEIGHTED ? dmul(w, av) : av;
            s[KN_Q2] += qv * qv;
            s[KN_APHI2] += aphi * aphi;
            const double r1 = dsub(aphi, wq);
            s[KN_PRIM1] += r1 * r1;
            s[KN_ALPHA2] += av * av;
            s[KN_FBB2] += fb * fb;
            const double r2 = dadd(fb, wa);
            s[KN_DUAL2] += r2 * r2;
            s[KN_QDOTA] += wq * av;
            if (momentum) {
                const double m = dmul(scD, wa);
                const double rb = dmul(dD, dmul(rho_avg, qv));
                const double d = dsub(m, rb);
                s[KN_MRHOB] += d * d;
                s[KN_M2] += m * m;
                s[KN_RHOB2] += rb * rb;
            }
        };
        const double rho_c = rho_at(p);
        if (up) edge(n, dadd(dmul(-sc.gt, ph), dmul(sc.gt, a.phi[n + g.P])), false, 0.0);
        if (hxp)
            edge(L + (i64)t * g.PBX + (i64)x * g.ny + y, dadd(dmul(-sc.gx, ph), dmul(sc.gx, a.phi[n + g.ny])), true,
                 dadd(rho_c, rho_at(p + g.ny)) / 2.0);
        if (hyp)
            edge(L + g.NBX + (i64)t * g.PBY + (i64)x * (g.ny - 1) + y, dadd(dmul(-sc.gy, ph), dmul(sc.gy, a.phi[n + 1])),
                 true, dadd(rho_c, rho_at(p + 1)) / 2.0);
        // dual residual A' alpha - c at the node (CSR-transpose row order)
        const double* al_bx = a.alpha + L;
        const double* al_by = al_bx + g.NBX;
        double acc = 0.0;
        bool first = true;
#define ADDTERM(val)                         \
    {                                        \
        const double tv_ = (val);            \
        acc = first ? tv_ : dadd(acc, tv_);  \
        first = false;                       \
    }
        if (dn) ADDTERM(dmul(sc.gt, a.alpha[n - g.P]));
        if (up) ADDTERM(dmul(-sc.gt, a.alpha[n]));
        const i64 ox = (i64)t * g.PBX + (i64)x * g.ny + y, oy = (i64)t * g.PBY + (i64)x * (g.ny - 1) + y;
        if (hxm) ADDTERM(dmul(sc.gx, al_bx[ox - g.ny]));
        if (hxp) ADDTERM(dmul(-sc.gx, al_bx[ox]));
        if (hym) ADDTERM(dmul(sc.gy, al_by[oy - 1]));
        if (hyp) ADDTERM(dmul(-sc.gy, al_by[oy]));
#undef ADDTERM
        double cv = 0.0;
        if (t == 0) cv = a.c0[p];
        else if (!up) cv = a.c1[p];
        const double rd = dsub(first ? 0.0 : acc, cv);
        s[KN_DUAL1] = rd * rd;
        s[KN_CPHI] = cv * ph;
        s[KN_PHI2] = ph * ph;
    }
    block_reduce_store<KN_COUNT, 256>(s, a.partial);
}

void launch_kkt_nodes(const KktArgs& a, bool weighted, cudaStream_t st)
{
    const int nl = a.tr.tn1 - a.tr.tn0;
    dim3 grid((unsigned)blocks_x(a.g), (unsigned)nl);
    if (weighted)
        k_kkt_nodes<true><<<grid, 256, 0, st>>>(a);
    else
        k_kkt_nodes<false><<<grid, 256, 0, st>>>(a);
    k_final_reduce<KN_COUNT><<<1, 256, 0, st>>>(a.partial, blocks_x(a.g) * nl, a.out);
}

// ---------------------------------------------------------------------------------------------------------------
// small streaming helpers
// ---------------------------------------------------------------------------------------------------------------
constexpr int SUMSQ_MAX_BLOCKS = 148 * 8;
__global__ void __launch_bounds__(256) k_sumsq(const double* __restrict__ x, i64 n, double* __restrict__ partial)
{
    double s[1] = {0.0};
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) s[0] += x[i] * x[i];
    block_reduce_store<1, 256>(s, partial);
}
int sumsq_blocks(i64 n)
{
    i64 b = (n + 255) / 256;
    if (b > SUMSQ_MAX_BLOCKS) b = SUMSQ_MAX_BLOCKS;
    if (b < 1) b = 1;
    return (int)b;
}
void launch_sumsq(const double* x, i64 n, double* partial, double* out, cudaStream_t st)
{
    const int nb = sumsq_blocks(n);
    k_sumsq<<<nb, 256, 0, st>>>(x, n, partial);
    k_final_reduce<1><<<1, 256, 0, st>>>(partial, nb, out);
}

__global__ void __launch_bounds__(256) k_scale(double* __restrict__ x, i64 n, double mul, double div)
{
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x)
        x[i] = dmul(x[i], mul) / div;
}
void launch_scale(double* x, i64 n, double mul, double div, cudaStream_t st)
{
    if (n <= 0) return;
    i64 b = (n + 255) / 256;
    if (b > 148 * 16) b = 148 * 16;
    k_scale<<<(unsigned)b, 256, 0, st>>>(x, n, mul, div);
}

__global__ void __launch_bounds__(256) k_halpern(double* __restrict__ x, double* __restrict__ xold, double* __restrict__ x0,
                                                 i64 n, double c1, double c2, double rho, int copy_anchor)
{
    const double omr = 1.0 - rho;
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        const double inner = dadd(dmul(omr, xold[i]), dmul(rho, x[i]));
        const double v = dadd(dmul(c1, x0[i]), dmul(c2, inner));
        x[i] = v;
        xold[i] = v;
        if (copy_anchor) x0[i] = v;
    }
}
void launch_halpern(double* x, double* xold, double* x0, i64 n, double c1, double c2, double rho, bool copy_anchor,
                    cudaStream_t st)
{
    if (n <= 0) return;
    i64 b = (n + 255) / 256;
    if (b > 148 * 16) b = 148 * 16;
    k_halpern<<<(unsigned)b, 256, 0, st>>>(x, xold, x0, n, c1, c2, rho, copy_anchor ? 1 : 0);
}

// acc-ADMM, general extrapolation (opts.theta != 2), solver_socp_accADMM.m:389-417:
//   hat = (1-rho)*old + rho*x ;  x = (1-c1)*old + c1*hat                      (k == 0)
//                                x = (1-c1)*old + (c1+c2)*hat - c2*hatOld      (k  > 0) ;  old = x ; hatOld = hat (unless restart)
__global__ void __launch_bounds__(256) k_accel3(double* __restrict__ x, double* __restrict__ xold, double* __restrict__ xhatold,
                                                i64 n, double rho, double a, double b, double c2, int first, int keep_hat)
{
    const double omr = 1.0 - rho;
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        const double o = xold[i];
        const double hat = dadd(dmul(omr, o), dmul(rho, x[i]));
        double v = dadd(dmul(a, o), dmul(b, hat));
        if (!first) v = dsub(v, dmul(c2, xhatold[i]));
        x[i] = v;
        xold[i] = v;
        if (keep_hat) xhatold[i] = hat;
    }
}
void launch_accel3(double* x, double* xold, double* xhatold, i64 n, double rho, double a, double b, double c2, bool first,
                   bool keep_hat, cudaStream_t st)
{
    if (n <= 0) return;
    i64 blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    k_accel3<<<(unsigned)blocks, 256, 0, st>>>(x, xold, xhatold, n, rho, a, b, c2, first ? 1 : 0, keep_hat ? 1 : 0);
}

__global__ void __launch_bounds__(256) k_cols6to10(const double* __restrict__ in6, double* __restrict__ out10, i64 L)
{
    const i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    if (i >= L) return;
#pragma unroll
    for (int j = 0; j < 5; j++) out10[(i64)j * L + i] = in6[(i64)j * L + i];
#pragma unroll
    for (int j = 5; j < 9; j++) out10[(i64)j * L + i] = 0.0;
    out10[9 * L + i] = in6[5 * L + i];
}
__global__ void __launch_bounds__(256) k_cols10to6(const double* __restrict__ in10, double* __restrict__ out6, i64 L)
{
    const i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    if (i >= L) return;
#pragma unroll
    for (int j = 0; j < 5; j++) out6[(i64)j * L + i] = in10[(i64)j * L + i];
    out6[5 * L + i] = in10[9 * L + i];
}
void launch_cols6to10(const double* in6, double* out10, i64 L, cudaStream_t st)
{
    k_cols6to10<<<(unsigned)((L + 255) / 256), 256, 0, st>>>(in6, out10, L);
}
void launch_cols10to6(const double* in10, double* out6, i64 L, cudaStream_t st)
{
    k_cols10to6<<<(unsigned)((L + 255) / 256), 256, 0, st>>>(in10, out6, L);
}

}  // namespace dsocp
