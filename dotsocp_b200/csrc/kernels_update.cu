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
#include "reduce.cuh"
#include "vmm.h"
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

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
    const i64 L = g.L, c = (i64)t * g.PC + (i64)x * g.py + y;
    if (zmat != nullptr) {
#pragma unroll
        for (int j = 0; j < 10; j++) z[j] = (ONE_D && j >= 5 && j <= 8) ? 0.0 : zmat[(i64)j * L + c];
        return;
    }
    const bool hxm = x > 0, hxp = x < g.nx - 1, hym = y > 0, hyp = y < g.ny - 1;
    const double* bx = q_old + L;
    const double* by = bx + g.NBX;
    const i64 ox = (i64)t * g.PBX + (i64)x * g.py + y, oy = (i64)t * g.PBY + (i64)x * g.pyb + y;
    CellQ cq;
    cq.q0 = q_old[c];
    cq.bxm = hxm ? bx[ox - g.py] : 0.0;
    cq.bx = hxp ? bx[ox] : 0.0;
    cq.bxm1 = hxm ? bx[ox + g.PBX - g.py] : 0.0;
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
    const PlanePos pp = plane_pos(g, p);
    if (!pp.ok) return;
    const int x = pp.x, y = pp.y;
    const i64 c = (i64)t * g.PC + p;
    const double* bx = q + g.L;
    const double* by = bx + g.NBX;
    const double pr = dmul(q[c], S);
    z[c] = dsub(DF, pr);
    z[9 * g.L + c] = dadd(pr, DF);
    if (x >= 1) {
        z[1 * g.L + c] = dmul(bx[(i64)t * g.PBX + (i64)(x - 1) * g.py + y], SF);
        z[3 * g.L + c] = dmul(bx[(i64)(t + 1) * g.PBX + (i64)(x - 1) * g.py + y], SF);
    }
    if (x <= g.nx - 2) {
        z[2 * g.L + c] = dmul(bx[(i64)t * g.PBX + (i64)x * g.py + y], SF);
        z[4 * g.L + c] = dmul(bx[(i64)(t + 1) * g.PBX + (i64)x * g.py + y], SF);
    }
    if (y >= 1) {
        z[5 * g.L + c] = dmul(by[(i64)t * g.PBY + (i64)x * g.pyb + (y - 1)], SF);
        z[7 * g.L + c] = dmul(by[(i64)(t + 1) * g.PBY + (i64)x * g.pyb + (y - 1)], SF);
    }
    if (y <= g.ny - 2) {
        z[6 * g.L + c] = dmul(by[(i64)t * g.PBY + (i64)x * g.pyb + y], SF);
        z[8 * g.L + c] = dmul(by[(i64)(t + 1) * g.PBY + (i64)x * g.pyb + y], SF);
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
    dim3 grid((unsigned)((g.PC + 255) / 256), (unsigned)(g.nt - 1));
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
    const PlanePos pp = plane_pos(g, p);
    if (!pp.ok) return;
    const int x = pp.x, y = pp.y;
    const i64 L = g.L;
    const i64 cu = (i64)t * g.PC + p;        // cell (t,x,y)
    const i64 cd = cu - g.PC;                // cell (t-1,x,y)
    const bool up = t < g.nt - 1, dn = t > 0;
    if (up) q2[cu] = dmul(dsub(Z(9 * L + cu), Z(cu)), S);
    if (x < g.nx - 1) {
        double s;
        if (!dn)
            s = dadd(Z(1 * L + cu + g.py), Z(2 * L + cu));
        else if (!up)
            s = dadd(Z(3 * L + cd + g.py), Z(4 * L + cd));
        else {
            s = dadd(Z(1 * L + cu + g.py), Z(2 * L + cu));
            s = dadd(s, Z(3 * L + cd + g.py));
            s = dadd(s, Z(4 * L + cd));
        }
        q2[L + (i64)t * g.PBX + p] = dmul(s, SF);
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
        q2[L + g.NBX + (i64)t * g.PBY + (i64)x * g.pyb + y] = dmul(s, SF);
    }
}

void launch_bfdconj(const Geo& g, double S, const double* z, double* q2, cudaStream_t st, const TRange* tr)
{
    const int tn0 = tr ? tr->tn0 : 0, tn1 = tr ? tr->tn1 : g.nt;
    dim3 grid((unsigned)((g.PC + 255) / 256), (unsigned)(tn1 - tn0));
    k_bfdconj<false><<<grid, 256, 0, st>>>(g, tn0, S, host_sf(S), z, nullptr, q2);
}
void launch_bfdconj_sum(const Geo& g, double S, const double* za, const double* zb, double* q2, cudaStream_t st)
{
    dim3 grid((unsigned)((g.PC + 255) / 256), (unsigned)g.nt);
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
template <bool WEIGHTED, bool ACC, bool KKT>
__device__ __forceinline__ void q_update(i64 e, double aphi, double dinv_plain, double s2term, const IterScal& sc,
                                         const double* __restrict__ q2, const double* __restrict__ weight,
                                         double* __restrict__ alpha, double* __restrict__ qout,
                                         const double* __restrict__ tmpq_in, double* __restrict__ tmpq_out, bool upd_alpha,
                                         double (&ks)[KQ_COUNT])
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
    double an;
    if (ACC)
        an = dsub(dadd(a, aphi), wq);
    else
        an = dadd(a, dmul(sc.tau, dsub(aphi, wq)));
    alpha[e] = an;
    if (KKT) {   // the terms of the check that live on this edge (solver_socp_inPALM.m:227-230,234,265)
        ks[KQ_Q2] += qn * qn;
        ks[KQ_APHI2] += aphi * aphi;
        const double r1 = dsub(aphi, wq);
        ks[KQ_PRIM1] += r1 * r1;
        ks[KQ_ALPHA2] += an * an;
        ks[KQ_QDOTA] += wq * an;
    }
}

template <bool WEIGHTED, bool ACC, bool KKT>
__global__ void __launch_bounds__(256) k_qstep(Geo g, int tn0, IterScal sc, const double* __restrict__ phi,
                                               const double* __restrict__ q2, const double* __restrict__ weight,
                                               double* __restrict__ alpha, double* __restrict__ qout,
                                               const double* __restrict__ tmpq_in, double* __restrict__ tmpq_out,
                                               bool upd_alpha, const double* __restrict__ c0, const double* __restrict__ c1,
                                               double* __restrict__ kpart, int kkt_t0)
{
    const i64 p = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    const int t = tn0 + blockIdx.y;
    double ks[KQ_COUNT];
#pragma unroll
    for (int k = 0; k < KQ_COUNT; k++) ks[k] = 0.0;
    const PlanePos pp = plane_pos(g, p);
    if (pp.ok) {
        const int x = pp.x, y = pp.y;
        const i64 n = (i64)t * g.P + pp.pn;
        const double ph = phi[n];
        const bool edge_t = (t == 0) || (t == g.nt - 1);
        if (t < g.nt - 1) {
            const double aphi = dadd(dmul(-sc.gt, ph), dmul(sc.gt, phi[n + g.P]));
            q_update<WEIGHTED, ACC, KKT>((i64)t * g.PC + p, aphi, sc.dinv1, sc.s2x2, sc, q2, weight, alpha, qout, tmpq_in, tmpq_out, upd_alpha, ks);
        }
        if (x < g.nx - 1) {
            const double aphi = dadd(dmul(-sc.gx, ph), dmul(sc.gx, phi[n + g.ny]));
            q_update<WEIGHTED, ACC, KKT>(g.L + (i64)t * g.PBX + p, aphi, edge_t ? sc.dinv2 : sc.dinv1,
                                         edge_t ? sc.s2x1 : sc.s2x2, sc, q2, weight, alpha, qout, tmpq_in, tmpq_out, upd_alpha, ks);
        }
        if (y < g.ny - 1) {
            const double aphi = dadd(dmul(-sc.gy, ph), dmul(sc.gy, phi[n + 1]));
            q_update<WEIGHTED, ACC, KKT>(g.L + g.NBX + (i64)t * g.PBY + (i64)x * g.pyb + y, aphi,
                                         edge_t ? sc.dinv2 : sc.dinv1, edge_t ? sc.s2x1 : sc.s2x2, sc, q2, weight, alpha, qout,
                                         tmpq_in, tmpq_out, upd_alpha, ks);
        }
        if (KKT) {
            double cv = 0.0;
            if (t == 0) cv = c0[pp.pn];
            else if (t == g.nt - 1) cv = c1[pp.pn];
            ks[KQ_CPHI] = cv * ph;
            ks[KQ_PHI2] = ph * ph;
        }
    }
    if (KKT) block_reduce_store<KQ_COUNT, 256>(ks, kpart, (i64)(t - kkt_t0) * gridDim.x + blockIdx.x);
}

void launch_qstep(const UpdateArgs& a, bool weighted, bool acc, cudaStream_t st, const double* tmpq_in, double* tmpq_out,
                  bool upd_alpha, const KktFused* kkt)
{
    dim3 grid((unsigned)((a.g.PC + 255) / 256), (unsigned)(a.tr.tn1 - a.tr.tn0));
    if (grid.y == 0) return;
#define QS(W, A, K)                                                                                                          \
    k_qstep<W, A, K><<<grid, 256, 0, st>>>(a.g, a.tr.tn0, a.sc, a.phi, a.q2, a.weight, a.alpha, a.q_new, tmpq_in, tmpq_out, \
                                           upd_alpha, a.c0, a.c1, kkt ? kkt->partial_q : nullptr, a.kkt_t0)
    if (kkt) {   // inPALM check iteration
        if (weighted) QS(true, false, true); else QS(false, false, true);
    } else if (weighted) {
        if (acc) QS(true, true, false); else QS(true, false, false);
    } else {
        if (acc) QS(false, true, false); else QS(false, false, false);
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
//
// A step has three stages: LOAD (the 25 values of the step, nothing else), phase 1 (cell-local: two projections, the
// multiplier update, publish w to shared memory) and phase 2 (gather the neighbours' w after the CTA barrier, emit q2 / rhs).
// ncu on the round-1 kernel: 55 % of the warp samples waited on the first use of a step's loads (long scoreboard), i.e. the
// HBM latency was exposed once per step.  PF = true keeps the loaded values of step t+1 in a second register set that is
// filled BEFORE step t is computed (software pipeline, one 255-register CTA per SM), so a step's loads have a whole step of
// arithmetic to arrive.  PF = 0 is the round-1 schedule (two 128-register CTAs per SM, loads then compute).  PF = 4 (the
// default on the aligned tiling): one thread per CTA issues the step's inputs as TMA tensor loads (cp.async.bulk.tensor, 10
// boxes per step, the ten beta planes in one) into a two-stage shared-memory ring guarded by mbarriers, one step and a half
// ahead; the arithmetic reads its operands from shared memory, which leaves 128 registers (two CTAs per SM) AND a deep
// prefetch -- 4.13 -> 3.50 ms at 512x512x256.  PF = 2 / 3 fill the same ring with one 256-byte cp.async.bulk per row and
// stream (155 per step): slower than PF = 1, the TMA unit cannot issue that many small copies.  DOTSOCP_KM_PF selects at run time.
//
// KKT (check iterations): the same march also accumulates every KKT term that lives on the data in registers
// (solver_socp_inPALM.m:225-244, compute_kkt_dot_complement.m) -- z, beta, z2, alpha and q are all there -- and leaves one
// partial sum per (time level, tile); s(BF)^* beta uses a second set of exchange planes.  The remaining terms (A*phi, q,
// alpha norms, <c,phi>) come from k_qstep<KKT>.
// ---------------------------------------------------------------------------------------------------------------
template <bool WEIGHTED>
__device__ __forceinline__ double uval(double q, double a, double w)
{
    return WEIGHTED ? dsub(dmul(w, q), a) : dsub(q, a);
}

#ifndef KM_TX
#define KM_TX 8          // tile rows (x) per CTA
#endif
#ifndef KM_TY
#define KM_TY 32         // tile columns (y, contiguous) per CTA: one warp per tile row
#endif
#ifndef KM_PF
#define KM_PF 4          // default of DOTSOCP_KM_PF: TMA tensor loads into a two-stage shared-memory ring (aligned tiling; else 1)
#endif

struct KktDev {          // scalars of the fused KKT terms
    double sigma, scD, dD, sE, dSD;   // sigma ; sigma*cScale*D ; dScale/D (momentum) ; dScale/E ; dScale/D (rhoFq)
};

// what phase 2 of a step needs from its phase 1
struct MultKeep {
    double w2, w4, w6, w8;            // own w columns of the (BF)^* sums
    double u0, uxm, ux, uym, uy;      // (w.*q - alpha) on the 5 edges of the rhs stencil that phase 1 can already form
    double cv;                        // c on the first / last time level
};
struct MultKeepK {                    // KKT extras
    double b2, b4, b6, b8, fb0;       // beta columns of s(BF)^* beta ; fb0 = S (b9 - b0)
    double a0, axm, ax, aym, ay;      // alpha on the node's edges (a0: t-edge above)
    double wa0;                       // w .* alpha0 (rho, dual residual 2)
    double qx, qy, wx, wy;            // q and weight on the node's own bx / by edge (momentum terms)
    double rhoc;                      // rho at the node: pair average in t of sigma*cScale*D*w.*alpha0
    double cell[KM_RHOFQ + 1];        // the 7 cell sums of this thread
};

// the values one time step reads from HBM
struct MultLoad {
    double b[10];                                 // beta of the cell
    double q0n, bxm1n, bx1n, bym1n, by1n;         // new q: q0 of the cell, bx / by at node level t+1
    double q0o, bxm1o, bx1o, bym1o, by1o;         // old q (UPDATE)
    double a0;                                    // alpha0 of the cell
    double al_xm, al_x, al_ym, al_y;              // alpha on the x / y edges of node level t (owner threads)
    double wt0, wt_xm, wt_x, wt_ym, wt_y;         // weights on the same edges (WEIGHTED)
    double cv;                                    // c on the first / last time level
};

// ---- TMA-engine bulk copies (PF == 2 of the aligned march): global -> shared, completion counted in bytes on an mbarrier ----
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 16-byte aligned source, destination and size
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// tensor (tiled) copies: one instruction moves a whole box of a tensor map; out-of-range parts arrive as zeros
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar)) : "memory");
}
// shared-memory image of one step's inputs (PF >= 2): rows of the tile, bx rows x0-1 .. x0+TX-1, by rows with columns y0-2 .. y0+31
template <int TX, bool WEIGHTED>
struct KmStage {
    static constexpr int NQ0 = 13 + (WEIGHTED ? 1 : 0);   // beta 0..9, q_new0, alpha0, q_old0 [, weight0]
    static constexpr int NXB = 3 + (WEIGHTED ? 1 : 0);    // q_new, q_old (level t+1), alpha [, weight] (level t)
    static constexpr int BYW = 34;                        // doubles per by row
    static constexpr int Q0 = 0, BX = NQ0 * TX * 32, BY = BX + NXB * (TX + 1) * 32, SIZE = BY + NXB * TX * BYW;
    enum { S_QN = 10, S_A0 = 11, S_QO = 12, S_W0 = 13 };
    enum { X_QN = 0, X_QO = 1, X_AL = 2, X_W = 3 };
};

// AL ("aligned", pitched layouts only): the tile is TX x TY cells that are ALL owned, with its first column on a 32-column
// (256-byte) boundary, so that every warp-wide load and store of a step is one aligned 256-byte run -- partially written
// 32-byte sectors are what limits the packed / haloed tiling (common.cuh, tools/stream_pattern3.cu).  Without the halo the
// last row / column of a tile cannot form its bx / by sums (they need w of the first row / column of the next tile):
// both sides leave their two w columns in a small side buffer instead (rows: straight from the warps, columns: through
// shared memory so that one warp writes full sectors) and k_q2_fix completes those edges after the march.
struct SideGeo {
    int t0;        // first cell layer backed by the buffer
    int nbx, nby;  // tiles along x / y
    int nxp;       // nx rounded up to TX
    i64 sx_t;      // doubles per cell layer of the row part   [t][kx][4][py]   : w1, w3 of the tile's first row; w2, w4 of its last
    i64 sy_off;    // offset of the column part                [t][ky][4][nxp]  : w5, w7 of the first column; w6, w8 of the last
    i64 sy_t;
};
__host__ __device__ __forceinline__ i64 side_sx(const SideGeo& sg, const Geo& g, int t, int kx, int comp, int y)
{
    return (i64)(t - sg.t0) * sg.sx_t + ((i64)kx * 4 + comp) * g.py + y;
}
__host__ __device__ __forceinline__ i64 side_sy(const SideGeo& sg, int t, int ky, int comp, int x)
{
    return sg.sy_off + (i64)(t - sg.t0) * sg.sy_t + ((i64)ky * 4 + comp) * sg.nxp + x;
}

template <int TX, int TY, int PF, bool WEIGHTED, bool ONE_D, bool UPDATE, bool EDGE, bool KKT, bool AL>
__device__ __forceinline__ void k_mult_body(const Geo& g, const TRange& tr, int kkt_t0, const IterScal& sc, const KktDev& kd,
                                            const SideGeo& sg, double* __restrict__ side, const KmMaps& maps,
                                            const double* __restrict__ qo, const double* __restrict__ qn,
                                            const double* __restrict__ alpha, const double* __restrict__ weight,
                                            const double* __restrict__ beta, double* __restrict__ beta_out,
                                            double* __restrict__ q2, double* __restrict__ rhs, const double* __restrict__ c0,
                                            const double* __restrict__ c1, double* __restrict__ kpart)
{
    static_assert(!KKT || UPDATE, "the KKT variant is an update kernel");
    static_assert(!AL || (!KKT && !ONE_D && TY == 32 && 4 * TX <= 32), "aligned tiling: plain 2-D update / prologue kernels");
    constexpr int TU = 1;
    constexpr int NPL = KKT ? 9 : 4;   // exchange planes: w1,w3,w5,w7 (+ b1,b3,b5,b7, rho)
    extern __shared__ __align__(128) double dyn_smem[];
    double (*sh)[TU][NPL][TX][TY] = reinterpret_cast<double (*)[TU][NPL][TX][TY]>(dyn_smem);
    double (*shy)[4][TX] = reinterpret_cast<double (*)[4][TX]>(dyn_smem + 2 * TU * NPL * TX * TY);   // AL: [2][4][TX]
    static_assert(PF < 2 || (AL && UPDATE && !KKT), "the bulk-copy ring exists for the aligned update kernel");
    // PF == 2: two stages, 128 registers, two CTAs per SM ; PF == 3: three stages, one 255-register CTA per SM ; both filled by
    // one 256-byte bulk copy per row and stream.  PF == 4: two stages filled by TENSOR copies -- one thread issues 10 (13
    // weighted) box loads per step (the 10 beta planes are one 20 KB box) instead of 155 row copies, which the TMA unit cannot
    // issue fast enough (profiles/README.md)
    constexpr bool RING = PF >= 2;
    constexpr bool TENSOR = PF == 4;
    constexpr int NSTG = PF == 3 ? 3 : 2;
    typedef KmStage<TX, WEIGHTED> ST;
    double* ring = dyn_smem + 2 * TU * NPL * TX * TY + 2 * 4 * TX;             // RING: [NSTG][ST::SIZE]
    unsigned long long* mbar = reinterpret_cast<unsigned long long*>(ring + NSTG * ST::SIZE);
    int rstage = 0;                                                            // ring stage of the step being computed
    const int ly = threadIdx.x, lx = threadIdx.y;
    // y tiles vary fastest over the grid so that CTAs running side by side stream adjacent pieces of the same rows
    const int x = AL ? blockIdx.y * TX + lx : blockIdx.y * (TX - 1) + lx, y = AL ? blockIdx.x * TY + ly : blockIdx.x * (TY - 1) + ly;
    // EDGE = false: the whole tile (halo included) lies strictly inside the domain, every neighbour exists and all the
    // boundary predicates fold away at compile time (most CTAs of a large grid)
    const bool valid = EDGE ? ((x < g.nx) && (y < g.ny)) : true;
    const bool owner = AL ? valid : (valid && (lx < TX - 1) && (ly < TY - 1));
    // AL: the x+1 / y+1 neighbour's w is inside the tile (else the edge is left to k_q2_fix)
    const bool in_x = AL ? (lx < TX - 1) : true, in_y = AL ? (ly < TY - 1) : true;
    const bool hxm = EDGE ? (valid && x > 0) : true, hxp = EDGE ? (valid && x < g.nx - 1) : true;
    const bool hym = EDGE ? (valid && y > 0) : true, hyp = EDGE ? (valid && y < g.ny - 1) : true;
    const i64 L = g.L;
    const i64 node = (i64)x * g.ny + y;            // inside a node plane (rhs, c)
    const i64 cellp = (i64)x * g.py + y;           // inside a cell / q0 plane
    const i64 ibx = cellp, ibxm = ibx - g.py;
    const i64 iby = (i64)x * g.pyb + y, ibym = iby - 1;
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
    co.q0 = 0.0;
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
    double wp3n = 0.0, wp4 = 0.0, wp7n = 0.0, wp8 = 0.0, u0p = 0.0;     // carried from the cell layer below
    double bp3n = 0.0, bp4 = 0.0, bp7n = 0.0, bp8 = 0.0, a0p = 0.0, wa0p = 0.0;   // KKT: same for beta / alpha0

    // ---- LOAD: everything step t reads from HBM, and nothing else ------------------------------------------------------
    auto load = [&](int t, MultLoad& ld) {
        const bool cell = t < g.nt - 1;
        const i64 cidx = (i64)t * g.PC + cellp;
        ld.al_xm = ld.al_x = ld.al_ym = ld.al_y = 0.0;
        ld.wt0 = ld.wt_xm = ld.wt_x = ld.wt_ym = ld.wt_y = 1.0;
        ld.cv = 0.0;
        if (owner) {
            const i64 ox = (i64)t * g.PBX, oy = (i64)t * g.PBY;
            if (hxm) ld.al_xm = al_bx[ox + ibxm];
            if (hxp) ld.al_x = al_bx[ox + ibx];
            if (hym) ld.al_ym = al_by[oy + ibym];
            if (hyp) ld.al_y = al_by[oy + iby];
            if (WEIGHTED) {
                if (hxm) ld.wt_xm = w_bx[ox + ibxm];
                if (hxp) ld.wt_x = w_bx[ox + ibx];
                if (hym) ld.wt_ym = w_by[oy + ibym];
                if (hyp) ld.wt_y = w_by[oy + iby];
            }
            if (t == 0) ld.cv = c0[node];
            else if (!cell) ld.cv = c1[node];
        }
        if (cell && valid) {
            const i64 o1x = (i64)(t + 1) * g.PBX, o1y = (i64)(t + 1) * g.PBY;
            if (WEIGHTED) ld.wt0 = weight[cidx];
#pragma unroll
            for (int j = 0; j < 10; j++) ld.b[j] = (ONE_D && j >= 5 && j <= 8) ? 0.0 : beta[(i64)j * L + cidx];
            ld.q0n = qn[cidx];
            ld.a0 = alpha[cidx];
            ld.bxm1n = hxm ? qn_bx[o1x + ibxm] : 0.0;
            ld.bx1n = hxp ? qn_bx[o1x + ibx] : 0.0;
            ld.bym1n = hym ? qn_by[o1y + ibym] : 0.0;
            ld.by1n = hyp ? qn_by[o1y + iby] : 0.0;
            if (UPDATE) {
                ld.q0o = qo[cidx];
                ld.bxm1o = hxm ? qo_bx[o1x + ibxm] : 0.0;
                ld.bx1o = hxp ? qo_bx[o1x + ibx] : 0.0;
                ld.bym1o = hym ? qo_by[o1y + ibym] : 0.0;
                ld.by1o = hyp ? qo_by[o1y + iby] : 0.0;
            }
        }
    };

    // ---- RING: every warp starts the bulk copies of its own row for step tt into ring stage `stg` (one copy per lane) -----
    auto issue = [&](int tt, int stg) {
        if (TENSOR) {
            if (lx == 0 && ly == 0) {
                const bool cellt_ = tt < g.nt - 1;
                double* stage_ = ring + (size_t)stg * ST::SIZE;
                const int x0_ = blockIdx.y * TX, y0_ = blockIdx.x * TY;
                constexpr unsigned BQ = TX * 32 * 8, BB = 10 * BQ, BXB = (TX + 1) * 32 * 8, BYB = TX * ST::BYW * 8;
                const unsigned nlev = 1 + (WEIGHTED ? 1 : 0);                       // alpha (+ weight) on node level tt
                const unsigned ncell = 3 + (WEIGHTED ? 1 : 0);                      // q_new0, alpha0, q_old0 (+ weight0)
                const unsigned total = nlev * (BXB + BYB) + (cellt_ ? BB + ncell * BQ + 2 * (BXB + BYB) : 0);
                mbar_arrive_expect_tx(&mbar[stg], total);
                unsigned long long* bar = &mbar[stg];
                // (time coordinates are relative to the first layer / level of the views: maps.toc, maps.ton)
                auto q0 = [&](int slot, const CUtensorMap* m) { tma_load_3d(stage_ + ST::Q0 + slot * TX * 32, m, y0_, x0_, tt - maps.toc, bar); };
                auto bxy = [&](int k, const CUtensorMap* m, int lvl) {
                    tma_load_3d(stage_ + ST::BX + k * (TX + 1) * 32, m + 1, y0_, x0_ - 1, lvl - maps.ton, bar);
                    tma_load_3d(stage_ + ST::BY + k * TX * ST::BYW, m + 2, y0_ - 2, x0_, lvl - maps.ton, bar);
                };
                if (cellt_) {
                    tma_load_4d(stage_ + ST::Q0, &maps.beta, y0_, x0_, tt - maps.toc, 0, bar);
                    q0(ST::S_QN, maps.qn); q0(ST::S_A0, maps.al); q0(ST::S_QO, maps.qo);
                    if (WEIGHTED) q0(ST::S_W0, maps.w);
                    bxy(ST::X_QN, maps.qn, tt + 1);
                    bxy(ST::X_QO, maps.qo, tt + 1);
                }
                bxy(ST::X_AL, maps.al, tt);
                if (WEIGHTED) bxy(ST::X_W, maps.w, tt);
            }
            return;
        }
        const bool cellt = tt < g.nt - 1;
        const int xr = blockIdx.y * TX + lx, y0 = blockIdx.x * TY;
        const bool row = xr < g.nx, rowx = xr < g.nx - 1;
        double* stage = ring + (size_t)stg * ST::SIZE;
        const double* src = nullptr;
        double* dst = nullptr;
        unsigned bytes = 0;
        const i64 crow = (i64)tt * g.PC + (i64)xr * g.py + y0;                          // cell row
        const i64 xrow1 = (i64)(tt + 1) * g.PBX + (i64)xr * g.py + y0, xrow0 = (i64)tt * g.PBX + (i64)xr * g.py + y0;
        const i64 yrow1 = (i64)(tt + 1) * g.PBY + (i64)xr * g.pyb + y0, yrow0 = (i64)tt * g.PBY + (i64)xr * g.pyb + y0;
        const int l = ly;
        if (row) {
            if (l < 10) { if (cellt) { src = beta + (i64)l * L + crow; dst = stage + ST::Q0 + (l * TX + lx) * 32; bytes = 256; } }
            else if (l == 10) { if (cellt) { src = qn + crow; dst = stage + ST::Q0 + (ST::S_QN * TX + lx) * 32; bytes = 256; } }
            else if (l == 11) { if (cellt) { src = alpha + crow; dst = stage + ST::Q0 + (ST::S_A0 * TX + lx) * 32; bytes = 256; } }
            else if (l == 12) { if (cellt) { src = qo + crow; dst = stage + ST::Q0 + (ST::S_QO * TX + lx) * 32; bytes = 256; } }
            else if (l == 13) { if (WEIGHTED && cellt) { src = weight + crow; dst = stage + ST::Q0 + (ST::S_W0 * TX + lx) * 32; bytes = 256; } }
            else if (l <= 17) {          // bx rows: own row goes to slot lx + 1
                const int k = l - 14;
                const bool lvl1 = k <= ST::X_QO;      // q: node level tt+1 ; alpha / weight: node level tt
                const double* base = k == ST::X_QN ? qn_bx : k == ST::X_QO ? qo_bx : k == ST::X_AL ? al_bx : w_bx;
                if (rowx && (k < 3 || WEIGHTED) && (!lvl1 || cellt)) {
                    src = base + (lvl1 ? xrow1 : xrow0);
                    dst = stage + ST::BX + (k * (TX + 1) + lx + 1) * 32;
                    bytes = 256;
                }
            } else if (l <= 21) {        // by rows: columns y0-2 .. y0+31 (the first tile has nothing to its left)
                const int k = l - 18;
                const bool lvl1 = k <= ST::X_QO;
                const double* base = k == ST::X_QN ? qn_by : k == ST::X_QO ? qo_by : k == ST::X_AL ? al_by : w_by;
                if ((k < 3 || WEIGHTED) && (!lvl1 || cellt)) {
                    const int sh2 = y0 == 0 ? 0 : 2;
                    src = base + (lvl1 ? yrow1 : yrow0) - sh2;
                    dst = stage + ST::BY + (k * TX + lx) * ST::BYW + 2 - sh2;
                    bytes = 256 + 8 * sh2;
                }
            } else if (l <= 25 && lx == 0 && xr > 0) {   // the bx row below the tile (x0 - 1), slot 0
                const int k = l - 22;
                const bool lvl1 = k <= ST::X_QO;
                const double* base = k == ST::X_QN ? qn_bx : k == ST::X_QO ? qo_bx : k == ST::X_AL ? al_bx : w_bx;
                if ((k < 3 || WEIGHTED) && (!lvl1 || cellt)) {
                    src = base + (lvl1 ? xrow1 : xrow0) - g.py;
                    dst = stage + ST::BX + (k * (TX + 1)) * 32;
                    bytes = 256;
                }
            }
        }
        const unsigned total = __reduce_add_sync(0xffffffffu, bytes);
        if (ly == 0) mbar_arrive_expect_tx(&mbar[stg], total);
        __syncwarp();
        if (bytes) bulk_g2s(dst, src, bytes, &mbar[stg]);
    };

    // ---- phase 1 of step t into exchange slot (buf, u) ---------------------------------------------------------------
    auto phase1 = [&](int t, int buf, int u, const MultLoad& ld, MultKeep& k, MultKeepK& kk) {
        const bool cell = t < g.nt - 1;
        const i64 cidx = (i64)t * g.PC + cellp;
        double w[10];
#pragma unroll
        for (int j = 0; j < 10; j++) w[j] = 0.0;
        double a0 = 0.0;
        // where the step's values come from: the prefetched register set (PF == 1) or HBM directly (PF == 0), every value
        // picked up where it is first needed
        const i64 o1x = (i64)(t + 1) * g.PBX, o1y = (i64)(t + 1) * g.PBY;
        const double* stq = ring + (size_t)rstage * ST::SIZE;   // RING: this step's stage
        auto F = [&](int slot, const double* gaddr, double regval) -> double {
            if (PF == 1) return regval;
            if (RING) {
                // slots: 0..9 beta, 10 q_new0, 11 alpha0, 12..15 q_new bx[x-1], bx[x], by[y-1], by[y], 16 q_old0, 17..20 q_old likewise
                if (slot <= 11) return stq[ST::Q0 + (slot * TX + lx) * 32 + ly];
                if (slot == 16) return stq[ST::Q0 + (ST::S_QO * TX + lx) * 32 + ly];
                const int k = slot >= 17 ? ST::X_QO : ST::X_QN, e = slot >= 17 ? slot - 17 : slot - 12;
                if (e == 0) return stq[ST::BX + (k * (TX + 1) + lx) * 32 + ly];
                if (e == 1) return stq[ST::BX + (k * (TX + 1) + lx + 1) * 32 + ly];
                if (e == 2) return stq[ST::BY + (k * TX + lx) * ST::BYW + ly + 1];
                return stq[ST::BY + (k * TX + lx) * ST::BYW + ly + 2];
            }
            return *gaddr;
        };
        double wt0 = 1.0, al_xm = 0.0, al_x = 0.0, al_ym = 0.0, al_y = 0.0, wt_xm = 1.0, wt_x = 1.0, wt_ym = 1.0, wt_y = 1.0;
        k.cv = 0.0;
        if (PF == 1) {
            wt0 = ld.wt0;
            al_xm = ld.al_xm; al_x = ld.al_x; al_ym = ld.al_ym; al_y = ld.al_y;
            wt_xm = ld.wt_xm; wt_x = ld.wt_x; wt_ym = ld.wt_ym; wt_y = ld.wt_y;
            k.cv = ld.cv;
        } else if (RING) {
            if (owner) {
                if (hxm) al_xm = stq[ST::BX + (ST::X_AL * (TX + 1) + lx) * 32 + ly];
                if (hxp) al_x = stq[ST::BX + (ST::X_AL * (TX + 1) + lx + 1) * 32 + ly];
                if (hym) al_ym = stq[ST::BY + (ST::X_AL * TX + lx) * ST::BYW + ly + 1];
                if (hyp) al_y = stq[ST::BY + (ST::X_AL * TX + lx) * ST::BYW + ly + 2];
                if (WEIGHTED) {
                    if (hxm) wt_xm = stq[ST::BX + (ST::X_W * (TX + 1) + lx) * 32 + ly];
                    if (hxp) wt_x = stq[ST::BX + (ST::X_W * (TX + 1) + lx + 1) * 32 + ly];
                    if (hym) wt_ym = stq[ST::BY + (ST::X_W * TX + lx) * ST::BYW + ly + 1];
                    if (hyp) wt_y = stq[ST::BY + (ST::X_W * TX + lx) * ST::BYW + ly + 2];
                }
                if (t == 0) k.cv = c0[node];
                else if (!cell) k.cv = c1[node];
            }
            if (WEIGHTED && cell && valid) wt0 = stq[ST::Q0 + (ST::S_W0 * TX + lx) * 32 + ly];
        } else {
            // level-t alpha (and weight) of the rhs stencil: direct loads, consumed behind the two projections
            if (owner) {
                const i64 ox = (i64)t * g.PBX, oy = (i64)t * g.PBY;
                if (hxm) al_xm = al_bx[ox + ibxm];
                if (hxp) al_x = al_bx[ox + ibx];
                if (hym) al_ym = al_by[oy + ibym];
                if (hyp) al_y = al_by[oy + iby];
                if (WEIGHTED) {
                    if (hxm) wt_xm = w_bx[ox + ibxm];
                    if (hxp) wt_x = w_bx[ox + ibx];
                    if (hym) wt_ym = w_by[oy + ibym];
                    if (hyp) wt_y = w_by[oy + iby];
                }
                if (t == 0) k.cv = c0[node];
                else if (!cell) k.cv = c1[node];
            }
            if (WEIGHTED && cell && valid) wt0 = weight[cidx];
        }
        if (KKT) {
#pragma unroll
            for (int j = 0; j <= KM_RHOFQ; j++) kk.cell[j] = 0.0;
            kk.b2 = kk.b4 = kk.b6 = kk.b8 = kk.fb0 = 0.0;
        }
        if (cell && valid) {
            double b[10];
#pragma unroll
            for (int j = 0; j < 10; j++) b[j] = (ONE_D && j >= 5 && j <= 8) ? 0.0 : F(j, beta + (i64)j * L + cidx, ld.b[j]);
            cn.q0 = F(10, qn + cidx, ld.q0n);
            a0 = F(11, alpha + cidx, ld.a0);
            cn.bxm1 = hxm ? F(12, qn_bx + o1x + ibxm, ld.bxm1n) : 0.0;
            cn.bx1 = hxp ? F(13, qn_bx + o1x + ibx, ld.bx1n) : 0.0;
            cn.bym1 = hym ? F(14, qn_by + o1y + ibym, ld.bym1n) : 0.0;
            cn.by1 = hyp ? F(15, qn_by + o1y + iby, ld.by1n) : 0.0;
            double z2n[10];
            cell_z2(cn, sc, hxm, hxp, hym, hyp, z2n);
            if (UPDATE) {
                co.q0 = F(16, qo + cidx, ld.q0o);
                co.bxm1 = hxm ? F(17, qo_bx + o1x + ibxm, ld.bxm1o) : 0.0;
                co.bx1 = hxp ? F(18, qo_bx + o1x + ibx, ld.bx1o) : 0.0;
                co.bym1 = hym ? F(19, qo_by + o1y + ibym, ld.bym1o) : 0.0;
                co.by1 = hyp ? F(20, qo_by + o1y + iby, ld.by1o) : 0.0;
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
                if (KKT) {
                    // cell terms of the check (:228,231,236,240-241 ; compute_kkt_dot_complement.m:2-8) on z, the new beta and z2n
                    double sz = 0.0, sb = 0.0, sp = 0.0, scp = 0.0, vv[10];
#pragma unroll
                    for (int j = 0; j < 10; j++) {
                        sz += v[j] * v[j];
                        sb += b[j] * b[j];
                        const double r = dsub(v[j], z2n[j]);
                        sp += r * r;
                        vv[j] = dsub(v[j], dmul(kd.sigma, b[j]));
                    }
                    proj_soc<ONE_D>(vv);
#pragma unroll
                    for (int j = 0; j < 10; j++) {
                        const double r = dsub(v[j], vv[j]);
                        scp += r * r;
                    }
                    const double al = WEIGHTED ? dmul(wt0, a0) : a0;
                    const double rhoT = dmul(kd.scD, al);
                    double ss = 0.0;
#pragma unroll
                    for (int j = 1; j <= 8; j++) {
                        const double e = dmul(kd.sE, z2n[j]);
                        ss = (j == 1) ? dmul(e, e) : dadd(ss, dmul(e, e));
                    }
                    double rhoFq = dadd(dadd(rhoT, dmul(kd.dSD, cn.q0)), ss / 4.0);
                    if (rhoFq < 0.0) rhoFq = 0.0;
                    const double dr = dsub(rhoT, rhoFq);
                    kk.cell[KM_Z2] = sz; kk.cell[KM_BETA2] = sb; kk.cell[KM_PRIM2] = sp; kk.cell[KM_COMPL] = scp;
                    kk.cell[KM_DOTC] = dr * dr; kk.cell[KM_RHOT] = rhoT * rhoT; kk.cell[KM_RHOFQ] = rhoFq * rhoFq;
                    sh[buf][u][4][lx][ly] = b[1];
                    sh[buf][u][5][lx][ly] = b[3];
                    sh[buf][u][6][lx][ly] = b[5];
                    sh[buf][u][7][lx][ly] = b[7];
                    kk.b2 = b[2]; kk.b4 = b[4]; kk.b6 = b[6]; kk.b8 = b[8];
                    kk.fb0 = dmul(dsub(b[9], b[0]), sc.S);
                }
            }
            // z-step input of the next iteration
#pragma unroll
            for (int j = 0; j < 10; j++) w[j] = dsub(z2n[j], b[j]);
            proj_soc<ONE_D>(w);
#pragma unroll
            for (int j = 0; j < 10; j++) w[j] = dadd(w[j], b[j]);
            sh[buf][u][0][lx][ly] = w[1];
            sh[buf][u][1][lx][ly] = w[3];
            sh[buf][u][2][lx][ly] = w[5];
            sh[buf][u][3][lx][ly] = w[7];
            if (AL) {
                // the tile's first / last row and column: what the neighbouring tile (resp. k_q2_fix) needs of this layer
                if (lx == 0) { side[side_sx(sg, g, t, blockIdx.y, 0, y)] = w[1]; side[side_sx(sg, g, t, blockIdx.y, 1, y)] = w[3]; }
                if (lx == TX - 1) { side[side_sx(sg, g, t, blockIdx.y, 2, y)] = w[2]; side[side_sx(sg, g, t, blockIdx.y, 3, y)] = w[4]; }
                if (ly == 0) { shy[buf][0][lx] = w[5]; shy[buf][1][lx] = w[7]; }
                if (ly == TY - 1) { shy[buf][2][lx] = w[6]; shy[buf][3][lx] = w[8]; }
            }
            if (owner && t >= tr.tn0) q2[cidx] = dmul(dsub(w[9], w[0]), sc.S);
        }
        k.w2 = w[2]; k.w4 = w[4]; k.w6 = w[6]; k.w8 = w[8];
        k.u0 = cell ? uval<WEIGHTED>(cn.q0, a0, wt0) : 0.0;
        k.uxm = uval<WEIGHTED>(cn.bxm, al_xm, wt_xm);
        k.ux = uval<WEIGHTED>(cn.bx, al_x, wt_x);
        k.uym = uval<WEIGHTED>(cn.bym, al_ym, wt_ym);
        k.uy = uval<WEIGHTED>(cn.by, al_y, wt_y);
        if (KKT) {
            kk.a0 = cell ? a0 : 0.0;
            kk.axm = al_xm; kk.ax = al_x; kk.aym = al_ym; kk.ay = al_y;
            kk.qx = cn.bx; kk.qy = cn.by; kk.wx = wt_x; kk.wy = wt_y;
            // rho at this node (compute_kkt_dot_complement.m:10): pair average in t of rhoT, zero beyond the two ends
            kk.wa0 = cell ? (WEIGHTED ? dmul(wt0, a0) : a0) : 0.0;
            const double lo = t > 0 ? dmul(kd.scD, wa0p) : 0.0;
            const double hi = cell ? dmul(kd.scD, kk.wa0) : 0.0;
            kk.rhoc = dadd(lo, hi) / 2.0;
            if (valid) sh[buf][u][8][lx][ly] = kk.rhoc;
            wa0p = kk.wa0;
        }
        // advance the carried node-level values
        cn.bxm = cn.bxm1; cn.bx = cn.bx1; cn.bym = cn.bym1; cn.by = cn.by1;
        if (UPDATE) { co.bxm = co.bxm1; co.bx = co.bx1; co.bym = co.bym1; co.by = co.by1; }
    };

    // ---- phase 2: gather the neighbours' columns, emit q2 (bx, by) and rhs of node level t -----------------------------
    auto phase2 = [&](int t, int buf, int u, const MultKeep& k, const MultKeepK& kk, double (&ks)[KM_COUNT]) {
        if (!owner) return;
        const bool cell = t < g.nt - 1;
        const bool emit = t >= tr.tn0;      // false only on the replayed ghost layer
        const i64 cidx = (i64)t * g.PC + cellp;
        const double w1n = (cell && in_x) ? sh[buf][u][0][lx + 1][ly] : 0.0, w3n = (cell && in_x) ? sh[buf][u][1][lx + 1][ly] : 0.0;
        const double w5n = (cell && in_y) ? sh[buf][u][2][lx][ly + 1] : 0.0, w7n = (cell && in_y) ? sh[buf][u][3][lx][ly + 1] : 0.0;
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
        if (cell) ADDTERM(dmul(-sc.gt, k.u0));
        if (hxm) ADDTERM(dmul(sc.gx, k.uxm));
        if (hxp) {
            ADDTERM(dmul(-sc.gx, k.ux));
            double s;
            if (t == 0)
                s = dadd(w1n, k.w2);
            else if (!cell)
                s = dadd(wp3n, wp4);
            else
                s = dadd(dadd(dadd(w1n, k.w2), wp3n), wp4);
            if (emit && in_x) q2_bx[(i64)t * g.PBX + ibx] = dmul(s, sc.SF);
            wp3n = w3n;
        }
        if (hym) ADDTERM(dmul(sc.gy, k.uym));
        if (hyp) {
            ADDTERM(dmul(-sc.gy, k.uy));
            double s;
            if (t == 0)
                s = dadd(w5n, k.w6);
            else if (!cell)
                s = dadd(wp7n, wp8);
            else
                s = dadd(dadd(dadd(w5n, k.w6), wp7n), wp8);
            if (emit && in_y) q2_by[(i64)t * g.PBY + iby] = dmul(s, sc.SF);
            wp7n = w7n;
        }
        if (emit) rhs[(i64)t * g.P + node] = dadd(first ? 0.0 : acc, k.cv);
        u0p = k.u0;
        wp4 = k.w4;
        wp8 = k.w8;
        if (KKT) {
            // node / edge terms (:225-238 ; compute_kkt_dot_complement.m:10-18) of node level t
            const double b1n = cell ? sh[buf][u][4][lx + 1][ly] : 0.0, b3n = cell ? sh[buf][u][5][lx + 1][ly] : 0.0;
            const double b5n = cell ? sh[buf][u][6][lx][ly + 1] : 0.0, b7n = cell ? sh[buf][u][7][lx][ly + 1] : 0.0;
            double fbb = 0.0, du2 = 0.0, mrb = 0.0, m2 = 0.0, rb2 = 0.0;
            auto edge = [&](double fb, double wa, bool momentum, double qv, double rho_avg) {   // wa = w .* alpha on the edge
                fbb += fb * fb;
                const double r2 = dadd(fb, wa);
                du2 += r2 * r2;
                if (momentum) {
                    const double m = dmul(kd.scD, wa);
                    const double rb = dmul(kd.dD, dmul(rho_avg, qv));
                    const double d = dsub(m, rb);
                    mrb += d * d;
                    m2 += m * m;
                    rb2 += rb * rb;
                }
            };
            // dual residual A' alpha - c (CSR-transpose row order)
            double ad = 0.0;
            bool f2 = true;
#define ADDA(val)                          \
    {                                      \
        const double tv_ = (val);          \
        ad = f2 ? tv_ : dadd(ad, tv_);     \
        f2 = false;                        \
    }
            if (t > 0) ADDA(dmul(sc.gt, a0p));
            if (cell) {
                ADDA(dmul(-sc.gt, kk.a0));
                edge(kk.fb0, kk.wa0, false, 0.0, 0.0);
            }
            if (hxm) ADDA(dmul(sc.gx, kk.axm));
            if (hxp) {
                ADDA(dmul(-sc.gx, kk.ax));
                double s;
                if (t == 0)
                    s = dadd(b1n, kk.b2);
                else if (!cell)
                    s = dadd(bp3n, bp4);
                else
                    s = dadd(dadd(dadd(b1n, kk.b2), bp3n), bp4);
                edge(dmul(s, sc.SF), WEIGHTED ? dmul(kk.wx, kk.ax) : kk.ax, true, kk.qx, dadd(kk.rhoc, sh[buf][u][8][lx + 1][ly]) / 2.0);
                bp3n = b3n;
            }
            if (hym) ADDA(dmul(sc.gy, kk.aym));
            if (hyp) {
                ADDA(dmul(-sc.gy, kk.ay));
                double s;
                if (t == 0)
                    s = dadd(b5n, kk.b6);
                else if (!cell)
                    s = dadd(bp7n, bp8);
                else
                    s = dadd(dadd(dadd(b5n, kk.b6), bp7n), bp8);
                edge(dmul(s, sc.SF), WEIGHTED ? dmul(kk.wy, kk.ay) : kk.ay, true, kk.qy, dadd(kk.rhoc, sh[buf][u][8][lx][ly + 1]) / 2.0);
                bp7n = b7n;
            }
#undef ADDA
            const double rd = dsub(f2 ? 0.0 : ad, k.cv);
            a0p = kk.a0;
            bp4 = kk.b4;
            bp8 = kk.b8;
            if (emit) {
#pragma unroll
                for (int j = 0; j <= KM_RHOFQ; j++) ks[j] = kk.cell[j];
                ks[KM_FBB2] = fbb; ks[KM_DUAL2] = du2; ks[KM_DUAL1] = rd * rd;
                ks[KM_MRHOB] = mrb; ks[KM_M2] = m2; ks[KM_RHOB2] = rb2;
            }
        }
#undef ADDTERM
    };

    MultKeep keep;
    MultKeepK keepk;
    // one step: phase 1 on the loaded values, barrier, phase 2 (+ the per-level KKT partial of the tile)
    auto step = [&](int t, int it, const MultLoad& ld) {
        const int buf = it & 1;
        phase1(t, buf, 0, ld, keep, keepk);
        __syncthreads();
        if (AL && lx == 0 && ly < 4 * TX && t < g.nt - 1) {
            // column part of the side buffer: one warp stores the 4 x TX values of this layer as full 64-byte runs
            const int comp = ly / TX, r = ly - comp * TX, xx = blockIdx.y * TX + r;
            if (xx < g.nx) side[side_sy(sg, t, blockIdx.x, comp, xx)] = shy[buf][comp][r];
        }
        // RING: every warp has read its inputs of step t (barrier above): refill the stage with step t + NSTG
        if (RING && t + NSTG < tr.tn1) issue(t + NSTG, rstage);
        double ks[KM_COUNT];
        if (KKT) {
#pragma unroll
            for (int j = 0; j < KM_COUNT; j++) ks[j] = 0.0;
        }
        phase2(t, buf, 0, keep, keepk, ks);
        if (KKT && t >= tr.tn0) {
            // one partial per (time level, tile): warp shuffles, then the TX warp sums in fixed order
            __shared__ double red[KM_COUNT][TX];
#pragma unroll
            for (int j = 0; j < KM_COUNT; j++) {
                double v = ks[j];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
                if (ly == 0) red[j][lx] = v;
            }
            __syncthreads();
            const int tid = lx * TY + ly;
            if (tid < KM_COUNT) {
                double v = 0.0;
#pragma unroll
                for (int i = 0; i < TX; i++) v += red[tid][i];
                const i64 tile = (i64)blockIdx.y * gridDim.x + blockIdx.x;
                kpart[((i64)(t - kkt_t0) * ((i64)gridDim.x * gridDim.y) + tile) * KM_COUNT + tid] = v;
            }
        }
    };
    int t = t_start, it = 0;
    if (RING) {
        if (lx == 0 && ly == 0) {
            for (int k = 0; k < NSTG; k++) mbar_init(&mbar[k], TENSOR ? 1 : TX);      // one arrival per warp (issuing thread) and phase
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        for (int k = 0; k < NSTG; k++)
            if (t + k < tr.tn1) issue(t + k, k);
        unsigned par = 0;
        for (; t < tr.tn1; t++, it++) {
            mbar_wait(&mbar[rstage], par);
            step(t, it, MultLoad());
            if (++rstage == NSTG) { rstage = 0; par ^= 1u; }
        }
    } else if (PF == 1) {
        // software pipeline, unrolled by two so that the two register sets swap roles without moves
        MultLoad la, lb;
        load(t, la);
        for (; t + 2 <= tr.tn1; t += 2, it += 2) {
            load(t + 1, lb);
            step(t, it, la);
            if (t + 2 < tr.tn1) load(t + 2, la);
            step(t + 1, it + 1, lb);
        }
        if (t < tr.tn1) step(t, it, la);
    } else {
        for (; t < tr.tn1; t++, it++) step(t, it, MultLoad());
    }
}

template <int TX, int TY, int PF, bool WEIGHTED, bool ONE_D, bool UPDATE, bool KKT, bool AL>
__global__ void __launch_bounds__(TX* TY, (PF == 1 || PF == 3 || KKT) ? 1 : 2)
k_mult(Geo g, TRange tr, int nchunk, IterScal sc, KktDev kd, SideGeo sg, double* __restrict__ side, const __grid_constant__ KmMaps maps,
       const double* __restrict__ qo,
       const double* __restrict__ qn, const double* __restrict__ alpha, const double* __restrict__ weight,
       const double* __restrict__ beta, double* __restrict__ beta_out, double* __restrict__ q2, double* __restrict__ rhs,
       const double* __restrict__ c0, const double* __restrict__ c1, double* __restrict__ kpart)
{
    const int kkt_t0 = tr.tn0;
    if (nchunk > 1) {
        // the time range is cut into nchunk pieces (blockIdx.z), each marched by its own CTA exactly like the slab of a
        // multi-GPU run: a piece that does not start at the range's first cell layer replays the layer below it
        const int i = blockIdx.z, nc = tr.tc1 - tr.tc0;
        const int a = tr.tc0 + (int)((i64)i * nc / nchunk), b = tr.tc0 + (int)((i64)(i + 1) * nc / nchunk);
        tr = TRange{a, b, i == 0 ? tr.tn0 : a, i == nchunk - 1 ? tr.tn1 : b};
    }
    const int x0 = blockIdx.y * (AL ? TX : TX - 1), y0 = blockIdx.x * (AL ? TY : TY - 1);
    const bool interior = !ONE_D && x0 >= 1 && x0 + TX - 1 <= g.nx - 2 && y0 >= 1 && y0 + TY - 1 <= g.ny - 2;
    if (interior)
        k_mult_body<TX, TY, PF, WEIGHTED, ONE_D, UPDATE, false, KKT, AL>(g, tr, kkt_t0, sc, kd, sg, side, maps, qo, qn, alpha, weight, beta,
                                                                         beta_out, q2, rhs, c0, c1, kpart);
    else
        k_mult_body<TX, TY, PF, WEIGHTED, ONE_D, UPDATE, true, KKT, AL>(g, tr, kkt_t0, sc, kd, sg, side, maps, qo, qn, alpha, weight, beta,
                                                                        beta_out, q2, rhs, c0, c1, kpart);
}

// The bx / by sums on the tile boundaries of the aligned march, from the side buffer (same operations, same order as
// phase 2 of k_mult: ((w1n + w2) + w3n') + w4' with ' = the cell layer below, times SF).  Node levels [tn0, tn1) of a slab.
template <int TX, int TY>
__global__ void __launch_bounds__(256) k_q2_fix(Geo g, int tn0, SideGeo sg, i64 nA, const double* __restrict__ side, double SF,
                                                double* __restrict__ q2)
{
    const int t = tn0 + blockIdx.y;
    const bool cell = t < g.nt - 1, below = t > 0;
    i64 idx = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    double a_n, a_o, b_n, b_o;      // neighbour / own value of layer t, neighbour / own value of layer t-1
    i64 dst;
    if (idx < nA) {                 // bx edge between the last row of tile kx and the first row of tile kx+1
        const int kx = (int)(idx / g.py), y = (int)(idx - (i64)kx * g.py);
        const int xb = kx * TX + TX - 1;
        if (xb >= g.nx - 1 || y >= g.ny) return;
        a_n = cell ? side[side_sx(sg, g, t, kx + 1, 0, y)] : 0.0;
        a_o = cell ? side[side_sx(sg, g, t, kx, 2, y)] : 0.0;
        b_n = below ? side[side_sx(sg, g, t - 1, kx + 1, 1, y)] : 0.0;
        b_o = below ? side[side_sx(sg, g, t - 1, kx, 3, y)] : 0.0;
        dst = g.L + (i64)t * g.PBX + (i64)xb * g.py + y;
    } else {                        // by edge between the last column of tile ky and the first column of tile ky+1
        idx -= nA;
        const int ky = (int)(idx / sg.nxp), x = (int)(idx - (i64)ky * sg.nxp);
        const int yb = ky * TY + TY - 1;
        if (ky >= sg.nby || yb >= g.ny - 1 || x >= g.nx) return;
        a_n = cell ? side[side_sy(sg, t, ky + 1, 0, x)] : 0.0;
        a_o = cell ? side[side_sy(sg, t, ky, 2, x)] : 0.0;
        b_n = below ? side[side_sy(sg, t - 1, ky + 1, 1, x)] : 0.0;
        b_o = below ? side[side_sy(sg, t - 1, ky, 3, x)] : 0.0;
        dst = g.L + g.NBX + (i64)t * g.PBY + (i64)x * g.pyb + yb;
    }
    double s;
    if (!below) s = dadd(a_n, a_o);
    else if (!cell) s = dadd(b_n, b_o);
    else s = dadd(dadd(dadd(a_n, a_o), b_n), b_o);
    q2[dst] = dmul(s, SF);
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

int mult_tiles(const Geo& g) { return ((g.ny + KM_TY - 2) / (KM_TY - 1)) * ((g.nx + KM_TX - 2) / (KM_TX - 1)); }

// side buffer of the aligned march for cell layers [t0, t1) (doubles); 0 when the layout / variant does not use it
static SideGeo side_geo(const Geo& g, int t0)
{
    SideGeo sg;
    sg.t0 = t0;
    sg.nbx = (g.nx + KM_TX - 1) / KM_TX;
    sg.nby = (g.ny + KM_TY - 1) / KM_TY;
    sg.nxp = sg.nbx * KM_TX;
    sg.sx_t = (i64)(sg.nbx + 1) * 4 * g.py;      // (+1: k_q2_fix addresses tile kx+1 before it tests the bounds)
    sg.sy_t = (i64)(sg.nby + 1) * 4 * sg.nxp;
    sg.sy_off = 0;
    return sg;
}
bool mult_aligned_ok(const Geo& g, bool one_d) { return !one_d && g.ny > 1 && (g.py % KM_TY) == 0 && (g.pyb % KM_TY) == 0; }
i64 mult_side_doubles(const Geo& g, bool one_d, int nlayers)
{
    if (!mult_aligned_ok(g, one_d)) return 0;
    const SideGeo sg = side_geo(g, 0);
    return (i64)nlayers * (sg.sx_t + sg.sy_t);
}

int make_stag_maps(const Geo& g, const double* base, CUtensorMap out[3], int c_lo, int c_hi, int n_lo, int n_hi)
{
    if (!mult_aligned_ok(g, false) || g.nt < 2 || g.nx < 2 || c_hi <= c_lo || n_hi <= n_lo) return -1;
    const unsigned long long py = g.py, pyb = g.pyb, nc = (unsigned long long)(c_hi - c_lo), nn = (unsigned long long)(n_hi - n_lo);
    const unsigned long long d0[3] = {py, (unsigned long long)g.nx, nc}, s0[2] = {py * 8, (unsigned long long)g.PC * 8};
    const unsigned long long d1[3] = {py, (unsigned long long)(g.nx - 1), nn}, s1[2] = {py * 8, (unsigned long long)g.PBX * 8};
    const unsigned long long d2[3] = {pyb, (unsigned long long)g.nx, nn}, s2[2] = {pyb * 8, (unsigned long long)g.PBY * 8};
    const unsigned b0[3] = {32, KM_TX, 1}, b1[3] = {32, KM_TX + 1, 1}, b2[3] = {KmStage<KM_TX, false>::BYW, KM_TX, 1};
    if (make_tensor_map_f64(&out[0], base + (i64)c_lo * g.PC, 3, d0, s0, b0)) return -1;
    if (make_tensor_map_f64(&out[1], base + g.L + (i64)n_lo * g.PBX, 3, d1, s1, b1)) return -1;
    if (make_tensor_map_f64(&out[2], base + g.L + g.NBX + (i64)n_lo * g.PBY, 3, d2, s2, b2)) return -1;
    return 0;
}
int make_beta_map(const Geo& g, const double* base, CUtensorMap* out, int c_lo, int c_hi)
{
    if (!mult_aligned_ok(g, false) || g.nt < 2 || c_hi <= c_lo) return -1;
    const unsigned long long d[4] = {(unsigned long long)g.py, (unsigned long long)g.nx, (unsigned long long)(c_hi - c_lo), 10};
    const unsigned long long s[3] = {(unsigned long long)g.py * 8, (unsigned long long)g.PC * 8, (unsigned long long)g.L * 8};
    const unsigned b[4] = {32, KM_TX, 1, 10};
    return make_tensor_map_f64(out, base + (i64)c_lo * g.PC, 4, d, s, b);
}

int launch_mult(const UpdateArgs& a, bool weighted, bool one_d, bool update, cudaStream_t st, const KktFused* kkt)
{
    int nlaunch = 1;
    constexpr int TX = KM_TX, TY = KM_TY;
    dim3 block(TY, TX);
    dim3 grid((unsigned)((a.g.ny + TY - 2) / (TY - 1)), (unsigned)((a.g.nx + TX - 2) / (TX - 1)));
    KktDev kd{0, 0, 0, 0, 0};
    double* kpart = nullptr;
    if (kkt) {
        kd.sigma = kkt->sigma;
        kd.scD = (kkt->sigma * kkt->cScale) * kkt->D;
        kd.dD = kkt->dScale / kkt->D;
        kd.sE = kkt->dScale / kkt->E;
        kd.dSD = kkt->dScale / kkt->D;
        kpart = kkt->partial_m;
    }
    // Every CTA marches the same number of time steps, so a grid that fills the machine 4.25 times runs for 5 full rounds.
    // Cutting the time range into pieces (grid.z) lets the CTA count land just below a whole number of rounds; each extra
    // piece costs one replayed cell layer.  DOTSOCP_KM_CHUNKS=n forces n pieces.
    static const int forced = [] { const char* e = getenv("DOTSOCP_KM_CHUNKS"); return e ? atoi(e) : 0; }();
    const char* pf_env = getenv("DOTSOCP_KM_PF");   // read per launch: tests and A/B runs switch it inside one process
    int pf = pf_env ? atoi(pf_env) : KM_PF;
    if (pf == 4 && a.maps == nullptr) pf = 1;
    static const KmMaps no_maps = KmMaps();
    const KmMaps& maps = a.maps ? *a.maps : no_maps;
    // aligned tiling (pitched layout, side buffer present, plain update kernel); DOTSOCP_KM_AL=0 keeps the haloed tiling
    const char* al_env = getenv("DOTSOCP_KM_AL");
    const bool al = update && !kkt && a.side != nullptr && mult_aligned_ok(a.g, one_d) && !(al_env && al_env[0] == '0');
    SideGeo sg = side_geo(a.g, a.side_t0);
    sg.sy_off = (i64)a.side_layers * sg.sx_t;
    if (al) grid = dim3((unsigned)((a.g.ny + TY - 1) / TY), (unsigned)((a.g.nx + TX - 1) / TX));
#define KM(PF, W, O, U, K, AL)                                                                                        \
    {                                                                                                                 \
        constexpr size_t smem = (size_t)2 * (K ? 9 : 4) * TX * TY * sizeof(double) + (AL ? (size_t)2 * 4 * TX * sizeof(double) : 0) + \
                                (PF >= 2 ? (size_t)(PF == 3 ? 3 : 2) * KmStage<TX, W>::SIZE * sizeof(double) + 32 + 128 : 0); \
        static int slots = 0;                                                                                         \
        if (!slots) {                                                                                                 \
            cudaFuncSetAttribute(k_mult<TX, TY, PF, W, O, U, K, AL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            int per_sm = 0, dev = 0, sms = 0;                                                                         \
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_mult<TX, TY, PF, W, O, U, K, AL>, TX * TY, smem); \
            cudaGetDevice(&dev);                                                                                      \
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);                                        \
            slots = per_sm > 0 && sms > 0 ? per_sm * sms : 1;                                                         \
        }                                                                                                             \
        const int nchunk = forced > 0 ? std::min(forced, std::max(1, a.tr.tc1 - a.tr.tc0))                            \
                                      : km_pick_chunks((long long)grid.x * grid.y, a.tr.tc1 - a.tr.tc0, slots);       \
        grid.z = (unsigned)nchunk;                                                                                    \
        k_mult<TX, TY, PF, W, O, U, K, AL><<<grid, block, smem, st>>>(a.g, a.tr, nchunk, a.sc, kd, sg, a.side, maps, a.q_old, a.q_new, \
                                                                      a.alpha, a.weight, a.beta_in, a.beta_out, a.q2, a.rhs, a.c0, a.c1, kpart); \
    }
    if (kkt && update) {
        if (one_d) KM(0, false, true, true, true, false) else if (weighted) KM(0, true, false, true, true, false) else KM(0, false, false, true, true, false)
    } else if (one_d) {
        if (update) KM(0, false, true, true, false, false) else KM(0, false, true, false, false, false)
    } else if (al) {
        if (weighted) { if (pf == 4) KM(4, true, false, true, false, true) else if (pf == 3) KM(3, true, false, true, false, true) else if (pf == 2) KM(2, true, false, true, false, true) else if (pf == 1) KM(1, true, false, true, false, true) else KM(0, true, false, true, false, true) }
        else { if (pf == 4) KM(4, false, false, true, false, true) else if (pf == 3) KM(3, false, false, true, false, true) else if (pf == 2) KM(2, false, false, true, false, true) else if (pf == 1) KM(1, false, false, true, false, true) else KM(0, false, false, true, false, true) }
        // the edges on the tile boundaries, from the side buffer
        const int nl = a.tr.tn1 - a.tr.tn0;
        const i64 nA = (i64)sg.nbx * a.g.py, nB = (i64)sg.nby * sg.nxp;
        if (nl > 0) {
            dim3 fg((unsigned)((nA + nB + 255) / 256), (unsigned)nl);
            k_q2_fix<TX, TY><<<fg, 256, 0, st>>>(a.g, a.tr.tn0, sg, nA, a.side, a.sc.SF, a.q2);
            nlaunch = 2;
        }
    } else if (weighted) {
        if (!update) KM(0, true, false, false, false, false)
        else if (pf == 1) KM(1, true, false, true, false, false) else KM(0, true, false, true, false, false)
    } else {
        if (!update) KM(0, false, false, false, false, false)
        else if (pf == 1) KM(1, false, false, true, false, false) else KM(0, false, false, true, false, false)
    }
#undef KM
    return nlaunch;
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
    const PlanePos pp = plane_pos(g, p);
    if (!pp.ok) return;
    const int x = pp.x, y = pp.y;
    const i64 L = g.L, n = (i64)t * g.PC + p;      // cell (t,x,y) == q0 edge above the node
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
    if (dn) ADDTERM(dmul(sc.gt, u(n - g.PC)));
    if (up) ADDTERM(dmul(-sc.gt, u(n)));
    const i64 ox = L + (i64)t * g.PBX + p, oy = L + g.NBX + (i64)t * g.PBY + (i64)x * g.pyb + y;
    if (x > 0) ADDTERM(dmul(sc.gx, u(ox - g.py)));
    if (x < g.nx - 1) ADDTERM(dmul(-sc.gx, u(ox)));
    if (y > 0) ADDTERM(dmul(sc.gy, u(oy - 1)));
    if (y < g.ny - 1) ADDTERM(dmul(-sc.gy, u(oy)));
#undef ADDTERM
    double cv = 0.0;
    if (t == 0) cv = c0[pp.pn];
    else if (!up) cv = c1[pp.pn];
    rhs[(i64)t * g.P + pp.pn] = dadd(first ? 0.0 : acc, cv);
}
void launch_rhs(const Geo& g, const IterScal& sc, bool weighted, const double* q, const double* alpha, const double* weight,
                const double* c0, const double* c1, double* rhs, cudaStream_t st)
{
    dim3 grid((unsigned)((g.PC + 255) / 256), (unsigned)g.nt);
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
    const PlanePos pp = plane_pos(g, p);
    if (!pp.ok) return;
    const int x = pp.x, y = pp.y;
    const bool hxm = x > 0, hxp = x < g.nx - 1, hym = y > 0, hyp = y < g.ny - 1;
    const i64 L = g.L, c = (i64)t * g.PC + p;
    const double* bx = q + L;
    const double* by = bx + g.NBX;
    const i64 ox = (i64)t * g.PBX + p, oy = (i64)t * g.PBY + (i64)x * g.pyb + y;
    CellQ cq;
    cq.q0 = q[c];
    cq.bxm = hxm ? bx[ox - g.py] : 0.0;
    cq.bx = hxp ? bx[ox] : 0.0;
    cq.bxm1 = hxm ? bx[ox + g.PBX - g.py] : 0.0;
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
    dim3 grid((unsigned)((g.PC + 255) / 256), (unsigned)(g.nt - 1));
#define CUK(O, M) k_cells_update<O, M><<<grid, 256, 0, st>>>(g, sc, q, z, beta)
    if (one_d) { if (mode == 0) CUK(true, 0); else if (mode == 1) CUK(true, 1); else CUK(true, 2); }
    else { if (mode == 0) CUK(false, 0); else if (mode == 1) CUK(false, 1); else CUK(false, 2); }
#undef CUK
}

// ---------------------------------------------------------------------------------------------------------------
// Deterministic reductions: every CTA writes its K partial sums, a single CTA adds them in a fixed tree order.
// ---------------------------------------------------------------------------------------------------------------
// stage 2: one CTA per time level adds the level's `nb` CTA partials (K sums each) in a fixed order and stores them in
// slots slot[0..K) of row t0 + blockIdx.x of the level table
struct SlotMap { int n; int slot[16]; };
__global__ void __launch_bounds__(256) k_level_reduce(const double* __restrict__ partial, int nb, int K, SlotMap sm, int t0,
                                                      double* __restrict__ lvl)
{
    __shared__ double red[256];
    const double* row = partial + (i64)blockIdx.x * nb * K;
    for (int k = 0; k < K; k++) {
        double v = 0.0;
        for (int i = threadIdx.x; i < nb; i += 256) v += row[(i64)i * K + k];
        red[threadIdx.x] = v;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) {
            if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
            __syncthreads();
        }
        if (threadIdx.x == 0) lvl[(i64)(t0 + blockIdx.x) * KSL + sm.slot[k]] = red[0];
        __syncthreads();
    }
}
// stage 3: out[k] = sum over the nt rows of the level table, fixed order (depends on nt only)
__global__ void __launch_bounds__(256) k_levels_total(const double* __restrict__ lvl, int nt, double* __restrict__ out)
{
    __shared__ double red[256];
    for (int k = 0; k < KSL; k++) {
        double v = 0.0;
        for (int t = threadIdx.x; t < nt; t += 256) v += lvl[(i64)t * KSL + k];
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
void launch_levels_total(const double* lvl, int nt, double* out, cudaStream_t st) { k_levels_total<<<1, 256, 0, st>>>(lvl, nt, out); }
void level_reduce(const double* partial, int nb, int K, const int* slots, int t0, int nlev, double* lvl, cudaStream_t st)
{
    if (nlev <= 0) return;
    SlotMap sm;
    sm.n = K;
    for (int k = 0; k < K; k++) sm.slot[k] = slots[k];
    k_level_reduce<<<nlev, 256, 0, st>>>(partial, nb, K, sm, t0, lvl);
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
    const PlanePos pp = plane_pos(g, p);
    if (pp.ok) {
        const int x = pp.x, y = pp.y;
        const bool hxm = x > 0, hxp = x < g.nx - 1, hym = y > 0, hyp = y < g.ny - 1;
        const i64 L = g.L, c = (i64)t * g.PC + p;
        const double* bx = a.q + L;
        const double* by = bx + g.NBX;
        CellQ cq;
        cq.q0 = a.q[c];
        const i64 ox = (i64)t * g.PBX + p, oy = (i64)t * g.PBY + (i64)x * g.pyb + y;
        cq.bxm = hxm ? bx[ox - g.py] : 0.0;
        cq.bx = hxp ? bx[ox] : 0.0;
        cq.bxm1 = hxm ? bx[ox + g.PBX - g.py] : 0.0;
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

int kkt_cells_blocks(const Geo& g) { return (int)((g.PC + 255) / 256) * (g.nt - 1); }
int kkt_nodes_blocks(const Geo& g) { return (int)((g.PC + 255) / 256) * g.nt; }
static int blocks_x(const Geo& g) { return (int)((g.PC + 255) / 256); }
int kkt_blocks_x(const Geo& g) { return blocks_x(g); }

void launch_kkt_cells(const KktArgs& a, bool weighted, bool one_d, cudaStream_t st)
{
    const int nl = a.tr.tc1 - a.tr.tc0;
    if (nl <= 0) return;
    dim3 grid((unsigned)blocks_x(a.g), (unsigned)nl);
    if (one_d)
        k_kkt_cells<false, true><<<grid, 256, 0, st>>>(a);
    else if (weighted)
        k_kkt_cells<true, false><<<grid, 256, 0, st>>>(a);
    else
        k_kkt_cells<false, false><<<grid, 256, 0, st>>>(a);
    int slots[KC_COUNT];
    for (int k = 0; k < KC_COUNT; k++) slots[k] = k;
    level_reduce(a.partial, blocks_x(a.g), KC_COUNT, slots, a.tr.tc0, nl, a.lvl, st);
}

// ---------------------------------------------------------------------------------------------------------------
// z = Pi_Q(d + BF q_old - beta_old) materialised (final output, :334) and/or its squared Frobenius norm (rescale
// block, :141).  zout may alias beta_old (cell-local read-then-write).
// ---------------------------------------------------------------------------------------------------------------
template <bool ONE_D>
__global__ void __launch_bounds__(256) k_zstep(Geo g, int tc0, IterScal sc, const double* q_old, const double* beta_old, double* zout)
{
    const i64 p = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    const int t = tc0 + blockIdx.y;
    const PlanePos pp = plane_pos(g, p);
    if (!pp.ok) return;
    const int x = pp.x, y = pp.y;
    double z[10];
    load_z<ONE_D>(g, sc, nullptr, q_old, beta_old, t, x, y, z);
    const i64 c = (i64)t * g.PC + p;
#pragma unroll
    for (int j = 0; j < 10; j++)
        if (!(ONE_D && j >= 5 && j <= 8)) zout[(i64)j * g.L + c] = z[j];
}

void launch_zstep(const Geo& g, const IterScal& sc, bool one_d, const double* q_old, const double* beta_old, double* zout,
                  cudaStream_t st, const TRange* tr)
{
    const int tc0 = tr ? tr->tc0 : 0, tc1 = tr ? tr->tc1 : g.nt - 1;
    if (tc1 <= tc0) return;
    dim3 grid((unsigned)blocks_x(g), (unsigned)(tc1 - tc0));
    if (one_d)
        k_zstep<true><<<grid, 256, 0, st>>>(g, tc0, sc, q_old, beta_old, zout);
    else
        k_zstep<false><<<grid, 256, 0, st>>>(g, tc0, sc, q_old, beta_old, zout);
}

// ---------------------------------------------------------------------------------------------------------------
// Rescale norms (solver_socp_inPALM.m:140-143): sum of squares of phi, q, z, alpha, beta in ONE pass, one thread per
// node (t,x,y) that owns phi[t,x,y], the edges q0/bx/by[t,x,y] and the cell (t,x,y); z is read or recomputed (load_z).
// ---------------------------------------------------------------------------------------------------------------
template <bool ONE_D>
__global__ void __launch_bounds__(256) k_norms(KktArgs a)
{
    const Geo& g = a.g;
    const i64 p = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    const int t = a.tr.tn0 + blockIdx.y;
    double s[NR_COUNT];
#pragma unroll
    for (int k = 0; k < NR_COUNT; k++) s[k] = 0.0;
    const PlanePos pp = plane_pos(g, p);
    if (pp.ok) {
        const int x = pp.x, y = pp.y;
        const i64 L = g.L, n = (i64)t * g.PC + p;      // cell / q0 index
        const double ph = a.phi[(i64)t * g.P + pp.pn];
        s[NR_PHI2] = ph * ph;
        auto edge = [&](i64 e) {
            const double qv = a.q[e], av = a.alpha[e];
            s[NR_Q2] += qv * qv;
            s[NR_ALPHA2] += av * av;
        };
        if (t < g.nt - 1) {
            edge(n);
            double z[10];
            load_z<ONE_D>(g, a.sc, a.z, a.q_old, a.beta_old, t, x, y, z);
#pragma unroll
            for (int j = 0; j < 10; j++) {
                const double b = (ONE_D && j >= 5 && j <= 8) ? 0.0 : a.beta[(i64)j * L + n];
                s[NR_Z2] += z[j] * z[j];
                s[NR_BETA2] += b * b;
            }
        }
        if (x < g.nx - 1) edge(L + (i64)t * g.PBX + p);
        if (y < g.ny - 1) edge(L + g.NBX + (i64)t * g.PBY + (i64)x * g.pyb + y);
    }
    block_reduce_store<NR_COUNT, 256>(s, a.partial);
}
void launch_norms(const KktArgs& a, bool one_d, cudaStream_t st)
{
    const int nl = a.tr.tn1 - a.tr.tn0;
    if (nl <= 0) return;
    dim3 grid((unsigned)blocks_x(a.g), (unsigned)nl);
    if (one_d) k_norms<true><<<grid, 256, 0, st>>>(a);
    else k_norms<false><<<grid, 256, 0, st>>>(a);
    int slots[NR_COUNT];
    for (int k = 0; k < NR_COUNT; k++) slots[k] = k;
    level_reduce(a.partial, blocks_x(a.g), NR_COUNT, slots, a.tr.tn0, nl, a.lvl, st);
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
    const PlanePos ppos = plane_pos(g, p);
    if (ppos.ok) {
        const int x = ppos.x, y = ppos.y;
        const i64 L = g.L, n = (i64)t * g.P + ppos.pn;   // node index
        const i64 ce = (i64)t * g.PC + p;                // cell / q0 index
        const bool up = t < g.nt - 1, dn = t > 0;
        const bool hxm = x > 0, hxp = x < g.nx - 1, hym = y > 0, hyp = y < g.ny - 1;
        const IterScal& sc = a.sc;
        const double ph = a.phi[n];
        const double scD = dmul(dmul(a.sigma, a.cScale), a.D);   // sigma*cScale*D
        const double dD = a.dScale / a.D;
        // Dalpha on the t-edges around this node and its x+1 / y+1 neighbours (for rho = pair-average in t)
        auto dal0 = [&](i64 c) -> double { return WEIGHTED ? dmul(a.weight[c], a.alpha[c]) : a.alpha[c]; };
        auto rho_at = [&](i64 pp) -> double {   // rho(t, node pp) = (rhoT[t-1] + rhoT[t]) / 2 with zero padding
            const double lo = dn ? dmul(scD, dal0((i64)(t - 1) * g.PC + pp)) : 0.0;
            const double hi = up ? dmul(scD, dal0((i64)t * g.PC + pp)) : 0.0;
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
        if (up) edge(ce, dadd(dmul(-sc.gt, ph), dmul(sc.gt, a.phi[n + g.P])), false, 0.0);
        if (hxp)
            edge(L + (i64)t * g.PBX + p, dadd(dmul(-sc.gx, ph), dmul(sc.gx, a.phi[n + g.ny])), true,
                 dadd(rho_c, rho_at(p + g.py)) / 2.0);
        if (hyp)
            edge(L + g.NBX + (i64)t * g.PBY + (i64)x * g.pyb + y, dadd(dmul(-sc.gy, ph), dmul(sc.gy, a.phi[n + 1])),
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
        if (dn) ADDTERM(dmul(sc.gt, a.alpha[ce - g.PC]));
        if (up) ADDTERM(dmul(-sc.gt, a.alpha[ce]));
        const i64 ox = (i64)t * g.PBX + p, oy = (i64)t * g.PBY + (i64)x * g.pyb + y;
        if (hxm) ADDTERM(dmul(sc.gx, al_bx[ox - g.py]));
        if (hxp) ADDTERM(dmul(-sc.gx, al_bx[ox]));
        if (hym) ADDTERM(dmul(sc.gy, al_by[oy - 1]));
        if (hyp) ADDTERM(dmul(-sc.gy, al_by[oy]));
#undef ADDTERM
        double cv = 0.0;
        if (t == 0) cv = a.c0[ppos.pn];
        else if (!up) cv = a.c1[ppos.pn];
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
    if (nl <= 0) return;
    dim3 grid((unsigned)blocks_x(a.g), (unsigned)nl);
    if (weighted)
        k_kkt_nodes<true><<<grid, 256, 0, st>>>(a);
    else
        k_kkt_nodes<false><<<grid, 256, 0, st>>>(a);
    int slots[KN_COUNT];
    for (int k = 0; k < KN_COUNT; k++) slots[k] = KC_COUNT + k;
    level_reduce(a.partial, blocks_x(a.g), KN_COUNT, slots, a.tr.tn0, nl, a.lvl, st);
}

// fused check: the partials left by k_qstep<KKT> and k_mult<KKT> -> the same slots of the level table
void launch_kkt_fused_reduce(const Geo& g, const TRange& tr, const KktFused& k, cudaStream_t st)
{
    const int nl = tr.tn1 - tr.tn0;
    static const int qs[KQ_COUNT] = {KC_COUNT + KN_Q2, KC_COUNT + KN_APHI2, KC_COUNT + KN_PRIM1, KC_COUNT + KN_ALPHA2,
                                     KC_COUNT + KN_QDOTA, KC_COUNT + KN_CPHI, KC_COUNT + KN_PHI2};
    static const int ms[KM_COUNT] = {KC_Z2, KC_BETA2, KC_PRIM2, KC_COMPL, KC_DOTC, KC_RHOT, KC_RHOFQ, KC_COUNT + KN_FBB2,
                                     KC_COUNT + KN_DUAL2, KC_COUNT + KN_DUAL1, KC_COUNT + KN_MRHOB, KC_COUNT + KN_M2,
                                     KC_COUNT + KN_RHOB2};
    level_reduce(k.partial_q, blocks_x(g), KQ_COUNT, qs, tr.tn0, nl, k.lvl, st);
    level_reduce(k.partial_m, mult_tiles(g), KM_COUNT, ms, tr.tn0, nl, k.lvl, st);
}

// ---------------------------------------------------------------------------------------------------------------
// small streaming helpers
// ---------------------------------------------------------------------------------------------------------------
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

__global__ void __launch_bounds__(256) k_fill(double* __restrict__ x, i64 n, double v)
{
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) x[i] = v;
}
void launch_fill(double* x, i64 n, double v, cudaStream_t st)
{
    if (n <= 0) return;
    i64 b = (n + 255) / 256;
    if (b > 148 * 16) b = 148 * 16;
    k_fill<<<(unsigned)b, 256, 0, st>>>(x, n, v);
}

// debugging aid (DOTSOCP_DEBUG_SUMS): sum of |x| and number of non-finite entries of a whole array
__global__ void __launch_bounds__(256) k_dbg_sum(const double* __restrict__ x, i64 n, double* __restrict__ out)
{
    double s = 0.0, bad = 0.0;
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        const double v = x[i];
        if (isfinite(v)) s += fabs(v); else bad += 1.0;
    }
    for (int o = 16; o > 0; o >>= 1) { s += __shfl_down_sync(0xffffffffu, s, o); bad += __shfl_down_sync(0xffffffffu, bad, o); }
    if ((threadIdx.x & 31) == 0) { atomicAdd(&out[0], s); atomicAdd(&out[1], bad); }
}
// no host synchronisation at the call (the sums land in a device log); debug_sum_flush prints what was collected
static double* g_dbg_log = nullptr;
static int g_dbg_n = 0;
static char g_dbg_names[256][48];
void debug_sum(const char* name, const double* x, i64 n, cudaStream_t st)
{
    if (!g_dbg_log) { cudaMalloc(&g_dbg_log, 256 * 16); cudaMemset(g_dbg_log, 0, 256 * 16); cudaDeviceSynchronize(); }
    if (g_dbg_n >= 256) return;
    snprintf(g_dbg_names[g_dbg_n], 48, "%s", name);
    if (n > 0) k_dbg_sum<<<148 * 8, 256, 0, st>>>(x, n, g_dbg_log + 2 * g_dbg_n);
    g_dbg_n++;
}
void debug_sum_flush(cudaStream_t st)
{
    if (!g_dbg_log || g_dbg_n == 0) return;
    std::vector<double> h(2 * g_dbg_n);
    cudaStreamSynchronize(st);
    cudaMemcpy(h.data(), g_dbg_log, h.size() * sizeof(double), cudaMemcpyDeviceToHost);
    for (int i = 0; i < g_dbg_n; i++)
        fprintf(stderr, "[dotsocp debug] %-40s sum|x|=%.15e nonfinite=%.0f\n", g_dbg_names[i], h[2 * i], h[2 * i + 1]);
    cudaMemset(g_dbg_log, 0, 256 * 16);
    g_dbg_n = 0;
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
