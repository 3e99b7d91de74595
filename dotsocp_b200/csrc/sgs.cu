// sgs.cu -- red-black symmetric Gauss-Seidel for the Neumann Poisson equation of the phi-step.
//
// Reference: mexsGS.mexa64 (binary only; internal name mexRBsGSscaling), called as mexsGS(phi, rhs, ep, scale, nt, nx, ny,
// its) at solver_socp_sGSinPALM.m:205 and solver_socp_accsGSADMM.m:256.  Semantics from the disassembly (mexFunction @0x23a0,
// RBGS_inside @0x1380, RBGS_face @0x1c80, RBGS_edge @0x17a0, RBGS_corner @0x1530; restated and checked bit for bit against
// the binary in oracle/kernels.py:_np_sGS):
//     H = (1/(nx-1))^2 / scale ;  C = ((nt-1)/(nx-1))^2 ;  eH = H * (scale*ep)
//     COE(node) = 1 / ((n_t*C + n_s) + eH),  n_t in {1,2} time neighbours (2C = C + C), n_s in {2,3,4} space neighbours
//     half sweep of one parity of t+x+y:  phi <- ((S + T) + H*rhs) * COE,
//         S = left-to-right sum of the existing neighbours in the order x-1, x+1, y-1, y+1
//         T = C*(phi[t-1] + phi[t+1])  or  C*phi[the one time neighbour]
//     order of the half sweeps: odd ; its x [ even ; odd ].
// The binary is only meaningful for nx == ny with odd node counts (its row walk toggles the start 1 <-> 2 and uses NY-based
// counts on all faces); the launcher refuses anything else.  A node's six neighbours all have the other parity, so the nodes
// of a half sweep are independent: one thread per node, each half sweep one launch, bit-identical to the sequential binary.
// Time slabs exchange one ghost plane per side between half sweeps -- no transposes, unlike the DCT solve.
// Compiled with -fmad=false.
#include "kernels.h"
#include "reduce.cuh"

namespace dsocp {

struct SgsCoef {
    double C, H;
    double coe[2][3];   // [n_t - 1][n_s - 2]
};

__global__ void __launch_bounds__(256) k_sgs_half(Geo g, int tn0, int parity, SgsCoef k, const double* __restrict__ rhs, double* phi)
{
    const i64 p = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    const int t = tn0 + blockIdx.y;
    if (p >= g.P) return;
    const int x = (int)(p / g.ny), y = (int)(p - (i64)x * g.ny);
    if (((t + x + y) & 1) != parity) return;
    const i64 n = (i64)t * g.P + p;
    const bool hxm = x > 0, hxp = x < g.nx - 1, hym = y > 0, hyp = y < g.ny - 1, dn = t > 0, up = t < g.nt - 1;
    double S = 0.0;
    bool first = true;
#define ADDN(cond, idx)                          \
    if (cond) {                                  \
        const double v_ = phi[idx];              \
        S = first ? v_ : dadd(S, v_);            \
        first = false;                           \
    }
    ADDN(hxm, n - g.ny)
    ADDN(hxp, n + g.ny)
    ADDN(hym, n - 1)
    ADDN(hyp, n + 1)
#undef ADDN
    double T;
    if (dn && up) T = dmul(dadd(phi[n - g.P], phi[n + g.P]), k.C);
    else T = dmul(dn ? phi[n - g.P] : phi[n + g.P], k.C);
    const int ns = (int)hxm + (int)hxp + (int)hym + (int)hyp;
    const double coe = k.coe[(dn && up) ? 1 : 0][ns - 2];
    phi[n] = dmul(dadd(dadd(S, T), dmul(rhs[n], k.H)), coe);
}

static SgsCoef sgs_coef(const Geo& g, double ep, double scale)
{
    SgsCoef k;
    const double hx = 1.0 / ((double)g.nx - 1.0);
    k.H = (hx * hx) / scale;
    const double c1 = ((double)g.nt - 1.0) / ((double)g.nx - 1.0);
    k.C = c1 * c1;
    const double eH = k.H * (scale * ep);
    const double twoC = k.C + k.C;
    for (int nt = 0; nt < 2; nt++)
        for (int ns = 0; ns < 3; ns++) k.coe[nt][ns] = 1.0 / (((nt ? twoC : k.C) + (double)(ns + 2)) + eH);
    return k;
}

bool sgs_supported(const Geo& g) { return g.nx == g.ny && (g.nx & 1) && (g.nt & 1) && g.nx >= 3 && g.nt >= 3; }

// one half sweep (parity of t+x+y) on node levels [tn0, tn1)
void launch_sgs_half(const Geo& g, double ep, double scale, int parity, const double* rhs, double* phi, int tn0, int tn1, cudaStream_t st)
{
    if (tn1 <= tn0) return;
    dim3 grid((unsigned)((g.P + 255) / 256), (unsigned)(tn1 - tn0));
    k_sgs_half<<<grid, 256, 0, st>>>(g, tn0, parity, sgs_coef(g, ep, scale), rhs, phi);
}

// ---------------------------------------------------------------------------------------------------------------
// Block residuals of the sGS loops (solver_socp_sGSinPALM.m:212-216, :322): per node  v = A'(A phi - q + alpha) - c
// (WITH_ALPHA_C) or  v = A'(A phi - q), summed as v^2 over the nodes with even linear index (1:2:end in MATLAB = even
// t+x+y on these odd grids) or over all nodes.  CSR row orders as everywhere: A phi = (-g) phi_i + g phi_{i+1},
// A' u = sum over (t-1, t, x-1, x, y-1, y) edges.
// ---------------------------------------------------------------------------------------------------------------
template <bool WITH_ALPHA_C>
__global__ void __launch_bounds__(256) k_sgs_resid(Geo g, int tn0, IterScal sc, const double* __restrict__ phi,
                                                   const double* __restrict__ q, const double* __restrict__ alpha,
                                                   const double* __restrict__ c0, const double* __restrict__ c1,
                                                   double* __restrict__ partial)
{
    const i64 p = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    const int t = tn0 + blockIdx.y;
    double s[1] = {0.0};
    if (p < g.P) {
        const int x = (int)(p / g.ny), y = (int)(p - (i64)x * g.ny);
        if (!WITH_ALPHA_C || (((t + x + y) & 1) == 0)) {
            const i64 L = g.L, n = (i64)t * g.P + p;
            const i64 ce = (i64)t * g.PC + (i64)x * g.py + y;   // q0 edge above the node (pitched staggered arrays)
            const double ph = phi[n];
            // u on edge e between node a (lower) and node b: ((-g) phi_a + g phi_b) - q_e [+ alpha_e]
            auto u = [&](i64 e, double gr, double pa, double pb) -> double {
                double v = dsub(dadd(dmul(-gr, pa), dmul(gr, pb)), q[e]);
                if (WITH_ALPHA_C) v = dadd(v, alpha[e]);
                return v;
            };
            double acc = 0.0;
            bool first = true;
#define ADDTERM(val)                         \
    {                                        \
        const double tv_ = (val);            \
        acc = first ? tv_ : dadd(acc, tv_);  \
        first = false;                       \
    }
            if (t > 0) ADDTERM(dmul(sc.gt, u(ce - g.PC, sc.gt, phi[n - g.P], ph)));
            if (t < g.nt - 1) ADDTERM(dmul(-sc.gt, u(ce, sc.gt, ph, phi[n + g.P])));
            const i64 ox = L + (i64)t * g.PBX + (i64)x * g.py + y, oy = L + g.NBX + (i64)t * g.PBY + (i64)x * g.pyb + y;
            if (x > 0) ADDTERM(dmul(sc.gx, u(ox - g.py, sc.gx, phi[n - g.ny], ph)));
            if (x < g.nx - 1) ADDTERM(dmul(-sc.gx, u(ox, sc.gx, ph, phi[n + g.ny])));
            if (y > 0) ADDTERM(dmul(sc.gy, u(oy - 1, sc.gy, phi[n - 1], ph)));
            if (y < g.ny - 1) ADDTERM(dmul(-sc.gy, u(oy, sc.gy, ph, phi[n + 1])));
#undef ADDTERM
            double v = first ? 0.0 : acc;
            if (WITH_ALPHA_C) {
                double cv = 0.0;
                if (t == 0) cv = c0[p];
                else if (t == g.nt - 1) cv = c1[p];
                v = dsub(v, cv);
            }
            s[0] = v * v;
        }
    }
    block_reduce_store<1, 256>(s, partial);
}

// rows tn0..tn1-1 of the level table, slot `slot`
void launch_sgs_resid(const Geo& g, const IterScal& sc, bool with_alpha_c, const double* phi, const double* q, const double* alpha,
                      const double* c0, const double* c1, double* partial, double* lvl, int slot, int tn0, int tn1, cudaStream_t st)
{
    if (tn1 <= tn0) return;
    dim3 grid((unsigned)((g.P + 255) / 256), (unsigned)(tn1 - tn0));
    if (with_alpha_c) k_sgs_resid<true><<<grid, 256, 0, st>>>(g, tn0, sc, phi, q, alpha, c0, c1, partial);
    else k_sgs_resid<false><<<grid, 256, 0, st>>>(g, tn0, sc, phi, q, alpha, c0, c1, partial);
    level_reduce(partial, (int)grid.x, 1, &slot, tn0, tn1 - tn0, lvl, st);
}

// sum of phi per level (slot 0) and phi -= shift  (phi = phi - integralL2(phi, h), solver_socp_sGSinPALM.m:142)
__global__ void __launch_bounds__(256) k_sum_nodes(Geo g, int tn0, const double* __restrict__ phi, double* __restrict__ partial)
{
    const i64 p = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    const int t = tn0 + blockIdx.y;
    double s[1] = {p < g.P ? phi[(i64)t * g.P + p] : 0.0};
    block_reduce_store<1, 256>(s, partial);
}
__global__ void __launch_bounds__(256) k_shift(double* __restrict__ x, i64 n, double shift)
{
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) x[i] = dsub(x[i], shift);
}
void launch_sum_nodes(const Geo& g, const double* phi, double* partial, double* lvl, int slot, int tn0, int tn1, cudaStream_t st)
{
    if (tn1 <= tn0) return;
    dim3 grid((unsigned)((g.P + 255) / 256), (unsigned)(tn1 - tn0));
    k_sum_nodes<<<grid, 256, 0, st>>>(g, tn0, phi, partial);
    level_reduce(partial, (int)grid.x, 1, &slot, tn0, tn1 - tn0, lvl, st);
}
void launch_shift(double* x, i64 n, double shift, cudaStream_t st)
{
    if (n <= 0) return;
    i64 b = (n + 255) / 256;
    if (b > 148 * 16) b = 148 * 16;
    k_shift<<<(unsigned)b, 256, 0, st>>>(x, n, shift);
}

}  // namespace dsocp
