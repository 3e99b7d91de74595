// kernels.h -- host-side launch wrappers of the CUDA kernels (definitions in kernels_update.cu / poisson.cu).
#pragma once
#include "common.cuh"

#include <vector>

namespace dsocp {

// owned part of a time slab: cell layers [tc0, tc1), node levels [tn0, tn1) (the last slab also owns level nt-1)
struct TRange { int tc0, tc1, tn0, tn1; };
inline TRange full_range(const Geo& g) { return TRange{0, g.nt - 1, 0, g.nt}; }

// ---- standalone counterparts of the reference's MEX kernels (device pointers) -------------------------------
void launch_bfd(const Geo& g, double S, double DF, const double* q, double* z, cudaStream_t st);
void launch_bfdconj(const Geo& g, double S, const double* z, double* q2, cudaStream_t st, const TRange* tr = nullptr);
void launch_bfdconj_sum(const Geo& g, double S, const double* za, const double* zb, double* q2, cudaStream_t st);
void launch_projsoc(i64 M, int N, const double* in, double* out, cudaStream_t st);

// ---- fused iteration kernels --------------------------------------------------------------------------------
struct UpdateArgs {
    Geo g;
    TRange tr;
    IterScal sc;
    const double* phi;
    const double* q_old;    // q of the previous iterate (input of the z-step), UPDATE only
    double* q_new;          // q-step output (k_qstep writes it, k_mult reads it)
    double* alpha;          // in/out
    const double* weight;   // NULL unless weighted
    const double* beta_in;  // L x 10, multiplier before the step
    double* beta_out;       // L x 10, multiplier after the step (ping-pong partner: the halo cells of neighbouring
                            // CTAs re-read beta_in, so the update cannot be in place)
    double* q2;             // in (k_qstep) / out (k_mult): s (BF)^*(z + beta) of the NEXT z-step
    double* rhs;            // out: A'(w.*q - alpha) + c for the next Poisson solve
    const double* c0;       // c on the first time level (nx*ny)
    const double* c1;       // c on the last time level
};
// q_new = ((A phi + alpha) + q2) .* diagQInv ; alpha += tau (A phi - q_new)      (solver_socp_inPALM.m:204-214)
// acc: alpha = (alpha + A phi) - q_new                                            (solver_socp_accADMM.m:237)
// tmpq_in != NULL: take A*phi from that buffer instead of phi; tmpq_out != NULL: also store A*phi; upd_alpha=false: q only
void launch_qstep(const UpdateArgs& a, bool weighted, bool acc, cudaStream_t st, const double* tmpq_in = nullptr,
                  double* tmpq_out = nullptr, bool upd_alpha = true);
void launch_rhs(const Geo& g, const IterScal& sc, bool weighted, const double* q, const double* alpha, const double* weight,
                const double* c0, const double* c1, double* rhs, cudaStream_t st);
// mode 0: beta += tau (z - z2(q)) ; mode 1: beta = (beta + z) - z2(q), z = Pi_Q(z2(q) - beta) ; mode 2: z = d + BF q
void launch_cells_update(const Geo& g, const IterScal& sc, bool one_d, int mode, const double* q, double* z, double* beta,
                         cudaStream_t st);
// z = Pi_Q(d + BF q_old - beta) ; beta += tau (z - (d + BF q_new)) ; then q2, rhs of the next iteration.
// update=false ("prologue"): no multiplier step, only q2/rhs from the current (q_new, alpha, beta).
void launch_mult(const UpdateArgs& a, bool weighted, bool one_d, bool update, cudaStream_t st);
// z = Pi_Q(d + BF q_old - beta_old): optional store (zout may alias beta_old) and out[0] = sum z^2
void launch_zstep(const Geo& g, const IterScal& sc, bool one_d, const double* q_old, const double* beta_old, double* zout,
                  double* partial, double* out, cudaStream_t st, const TRange* tr = nullptr);

// ---- KKT / norms --------------------------------------------------------------------------------------------
enum { KC_Z2 = 0, KC_BETA2, KC_PRIM2, KC_COMPL, KC_DOTC, KC_RHOT, KC_RHOFQ, KC_COUNT };
enum { KN_Q2 = 0, KN_APHI2, KN_PRIM1, KN_ALPHA2, KN_FBB2, KN_DUAL2, KN_QDOTA, KN_MRHOB, KN_M2, KN_RHOB2,
       KN_DUAL1, KN_CPHI, KN_PHI2, KN_COUNT };
struct KktArgs {
    Geo g;
    TRange tr;
    IterScal sc;
    double sigma, cScale, dScale, D, E;
    const double* phi;
    const double* q;
    const double* alpha;
    const double* weight;
    const double* beta;     // current multiplier
    const double* z;        // materialised z, or NULL: recompute from (q_old, beta_old)
    const double* q_old;
    const double* beta_old;
    const double* q2b;      // s (BF)^* beta (from launch_bfdconj)
    const double* c0;
    const double* c1;
    double* partial;        // scratch: [nblocks][K]
    double* out;            // [K] results (device)
};
int  kkt_cells_blocks(const Geo& g);
int  kkt_nodes_blocks(const Geo& g);
void launch_kkt_cells(const KktArgs& a, bool weighted, bool one_d, cudaStream_t st);
void launch_kkt_nodes(const KktArgs& a, bool weighted, cudaStream_t st);
// out[0] = sum x[i]^2 (deterministic two-stage reduction); partial needs sumsq_blocks(n) doubles
int  sumsq_blocks(i64 n);
void launch_sumsq(const double* x, i64 n, double* partial, double* out, cudaStream_t st);
// x = (x * mul) / div, elementwise (mul == 1 and div == 1 are exact no-ops)
// level transfer on the device (prolong.cu): coarse (phi, beta) of a finished level -> fine (phi, q, alpha, beta) of the next,
// with the recoverOrgVar / InitialScaling factors folded in exactly where the host path rounds them
struct ProlongScal {
    double phi_recover, beta_recover;   // dScale, cScale*E of the coarse level (recoverOrgVar)
    double grad_t, grad_x, grad_y;      // unscaled forward-difference weights of the fine grid: 1/ht, 1/hx, 1/hy
    double phi_scale, q_scale, alpha_scale, beta_scale;   // 1/dScale, D/dScale, 1/cScale/D, 1/cScale/E of the fine level
};
int launch_prolong(const Geo& gc, const Geo& gf, const ProlongScal& s, const double* phi_c, const double* beta_c, double* phi_f,
                   double* q_f, double* alpha_f, double* beta_f, const double* weight_f, cudaStream_t st);
void launch_scale(double* x, i64 n, double mul, double div, cudaStream_t st);
// Halpern / affine extrapolation of solver_socp_accADMM.m:371-388:
//   x = c1*x0 + c2*((1-rho)*xold + rho*x) ; xold = x ; if (copy_anchor) x0 = x
void launch_halpern(double* x, double* xold, double* x0, i64 n, double c1, double c2, double rho, bool copy_anchor,
                    cudaStream_t st);
// general acc-ADMM extrapolation (opts.theta != 2): x, old, hatOld updated in place (k_accel3)
void launch_accel3(double* x, double* xold, double* xhatold, i64 n, double rho, double a, double b, double c2, bool first,
                   bool keep_hat, cudaStream_t st);
// 6 <-> 10 column conversion of the 1-D variant's z/beta (cols 0..4 -> 0..4, col 5 -> 9; 5..8 zero)
void launch_cols6to10(const double* in6, double* out10, i64 L, cudaStream_t st);
void launch_cols10to6(const double* in10, double* out6, i64 L, cudaStream_t st);

// ---- Poisson / DCT (poisson.cu) -----------------------------------------------------------------------------
struct DctPlan;   // per-length tables (chirps, twiddles, dense matrices), device resident
struct PoissonPlan {
    Geo g;
    DctPlan* py;
    DctPlan* px;
    DctPlan* pt;
    double* lam_t;  // (2 (n-1)^2)(1 - cos(pi k / n))   initialize_FFTkernel.m:6-8
    double* lam_x;
    double* lam_y;
    // tridiagonal t-solve (default; DOTSOCP_TSOLVE=dct selects the fused DCT_t / divide / IDCT_t kernel instead)
    bool use_thomas;
    // pivot reciprocals g_t(mode), [nt][lines] for modes [p0, p0+lines); t_fix[mode] = first level from which g_t no longer
    // changes (the recurrence has reached its fixed point bit for bit), so levels t_fix .. nt-2 need no table read
    struct GTab { i64 p0, lines; double* tab; int* t_fix; };
    std::vector<GTab> gtabs;
    double* cmat_t;     // dense nt x nt orthonormal DCT-II matrix for the singular mode kx = ky = 0
};
PoissonPlan* poisson_plan_create(int nt, int nx, int ny);
void poisson_plan_destroy(PoissonPlan* p);
// a <- idctn( dctn(rhs) ./ (D2 * kernel) ); rhs is only read (rhs == a is allowed)
void poisson_solve(PoissonPlan* p, const double* rhs, double* a, double D2, cudaStream_t st, double* launches);
// time-slab pieces: forward (y then x) / inverse (x then y) transforms of node levels [tn0, tn0+nlev) of the global
// array (src is only read; src == a allowed), and the t-pass on a transposed [nt][chunk] buffer of modes p0..p0+chunk-1
void poisson_xy(PoissonPlan* p, const double* src, double* a, int tn0, int nlev, bool inverse, cudaStream_t st, double* launches,
                double* packed = nullptr, int world = 1, int slab_nlev = 0, int slab_t0 = 0, double* const* push_tab = nullptr);
bool poisson_can_pack(const PoissonPlan* p);   // the x passes can read/write the packed all-to-all buffer themselves
// push_tab / tcut (device tables, see Slab::d_bwd): store the solution in the buffers of the owners of the time levels
void poisson_t_chunk(PoissonPlan* p, double* buf, i64 chunk, i64 p0, double D2, cudaStream_t st, double* launches,
                     double* const* push_tab = nullptr, const int* tcut = nullptr, int world = 1);
// in-place orthonormal DCT-II (or inverse) along all axes
void poisson_dctn(PoissonPlan* p, double* a, bool inverse, cudaStream_t st, double* launches);

}  // namespace dsocp
