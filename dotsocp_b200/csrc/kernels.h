// kernels.h -- host-side launch wrappers of the CUDA kernels (definitions in kernels_update.cu / poisson.cu).
#pragma once
#include "common.cuh"

#include <cuda.h>

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
struct KktFused;
// tensor maps (TMA descriptors) of the arrays one step of the aligned k_mult reads (DOTSOCP_KM_PF=4): the 10 planes of beta as
// one 4-D view, and q0 / bx / by views of q_new, q_old, alpha and the weight
struct KmMaps {
    CUtensorMap beta;
    CUtensorMap qn[3], qo[3], al[3], w[3];
    int toc, ton;   // first cell layer / node level the views start at (time coordinate of a box = level - offset)
};
// q0 / bx / by views of one staggered array (q0: box 32 x TX cells; bx: TX + 1 rows from x0 - 1; by: 34 columns from y0 - 2)
// and the 4-D view of a 10-plane array; 0 on success
// The views cover cell layers [c_lo, c_hi) and node levels [n_lo, n_hi) only -- the window a time slab backs -- so that no part
// of a tensor lies in unmapped address space (a view of the whole array faults in slab sessions of large grids, where the
// windows of a slab are whole 2 MB pages away from the start of the array)
int make_stag_maps(const Geo& g, const double* base, CUtensorMap out[3], int c_lo, int c_hi, int n_lo, int n_hi);
int make_beta_map(const Geo& g, const double* base, CUtensorMap* out, int c_lo, int c_hi);
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
    int kkt_t0;             // first node level of the slab (row 0 of the fused KKT partials)
    // side buffer of the aligned march (mult_side_doubles), NULL: haloed tiling; covers cell layers [side_t0, side_t0 + side_layers)
    double* side;
    int side_t0, side_layers;
    const KmMaps* maps;     // NULL: no tensor maps (DOTSOCP_KM_PF=4 then falls back to the register prefetch)
};
// q_new = ((A phi + alpha) + q2) .* diagQInv ; alpha += tau (A phi - q_new)      (solver_socp_inPALM.m:204-214)
// acc: alpha = (alpha + A phi) - q_new                                            (solver_socp_accADMM.m:237)
// tmpq_in != NULL: take A*phi from that buffer instead of phi; tmpq_out != NULL: also store A*phi; upd_alpha=false: q only
void launch_qstep(const UpdateArgs& a, bool weighted, bool acc, cudaStream_t st, const double* tmpq_in = nullptr,
                  double* tmpq_out = nullptr, bool upd_alpha = true, const KktFused* kkt = nullptr);
void launch_rhs(const Geo& g, const IterScal& sc, bool weighted, const double* q, const double* alpha, const double* weight,
                const double* c0, const double* c1, double* rhs, cudaStream_t st);
// mode 0: beta += tau (z - z2(q)) ; mode 1: beta = (beta + z) - z2(q), z = Pi_Q(z2(q) - beta) ; mode 2: z = d + BF q
void launch_cells_update(const Geo& g, const IterScal& sc, bool one_d, int mode, const double* q, double* z, double* beta,
                         cudaStream_t st);
// z = Pi_Q(d + BF q_old - beta) ; beta += tau (z - (d + BF q_new)) ; then q2, rhs of the next iteration.
// update=false ("prologue"): no multiplier step, only q2/rhs from the current (q_new, alpha, beta).
// kkt != NULL (update only): the launch also leaves the per-(time level, tile) partial sums of the KKT terms that live on
// the data it streams anyway (see KktFused) -- a check then costs no extra pass over the 10-column arrays.
int  launch_mult(const UpdateArgs& a, bool weighted, bool one_d, bool update, cudaStream_t st, const KktFused* kkt = nullptr);   // returns the number of kernels launched
int  mult_tiles(const Geo& g);   // CTAs per time level of launch_mult (rows of its KKT partials)
bool mult_aligned_ok(const Geo& g, bool one_d);                 // the layout allows the aligned (halo-free) tiling of launch_mult
i64  mult_side_doubles(const Geo& g, bool one_d, int nlayers);  // size of its side buffer for `nlayers` cell layers (0: not used)
// z = Pi_Q(d + BF q_old - beta_old): optional store (zout may alias beta_old)
void launch_zstep(const Geo& g, const IterScal& sc, bool one_d, const double* q_old, const double* beta_old, double* zout,
                  cudaStream_t st, const TRange* tr = nullptr);

// ---- KKT / norms --------------------------------------------------------------------------------------------
// All sums are reduced in three fixed-order stages so that the result does not depend on how the time axis is cut into
// slabs (1, 2, 4 or 8 GPUs give the same bits): CTA partial -> one sum per TIME LEVEL (k_level_reduce: row t of the
// level table, KSL slots per row) -> sum over the nt levels (k_levels_total).  Between the last two stages the rows owned
// by other ranks arrive by an all-reduce of the table (every row has exactly one non-zero contributor, so it is exact).
enum { KC_Z2 = 0, KC_BETA2, KC_PRIM2, KC_COMPL, KC_DOTC, KC_RHOT, KC_RHOFQ, KC_COUNT };
enum { KN_Q2 = 0, KN_APHI2, KN_PRIM1, KN_ALPHA2, KN_FBB2, KN_DUAL2, KN_QDOTA, KN_MRHOB, KN_M2, KN_RHOB2,
       KN_DUAL1, KN_CPHI, KN_PHI2, KN_COUNT };
constexpr int KSL = 24;                       // slots per level row: KC_* at 0.., KN_* at KC_COUNT.., then KS_*
constexpr int KS_ELAPSED = KC_COUNT + KN_COUNT;   // host clock of slab 0 (row 0 only), so that all ranks decide alike
constexpr int KS_SGS_BLOCKS = KS_ELAPSED + 1;     // sGS loops: sum (A'(A phi - q + alpha) - c)^2 over the even nodes (:212-216)
constexpr int KS_SGS_KKT = KS_ELAPSED + 2;        //            sum (A'(A phi - q))^2 over all nodes (:322)
// rescale norms (solver_socp_inPALM.m:140-143): one fused pass, slots 0..4 of a level row
enum { NR_PHI2 = 0, NR_Q2, NR_Z2, NR_ALPHA2, NR_BETA2, NR_COUNT };
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
    double* partial;        // scratch: [level][block][K]
    double* lvl;            // level table [nt][KSL] (device)
};
// fused KKT (inPALM check iterations): k_qstep leaves KQ_COUNT sums per (level, block), k_mult KM_COUNT per (level, tile)
enum { KQ_Q2 = 0, KQ_APHI2, KQ_PRIM1, KQ_ALPHA2, KQ_QDOTA, KQ_CPHI, KQ_PHI2, KQ_COUNT };
enum { KM_Z2 = 0, KM_BETA2, KM_PRIM2, KM_COMPL, KM_DOTC, KM_RHOT, KM_RHOFQ, KM_FBB2, KM_DUAL2, KM_DUAL1, KM_MRHOB, KM_M2,
       KM_RHOB2, KM_COUNT };
struct KktFused {
    double sigma, cScale, dScale, D, E;
    double* partial_q;      // [owned node levels][blocks_x][KQ_COUNT]
    double* partial_m;      // [owned node levels][mult_tiles][KM_COUNT]
    double* lvl;            // level table [nt][KSL]
};
int  kkt_cells_blocks(const Geo& g);
int  kkt_nodes_blocks(const Geo& g);
int  kkt_blocks_x(const Geo& g);
void launch_kkt_cells(const KktArgs& a, bool weighted, bool one_d, cudaStream_t st);   // -> rows tc0..tc1-1, slots KC_*
void launch_kkt_nodes(const KktArgs& a, bool weighted, cudaStream_t st);               // -> rows tn0..tn1-1, slots KC_COUNT + KN_*
void launch_kkt_fused_reduce(const Geo& g, const TRange& tr, const KktFused& k, cudaStream_t st);   // partial_q / partial_m -> rows
void launch_norms(const KktArgs& a, bool one_d, cudaStream_t st);                      // -> rows tn0..tn1-1, slots NR_*
void launch_levels_total(const double* lvl, int nt, double* out, cudaStream_t st);     // out[KSL] = fixed-order sum over the rows
// stage 2 for any producer: partial[level][block][K] -> slots[k] of rows t0 .. t0+nlev-1
void level_reduce(const double* partial, int nb, int K, const int* slots, int t0, int nlev, double* lvl, cudaStream_t st);
// ---- output recovery on the device (recover.cu): recover_RhoE.m:13-25, recover_q.m:12-22, check_massConservation.m:16-34
enum { RC_RHO = 0, RC_EX, RC_EY, RC_Q0, RC_BX, RC_BY };
enum { RS_SUMRHO = 0, RS_SUMNEG, RS_W2, RS_COUNT };
struct RecoverArgs {
    Geo g;
    TRange tr;
    double arec, qrec;        // cScale*D (var.alpha = (cScale*D)*alpha), dScale/D (var.q = (dScale/D)*q): recoverOrgVar
    const double* alpha;
    const double* q;
    const double* weight;     // NULL unless weighted
    const double* rho0;       // nx*ny planes (device): model.rho0(:), model.rho1(:)
    const double* rho1;
    double* out;              // node-indexed scratch (global index space), levels of the slab
    double* partial;
    double* lvl;
};
void launch_recover(const RecoverArgs& a, int which, bool weighted, cudaStream_t st);      // one field into a.out
void launch_recover_stats(const RecoverArgs& a, bool weighted, bool one_d, cudaStream_t st);   // rows tn0..tn1-1, slots RS_*
// x = (x * mul) / div, elementwise (mul == 1 and div == 1 are exact no-ops)
// level transfer on the device (prolong.cu): coarse (phi, beta) of a finished level -> fine (phi, q, alpha, beta) of the next,
// with the recoverOrgVar / InitialScaling factors folded in exactly where the host path rounds them
struct ProlongScal {
    double phi_recover, beta_recover;   // dScale, cScale*E of the coarse level (recoverOrgVar)
    double grad_t, grad_x, grad_y;      // unscaled forward-difference weights of the fine grid: 1/ht, 1/hx, 1/hy
    double phi_scale, q_scale, alpha_scale, beta_scale;   // 1/dScale, D/dScale, 1/cScale/D, 1/cScale/E of the fine level
};
// stages of the transfer, each on a range of fine time levels (prolong.cu); dotsocp_prolong strings them together per slab
void launch_prolong_phi(const Geo& gc, const Geo& gf, double rec, const double* phi_c, double* phi_f, int t0, int t1, cudaStream_t st);
void launch_prolong_q(const Geo& gf, const ProlongScal& s, const double* phi_f, const double* weight_f, double* q_f, int t0, int t1,
                      cudaStream_t st);
void launch_prolong_beta(const Geo& gc, const Geo& gf, double rec, const double* beta_c, double* beta_f, int c0, int c1, cudaStream_t st);
void launch_mul_inplace(double* x, i64 n, double s, cudaStream_t st);
void launch_prolong_alpha(double* alpha, const double* weight, i64 n, double scale, cudaStream_t st);
void launch_scale(double* x, i64 n, double mul, double div, cudaStream_t st);
void launch_fill(double* x, i64 n, double v, cudaStream_t st);   // x[0..n) = v
void debug_sum(const char* name, const double* x, i64 n, cudaStream_t st);   // debugging aid: sum |x| / non-finite count into a device log
void debug_sum_flush(cudaStream_t st);                                        // ... printed here (the only synchronisation)
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

// ---- red-black symmetric Gauss-Seidel (sgs.cu): mexsGS.mexa64 -----------------------------------------------------
bool sgs_supported(const Geo& g);   // nx == ny, odd node counts (the only grids the reference binary handles)
void launch_sgs_half(const Geo& g, double ep, double scale, int parity, const double* rhs, double* phi, int tn0, int tn1,
                     cudaStream_t st);
// sum over nodes of (A'(A phi - q + alpha) - c)^2 on the even nodes (with_alpha_c) or of (A'(A phi - q))^2 on all nodes
void launch_sgs_resid(const Geo& g, const IterScal& sc, bool with_alpha_c, const double* phi, const double* q, const double* alpha,
                      const double* c0, const double* c1, double* partial, double* lvl, int slot, int tn0, int tn1, cudaStream_t st);
void launch_sum_nodes(const Geo& g, const double* phi, double* partial, double* lvl, int slot, int tn0, int tn1, cudaStream_t st);
void launch_shift(double* x, i64 n, double shift, cudaStream_t st);

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
// All poisson_* launchers return 0, or a negative DOTSOCP_E* code with dotsocp_last_error() set (unsupported geometry).
// a <- idctn( dctn(rhs) ./ (D2 * kernel) ); rhs is only read (rhs == a is allowed)
int  poisson_solve(PoissonPlan* p, const double* rhs, double* a, double D2, cudaStream_t st, double* launches);
// time-slab pieces: forward (y then x) / inverse (x then y) transforms of node levels [tn0, tn0+nlev) of the global
// array (src is only read; src == a allowed), and the t-pass on a transposed [nt][chunk] buffer of modes p0..p0+chunk-1
int  poisson_xy(PoissonPlan* p, const double* src, double* a, int tn0, int nlev, bool inverse, cudaStream_t st, double* launches,
                double* packed = nullptr, int world = 1, int slab_nlev = 0, int slab_t0 = 0, double* const* push_tab = nullptr);
bool poisson_can_pack(const PoissonPlan* p);   // the x passes can read/write the packed all-to-all buffer themselves
// push_tab / tcut (device tables, see Slab::d_bwd): store the solution in the buffers of the owners of the time levels
int  poisson_t_chunk(PoissonPlan* p, double* buf, i64 chunk, i64 p0, double D2, cudaStream_t st, double* launches,
                     double* const* push_tab = nullptr, const int* tcut = nullptr, int world = 1);
// pipelined slab Thomas: the t-solve with the time axis cut into slabs, one carry plane per slab boundary and direction
// instead of the two transposes; bit-identical to the single-slab solve (poisson.cu)
// builds the tables of the session's t-solve now instead of inside its first solve (lines > 0: modes p0 .. p0+lines-1 of the
// whole time axis; lines == 0: the slab of levels [t0, t1) for the pipelined sweeps)
int  poisson_prepare(PoissonPlan* p, i64 p0, i64 lines, int t0, int t1, cudaStream_t st);
bool poisson_slab_thomas_ok(const PoissonPlan* p);
// sync != NULL: the hand-off with the neighbouring GPUs happens inside the kernel (peer-memory stores + epoch flags) instead of
// NCCL send / recv around it: wait until *wait_flag >= wait_value before reading carry_in (NULL: no wait); after the last CTA has
// stored its part of carry_out (which then points into the neighbour's buffer) write signal_value to *signal_flag (NULL: no signal)
struct SlabSync {
    const int* wait_flag;
    int wait_value;
    int* done;            // CTA counter of this chunk (local, zero between launches)
    int* signal_flag;     // the neighbour's flag (peer memory)
    int signal_value;
};
int  poisson_thomas_slab(PoissonPlan* p, double* a, int t0, int t1, i64 m0, i64 m1, double D2, bool backward, const double* carry_in,
                         double* carry_out, cudaStream_t st, double* launches, const SlabSync* sync = nullptr);
void poisson_line0_gather(PoissonPlan* p, const double* a, double* line, int t0, int t1, cudaStream_t st);
void poisson_line0_solve(PoissonPlan* p, double* line, double D2, cudaStream_t st);
void poisson_line0_scatter(PoissonPlan* p, double* a, const double* line, int t0, int t1, cudaStream_t st);
// in-place orthonormal DCT-II (or inverse) along all axes
int  poisson_dctn(PoissonPlan* p, double* a, bool inverse, cudaStream_t st, double* launches);

// ---- level weights resident on the device (weights.cu): one packed array [q0 | bx | by] per level, finest first
struct WeightLevel {
    int nt, nx, ny;        // nodes
    i64 L, NBX, NBY, Q;    // packed (reference) sizes
    double* w;             // device
};
// window [b, e) of a session's staggered array (device layout g) <- the packed level array; pad entries become 1
void launch_weight_scatter(const Geo& g, i64 b, i64 e, const double* packed, double* dst, cudaStream_t st);

}  // namespace dsocp

struct dotsocp_weights {
    std::vector<dsocp::WeightLevel> lv;
    int device = 0;
    int filled = 0;            // levels 0 .. filled-1 hold values
    double* scratch = nullptr;
    double launches = 0.0;
    ~dotsocp_weights();
};
