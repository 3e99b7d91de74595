/* dotsocp.h -- C ABI of libdotsocp.so, the B200 (sm_100a) implementation of DOTSOCP's iteration hot path.
 *
 * Everything here is plain C: pointers, sizes and POD structs; no torch / CUDA types cross the boundary.
 * All arrays are FP64 in MATLAB column-major order (y fastest, then x, then t), exactly as the reference
 * hands them to its own MEX kernels:
 *     q, alpha, weight : [ q0 (nt-1,nx,ny) | bx (nt,nx-1,ny) | by (nt,nx,ny-1) ]      Q doubles
 *     z, beta          : L x 10 column-major (L x 6 for the 1-D variant),  L = (nt-1)*nx*ny
 *     phi, c           : N = nt*nx*ny doubles
 * Unless stated otherwise pointers are HOST pointers; the library owns all device memory.
 * Every function returns 0 on success, a negative DOTSOCP_E* code otherwise; dotsocp_last_error() gives text.
 * There is NO CPU fallback: without a usable CUDA device every compute entry point fails with DOTSOCP_ENODEV.
 *
 * Reference interfaces replaced (paths relative to the reference checkout):
 *   kernel level  (FFI the reference binds today: mexFunction of the pre-built MEX files)
 *     dotsocp_mexBFd        <- mexBFd(z2,q,nt,nx,ny,scaleBF,scaleD)     socp/dot2d/algorithms/solver_socp_inPALM.m:133,187,212,242
 *     dotsocp_mexBFdConj    <- mexBFdConj(q2,z,nt,nx,ny,scaleBF)        socp/dot2d/algorithms/solver_socp_inPALM.m:205,225 ; utils/jump_nextLevel.m:16
 *     dotsocp_mexProjSoc    <- mexProjSoc(out,in)                        socp/dot2d/algorithms/solver_socp_inPALM.m:199,240
 *     dotsocp_mexsGS        <- mexsGS(phi,rhs,ep,scale,nt,nx,ny,its)      socp/dot2d/algorithms/solver_socp_sGSinPALM.m:205
 *     dotsocp_mexBFd1d      <- mexBFd1d(z,q,nt,nx,scale,dFactor)         socp/dot1d/algorithms/solver_socp_inPALM.m:132,186,211,241
 *     dotsocp_mexBFdConj1d  <- mexBFdConj1d(q,z,nt,nx,scale)             socp/dot1d/algorithms/solver_socp_inPALM.m:204,224
 *     dotsocp_poisson       <- oper_poisson3dim(D^2*initialize_FFTkernel(nt,nx,ny), rhs)   socp/dot2d/utils/oper_poisson3dim.m:4,
 *                              initialize_FFTkernel.m:6-15 ; 1-D: socp/dot1d/utils/oper_poisson.m:4  (ny = 1)
 *   solver level  (the boundary the north star names)
 *     dotsocp_solve_level   <- [runHist,sigma] = solver_socp_inPALM(var,opts,model)   socp/dot2d/algorithms/solver_socp_inPALM.m:1
 *                              solver_socp_PALM.m:1, solver_socp_accADMM.m:1, solver_socp_sGSinPALM.m:1, solver_socp_accsGSADMM.m:1,
 *                              socp/wdot2d/algorithms/solver_wsocp_inPALM.m:1, solver_wsocp_accADMM.m:1,
 *                              socp/dot1d/algorithms/solver_socp_inPALM.m:1
 *   device-resident session (same loop, state stays in HBM; used by the multilevel driver and the benchmark)
 *     dotsocp_create / _upload / _run / _iterate / _download / _destroy
 */
#ifndef DOTSOCP_H
#define DOTSOCP_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DOTSOCP_OK        0
#define DOTSOCP_EINVAL   -1   /* bad argument (sizes, NULL pointer, unknown variant/method) */
#define DOTSOCP_ENODEV   -2   /* no CUDA device / driver: the library never falls back to the CPU */
#define DOTSOCP_ECUDA    -3   /* a CUDA runtime call or kernel failed */
#define DOTSOCP_ENOMEM   -4   /* device or host allocation failed */
#define DOTSOCP_ENCCL    -5   /* NCCL failure (multi-GPU sessions) */
#define DOTSOCP_ESTATE   -6   /* call sequence error (e.g. run before upload) */

/* variants: which reference sub-tree the call mirrors */
#define DOTSOCP_VARIANT_DOT2D   0   /* socp/dot2d  */
#define DOTSOCP_VARIANT_WDOT2D  1   /* socp/wdot2d (needs weight) */
#define DOTSOCP_VARIANT_DOT1D   2   /* socp/dot1d  (ny must be 1, z/beta have 6 columns at the boundary) */

/* methods: which algorithms/*.m loop */
#define DOTSOCP_METHOD_INPALM   0   /* solver_*socp_inPALM.m ; "ALG2" is the same loop with tau = 1 */
#define DOTSOCP_METHOD_PALM     1   /* solver_socp_PALM.m (dot2d only) */
#define DOTSOCP_METHOD_ACCADMM  2   /* solver_*socp_accADMM.m (dot2d, wdot2d) */
#define DOTSOCP_METHOD_SGSINPALM 3  /* solver_socp_sGSinPALM.m (dot2d, nx == ny, odd node counts: the grids mexsGS handles) */
#define DOTSOCP_METHOD_ACCSGSADMM 4 /* solver_socp_accsGSADMM.m (same grids, single GPU)                                       */

#define DOTSOCP_NTIMES 8

/* opts/var/model scalars of one level solve -- field meaning follows the reference names exactly */
typedef struct dotsocp_level_opts {
    int32_t variant;            /* DOTSOCP_VARIANT_*                                                   */
    int32_t method;             /* DOTSOCP_METHOD_*                                                    */
    int32_t nt, nx, ny;         /* model.nt, model.nx, model.ny (nodes); ny = 1 for the 1-D variant     */
    int32_t maxit;              /* opts.maxit                                                          */
    int32_t ifCheckStepByStep;  /* opts.ifCheckStepByStep                                              */
    int32_t scaling;            /* opts.scaling (rescale = 1 when non-zero), solver_socp_inPALM.m:64-68 */
    int32_t checkPrimDualFeas;  /* opts.checkPrimDualFeas; -1 = absent (default true, weighted: false)  */
    int32_t restart;            /* accADMM opts.restart; <=0 = absent (100)                             */
    double  tau;                /* opts.tau (inPALM/PALM); ignored by accADMM                           */
    double  sigma;              /* opts.sigma                                                          */
    double  tol;                /* opts.tol                                                            */
    double  time_limit;         /* opts.time_limit; NaN = absent (3600 s).  <= 0 = budget already spent: the loop runs one
                                   iteration and one check and stops, as the reference does when the driver hands down
                                   time_limit - Total_Time (solver_dotsocp2d.m:244, solver_socp_inPALM.m:287-289)  */
    double  rho;                /* accADMM opts.rho;   <=0 = absent (2)                                 */
    double  theta;              /* accADMM opts.theta; <=0 = absent (2 => Halpern)                      */
    double  cScale, dScale, D, E;   /* var.cScale, var.dScale, var.D, var.E  (InitialScaling)          */
    double  normc, normd;       /* model.normc, model.normd                                            */
    double  grad_t, grad_x, grad_y; /* the three distinct magnitudes of model.grad AFTER InitialScaling:
                                       D*(1/ht), D*(1/hx), D*(1/hy)  (initialize.m:67-87, solver_dotsocp2d.m:338) */
} dotsocp_level_opts;

/* what the loop returns besides the mutated arrays */
typedef struct dotsocp_level_result {
    int32_t iters;              /* `it` at exit (var.time.Iters)                                        */
    int32_t hist_len;           /* runHist.len                                                         */
    double  sigma;              /* returned sigma (= sigma / sigmaScale, solver_socp_inPALM.m:357)      */
    double  cScale, dScale, D, E;   /* var.* after rescaling (:344-347)                                 */
    double  times[DOTSOCP_NTIMES];  /* seconds, device-measured; order per method:
                                       inPALM : FFT, ProjSOC, Q_Step, Multiplier, KKT, Total, 0, 0      (:339-340)
                                       PALM   : Q_Step(1), FFT, ProjSOC, Q_Step(3), Multiplier, KKT, Total, 0
                                       accADMM: Q_Step, Multiplier, FFT, ProjSOC, KKT, Interp, Total, 0
                                       sGSinPALM: sGS, ProjSOC, Q_Step, Multiplier, KKT, Total, 0, 0  (:419-420)
                                       accsGSADMM: sGS, ProjSOC, Multiplier, Q_Step, Interp, KKT, Total, 0  (:512-513)
                                       The fused kernels do not separate ProjSOC from the multiplier step; the fused
                                       time is booked under the step that dominates it.             */
    double  gpu_launches;       /* number of kernels launched by this call                              */
} dotsocp_level_result;

/* runHist buffers, caller-allocated with room for `cap` checks (cap = maxit is always enough) */
typedef struct dotsocp_hist {
    int32_t cap;
    double *kkt;                /* cap x 7, ROW-major here: kkt[i*7 + j] = runHist.kkt(i+1, j+1)       */
    double *time;               /* cap                                                                 */
    double *iter;               /* cap                                                                 */
    double *pdGap;              /* cap                                                                 */
    double *priVal;             /* cap -- extra: priVal of solver_socp_inPALM.m:265 (not in the reference struct) */
    double *dualVal;            /* cap -- extra: dualVal of :266                                       */
} dotsocp_hist;

const char *dotsocp_last_error(void);
int  dotsocp_version(void);
int  dotsocp_device_count(void);          /* >= 0, or DOTSOCP_ENODEV */
int  dotsocp_set_device(int device);

/* ------------------------------------------------------------------ kernel level (host buffers, in place into arg 1) */
int dotsocp_mexBFd(double *z2, const double *q, int nt, int nx, int ny, double scaleBF, double scaleD);
int dotsocp_mexBFdConj(double *q2, const double *z, int nt, int nx, int ny, double scaleBF);
int dotsocp_mexProjSoc(double *out, const double *in, int64_t M, int N);
/* mexsGS(phi, rhs, ep, scale, nt, nx, ny, its): `its` symmetric red-black Gauss-Seidel sweeps, in place into phi
 * (mexsGS.mexa64; solver_socp_sGSinPALM.m:205, solver_socp_accsGSADMM.m:256).  nx == ny, odd node counts only.        */
int dotsocp_mexsGS(double *phi, const double *rhs, double ep, double scale, int nt, int nx, int ny, int its);
int dotsocp_mexBFd1d(double *z, const double *q, int nt, int nx, double scale, double dFactor);
int dotsocp_mexBFdConj1d(double *q, const double *z, int nt, int nx, double scale);
/* phi = idctn( dctn(rhs) ./ (D^2 * kernel) ), Neumann eigenvalues, zero mode := 1 */
int dotsocp_poisson(double *phi, const double *rhs, int nt, int nx, int ny, double D);
/* orthonormal DCT-II (inverse != 0: its inverse) along every axis of a (nt,nx,ny) array: mirt_dctn / mirt_idctn */
int dotsocp_dctn(double *a, int nt, int nx, int ny, int inverse);

/* ------------------------------------------------------------------ solver level (host buffers, mutated in place) */
int dotsocp_solve_level(const dotsocp_level_opts *opts,
                        double *phi, double *q, double *z, double *alpha, double *beta,
                        const double *c, const double *weight /* NULL unless WDOT2D */,
                        dotsocp_hist *hist, dotsocp_level_result *res);
/* dotsocp_solve_level keeps its device session alive between calls (process lifetime, keyed by variant and grid; the MEX
 * gateway registers this with mexAtExit).  Frees it; a no-op when nothing is cached.  DOTSOCP_CACHE_CTX=0 disables caching. */
void dotsocp_release_cached(void);

/* ------------------------------------------------------------------ device-resident session */
typedef struct dotsocp_ctx dotsocp_ctx;
/* world > 1 : time-slab partition over `world` ranks (one process per GPU); nccl_id is the 128-byte
 * ncclUniqueId produced by dotsocp_nccl_unique_id() on rank 0 and broadcast by the caller.  The communicator is
 * process-wide: passing 128 zero bytes re-uses the communicator of an earlier session of this process (same rank/world),
 * whose peer connections are already established.  world > 1 with nccl_id == NULL runs all slabs in this process on
 * the current device (single-GPU emulation of the slab code path, used by the tests).                               */
int  dotsocp_nccl_unique_id(char id128[128]);
int  dotsocp_create(dotsocp_ctx **ctx, int variant, int nt, int nx, int ny, int rank, int world, const char *nccl_id);
void dotsocp_destroy(dotsocp_ctx *ctx);
/* full (global) host arrays in, each rank keeps its slab.  z may be NULL when the next run is inPALM with maxit >= 1:
 * that loop overwrites z (solver_socp_inPALM.m:199) before reading it, so its incoming value need not cross PCIe
 * (dotsocp_solve_level does this itself); any other run, or a download of z before a run, then fails with DOTSOCP_ESTATE. */
int  dotsocp_upload(dotsocp_ctx *ctx, const double *phi, const double *q, const double *z,
                    const double *alpha, const double *beta, const double *c, const double *weight);
int  dotsocp_download(dotsocp_ctx *ctx, double *phi, double *q, double *z, double *alpha, double *beta);
/* Output recovery and checks on the device after the last level, instead of downloading the whole state: what
 * solver_dotsocp2d.m:268-287 computes -- recoverOrgVar (:368-386), recover_RhoE (utils/recover_RhoE.m:13-25; wdot2d
 * multiplies alpha by the weight), recover_q (utils/recover_q.m:12-22), check_massConservation (utils/
 * check_massConservation.m:16-34) -- plus the transport cost mean(|m|^2/rho) over the space-time nodes with rho > 1e-12.
 * rho0 / rho1: model.rho0(:), model.rho1(:) (nx*ny doubles).  Outputs (any may be NULL): rho, Ex, Ey on the nt node
 * levels (N doubles, MATLAB order (ny,nx,nt)); q0, bx, by on the nt-1 cell layers (L doubles); sumRho / sumNegRho: nt
 * doubles each; w2cost: one double.  One-process-per-GPU sessions pass / receive the slab's own levels for the fields and
 * the full-length vectors for the per-level sums.  The fields are bit-identical to the host path.                     */
typedef struct dotsocp_recover_scal {
    double alpha_recover;           /* var.cScale * var.D   (var.alpha = (cScale*D) * var.alpha) */
    double q_recover;               /* var.dScale / var.D   (var.q = (dScale/D) * var.q)         */
} dotsocp_recover_scal;
int  dotsocp_recover(dotsocp_ctx *ctx, const dotsocp_recover_scal *s, const double *rho0, const double *rho1,
                     double *rho, double *Ex, double *Ey, double *q0, double *bx, double *by,
                     double *sumRho, double *sumNegRho, double *w2cost);
/* Level transfer of the multilevel drivers with the state resident in HBM: what solver_dotsocp2d.m:230-250 does between two
 * levels -- recoverOrgVar (:368-386), interpolate (utils/interpolate.m:46-84), jump_nextLevel (utils/jump_nextLevel.m:5-16:
 * z = 0, q = A phi, alpha = (BF)^*(-beta)), InitialScaling (:304-365) -- from the finished coarse session into a fresh
 * session of the refined grid (2n-1 nodes per refined axis).  The scalars are the ones the driver computes on the host;
 * every value is rounded exactly where the host path rounds it.  c_first / c_last: the two non-zero planes of the fine
 * model.c (nx*ny doubles each, already divided by cScale); weight: fine weight (Q doubles) for WDOT2D, else NULL.
 * Time slabs: `fine` comes from dotsocp_create_refined (slab r of the fine grid refines slab r of the coarse one), the
 * host arrays follow the upload convention (global arrays, or the slab's own part in a one-process-per-GPU run).       */
typedef struct dotsocp_prolong_scal {
    double phi_recover;             /* coarse var.dScale                 (var.phi  = dScale * var.phi)             */
    double beta_recover;            /* coarse var.cScale * var.E         (var.beta = (cScale*E) * var.beta)        */
    double grad_t, grad_x, grad_y;  /* fine, UNSCALED: 1/ht, 1/hx, 1/hy  (initialize.m:67-87)                      */
    double phi_scale;               /* fine 1/dScale                                                               */
    double q_scale;                 /* fine D/dScale                                                               */
    double alpha_scale;             /* fine 1/cScale/D                                                             */
    double beta_scale;              /* fine 1/cScale/E                                                             */
} dotsocp_prolong_scal;
int  dotsocp_create_refined(dotsocp_ctx **fine, const dotsocp_ctx *coarse);
int  dotsocp_prolong(dotsocp_ctx *coarse, dotsocp_ctx *fine, const dotsocp_prolong_scal *s,
                     const double *c_first, const double *c_last, const double *weight);
/* ------------------------------------------------------------------ level weights resident on the device (WDOT2D)
 * What the weighted drivers do with the weight before the first level starts -- examples/wdot2d/gene_weight_circle.m:6-27 and
 * get_weight_by_barrier.m:12-33 (finest weight = [ones ; repmat(weightX, nt) ; repmat(weightY, nt)]), the restriction chain
 * socp/wdot2d/utils/downSample_q.m:4-19 / downSample_barrier.m:4-24 (solver_wdotsocp2d.m:179-187) and mean(log10(weight +
 * 1e-10)) of solver_wdotsocp2d.m:312-316 -- without a Q-sized host array per level: a pyramid of `levels` packed arrays
 * [q0 | bx | by] in HBM, level 0 = the finest (nt, nx, ny) grid, level l = (n+1)/2 nodes per axis of level l-1.
 *   _set        : finest level from a host array (Q doubles, the reference layout)
 *   _set_planes : finest level from the two (x,y) planes of the generators: weightX (nx-1)*ny doubles, weightY nx*(ny-1)
 *                 doubles, both in C order (x, y) = MATLAB (ny, nx-1)(:) / (ny-1, nx)(:); replicated over the time levels,
 *                 ones on the q0 part
 *   _restrict   : levels 1 .. levels-1 from level 0; geometric != 0: exp(restrict(log w)) (downSample_barrier), else downSample_q
 *   _get / _log10_mean / _dims : read a level back (tests), the mean for `adjust`, the node counts of a level
 *   dotsocp_set_weight : level `level` becomes the weight of a WDOT2D session (any slab layout; device to device);
 *                 dotsocp_upload / dotsocp_prolong then take weight == NULL.  In a one-process-per-GPU run every rank builds
 *                 the pyramid on its own GPU (no communication) and its session takes only the slab's part.               */
typedef struct dotsocp_weights dotsocp_weights;
int  dotsocp_weights_create(dotsocp_weights **w, int nt, int nx, int ny, int levels);
void dotsocp_weights_destroy(dotsocp_weights *w);
int  dotsocp_weights_set(dotsocp_weights *w, const double *weight);
int  dotsocp_weights_set_planes(dotsocp_weights *w, const double *weightX, const double *weightY);
int  dotsocp_weights_restrict(dotsocp_weights *w, int geometric);
int  dotsocp_weights_get(const dotsocp_weights *w, int level, double *weight);
int  dotsocp_weights_dims(const dotsocp_weights *w, int level, int *nt, int *nx, int *ny);
int  dotsocp_weights_log10_mean(dotsocp_weights *w, int level, double *mean);
double dotsocp_weights_launch_count(const dotsocp_weights *w);
int  dotsocp_set_weight(dotsocp_ctx *ctx, const dotsocp_weights *w, int level);

/* the reference loop on the resident state (sigma folding at entry, un-folding at exit, like :102-104, :335-336) */
int  dotsocp_run(dotsocp_ctx *ctx, const dotsocp_level_opts *opts, dotsocp_hist *hist, dotsocp_level_result *res);
/* benchmark primitive: begin (sigma folding + prologue), n iterations, elapsed device ms.  with_kkt_every = k > 0 makes every
 * k-th iteration a check iteration (KKT sums fused into the update kernels + reductions + read-back); 0 = no checks.      */
int  dotsocp_iter_begin(dotsocp_ctx *ctx, const dotsocp_level_opts *opts);
int  dotsocp_iterate(dotsocp_ctx *ctx, int n_iters, int with_kkt_every, float *elapsed_ms, float *ms_by_kernel /* [4] or NULL */);
int  dotsocp_iter_end(dotsocp_ctx *ctx);
/* number of kernels launched by this context since creation */
double dotsocp_launch_count(const dotsocp_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* DOTSOCP_H */
