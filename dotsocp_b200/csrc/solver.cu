// solver.cu -- device-resident session, host control loop and the C ABI of libdotsocp.so.
//
// The control flow mirrors socp/dot2d/algorithms/solver_socp_inPALM.m line by line (rescaling :138-190, iteration
// :192-216, KKT :218-324, output :328-357); only the order of the cell-local steps inside one iteration is fused
// differently (see kernels_update.cu / DESIGN.md): the z-step of iteration i+1 does not depend on phi_{i+1}, so it is
// evaluated inside the multiplier kernel of iteration i.  All scalar decisions (sigma rule, rescale triggers, stop
// test) stay on the host in double precision, exactly as in the reference.
#include "../../include/dotsocp.h"
#include "kernels.h"

#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

using namespace dsocp;

// ------------------------------------------------------------------------------------------------ error plumbing
static thread_local char g_err[1024] = "";
static int set_err(int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}
#define CU(call)                                                                                              \
    do {                                                                                                      \
        cudaError_t e_ = (call);                                                                              \
        if (e_ != cudaSuccess)                                                                                \
            return set_err(e_ == cudaErrorMemoryAllocation ? DOTSOCP_ENOMEM : DOTSOCP_ECUDA, "%s:%d %s: %s", __FILE__, \
                           __LINE__, #call, cudaGetErrorString(e_));                                          \
    } while (0)

extern "C" const char* dotsocp_last_error(void) { return g_err; }
extern "C" int dotsocp_version(void) { return 100; }

static int require_device()
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return set_err(DOTSOCP_ENODEV, "no usable CUDA device (%s); libdotsocp has no CPU fallback",
                       e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    return DOTSOCP_OK;
}
extern "C" int dotsocp_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return set_err(DOTSOCP_ENODEV, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    }
    return n;
}
extern "C" int dotsocp_set_device(int device)
{
    int rc = require_device();
    if (rc) return rc;
    CU(cudaSetDevice(device));
    return DOTSOCP_OK;
}

// ------------------------------------------------------------------------------------------------ context
struct EvPool {
    std::vector<cudaEvent_t> ev;
    size_t used = 0;
    cudaEvent_t get()
    {
        if (used == ev.size()) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            ev.push_back(e);
        }
        return ev[used++];
    }
    void reset() { used = 0; }
    ~EvPool() { for (auto e : ev) cudaEventDestroy(e); }
};

struct dotsocp_ctx {
    int variant = 0;
    bool one_d = false, weighted = false;
    int rank = 0, world = 1;
    Geo g;
    cudaStream_t st = nullptr;
    double *phi = nullptr, *rhs = nullptr;
    double* q[2] = {nullptr, nullptr};
    int qcur = 0;
    double *alpha = nullptr, *q2 = nullptr, *qtmp = nullptr, *weight = nullptr;
    double* beta[2] = {nullptr, nullptr};   // beta[bcur] = multiplier; beta[1-bcur] = previous multiplier / materialised z
    int bcur = 0;
    bool z_materialised = true;             // beta[1-bcur] holds z itself (after upload / at exit) instead of beta_old
    double *c0 = nullptr, *c1 = nullptr;
    double *partial = nullptr, *dsums = nullptr;
    double* hsums = nullptr;                // pinned
    PoissonPlan* pp = nullptr;
    double launches = 0;
    bool uploaded = false;
    // acc-ADMM / PALM state (allocated on demand)
    double *zmat = nullptr;
    double *old_[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // phi, z, q, alpha, beta
    double *anc_[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    // benchmark session (dotsocp_iter_begin .. _end)
    bool iter_open = false;
    IterScal sc;
    double sigma_fold = 1.0;
    double D2 = 1.0;
    EvPool evs;
};

static size_t partial_doubles(const Geo& g)
{
    size_t a = (size_t)kkt_nodes_blocks(g) * KN_COUNT;
    size_t b = (size_t)kkt_cells_blocks(g) * KC_COUNT;
    size_t c = (size_t)sumsq_blocks(10 * g.L + g.Q);
    size_t m = a > b ? a : b;
    return (m > c ? m : c) + 64;
}

extern "C" int dotsocp_nccl_unique_id(char id128[128])
{
    (void)id128;
    return set_err(DOTSOCP_ENCCL, "multi-GPU sessions are not built into this library version");
}

extern "C" void dotsocp_destroy(dotsocp_ctx* c)
{
    if (!c) return;
    cudaFree(c->phi); cudaFree(c->rhs); cudaFree(c->q[0]); cudaFree(c->q[1]); cudaFree(c->alpha); cudaFree(c->q2);
    cudaFree(c->qtmp); cudaFree(c->weight); cudaFree(c->beta[0]); cudaFree(c->beta[1]); cudaFree(c->c0); cudaFree(c->c1);
    cudaFree(c->partial); cudaFree(c->dsums); cudaFree(c->zmat);
    for (int i = 0; i < 5; i++) { cudaFree(c->old_[i]); cudaFree(c->anc_[i]); }
    if (c->hsums) cudaFreeHost(c->hsums);
    poisson_plan_destroy(c->pp);
    if (c->st) cudaStreamDestroy(c->st);
    delete c;
}

extern "C" int dotsocp_create(dotsocp_ctx** out, int variant, int nt, int nx, int ny, int rank, int world, const char* nccl_id)
{
    (void)nccl_id;
    if (!out) return set_err(DOTSOCP_EINVAL, "ctx pointer is NULL");
    *out = nullptr;
    if (variant < 0 || variant > 2) return set_err(DOTSOCP_EINVAL, "unknown variant %d", variant);
    if (nt < 2 || nx < 2 || ny < 1) return set_err(DOTSOCP_EINVAL, "grid %d x %d x %d too small (need nt,nx >= 2)", nt, nx, ny);
    if (variant == DOTSOCP_VARIANT_DOT1D && ny != 1) return set_err(DOTSOCP_EINVAL, "1-D variant needs ny == 1");
    if (variant != DOTSOCP_VARIANT_DOT1D && ny < 2) return set_err(DOTSOCP_EINVAL, "2-D variants need ny >= 2");
    if (world != 1 || rank != 0) return set_err(DOTSOCP_EINVAL, "world=%d: time-slab multi-GPU sessions are not available in this build", world);
    int rc = require_device();
    if (rc) return rc;
    dotsocp_ctx* c = new dotsocp_ctx();
    c->variant = variant;
    c->one_d = variant == DOTSOCP_VARIANT_DOT1D;
    c->weighted = variant == DOTSOCP_VARIANT_WDOT2D;
    c->rank = rank;
    c->world = world;
    c->g = make_geo(nt, nx, ny);
    const Geo& g = c->g;
#define ALLOC(ptr, count)                                                                              \
    do {                                                                                               \
        cudaError_t e_ = cudaMalloc(&(ptr), (size_t)(count) * sizeof(double));                          \
        if (e_ != cudaSuccess) {                                                                       \
            cudaGetLastError();                                                                        \
            int code_ = set_err(DOTSOCP_ENOMEM, "cudaMalloc of %zu bytes failed: %s",                   \
                                (size_t)(count) * sizeof(double), cudaGetErrorString(e_));              \
            dotsocp_destroy(c);                                                                        \
            return code_;                                                                              \
        }                                                                                              \
    } while (0)
    ALLOC(c->phi, g.N); ALLOC(c->rhs, g.N);
    ALLOC(c->q[0], g.Q); ALLOC(c->q[1], g.Q); ALLOC(c->alpha, g.Q); ALLOC(c->q2, g.Q); ALLOC(c->qtmp, g.Q);
    if (c->weighted) ALLOC(c->weight, g.Q);
    ALLOC(c->beta[0], 10 * g.L); ALLOC(c->beta[1], 10 * g.L);
    ALLOC(c->c0, g.P); ALLOC(c->c1, g.P);
    ALLOC(c->partial, partial_doubles(g)); ALLOC(c->dsums, 64);
#undef ALLOC
    if (cudaMallocHost(&c->hsums, 64 * sizeof(double)) != cudaSuccess) {
        dotsocp_destroy(c);
        return set_err(DOTSOCP_ENOMEM, "cudaMallocHost failed");
    }
    if (cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking) != cudaSuccess) {
        dotsocp_destroy(c);
        return set_err(DOTSOCP_ECUDA, "cudaStreamCreate failed");
    }
    c->pp = poisson_plan_create(nt, nx, ny);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        dotsocp_destroy(c);
        return set_err(DOTSOCP_ECUDA, "plan creation: %s", cudaGetErrorString(e));
    }
    *out = c;
    return DOTSOCP_OK;
}

extern "C" double dotsocp_launch_count(const dotsocp_ctx* c) { return c ? c->launches : 0.0; }

static int upload_cols(dotsocp_ctx* c, double* dst10, const double* src)
{
    const Geo& g = c->g;
    if (!c->one_d) {
        CU(cudaMemcpyAsync(dst10, src, (size_t)10 * g.L * sizeof(double), cudaMemcpyHostToDevice, c->st));
        return 0;
    }
    // 6 columns at the boundary: stage through qtmp/rhs-sized scratch is too small in general, use a temporary
    double* tmp = nullptr;
    CU(cudaMalloc(&tmp, (size_t)6 * g.L * sizeof(double)));
    CU(cudaMemcpyAsync(tmp, src, (size_t)6 * g.L * sizeof(double), cudaMemcpyHostToDevice, c->st));
    launch_cols6to10(tmp, dst10, g.L, c->st);
    c->launches += 1;
    CU(cudaStreamSynchronize(c->st));
    cudaFree(tmp);
    return 0;
}

static int download_cols(dotsocp_ctx* c, double* dst, const double* src10)
{
    const Geo& g = c->g;
    if (!c->one_d) {
        CU(cudaMemcpyAsync(dst, src10, (size_t)10 * g.L * sizeof(double), cudaMemcpyDeviceToHost, c->st));
        return 0;
    }
    double* tmp = nullptr;
    CU(cudaMalloc(&tmp, (size_t)6 * g.L * sizeof(double)));
    launch_cols10to6(src10, tmp, g.L, c->st);
    c->launches += 1;
    CU(cudaMemcpyAsync(dst, tmp, (size_t)6 * g.L * sizeof(double), cudaMemcpyDeviceToHost, c->st));
    CU(cudaStreamSynchronize(c->st));
    cudaFree(tmp);
    return 0;
}

extern "C" int dotsocp_upload(dotsocp_ctx* c, const double* phi, const double* q, const double* z, const double* alpha,
                              const double* beta, const double* cvec, const double* weight)
{
    if (!c || !phi || !q || !z || !alpha || !beta || !cvec) return set_err(DOTSOCP_EINVAL, "NULL array");
    if (c->weighted && !weight) return set_err(DOTSOCP_EINVAL, "weighted variant needs weight");
    const Geo& g = c->g;
    // c is -rho0/ht on the first time level, +rho1/ht on the last and zero in between (initialize.m:41-44); only the two
    // planes are kept on the device.
    for (i64 i = g.P; i < g.N - g.P; i++)
        if (cvec[i] != 0.0) return set_err(DOTSOCP_EINVAL, "model.c has a non-zero interior entry at %lld: unsupported", (long long)i);
    c->qcur = 0;
    c->bcur = 0;
    CU(cudaMemcpyAsync(c->phi, phi, g.N * sizeof(double), cudaMemcpyHostToDevice, c->st));
    CU(cudaMemcpyAsync(c->q[0], q, g.Q * sizeof(double), cudaMemcpyHostToDevice, c->st));
    CU(cudaMemcpyAsync(c->alpha, alpha, g.Q * sizeof(double), cudaMemcpyHostToDevice, c->st));
    if (c->weighted) CU(cudaMemcpyAsync(c->weight, weight, g.Q * sizeof(double), cudaMemcpyHostToDevice, c->st));
    CU(cudaMemcpyAsync(c->c0, cvec, g.P * sizeof(double), cudaMemcpyHostToDevice, c->st));
    CU(cudaMemcpyAsync(c->c1, cvec + (g.N - g.P), g.P * sizeof(double), cudaMemcpyHostToDevice, c->st));
    int rc = upload_cols(c, c->beta[0], beta);
    if (rc) return rc;
    rc = upload_cols(c, c->beta[1], z);
    if (rc) return rc;
    c->z_materialised = true;
    CU(cudaStreamSynchronize(c->st));
    c->uploaded = true;
    return DOTSOCP_OK;
}

extern "C" int dotsocp_download(dotsocp_ctx* c, double* phi, double* q, double* z, double* alpha, double* beta)
{
    if (!c) return set_err(DOTSOCP_EINVAL, "NULL ctx");
    if (!c->uploaded) return set_err(DOTSOCP_ESTATE, "download before upload");
    if (!c->z_materialised && z) return set_err(DOTSOCP_ESTATE, "z is not materialised (session still open)");
    const Geo& g = c->g;
    if (phi) CU(cudaMemcpyAsync(phi, c->phi, g.N * sizeof(double), cudaMemcpyDeviceToHost, c->st));
    if (q) CU(cudaMemcpyAsync(q, c->q[c->qcur], g.Q * sizeof(double), cudaMemcpyDeviceToHost, c->st));
    if (alpha) CU(cudaMemcpyAsync(alpha, c->alpha, g.Q * sizeof(double), cudaMemcpyDeviceToHost, c->st));
    if (beta) { int rc = download_cols(c, beta, c->beta[c->bcur]); if (rc) return rc; }
    if (z) { int rc = download_cols(c, z, c->beta[1 - c->bcur]); if (rc) return rc; }
    CU(cudaStreamSynchronize(c->st));
    return DOTSOCP_OK;
}

// ------------------------------------------------------------------------------------------------ loop helpers
static const double UPDATE_RULE[11][2] = {   // solver_socp_inPALM.m:39-51
    {1.1, 1.10}, {1.2, 1.15}, {1.5, 1.20}, {2, 1.26}, {2.5, 1.28}, {3.33, 1.32},
    {5, 1.35}, {10, 1.40}, {20, 1.60}, {40, 1.80}, {50, 2.00}};

static double get_factor(double xi)   // adjust_lagrangianParam.m:47-59
{
    double factor = 1;
    for (int i = 0; i < 11; i++) {
        if (xi >= UPDATE_RULE[i][0]) factor = UPDATE_RULE[i][1];
        else break;
    }
    return factor;
}
static void adjust_lagrangianParam(double& sigma, double xi, double& factor)   // adjust_lagrangianParam.m:14-39
{
    const double lower = 1e-3, upper = 1e3;
    if (xi >= 1) factor = get_factor(xi);
    else if (xi < 1) factor = 1 / get_factor(1 / xi);
    else factor = 1;   // NaN ratio: MATLAB would raise; keep sigma
    if (factor != 1) {
        const double sigmaOld = sigma;
        sigma = fmax(fmin(sigma * factor, upper), lower);
        factor = sigma / sigmaOld;
    }
}
static bool IfAdjustSigma(double it, double last)   // solver_socp_inPALM.m:361-379
{
    const double passed = it - last;
    if (it < 20 && passed >= 3) return true;
    if (it < 50 && passed >= 6) return true;
    if (it < 100 && passed >= 10) return true;
    if (it < 200 && passed >= 15) return true;
    if (it < 500 && passed >= 25) return true;
    return passed >= 40;
}
static double mmax(std::initializer_list<double> v)   // MATLAB max ignores NaN
{
    double m = std::numeric_limits<double>::quiet_NaN();
    for (double x : v)
        if (!std::isnan(x) && (std::isnan(m) || x > m)) m = x;
    return m;
}

static IterScal make_scal(const dotsocp_level_opts& o, double D, double E, double dScale, double tau)
{
    IterScal sc;
    uint64_t b = DSOCP_INV_SQRT2_BITS;
    double lit;
    memcpy(&lit, &b, 8);
    sc.S = E / D;
    sc.SF = lit * sc.S;
    sc.DF = E / dScale;
    sc.tau = tau;
    sc.gt = o.grad_t; sc.gx = o.grad_x; sc.gy = o.grad_y;
    const double tmp = (E / D) * (E / D);      // (E / D)^2, oper_q.m:17
    sc.dinv1 = 1.0 / (1 + 2 * tmp);
    sc.dinv2 = 1.0 / (1 + tmp);
    sc.s2x2 = 2 * tmp;
    sc.s2x1 = tmp;
    return sc;
}

struct Loop {
    dotsocp_ctx* c;
    const dotsocp_level_opts* o;
    IterScal sc;
    UpdateArgs ua() const
    {
        UpdateArgs a;
        a.g = c->g; a.sc = sc; a.phi = c->phi; a.q_old = c->q[c->qcur]; a.q_new = c->q[1 - c->qcur];
        a.alpha = c->alpha; a.weight = c->weight; a.beta_in = c->beta[c->bcur]; a.beta_out = c->beta[1 - c->bcur];
        a.q2 = c->q2; a.rhs = c->rhs; a.c0 = c->c0; a.c1 = c->c1;
        return a;
    }
    // q2, rhs from the current (q, alpha, beta): the z-step part of the first iteration / after any rescaling
    void prologue()
    {
        UpdateArgs a = ua();
        a.q_old = nullptr;
        a.q_new = c->q[c->qcur];
        a.beta_out = nullptr;
        launch_mult(a, c->weighted, c->one_d, false, c->st);
        c->launches += 1;
    }
    void step_phi()
    {
        poisson_solve(c->pp, c->rhs, c->phi, sc_D2, c->st, &c->launches);
    }
    void step_q(bool acc)
    {
        launch_qstep(ua(), c->weighted, acc, c->st);
        c->launches += 1;
    }
    void step_mult()
    {
        launch_mult(ua(), c->weighted, c->one_d, true, c->st);
        c->launches += 1;
        c->qcur ^= 1;
        c->bcur ^= 1;
        c->z_materialised = false;
    }
    double sc_D2 = 1.0;
};

static int fetch_sums(dotsocp_ctx* c, int count)
{
    CU(cudaMemcpyAsync(c->hsums, c->dsums, count * sizeof(double), cudaMemcpyDeviceToHost, c->st));
    CU(cudaStreamSynchronize(c->st));
    return 0;
}

static int sumsq_host(dotsocp_ctx* c, const double* x, i64 n, double* out)
{
    launch_sumsq(x, n, c->partial, c->dsums, c->st);
    c->launches += 2;
    int rc = fetch_sums(c, 1);
    if (rc) return rc;
    *out = c->hsums[0];
    return 0;
}

static double now_s()
{
    using namespace std::chrono;
    return duration<double>(steady_clock::now().time_since_epoch()).count();
}

// ------------------------------------------------------------------------------------------------ the level loops
// One function for the three reference loops; the shared parts (rescaling, KKT, sigma rule, output) are literally the
// same code in solver_socp_inPALM.m, solver_socp_PALM.m and solver_socp_accADMM.m, only the iteration body differs.
//   inPALM / ALG2 : fused kernels, z never stored (recomputed from (q_old, beta_old) where the reference reads it)
//   PALM, acc-ADMM: z is genuine state (it enters the first q-step / the extrapolation), kept in beta[1-bcur]
static int ensure_alloc(double*& p, i64 n)
{
    if (p) return 0;
    cudaError_t e = cudaMalloc(&p, (size_t)n * sizeof(double));
    if (e != cudaSuccess) { cudaGetLastError(); return set_err(DOTSOCP_ENOMEM, "cudaMalloc(%lld doubles): %s", (long long)n, cudaGetErrorString(e)); }
    return 0;
}

static int run_level(dotsocp_ctx* c, const dotsocp_level_opts& o, dotsocp_hist* hist, dotsocp_level_result* res)
{
    const Geo& g = c->g;
    const int method = o.method;
    const bool inpalm = method == DOTSOCP_METHOD_INPALM, palm = method == DOTSOCP_METHOD_PALM, acc = method == DOTSOCP_METHOD_ACCADMM;
    const bool weighted = c->weighted;
    const bool checkPD = o.checkPrimDualFeas < 0 ? !weighted : (o.checkPrimDualFeas != 0);   // :20-24 / wsocp :25-29
    const double time_limit = o.time_limit > 0 ? o.time_limit : 3600;
    const double tau = o.tau;
    double sigma = o.sigma;
    const int maxit = o.maxit;
    const double tol = o.tol;
    const bool checkSByS = o.ifCheckStepByStep != 0;
    double lastSigmaIt = -std::numeric_limits<double>::infinity();
    double cScale = o.cScale, dScale = o.dScale;
    const double D = o.D, E = o.E;
    int use_feasOrg = 0;
    const double tol_feasOrg = 5 * tol;
    int rescale = o.scaling ? 1 : 0;
    const int firstScaleIter = 10, SecondScaleIter = 50, checkRescaleIters = acc ? 200 : 100;   // accADMM :96
    const double ratioThreshold = 1.2;
    double maxFeas = INFINITY, relGap = INFINITY;
    const double h = 1.0 / (double)g.N;
    double norm_c = o.normc, norm_d = o.normd;
    const double kktConst = 1;
    double sigmaScale = 1;
    // acc-ADMM parameters (:11-34)
    const int restart = o.restart > 0 ? o.restart : 100;
    const double stepRho = o.rho > 0 ? o.rho : 2;
    const double stepAlpha = o.theta > 0 ? o.theta : 2;
    if (acc && stepAlpha != 2) return set_err(DOTSOCP_EINVAL, "acc-ADMM: only the Halpern iteration (opts.theta == 2, the default) is available");
    if (palm && (weighted || c->one_d)) return set_err(DOTSOCP_EINVAL, "PALM exists only for socp/dot2d");
    if (acc && c->one_d) return set_err(DOTSOCP_EINVAL, "acc-ADMM does not exist for socp/dot1d");
    int kacc = 0;

    Loop L;
    L.c = c; L.o = &o;
    L.sc = make_scal(o, D, E, dScale, tau);
    L.sc_D2 = D * D;
    const i64 nB = 10 * g.L;
    double* zmat = nullptr;      // PALM / acc: z lives here
    double* tmpq = nullptr;      // PALM: stored A*phi
    i64 vn[5] = {g.N, nB, g.Q, g.Q, nB};
    double* cur[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    if (!inpalm) {
        if (!c->z_materialised) return set_err(DOTSOCP_ESTATE, "z is not materialised");
        zmat = c->beta[1 - c->bcur];
    }
    if (palm) { int rc = ensure_alloc(c->zmat, g.Q); if (rc) return rc; tmpq = c->zmat; }
    if (acc) {
        for (int i = 0; i < 5; i++) {
            int rc = ensure_alloc(c->old_[i], vn[i]); if (rc) return rc;
            rc = ensure_alloc(c->anc_[i], vn[i]); if (rc) return rc;
        }
    }
    auto refresh_cur = [&]() { cur[0] = c->phi; cur[1] = zmat; cur[2] = c->q[c->qcur]; cur[3] = c->alpha; cur[4] = c->beta[c->bcur]; };
    auto copy_to = [&](double** dst) {
        refresh_cur();
        for (int i = 0; i < 5; i++) cudaMemcpyAsync(dst[i], cur[i], (size_t)vn[i] * sizeof(double), cudaMemcpyDeviceToDevice, c->st);
    };

    // alpha, beta, c <- ./sigma  (:102-104)
    launch_scale(c->alpha, g.Q, 1.0, sigma, c->st);
    launch_scale(c->beta[c->bcur], nB, 1.0, sigma, c->st);
    launch_scale(c->c0, g.P, 1.0, sigma, c->st);
    launch_scale(c->c1, g.P, 1.0, sigma, c->st);
    c->launches += 4;
    if (inpalm) L.prologue();   // z2 := d + BF q (:133) folded into the first z-step
    if (palm) {                 // tmp_q = A*phi ; mexBFd(z, tmp_q, ...)   (PALM :137-138)
        UpdateArgs a = L.ua();
        a.q_new = c->q[1 - c->qcur];   // scratch: only tmpq_out matters here
        launch_qstep(a, false, false, c->st, nullptr, tmpq, false);
        launch_cells_update(g, L.sc, false, 2, tmpq, zmat, nullptr, c->st);
        c->launches += 2;
    }
    if (acc) { copy_to(c->old_); copy_to(c->anc_); }   // :157-163

    cudaEvent_t ev_begin, ev_end;
    cudaEventCreate(&ev_begin);
    cudaEventCreate(&ev_end);
    cudaEventRecord(ev_begin, c->st);
    struct Seg { cudaEvent_t a, b; int kind; };
    std::vector<Seg> segs;
    double T[7] = {0, 0, 0, 0, 0, 0, 0};   // 0 lineq, 1 proj, 2 q, 3 mult, 4 kkt, 5 q0 (PALM) / interp (acc)
    auto flush_segs = [&]() {
        for (auto& sg : segs) {
            float ms = 0;
            cudaEventElapsedTime(&ms, sg.a, sg.b);
            T[sg.kind] += ms * 1e-3;
        }
        segs.clear();
        c->evs.reset();
    };
    auto mark = [&]() { cudaEvent_t e = c->evs.get(); cudaEventRecord(e, c->st); return e; };

    const double clock_total = now_s();
    int it = 0, hist_len = 0;
    bool z_ever = false;
    for (it = 1; it <= maxit; it++) {
        // ---------------------------------------------------------------- rescaling :138-190
        bool scaleYes = false;
        double normPhis = 0, normAlps = 0;
        auto rescale_norms = [&]() -> int {
            double s_phi, s_q, s_z, s_a, s_b;
            int rc;
            if ((rc = sumsq_host(c, c->phi, g.N, &s_phi))) return rc;
            if ((rc = sumsq_host(c, c->q[c->qcur], g.Q, &s_q))) return rc;
            if (c->z_materialised) {
                if ((rc = sumsq_host(c, c->beta[1 - c->bcur], nB, &s_z))) return rc;
            } else {
                launch_zstep(g, L.sc, c->one_d, c->q[1 - c->qcur], c->beta[1 - c->bcur], nullptr, c->partial, c->dsums, c->st);
                c->launches += 2;
                if ((rc = fetch_sums(c, 1))) return rc;
                s_z = c->hsums[0];
            }
            if ((rc = sumsq_host(c, c->alpha, g.Q, &s_a))) return rc;
            if ((rc = sumsq_host(c, c->beta[c->bcur], nB, &s_b))) return rc;
            const double normPhi = sqrt(h) * sqrt(s_phi), normQ = sqrt(h) * sqrt(s_q), normZ = sqrt(h) * sqrt(s_z);
            const double normAlpha = sigma * (sqrt(h) * sqrt(s_a)), normBeta = sigma * (sqrt(h) * sqrt(s_b));
            normPhis = mmax({normPhi, normQ, normZ});
            normAlps = mmax({normAlpha, normBeta});
            return 0;
        };
        if (rescale >= 3 && it % checkRescaleIters == 0) {
            int rc = rescale_norms();
            if (rc) return rc;
            const double ratio = fmax(normAlps, normPhis) / fmin(normAlps, normPhis);
            if (ratio > ratioThreshold) scaleYes = true;
        }
        if ((rescale == 1 && maxFeas < 2e-2 && it >= firstScaleIter && relGap < 5e-2) ||
            (rescale == 2 && maxFeas < 5e-3 && it >= SecondScaleIter && relGap < 1e-2) || scaleYes) {
            if (!scaleYes) {
                int rc = rescale_norms();
                if (rc) return rc;
            }
            const double dScale2 = normPhis, cScale2 = normAlps;
            sigma = sigma * (cScale2 / dScale2);
            const double cs2 = cScale2 * cScale2;
            // c, alpha, beta <- x * dScale2 / cScale2^2 ; q, z <- ./dScale2 (inPALM: z is recomputed, never stored)
            launch_scale(c->c0, g.P, dScale2, cs2, c->st);
            launch_scale(c->c1, g.P, dScale2, cs2, c->st);
            norm_c = norm_c / cScale2;
            if (!weighted) norm_d = norm_d / dScale2;
            launch_scale(c->alpha, g.Q, dScale2, cs2, c->st);
            launch_scale(c->beta[c->bcur], nB, dScale2, cs2, c->st);
            if (acc) launch_scale(c->phi, g.N, 1.0, dScale2, c->st);                 // accADMM :207
            if (!palm) launch_scale(c->q[c->qcur], g.Q, 1.0, dScale2, c->st);        // :177 (absent in PALM)
            if (c->z_materialised) launch_scale(c->beta[1 - c->bcur], nB, 1.0, dScale2, c->st);
            if (palm) launch_scale(tmpq, g.Q, 1.0, dScale2, c->st);                  // PALM :191
            c->launches += 7;
            dScale = dScale2 * dScale;
            cScale = cScale2 * cScale;
            L.sc.DF = E / dScale;                                   // scaleD
            sigmaScale = sigmaScale * (cScale2 / dScale2);
            if (inpalm) L.prologue();                               // mexBFd(z2, q, ...) refresh (:187)
            if (acc) { kacc = 0; copy_to(c->old_); copy_to(c->anc_); }   // accADMM :217-222
            rescale += 1;
        }

        // ---------------------------------------------------------------- iteration
        cudaEvent_t e_last;
        if (inpalm) {   // :192-216, fused order
            cudaEvent_t e0 = mark();
            L.step_phi();
            cudaEvent_t e1 = mark();
            L.step_q(false);
            cudaEvent_t e2 = mark();
            L.step_mult();
            cudaEvent_t e3 = mark();
            segs.push_back({e0, e1, 0});
            segs.push_back({e1, e2, 2});
            segs.push_back({e2, e3, 3});
            e_last = e3;
            z_ever = true;
        } else if (palm) {   // solver_socp_PALM.m:196-224
            double* q = c->q[c->qcur];
            double* beta = c->beta[c->bcur];
            UpdateArgs a = L.ua();
            a.q_new = q;
            cudaEvent_t e0 = mark();
            launch_bfdconj_sum(g, L.sc.S, zmat, beta, c->q2, c->st);
            launch_qstep(a, false, false, c->st, tmpq, nullptr, false);              // q = (tmp_q + alpha + q2).*diagQInv
            cudaEvent_t e1 = mark();
            launch_rhs(g, L.sc, false, q, c->alpha, nullptr, c->c0, c->c1, c->rhs, c->st);
            c->launches += 3;
            L.step_phi();
            cudaEvent_t e2 = mark();
            launch_zstep(g, L.sc, false, q, beta, zmat, c->partial, c->dsums, c->st);   // mexBFd + mexProjSoc (:209-210)
            cudaEvent_t e3 = mark();
            launch_bfdconj_sum(g, L.sc.S, zmat, beta, c->q2, c->st);
            launch_qstep(a, false, false, c->st, nullptr, tmpq, true);               // tmp_q = A*phi ; q ; alpha
            cudaEvent_t e4 = mark();
            launch_cells_update(g, L.sc, false, 0, q, zmat, beta, c->st);            // beta += tau (z - z2)
            cudaEvent_t e5 = mark();
            c->launches += 5;
            segs.push_back({e0, e1, 5});
            segs.push_back({e1, e2, 0});
            segs.push_back({e2, e3, 1});
            segs.push_back({e3, e4, 2});
            segs.push_back({e4, e5, 3});
            e_last = e5;
        } else {   // solver_socp_accADMM.m:227-249
            double* q = c->q[c->qcur];
            double* beta = c->beta[c->bcur];
            UpdateArgs a = L.ua();
            a.q_new = q;
            cudaEvent_t e0 = mark();
            launch_bfdconj_sum(g, L.sc.S, zmat, beta, c->q2, c->st);
            launch_qstep(a, weighted, true, c->st);                                  // q ; alpha = (alpha + A phi) - w.*q
            cudaEvent_t e1 = mark();
            launch_rhs(g, L.sc, weighted, q, c->alpha, c->weight, c->c0, c->c1, c->rhs, c->st);
            L.step_phi();
            cudaEvent_t e2 = mark();
            launch_cells_update(g, L.sc, false, 1, q, zmat, beta, c->st);            // beta = (beta + z) - z2 ; z = Pi_Q(z2 - beta)
            cudaEvent_t e3 = mark();
            c->launches += 4;
            segs.push_back({e0, e1, 2});
            segs.push_back({e1, e2, 0});
            segs.push_back({e2, e3, 3});
            e_last = e3;
        }

        // ---------------------------------------------------------------- kkt :218-324
        const bool adjustSigmaYes = IfAdjustSigma(it, lastSigmaIt);
        bool over_time = (now_s() - clock_total) > time_limit;
        if (over_time) {   // the host runs ahead of the device: confirm against completed work
            cudaStreamSynchronize(c->st);
            over_time = (now_s() - clock_total) > time_limit;
        }
        const bool check = checkSByS || adjustSigmaYes || it == maxit || over_time;
        bool stop = false;
        if (check) {
            launch_bfdconj(g, L.sc.S, c->beta[c->bcur], c->qtmp, c->st);   // q2 = s (BF)^* beta   (:225)
            KktArgs ka;
            ka.g = g; ka.sc = L.sc; ka.sigma = sigma; ka.cScale = cScale; ka.dScale = dScale; ka.D = D; ka.E = E;
            ka.phi = c->phi; ka.q = c->q[c->qcur]; ka.alpha = c->alpha; ka.weight = c->weight;
            ka.beta = c->beta[c->bcur]; ka.z = zmat; ka.q_old = c->q[1 - c->qcur]; ka.beta_old = c->beta[1 - c->bcur];
            ka.q2b = c->qtmp; ka.c0 = c->c0; ka.c1 = c->c1; ka.partial = c->partial;
            ka.out = c->dsums;
            launch_kkt_cells(ka, weighted, c->one_d, c->st);
            ka.out = c->dsums + KC_COUNT;
            launch_kkt_nodes(ka, weighted, c->st);
            c->launches += 5;
            cudaEvent_t e4 = mark();
            segs.push_back({e_last, e4, 4});
            int rc = fetch_sums(c, KC_COUNT + KN_COUNT);
            if (rc) return rc;
            flush_segs();
            const double* sc_ = c->hsums;
            const double* sn = c->hsums + KC_COUNT;
            auto nrm = [&](double v) { return sqrt(h) * sqrt(v); };
            const double norm_q = nrm(sn[KN_Q2]);
            const double norm_z = nrm(sc_[KC_Z2]);
            const double norm_Aphi = nrm(sn[KN_APHI2]);
            const double norm_alpha = sigma * nrm(sn[KN_ALPHA2]);
            const double norm_beta = sigma * nrm(sc_[KC_BETA2]);
            const double norm_FBbeta = sigma * nrm(sn[KN_FBB2]);
            const double primFea1 = nrm(sn[KN_PRIM1]);
            const double primFea2 = nrm(sc_[KC_PRIM2]);
            const double dualFea1 = sigma * nrm(sn[KN_DUAL1]);
            const double dualFea2 = sigma * nrm(sn[KN_DUAL2]);
            const double complem = nrm(sc_[KC_COMPL]);
            const double dotcomplem = nrm(sc_[KC_DOTC]);
            const double normRho = nrm(sc_[KC_RHOT]);
            const double norm_rhoFq = nrm(sc_[KC_RHOFQ]);
            const double mRhoB = nrm(sn[KN_MRHOB]);    // sqrt(normL2(mx-rhoBx)^2 + normL2(my-rhoBy)^2)
            const double normM = nrm(sn[KN_M2]);
            const double normRhoB = nrm(sn[KN_RHOB2]);
            const double den2o = weighted ? (kktConst * E / dScale + norm_q + norm_z) : (kktConst * E / dScale + norm_d);
            const double den2 = weighted ? (kktConst + norm_q + norm_z) : (kktConst + norm_d);
            const double KO[7] = {primFea1 / (kktConst * D / dScale + norm_Aphi + norm_q),
                                  primFea2 / den2o,
                                  dualFea1 / (kktConst / cScale + norm_c),
                                  complem / (kktConst * E / dScale + norm_z + norm_beta),
                                  dualFea2 / (kktConst / cScale / D + norm_FBbeta + norm_alpha),
                                  dotcomplem / (kktConst + normRho + norm_rhoFq),
                                  mRhoB / (kktConst + normM + normRhoB)};
            const double KR[5] = {primFea1 / (kktConst + norm_Aphi + norm_q), primFea2 / den2,
                                  dualFea1 / (kktConst + norm_c), complem / (kktConst + norm_z + norm_beta),
                                  dualFea2 / (kktConst + norm_FBbeta + norm_alpha)};
            const double priVal = (sigma * cScale * dScale * h) * sn[KN_QDOTA];
            const double dualVal = (sigma * cScale * dScale * h) * sn[KN_CPHI];
            const double pdGap = fabs(priVal - dualVal) / (1 + fabs(priVal) + fabs(dualVal));
            if (hist && hist_len < hist->cap) {
                if (hist->kkt) for (int j = 0; j < 7; j++) hist->kkt[(size_t)hist_len * 7 + j] = KO[j];
                if (hist->time) hist->time[hist_len] = now_s() - clock_total;
                if (hist->iter) hist->iter[hist_len] = it;
                if (hist->pdGap) hist->pdGap[hist_len] = pdGap;
                if (hist->priVal) hist->priVal[hist_len] = priVal;
                if (hist->dualVal) hist->dualVal[hist_len] = dualVal;
            }
            hist_len++;
            const double stopv = checkPD ? mmax({KO[0], KO[2], KO[5], KO[6]}) : mmax({KO[0], KO[2], KO[5]});
            if (stopv < tol || (now_s() - clock_total) > time_limit) {
                stop = true;
            } else {
                if (mmax({KR[0], KR[1], KR[2], KR[3], KR[4]}) < tol_feasOrg) use_feasOrg = 1;
                if (adjustSigmaYes) {
                    lastSigmaIt = it;
                    double resiPri, resiDual;
                    if (use_feasOrg) { resiPri = mmax({KO[0], KO[1]}); resiDual = mmax({KO[2], KO[4]}); }
                    else { resiPri = mmax({KR[0], KR[1]}); resiDual = mmax({KR[2], KR[4]}); }
                    double factor = 1;
                    adjust_lagrangianParam(sigma, resiPri / resiDual, factor);
                    if (factor != 1) {
                        launch_scale(c->alpha, g.Q, 1.0, factor, c->st);
                        launch_scale(c->beta[c->bcur], nB, 1.0, factor, c->st);
                        launch_scale(c->c0, g.P, 1.0, factor, c->st);
                        launch_scale(c->c1, g.P, 1.0, factor, c->st);
                        c->launches += 4;
                        // inPALM: q2, rhs were computed with the old alpha/beta: refresh.  The prologue reads only the
                        // current buffers, so the (q_old, beta_old) pair that defines z survives.
                        if (inpalm) L.prologue();
                        if (acc) {   // accADMM :346-358
                            launch_scale(c->old_[3], g.Q, 1.0, factor, c->st);
                            launch_scale(c->old_[4], nB, 1.0, factor, c->st);
                            c->launches += 2;
                            kacc = 0;
                            copy_to(c->anc_);
                        }
                    }
                }
                if (rescale > 0) {
                    maxFeas = mmax({KR[0], KR[1], KR[2], KR[3], KR[4]});
                    relGap = pdGap;
                }
            }
        }
        if (stop) break;
        if (acc) {   // Halpern iteration, accADMM :371-388
            cudaEvent_t e5 = mark();
            const double c1 = 1.0 / (kacc + 2), c2 = (double)(kacc + 1) / (kacc + 2);
            kacc += 1;
            const bool anchor = kacc >= restart;
            refresh_cur();
            for (int i = 0; i < 5; i++) launch_halpern(cur[i], c->old_[i], c->anc_[i], vn[i], c1, c2, stepRho, anchor, c->st);
            c->launches += 5;
            if (anchor) kacc = 0;
            cudaEvent_t e6 = mark();
            segs.push_back({e5, e6, 5});
        }
    }
    if (it > maxit) it = maxit;
    // ---------------------------------------------------------------- output :328-357
    if (!c->z_materialised && z_ever) {
        // z = Pi_Q(d + BF q_old - beta_old), written over beta_old (cell-local, safe in place)
        launch_zstep(g, L.sc, c->one_d, c->q[1 - c->qcur], c->beta[1 - c->bcur], c->beta[1 - c->bcur], c->partial, c->dsums, c->st);
        c->launches += 2;
        c->z_materialised = true;
    }
    launch_scale(c->alpha, g.Q, sigma, 1.0, c->st);               // var.alpha = sigma*alpha
    launch_scale(c->beta[c->bcur], nB, sigma, 1.0, c->st);        // var.beta  = sigma*beta
    launch_scale(c->c0, g.P, sigma, 1.0, c->st);                  // undo the folding of model.c (the reference never
    launch_scale(c->c1, g.P, sigma, 1.0, c->st);                  // writes its local copy back; keeps the session reusable)
    c->launches += 4;
    cudaEventRecord(ev_end, c->st);
    CU(cudaStreamSynchronize(c->st));
    flush_segs();
    float total_ms = 0;
    cudaEventElapsedTime(&total_ms, ev_begin, ev_end);
    cudaEventDestroy(ev_begin);
    cudaEventDestroy(ev_end);
    CU(cudaGetLastError());
    if (res) {
        memset(res, 0, sizeof(*res));
        res->iters = it;
        res->hist_len = hist_len;
        res->sigma = sigma / sigmaScale;
        res->cScale = cScale; res->dScale = dScale; res->D = D; res->E = E;
        const double tot = total_ms * 1e-3;
        if (inpalm) { res->times[0] = T[0]; res->times[1] = T[1]; res->times[2] = T[2]; res->times[3] = T[3]; res->times[4] = T[4]; res->times[5] = tot; }
        else if (palm) { res->times[0] = T[5]; res->times[1] = T[0]; res->times[2] = T[1]; res->times[3] = T[2]; res->times[4] = T[3]; res->times[5] = T[4]; res->times[6] = tot; }
        else { res->times[0] = T[2]; res->times[1] = T[3]; res->times[2] = T[0]; res->times[3] = T[1]; res->times[4] = T[4]; res->times[5] = T[5]; res->times[6] = tot; }
        res->gpu_launches = c->launches;
    }
    return DOTSOCP_OK;
}

extern "C" int dotsocp_run(dotsocp_ctx* c, const dotsocp_level_opts* o, dotsocp_hist* hist, dotsocp_level_result* res)
{
    if (!c || !o) return set_err(DOTSOCP_EINVAL, "NULL ctx/opts");
    if (!c->uploaded) return set_err(DOTSOCP_ESTATE, "run before upload");
    if (c->iter_open) return set_err(DOTSOCP_ESTATE, "a benchmark session is open");
    if (o->variant != c->variant || o->nt != c->g.nt || o->nx != c->g.nx || o->ny != c->g.ny)
        return set_err(DOTSOCP_EINVAL, "opts do not match the context (variant/grid)");
    if (o->maxit < 1) return set_err(DOTSOCP_EINVAL, "maxit must be >= 1");
    if (o->method < 0 || o->method > 2) return set_err(DOTSOCP_EINVAL, "unknown method %d", o->method);
    return run_level(c, *o, hist, res);
}

extern "C" int dotsocp_solve_level(const dotsocp_level_opts* o, double* phi, double* q, double* z, double* alpha, double* beta,
                                   const double* cvec, const double* weight, dotsocp_hist* hist, dotsocp_level_result* res)
{
    if (!o) return set_err(DOTSOCP_EINVAL, "NULL opts");
    dotsocp_ctx* c = nullptr;
    int rc = dotsocp_create(&c, o->variant, o->nt, o->nx, o->ny, 0, 1, nullptr);
    if (rc) return rc;
    rc = dotsocp_upload(c, phi, q, z, alpha, beta, cvec, weight);
    if (!rc) rc = dotsocp_run(c, o, hist, res);
    if (!rc) rc = dotsocp_download(c, phi, q, z, alpha, beta);
    dotsocp_destroy(c);
    return rc;
}

// ------------------------------------------------------------------------------------------------ benchmark session
extern "C" int dotsocp_iter_begin(dotsocp_ctx* c, const dotsocp_level_opts* o)
{
    if (!c || !o) return set_err(DOTSOCP_EINVAL, "NULL ctx/opts");
    if (!c->uploaded) return set_err(DOTSOCP_ESTATE, "iter_begin before upload");
    const Geo& g = c->g;
    c->sc = make_scal(*o, o->D, o->E, o->dScale, o->tau);
    c->sigma_fold = o->sigma;
    c->D2 = o->D * o->D;
    launch_scale(c->alpha, g.Q, 1.0, o->sigma, c->st);
    launch_scale(c->beta[c->bcur], 10 * g.L, 1.0, o->sigma, c->st);
    launch_scale(c->c0, g.P, 1.0, o->sigma, c->st);
    launch_scale(c->c1, g.P, 1.0, o->sigma, c->st);
    c->launches += 4;
    Loop L; L.c = c; L.o = o; L.sc = c->sc; L.sc_D2 = o->D * o->D;
    L.prologue();
    CU(cudaStreamSynchronize(c->st));
    c->iter_open = true;
    return DOTSOCP_OK;
}

extern "C" int dotsocp_iterate(dotsocp_ctx* c, int n_iters, int with_kkt_every, float* elapsed_ms, float* ms_by_kernel)
{
    (void)with_kkt_every;
    if (!c || !c->iter_open) return set_err(DOTSOCP_ESTATE, "iterate without iter_begin");
    Loop L; L.c = c; L.o = nullptr; L.sc = c->sc;
    L.sc_D2 = c->D2;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    std::vector<cudaEvent_t> ev;
    if (ms_by_kernel) {
        ev.resize((size_t)n_iters * 4);
        for (auto& e : ev) cudaEventCreate(&e);
    }
    cudaEventRecord(a, c->st);
    for (int i = 0; i < n_iters; i++) {
        if (ms_by_kernel) cudaEventRecord(ev[4 * i + 0], c->st);
        L.step_phi();
        if (ms_by_kernel) cudaEventRecord(ev[4 * i + 1], c->st);
        L.step_q(false);
        if (ms_by_kernel) cudaEventRecord(ev[4 * i + 2], c->st);
        L.step_mult();
        if (ms_by_kernel) cudaEventRecord(ev[4 * i + 3], c->st);
    }
    cudaEventRecord(b, c->st);
    CU(cudaStreamSynchronize(c->st));
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    if (elapsed_ms) *elapsed_ms = ms;
    if (ms_by_kernel) {
        ms_by_kernel[0] = ms_by_kernel[1] = ms_by_kernel[2] = ms_by_kernel[3] = 0;
        for (int i = 0; i < n_iters; i++)
            for (int k = 0; k < 3; k++) {
                float t = 0;
                cudaEventElapsedTime(&t, ev[4 * i + k], ev[4 * i + k + 1]);
                ms_by_kernel[k] += t;
            }
        for (auto& e : ev) cudaEventDestroy(e);
    }
    cudaEventDestroy(a); cudaEventDestroy(b);
    CU(cudaGetLastError());
    return DOTSOCP_OK;
}

extern "C" int dotsocp_iter_end(dotsocp_ctx* c)
{
    if (!c || !c->iter_open) return set_err(DOTSOCP_ESTATE, "iter_end without iter_begin");
    const Geo& g = c->g;
    if (!c->z_materialised) {
        launch_zstep(g, c->sc, c->one_d, c->q[1 - c->qcur], c->beta[1 - c->bcur], c->beta[1 - c->bcur], c->partial, c->dsums, c->st);
        c->launches += 2;
        c->z_materialised = true;
    }
    launch_scale(c->alpha, g.Q, c->sigma_fold, 1.0, c->st);
    launch_scale(c->beta[c->bcur], 10 * g.L, c->sigma_fold, 1.0, c->st);
    launch_scale(c->c0, g.P, c->sigma_fold, 1.0, c->st);
    launch_scale(c->c1, g.P, c->sigma_fold, 1.0, c->st);
    c->launches += 4;
    CU(cudaStreamSynchronize(c->st));
    c->iter_open = false;
    return DOTSOCP_OK;
}

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
    if (what == 2) poisson_solve(pp, a.p, b.p, D * D, 0, nullptr);
    else { CU(cudaMemcpy(b.p, a.p, (size_t)N * sizeof(double), cudaMemcpyDeviceToDevice)); poisson_dctn(pp, b.p, what == 1, 0, nullptr); }
    cudaError_t e = cudaDeviceSynchronize();
    poisson_plan_destroy(pp);
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
