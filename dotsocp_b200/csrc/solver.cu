// solver.cu -- device-resident session (one or several time slabs), host control loop and the session part of the C ABI.
//
// The control flow mirrors socp/dot2d/algorithms/solver_socp_inPALM.m line by line (rescaling :138-190, iteration
// :192-216, KKT :218-324, output :328-357); only the order of the cell-local steps inside one iteration is fused
// differently (see kernels_update.cu / DESIGN.md): the z-step of iteration i+1 does not depend on phi_{i+1}, so it is
// evaluated inside the multiplier kernel of iteration i.  All scalar decisions (sigma rule, rescale triggers, stop
// test) stay on the host in double precision, exactly as in the reference.
//
// Multi-GPU: the grid is cut into `world` time slabs.  Every array keeps its global index space (vmm.h) and each slab
// backs its own levels plus one ghost level per side; per iteration the slabs exchange one ghost plane of phi, the
// ghost planes of the new q and of alpha_0 (neighbour send/recv) and transpose the spectrum twice (all-to-all) around
// the t-pass of the Poisson solve; KKT sums are all-reduced.  A process owns either one slab (NCCL between processes,
// one process per GPU) or -- world > 1 without an NCCL id -- all slabs on its own device ("emulation": the same code
// path with device-to-device copies instead of NCCL, used to test the slab logic on a single GPU).
#include "../../include/dotsocp.h"
#include "errs.h"
#include "kernels.h"
#include "vmm.h"
#include "hostcopy.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <mutex>
#include <string>
#include <vector>

using namespace dsocp;

// ------------------------------------------------------------------------------------------------ error plumbing
static thread_local char g_err[1024] = "";
int dsocp_set_err(int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}
extern "C" const char* dotsocp_last_error(void) { return g_err; }
extern "C" int dotsocp_version(void) { return 101; }

int dsocp_require_device()
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
#define NC(call)                                                                                       \
    do {                                                                                               \
        int r_ = (call);                                                                               \
        if (r_ != 0) return set_err(DOTSOCP_ENCCL, "%s:%d %s: %s", __FILE__, __LINE__, #call, nccl_api().GetErrorString(r_)); \
    } while (0)

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

struct Range { i64 b, e; };
typedef std::vector<Range> Ranges;

static void add_range(Ranges& r, i64 b, i64 e)
{
    if (e <= b) return;
    if (!r.empty() && r.back().e == b) r.back().e = e;   // merge adjacent pieces (one slab => whole array)
    else r.push_back({b, e});
}

struct Slab {
    int id = 0;
    TRange tr;
    int lo_c = 0, hi_c = 0, lo_n = 0, hi_n = 0;   // backed cell layers / node levels (owned + ghosts)
    i64 p0 = 0, p1 = 0;                           // (x,y) modes this slab solves along t
    SparseArray a_phi, a_rhs, a_q[2], a_alpha, a_q2, a_qtmp, a_weight, a_beta[2];
    double *phi = nullptr, *rhs = nullptr, *q[2] = {nullptr, nullptr}, *alpha = nullptr, *q2 = nullptr, *qtmp = nullptr,
           *weight = nullptr, *beta[2] = {nullptr, nullptr};
    double *c0 = nullptr, *c1 = nullptr, *partial = nullptr;
    double* side = nullptr;                              // side buffer of the aligned k_mult (kernels.h: mult_side_doubles)
    // tensor maps of the arrays k_mult streams (DOTSOCP_KM_PF=4); km = the set of the current launch
    bool maps_ok = false;
    CUtensorMap tm_beta[2], tm_q[2][3], tm_alpha[3], tm_w[3];
    KmMaps km;
    double *partial_q = nullptr, *partial_m = nullptr;   // fused KKT partials (allocated at the first fused check)
    double *tsend = nullptr, *trecv = nullptr;   // transposed t-solve (DOTSOCP_TSOLVE=transpose)
    double* carry[4] = {nullptr, nullptr, nullptr, nullptr};   // pipelined t-solve: forward in / out, backward in / out (P doubles each)
    // ... with the hand-off inside the kernels (one process per GPU): xbuf = [forward in | backward in | flags] is ONE allocation
    // (one CUDA-IPC handle) that the two neighbours map; carry[0] / carry[2] then point into it
    double* xbuf = nullptr;
    int *tflags = nullptr, *tdone = nullptr;        // epoch flags written by the neighbours (in xbuf) / local CTA counters
    double *peer_up = nullptr, *peer_dn = nullptr;  // xbuf of slab id+1 / id-1 (peer memory)
    std::vector<double*> peer_tsend, peer_trecv;   // CUDA-IPC mappings of the other ranks' transpose buffers (NCCL mode)
    // direct exchange (kernels store into the destination slab's buffer): device tables of `world` pointers
    //   d_fwd[b] = row tn0 of the t-solve buffer (trecv) of slab b   -- written by this slab's forward x pass
    //   d_bwd[a] = block of this slab's mode chunk in the packed buffer (tsend) of slab a   -- written by this slab's t-solve
    double **d_fwd = nullptr, **d_bwd = nullptr;
    Ranges n_own, n_all, q_own, q_all, b_own, b_all;   // element ranges of node / staggered / 10-column arrays
    // acc-ADMM / PALM state (single slab only)
    double* tmpq = nullptr;
    double* old_[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    double* anc_[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    ~Slab()
    {
        for (double* p : peer_tsend) if (p) cudaIpcCloseMemHandle(p);
        for (double* p : peer_trecv) if (p) cudaIpcCloseMemHandle(p);
        cudaFree(c0); cudaFree(c1); cudaFree(partial); cudaFree(side); cudaFree(partial_q); cudaFree(partial_m); cudaFree(tsend); cudaFree(trecv);
        cudaFree(tmpq); cudaFree(d_fwd); cudaFree(d_bwd);
        if (peer_up) cudaIpcCloseMemHandle(peer_up);
        if (peer_dn) cudaIpcCloseMemHandle(peer_dn);
        if (xbuf) { carry[0] = carry[2] = nullptr; cudaFree(xbuf); }
        cudaFree(tdone);
        for (double* p : carry) cudaFree(p);
        for (int i = 0; i < 5; i++) { cudaFree(old_[i]); cudaFree(anc_[i]); }
    }
};

struct dotsocp_ctx {
    int variant = 0;
    bool one_d = false, weighted = false;
    bool weight_set = false;    // the weight came from a device pyramid (dotsocp_set_weight): upload / prolong may pass NULL
    int world = 1;
    bool emulate = false;       // all slabs in this process
    int my = 0;                 // first (NCCL mode: only) local slab id
    void* comm = nullptr;       // ncclComm_t
    int device = 0;
    Geo g;                      // device layout: rows of the staggered / 10-column arrays pitched to 256 bytes (common.cuh)
    Geo gh;                     // the reference's packed layout: what the host arrays of the C ABI use
    cudaStream_t st = nullptr;
    cudaStream_t st2 = nullptr;  // communication stream (time slabs): transposes and ghost planes overlap with compute
    std::vector<cudaEvent_t> cev;   // reusable events for the st <-> st2 hand-offs
    size_t cev_used = 0;
    cudaEvent_t comm_event()
    {
        if (cev_used == cev.size()) { cudaEvent_t e; cudaEventCreateWithFlags(&e, cudaEventDisableTiming); cev.push_back(e); }
        return cev[cev_used++];
    }
    // DOTSOCP_TRACE=1: device time of the phases of the distributed Poisson solve (printed by rank/slab 0 at destroy)
    bool trace = false;
    bool tpipe = false;         // t-solve of the slabs by the pipelined Thomas sweeps (default) instead of transposes
    bool tpush = false;         // ... whose carry planes go GPU to GPU inside the kernels (peer stores + epoch flags), not by NCCL
    int tepoch = 0;             // ... number of the current solve (the value the flags are compared with)
    int tchunks = 1;            // ... with the modes cut into this many chunks so that consecutive slabs overlap
    double* d_line0 = nullptr;  // ... the singular mode's nt values
    bool ipc = false;           // transposes by peer-to-peer copies (copy engines over NVLink) instead of NCCL send/recv
    bool direct = false;        // ... or by the transform / t-solve kernels storing straight into the destination slab's buffer
    int* d_tcut = nullptr;      // first node level of every slab (world + 1 entries), device copy
    double* barrier_buf = nullptr;
    std::vector<cudaStream_t> cps;   // one copy stream per peer so that the pushes use several copy engines at once
    std::vector<cudaEvent_t> tev;
    double tacc[6] = {0, 0, 0, 0, 0, 0};
    long tcount = 0, tseen = 0;
    std::vector<Slab*> slabs;   // local slabs
    std::vector<TRange> part;   // partition of all `world` slabs
    std::vector<i64> pcut;      // mode chunks [pcut[r], pcut[r+1])
    int qcur = 0, bcur = 0;
    bool z_materialised = true; // beta[1-bcur] holds z itself (after upload / at exit) instead of beta_old
    bool z_absent = false;      // upload was given no z: legal only for inPALM with maxit >= 1, which overwrites z before any read
    PoissonPlan* pp = nullptr;
    // KKT / norm sums: level table [nt][KSL] (every slab fills its own rows), totals, pinned host copies
    double *d_lvl = nullptr, *d_tot = nullptr, *h_tot = nullptr, *h_elapsed = nullptr;
    bool fuse_kkt = true;        // DOTSOCP_KKT=separate: always use the stand-alone KKT kernels (A/B and tests)
    double launches = 0;
    bool uploaded = false;
    bool iter_open = false;
    IterScal sc;
    double sigma_fold = 1.0, D2 = 1.0;
    double it_cScale = 1.0, it_dScale = 1.0, it_D = 1.0, it_E = 1.0;   // scalars of the open benchmark session
    EvPool evs;
    Slab* local(int id) const
    {
        for (Slab* s : slabs) if (s->id == id) return s;
        return nullptr;
    }
};

// CTA partials of the stand-alone KKT / norm kernels for `nlev` time levels
static size_t partial_doubles(const Geo& g, int nlev)
{
    const int k = std::max(std::max((int)KN_COUNT, (int)KC_COUNT), (int)NR_COUNT);
    return (size_t)kkt_blocks_x(g) * (size_t)nlev * k + 64;
}

static void* g_comm = nullptr;   // process-wide NCCL communicator (one process per GPU)
static int g_comm_world = 0, g_comm_rank = -1;

extern "C" int dotsocp_nccl_unique_id(char id128[128])
{
    if (!id128) return set_err(DOTSOCP_EINVAL, "NULL id buffer");
    const NcclApi& n = nccl_api();
    if (!n.ok) return set_err(DOTSOCP_ENCCL, "NCCL unavailable: %s", n.why);
    NcclId id;
    NC(n.GetUniqueId(&id));
    memcpy(id128, id.internal, 128);
    return DOTSOCP_OK;
}

extern "C" void dotsocp_destroy(dotsocp_ctx* c)
{
    if (!c) return;
    if (c->st) cudaStreamSynchronize(c->st);
    if (c->st2) cudaStreamSynchronize(c->st2);
    if (c->trace && c->tcount > 0 && c->my == 0)
        fprintf(stderr, "[dotsocp trace] distributed Poisson, mean ms over %ld solves: (y,x) forward %.3f | wait for all-to-all %.3f | "
                "t-solve %.3f | all-to-all back + (x,y) inverse %.3f | phi ghost %.3f\n", c->tcount, c->tacc[0] / c->tcount,
                c->tacc[1] / c->tcount, c->tacc[2] / c->tcount, c->tacc[3] / c->tcount, c->tacc[4] / c->tcount);
    for (auto e : c->tev) cudaEventDestroy(e);
    if ((c->ipc || c->tpush) && c->comm && c->barrier_buf) {   // nobody may unmap / free while a peer can still touch the buffers
        nccl_api().AllReduce(c->barrier_buf, c->barrier_buf, 1, NCCL_FLOAT64, NCCL_SUM, c->comm, c->st);
        cudaStreamSynchronize(c->st);
    }
    cudaFree(c->barrier_buf);
    cudaFree(c->d_tcut);
    cudaFree(c->d_lvl);
    cudaFree(c->d_line0);
    cudaFree(c->d_tot);
    if (c->h_tot) cudaFreeHost(c->h_tot);
    if (c->h_elapsed) cudaFreeHost(c->h_elapsed);
    for (auto e : c->cev) cudaEventDestroy(e);
    for (auto cs : c->cps) cudaStreamDestroy(cs);
    if (c->st2) cudaStreamDestroy(c->st2);
    for (Slab* s : c->slabs) delete s;
    // the communicator is process-wide (g_comm) and survives the session so that later sessions can reuse it
    poisson_plan_destroy(c->pp);
    if (c->st) cudaStreamDestroy(c->st);
    delete c;
}

static int make_slab(dotsocp_ctx* c, int id)
{
    const Geo& g = c->g;
    Slab* s = new Slab();
    c->slabs.push_back(s);
    s->id = id;
    s->tr = c->part[id];
    const TRange& tr = s->tr;
    const bool dense = c->world == 1;
    s->lo_c = tr.tc0 > 0 ? tr.tc0 - 1 : 0;
    s->hi_c = tr.tc1 < g.nt - 1 ? tr.tc1 + 1 : g.nt - 1;
    s->lo_n = tr.tn0 > 0 ? tr.tn0 - 1 : 0;
    s->hi_n = tr.tn1 < g.nt ? tr.tn1 + 1 : g.nt;
    s->p0 = c->pcut[id];
    s->p1 = c->pcut[id + 1];
    add_range(s->n_own, tr.tn0 * g.P, tr.tn1 * g.P);
    add_range(s->n_all, s->lo_n * g.P, s->hi_n * g.P);
    add_range(s->q_own, tr.tc0 * g.PC, tr.tc1 * g.PC);
    add_range(s->q_own, g.L + tr.tn0 * g.PBX, g.L + tr.tn1 * g.PBX);
    add_range(s->q_own, g.L + g.NBX + tr.tn0 * g.PBY, g.L + g.NBX + tr.tn1 * g.PBY);
    add_range(s->q_all, s->lo_c * g.PC, s->hi_c * g.PC);
    add_range(s->q_all, g.L + s->lo_n * g.PBX, g.L + s->hi_n * g.PBX);
    add_range(s->q_all, g.L + g.NBX + s->lo_n * g.PBY, g.L + g.NBX + s->hi_n * g.PBY);
    for (int j = 0; j < 10; j++) {
        add_range(s->b_own, j * g.L + tr.tc0 * g.PC, j * g.L + tr.tc1 * g.PC);
        add_range(s->b_all, j * g.L + s->lo_c * g.PC, j * g.L + s->hi_c * g.PC);
    }
    auto win = [](const Ranges& r) {
        std::vector<std::pair<long long, long long>> w;
        for (auto& x : r) w.emplace_back(x.b, x.e);
        return w;
    };
    const char* why = "";
#define MK(arr, count, ranges)                                                                        \
    do {                                                                                              \
        int e_ = arr.create((count), win(ranges), dense, c->device, &why);                            \
        if (e_) return set_err(e_ == 2 ? DOTSOCP_ENOMEM : DOTSOCP_ECUDA, "device array of %lld doubles: %s", (long long)(count), why); \
    } while (0)
    MK(s->a_phi, g.N, s->n_all); MK(s->a_rhs, g.N, s->n_all);
    MK(s->a_q[0], g.Q, s->q_all); MK(s->a_q[1], g.Q, s->q_all); MK(s->a_alpha, g.Q, s->q_all); MK(s->a_q2, g.Q, s->q_all);
    // (a_qtmp, the scratch of the stand-alone KKT check, is created on first use: ensure_qtmp)
    if (c->weighted) MK(s->a_weight, g.Q, s->q_all);
    MK(s->a_beta[0], 10 * g.L, s->b_all); MK(s->a_beta[1], 10 * g.L, s->b_all);
#undef MK
    s->phi = s->a_phi.ptr(); s->rhs = s->a_rhs.ptr(); s->q[0] = s->a_q[0].ptr(); s->q[1] = s->a_q[1].ptr();
    s->alpha = s->a_alpha.ptr(); s->q2 = s->a_q2.ptr(); s->qtmp = nullptr;
    s->weight = c->weighted ? s->a_weight.ptr() : nullptr;
    s->beta[0] = s->a_beta[0].ptr(); s->beta[1] = s->a_beta[1].ptr();
    if (!g.packed()) {
        // pad columns of the pitched rows are never written by the kernels: they must hold zeros (weight: ones, see
        // dotsocp_upload) so that the flat streaming kernels (scalings, extrapolations) keep them finite
        double* qa[] = {s->q[0], s->q[1], s->alpha, s->q2, s->qtmp, s->weight};
        for (double* a : qa)
            if (a) for (auto& x : s->q_all) CU(cudaMemsetAsync(a + x.b, 0, (size_t)(x.e - x.b) * sizeof(double), c->st));
        for (int k = 0; k < 2; k++)
            for (auto& x : s->b_all) CU(cudaMemsetAsync(s->beta[k] + x.b, 0, (size_t)(x.e - x.b) * sizeof(double), c->st));
    }
    CU(cudaMalloc(&s->c0, g.P * sizeof(double)));
    CU(cudaMalloc(&s->c1, g.P * sizeof(double)));
    CU(cudaMemsetAsync(s->c0, 0, g.P * sizeof(double), c->st));
    CU(cudaMemsetAsync(s->c1, 0, g.P * sizeof(double), c->st));
    CU(cudaMalloc(&s->partial, partial_doubles(g, tr.tn1 - tr.tn0) * sizeof(double)));
    if (mult_aligned_ok(g, c->one_d)) {
        // views of the slab's own window (backed cell layers / node levels), not of the whole index space
        const int cl = s->lo_c, ch = s->hi_c, nl = s->lo_n, nh = s->hi_n;
        s->maps_ok = make_beta_map(g, s->beta[0], &s->tm_beta[0], cl, ch) == 0 && make_beta_map(g, s->beta[1], &s->tm_beta[1], cl, ch) == 0 &&
                     make_stag_maps(g, s->q[0], s->tm_q[0], cl, ch, nl, nh) == 0 && make_stag_maps(g, s->q[1], s->tm_q[1], cl, ch, nl, nh) == 0 &&
                     make_stag_maps(g, s->alpha, s->tm_alpha, cl, ch, nl, nh) == 0 &&
                     (!c->weighted || make_stag_maps(g, s->weight, s->tm_w, cl, ch, nl, nh) == 0);
        cudaGetLastError();
    }
    {
        const i64 nside = mult_side_doubles(g, c->one_d, tr.tc1 - s->lo_c);
        if (nside > 0) {
            cudaError_t e_ = cudaMalloc(&s->side, (size_t)nside * sizeof(double));
            if (e_ != cudaSuccess) { cudaGetLastError(); return set_err(DOTSOCP_ENOMEM, "side buffer of %lld doubles: %s", (long long)nside, cudaGetErrorString(e_)); }
        }
    }
    if (c->world > 1 && c->tpipe) {
        for (double*& p : s->carry) CU(cudaMalloc(&p, (size_t)g.P * sizeof(double)));
    } else if (c->world > 1) {
        CU(cudaMalloc(&s->tsend, (size_t)(tr.tn1 - tr.tn0) * g.P * sizeof(double)));
        CU(cudaMalloc(&s->trecv, (size_t)g.nt * (s->p1 - s->p0) * sizeof(double)));
    }
    if (const char* poison = getenv("DOTSOCP_POISON")) {
        // debugging aid: everything a kernel must write before anybody reads it starts as NaN, so that a read of
        // uninitialised memory shows up in the KKT rows of a small case instead of depending on what the allocator returned
        if (poison[0] == '1') {
            const double nan_ = std::numeric_limits<double>::quiet_NaN();
            for (auto& x : s->n_all) { launch_fill(s->phi + x.b, x.e - x.b, nan_, c->st); launch_fill(s->rhs + x.b, x.e - x.b, nan_, c->st); }
            if (s->side) launch_fill(s->side, mult_side_doubles(g, c->one_d, tr.tc1 - s->lo_c), nan_, c->st);
            launch_fill(s->partial, (i64)partial_doubles(g, tr.tn1 - tr.tn0), nan_, c->st);
            if (g.packed()) {
                double* qa[] = {s->q[0], s->q[1], s->alpha, s->q2, s->qtmp};
                for (double* a : qa) if (a) for (auto& x : s->q_all) launch_fill(a + x.b, x.e - x.b, nan_, c->st);
                for (int k = 0; k < 2; k++) for (auto& x : s->b_all) launch_fill(s->beta[k] + x.b, x.e - x.b, nan_, c->st);
            }
            CU(cudaGetLastError());
        }
    }
    return 0;
}

// Q doubles of scratch for s (BF)^* beta of the stand-alone KKT check (PALM / acc-ADMM, DOTSOCP_KKT=separate, or a check the
// schedule did not announce): 13 GB at 1024x1024x512 that the fused inPALM loop never touches, so it is not part of a session
// until somebody asks -- which is what lets the last level transfer of that grid fit one 180 GB GPU
static int ensure_qtmp(dotsocp_ctx* c, Slab* s)
{
    if (s->qtmp) return 0;
    std::vector<std::pair<long long, long long>> w;
    for (auto& x : s->q_all) w.emplace_back(x.b, x.e);
    const char* why = "";
    const int e_ = s->a_qtmp.create(c->g.Q, w, c->world == 1, c->device, &why);
    if (e_) return set_err(e_ == 2 ? DOTSOCP_ENOMEM : DOTSOCP_ECUDA, "KKT scratch of %lld doubles: %s", (long long)c->g.Q, why);
    s->qtmp = s->a_qtmp.ptr();
    if (!c->g.packed())
        for (auto& x : s->q_all) CU(cudaMemsetAsync(s->qtmp + x.b, 0, (size_t)(x.e - x.b) * sizeof(double), c->st));
    return 0;
}

// cuts: first cell layer of every slab (world + 1 entries, 0 .. nt-1, strictly increasing), or NULL for the even partition
static int create_impl(dotsocp_ctx** out, int variant, int nt, int nx, int ny, int rank, int world, const char* nccl_id,
                       const int* cuts);
static void release_cached_unlocked_if_free();
static int create_retry(dotsocp_ctx** out, int variant, int nt, int nx, int ny, int rank, int world, const char* nccl_id,
                        const int* cuts)
{
    int rc = create_impl(out, variant, nt, nx, ny, rank, world, nccl_id, cuts);
    if (rc == DOTSOCP_ENOMEM && !(world > 1 && nccl_id)) {   // the session cached by dotsocp_solve_level may hold the memory
        release_cached_unlocked_if_free();
        rc = create_impl(out, variant, nt, nx, ny, rank, world, nccl_id, cuts);
    }
    return rc;
}
extern "C" int dotsocp_create(dotsocp_ctx** out, int variant, int nt, int nx, int ny, int rank, int world, const char* nccl_id)
{
    return create_retry(out, variant, nt, nx, ny, rank, world, nccl_id, nullptr);
}
// Session of the next finer level (2n-1 nodes per refined axis) with the same variant, rank, world and communicator as
// `coarse` and the SAME partition in physical time: every cut of the coarse partition doubled, so that slab r of the fine
// grid covers exactly the cells of slab r of the coarse grid and dotsocp_prolong needs no data from another rank.
extern "C" int dotsocp_create_refined(dotsocp_ctx** out, const dotsocp_ctx* coarse)
{
    if (!out || !coarse) return set_err(DOTSOCP_EINVAL, "NULL argument");
    const Geo& gc = coarse->g;
    std::vector<int> cuts(coarse->world + 1);
    for (int r = 0; r < coarse->world; r++) cuts[r] = 2 * coarse->part[r].tc0;
    cuts[coarse->world] = 2 * (gc.nt - 1);
    static const char zero_id[128] = {0};
    const char* id = coarse->comm ? zero_id : nullptr;   // re-use the process-wide communicator / stay in emulation
    return create_retry(out, coarse->variant, 2 * gc.nt - 1, 2 * gc.nx - 1, gc.ny > 1 ? 2 * gc.ny - 1 : 1, coarse->my, coarse->world, id,
                        cuts.data());
}
static int create_impl(dotsocp_ctx** out, int variant, int nt, int nx, int ny, int rank, int world, const char* nccl_id,
                       const int* cuts)
{
    if (!out) return set_err(DOTSOCP_EINVAL, "ctx pointer is NULL");
    *out = nullptr;
    if (variant < 0 || variant > 2) return set_err(DOTSOCP_EINVAL, "unknown variant %d", variant);
    if (nt < 2 || nx < 2 || ny < 1) return set_err(DOTSOCP_EINVAL, "grid %d x %d x %d too small (need nt,nx >= 2)", nt, nx, ny);
    if (variant == DOTSOCP_VARIANT_DOT1D && ny != 1) return set_err(DOTSOCP_EINVAL, "1-D variant needs ny == 1");
    if (variant != DOTSOCP_VARIANT_DOT1D && ny < 2) return set_err(DOTSOCP_EINVAL, "2-D variants need ny >= 2");
    if (world < 1 || rank < 0 || rank >= world) return set_err(DOTSOCP_EINVAL, "bad rank %d / world %d", rank, world);
    if (world > nt - 1) return set_err(DOTSOCP_EINVAL, "world %d exceeds the %d cell layers", world, nt - 1);
    int rc = require_device();
    if (rc) return rc;
    dotsocp_ctx* c = new dotsocp_ctx();
    c->variant = variant;
    c->one_d = variant == DOTSOCP_VARIANT_DOT1D;
    c->weighted = variant == DOTSOCP_VARIANT_WDOT2D;
    c->world = world;
    c->emulate = world > 1 && nccl_id == nullptr;
    { const char* tr_ = getenv("DOTSOCP_TRACE"); c->trace = tr_ && tr_[0] == '1'; }
    c->my = c->emulate ? 0 : rank;
    {   // DOTSOCP_LAYOUT=packed keeps the reference's packed rows on the device as well (A/B measurements, tests)
        const char* lay = getenv("DOTSOCP_LAYOUT");
        c->gh = make_geo(nt, nx, ny);
        c->g = (lay && strcmp(lay, "packed") == 0) ? c->gh : make_geo_padded(nt, nx, ny);
    }
    const Geo& g = c->g;
    cudaGetDevice(&c->device);
    for (int r = 0; r < world; r++) {
        TRange tr;
        tr.tc0 = cuts ? cuts[r] : (int)((i64)r * (nt - 1) / world);
        tr.tc1 = cuts ? cuts[r + 1] : (int)((i64)(r + 1) * (nt - 1) / world);
        tr.tn0 = tr.tc0;
        tr.tn1 = (r == world - 1) ? nt : tr.tc1;
        if (tr.tc1 <= tr.tc0 || tr.tc0 < 0 || tr.tc1 > nt - 1 || (r == 0 && tr.tc0 != 0) || (r == world - 1 && tr.tc1 != nt - 1)) {
            delete c;
            return set_err(DOTSOCP_EINVAL, "bad time partition: slab %d = cell layers [%d, %d) of %d", r, tr.tc0, tr.tc1, nt - 1);
        }
        c->part.push_back(tr);
    }
    {   // mode chunks of ceil(P/world) (the last one shorter): cheap owner arithmetic inside the fused-pack DCT kernels
        const i64 C = (g.P + world - 1) / world;
        for (int r = 0; r <= world; r++) c->pcut.push_back(std::min(g.P, (i64)r * C));
    }
    {   // t-solve across slabs: pipelined Thomas sweeps (one carry plane per boundary and direction) unless DOTSOCP_TSOLVE asks
        // for the transposes ("transpose") or for the transform-based t pass ("dct", which needs whole lines on one slab)
        const char* ts = getenv("DOTSOCP_TSOLVE");
        c->tpipe = world > 1 && nt >= 3 && !(ts && (strcmp(ts, "transpose") == 0 || strcmp(ts, "dct") == 0));
        const char* tc = getenv("DOTSOCP_TCHUNKS");
        const int want = tc ? atoi(tc) : 0;
        c->tchunks = want > 0 ? want : (nccl_id ? 8 : 1);
    }
    // the communication stream gets the highest priority so that the NCCL copy kernels are scheduled as soon as SM slots
    // free up instead of queueing behind the (much larger) compute grids they are meant to overlap with
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if (cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking) != cudaSuccess ||
        (world > 1 && cudaStreamCreateWithPriority(&c->st2, cudaStreamNonBlocking, prio_hi) != cudaSuccess)) {
        dotsocp_destroy(c);
        return set_err(DOTSOCP_ECUDA, "cudaStreamCreate failed");
    }
    if (world > 1 && !c->emulate) {
        const NcclApi& n = nccl_api();
        if (!n.ok) { dotsocp_destroy(c); return set_err(DOTSOCP_ENCCL, "NCCL unavailable: %s", n.why); }
        bool zero_id = true;
        for (int i = 0; i < 128; i++) zero_id = zero_id && nccl_id[i] == 0;
        if (zero_id) {   // reuse the process-wide communicator of an earlier session (its peer connections are warm)
            if (!g_comm || g_comm_world != world || g_comm_rank != rank) {
                dotsocp_destroy(c);
                return set_err(DOTSOCP_ESTATE, "no reusable communicator for rank %d / world %d", rank, world);
            }
        } else {
            if (g_comm) { n.CommDestroy(g_comm); g_comm = nullptr; }
            NcclId id;
            memcpy(id.internal, nccl_id, 128);
            int r_ = n.CommInitRank(&g_comm, world, id, rank);
            if (r_ != 0) { g_comm = nullptr; dotsocp_destroy(c); return set_err(DOTSOCP_ENCCL, "ncclCommInitRank: %s", n.GetErrorString(r_)); }
            g_comm_world = world;
            g_comm_rank = rank;
        }
        c->comm = g_comm;
    }
    const auto t_c0 = std::chrono::steady_clock::now();
    auto since = [&](std::chrono::steady_clock::time_point a) { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - a).count(); };
    if (c->emulate) {
        for (int r = 0; r < world; r++)
            if ((rc = make_slab(c, r))) { dotsocp_destroy(c); return rc; }
    } else if ((rc = make_slab(c, rank))) {
        dotsocp_destroy(c);
        return rc;
    }
    if (c->comm && c->tpipe) {
        // Pipelined sweeps, one process per GPU: map the neighbours' carry-in planes and flags (CUDA IPC) so that the Thomas kernels
        // hand their boundary plane over themselves -- a peer store and a flag instead of an ncclSend / ncclRecv pair per chunk
        // and direction (2 x 8 x ~25 us per solve on 8 GPUs).  DOTSOCP_TPUSH=0, or any failure here, keeps the NCCL hand-off.
        const char* tp = getenv("DOTSOCP_TPUSH");
        Slab* s = c->slabs[0];
        const size_t xdoubles = (size_t)2 * g.P + 64;
        cudaIpcMemHandle_t mine;
        memset(&mine, 0, sizeof mine);
        bool ok = !(tp && tp[0] == '0') && cudaMalloc(&c->barrier_buf, 64) == cudaSuccess && cudaMalloc(&s->xbuf, xdoubles * sizeof(double)) == cudaSuccess &&
                  cudaMalloc(&s->tdone, 128 * sizeof(int)) == cudaSuccess;
        if (ok) {
            cudaMemset(c->barrier_buf, 0, 64);
            cudaMemset(s->xbuf, 0, xdoubles * sizeof(double));
            cudaMemset(s->tdone, 0, 128 * sizeof(int));
            cudaDeviceSynchronize();
            ok = cudaIpcGetMemHandle(&mine, s->xbuf) == cudaSuccess;
        }
        cudaGetLastError();
        // every rank takes part in the collectives even if its own export failed (flag byte), to stay in lock-step
        const size_t hb = sizeof(cudaIpcMemHandle_t), slot = hb + 8;
        std::vector<char> sendb(slot, 0), recvb(slot * world);
        memcpy(sendb.data(), &mine, hb);
        sendb[hb] = ok ? 1 : 0;
        char* dall = nullptr;
        bool all_ok = false;
        if (cudaMalloc(&dall, slot * (world + 1)) == cudaSuccess) {
            cudaMemcpyAsync(dall + slot * world, sendb.data(), slot, cudaMemcpyHostToDevice, c->st);
            int r_ = nccl_api().AllGather(dall + slot * world, dall, slot, NCCL_INT8, c->comm, c->st);
            cudaMemcpyAsync(recvb.data(), dall, slot * world, cudaMemcpyDeviceToHost, c->st);
            cudaStreamSynchronize(c->st);
            cudaFree(dall);
            all_ok = r_ == 0;
            for (int r = 0; r < world; r++) all_ok = all_ok && recvb[slot * r + hb] == 1;
            if (all_ok) {
                for (int nb = 0; nb < 2 && all_ok; nb++) {
                    const int r = nb == 0 ? rank + 1 : rank - 1;
                    if (r < 0 || r >= world) continue;
                    cudaIpcMemHandle_t hh;
                    memcpy(&hh, recvb.data() + slot * r, hb);
                    void* pp_ = nullptr;
                    if (cudaIpcOpenMemHandle(&pp_, hh, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { all_ok = false; cudaGetLastError(); }
                    (nb == 0 ? s->peer_up : s->peer_dn) = (double*)pp_;
                }
            }
            // agree collectively (a rank that failed to open a handle disables the path for everybody)
            if (c->barrier_buf) {
                double flag = all_ok ? 0.0 : 1.0;
                cudaMemcpyAsync(c->barrier_buf, &flag, sizeof(double), cudaMemcpyHostToDevice, c->st);
                nccl_api().AllReduce(c->barrier_buf, c->barrier_buf, 1, NCCL_FLOAT64, NCCL_SUM, c->comm, c->st);
                cudaMemcpyAsync(&flag, c->barrier_buf, sizeof(double), cudaMemcpyDeviceToHost, c->st);
                cudaStreamSynchronize(c->st);
                all_ok = all_ok && flag == 0.0;
            } else
                all_ok = false;
        }
        cudaGetLastError();
        c->tpush = all_ok;
        if (c->trace) fprintf(stderr, "[dotsocp trace] rank %d: in-kernel carry hand-off %s (xbuf %p, up %p, down %p)\n", rank, c->tpush ? "on" : "off",
                              (void*)s->xbuf, (void*)s->peer_up, (void*)s->peer_dn);
        if (c->tpush) {
            cudaFree(s->carry[0]); cudaFree(s->carry[2]);
            s->carry[0] = s->xbuf;
            s->carry[2] = s->xbuf + g.P;
            s->tflags = reinterpret_cast<int*>(s->xbuf + 2 * g.P);
        } else {
            if (s->peer_up) { cudaIpcCloseMemHandle(s->peer_up); s->peer_up = nullptr; }
            if (s->peer_dn) { cudaIpcCloseMemHandle(s->peer_dn); s->peer_dn = nullptr; }
            cudaFree(s->xbuf); s->xbuf = nullptr;
            cudaGetLastError();
        }
    }
    if (c->comm && !c->tpipe) {
        // exchange CUDA-IPC handles of the transpose buffers; any failure just keeps the NCCL send/recv path
        const char* noipc = getenv("DOTSOCP_NO_IPC");
        Slab* s = c->slabs[0];
        cudaIpcMemHandle_t mine[2];
        bool ok = !(noipc && noipc[0] == '1') && cudaMalloc(&c->barrier_buf, 64) == cudaSuccess &&
                  cudaIpcGetMemHandle(&mine[0], s->tsend) == cudaSuccess && cudaIpcGetMemHandle(&mine[1], s->trecv) == cudaSuccess;
        cudaGetLastError();
        if (c->barrier_buf) cudaMemset(c->barrier_buf, 0, 64);
        const size_t hb = sizeof(cudaIpcMemHandle_t) * 2;
        char* dall = nullptr;
        std::vector<char> hall(hb * world);
        // every rank takes part in the collective even if its own export failed (flag byte), to stay in lock-step
        std::vector<char> sendb(hb + 8, 0);
        if (ok) memcpy(sendb.data(), mine, hb);
        sendb[hb] = ok ? 1 : 0;
        std::vector<char> recvb((hb + 8) * world);
        if (cudaMalloc(&dall, (hb + 8) * (world + 1)) == cudaSuccess) {
            cudaMemcpyAsync(dall + (hb + 8) * world, sendb.data(), hb + 8, cudaMemcpyHostToDevice, c->st);
            int r_ = nccl_api().AllGather(dall + (hb + 8) * world, dall, hb + 8, NCCL_INT8, c->comm, c->st);
            cudaMemcpyAsync(recvb.data(), dall, (hb + 8) * world, cudaMemcpyDeviceToHost, c->st);
            cudaStreamSynchronize(c->st);
            cudaFree(dall);
            bool all_ok = r_ == 0;
            for (int r = 0; r < world; r++) all_ok = all_ok && recvb[(hb + 8) * r + hb] == 1;
            if (all_ok) {
                s->peer_tsend.assign(world, nullptr);
                s->peer_trecv.assign(world, nullptr);
                for (int r = 0; r < world && all_ok; r++) {
                    if (r == rank) continue;
                    cudaIpcMemHandle_t hh[2];
                    memcpy(hh, recvb.data() + (hb + 8) * r, hb);
                    void *p0 = nullptr, *p1 = nullptr;
                    if (cudaIpcOpenMemHandle(&p0, hh[0], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
                        cudaIpcOpenMemHandle(&p1, hh[1], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                        all_ok = false;
                        cudaGetLastError();
                    }
                    s->peer_tsend[r] = (double*)p0;
                    s->peer_trecv[r] = (double*)p1;
                }
            }
            // agree collectively (a rank that failed to open a handle disables the path for everybody)
            double flag = all_ok ? 0.0 : 1.0, *dflag = c->barrier_buf;
            if (dflag) {
                cudaMemcpyAsync(dflag, &flag, sizeof(double), cudaMemcpyHostToDevice, c->st);
                nccl_api().AllReduce(dflag, dflag, 1, NCCL_FLOAT64, NCCL_SUM, c->comm, c->st);
                cudaMemcpyAsync(&flag, dflag, sizeof(double), cudaMemcpyDeviceToHost, c->st);
                cudaStreamSynchronize(c->st);
                c->ipc = flag == 0.0;
            }
            if (c->ipc) {
                int plo = 0, phi_ = 0;
                cudaDeviceGetStreamPriorityRange(&plo, &phi_);
                c->cps.assign(world, nullptr);
                for (int r = 0; r < world; r++) cudaStreamCreateWithPriority(&c->cps[r], cudaStreamNonBlocking, phi_);
            }
        }
        cudaGetLastError();
    }
    if (world > 1 && !c->tpipe && (c->ipc || !c->comm)) {
        // DOTSOCP_XCHG=direct: the kernels store straight into the destination slab's buffers instead of the copy-engine
        // pushes.  Off by default: on 8 B200s the t-solve of every GPU then writes to the SAME peer at the same time (the owner
        // of the current time level), and the 32-byte runs of the x pass cost more than the copies they replace
        // (Poisson solve 8.99 ms against 6.59 ms at 1024x1024x512, profiles/README.md).
        const char* xm = getenv("DOTSOCP_XCHG");
        c->direct = xm && strcmp(xm, "direct") == 0;
        if (c->direct) {
            std::vector<int> tc(world + 1);
            for (int r = 0; r < world; r++) tc[r] = c->part[r].tn0;
            tc[world] = c->part[world - 1].tn1;
            cudaError_t de = cudaMalloc(&c->d_tcut, (world + 1) * sizeof(int));
            if (de == cudaSuccess) de = cudaMemcpy(c->d_tcut, tc.data(), (world + 1) * sizeof(int), cudaMemcpyHostToDevice);
            for (Slab* s : c->slabs) {
                std::vector<double*> fwd(world), bwd(world);
                for (int r = 0; r < world; r++) {
                    Slab* o = c->local(r);
                    double* o_trecv = o ? o->trecv : s->peer_trecv[r];
                    double* o_tsend = o ? o->tsend : s->peer_tsend[r];
                    fwd[r] = o_trecv + (i64)s->tr.tn0 * (c->pcut[r + 1] - c->pcut[r]);
                    bwd[r] = o_tsend + (i64)(c->part[r].tn1 - c->part[r].tn0) * c->pcut[s->id];
                }
                if (de == cudaSuccess) de = cudaMalloc(&s->d_fwd, world * sizeof(double*));
                if (de == cudaSuccess) de = cudaMalloc(&s->d_bwd, world * sizeof(double*));
                if (de == cudaSuccess) de = cudaMemcpy(s->d_fwd, fwd.data(), world * sizeof(double*), cudaMemcpyHostToDevice);
                if (de == cudaSuccess) de = cudaMemcpy(s->d_bwd, bwd.data(), world * sizeof(double*), cudaMemcpyHostToDevice);
            }
            if (de != cudaSuccess) {
                cudaGetLastError();
                dotsocp_destroy(c);
                return set_err(DOTSOCP_ECUDA, "direct-exchange tables: %s", cudaGetErrorString(de));
            }
        }
    }
    const double ms_slabs = since(t_c0);
    const auto t_c1 = std::chrono::steady_clock::now();
    c->pp = poisson_plan_create(nt, nx, ny);
    const double ms_plan = since(t_c1);
    const auto t_c2 = std::chrono::steady_clock::now();
    {
        // Nothing is left to be built lazily inside the first Poisson solve, and the session starts with an idle device.  With the
        // lazy construction (DOTSOCP_PREP=0) the first solve of a level occasionally returned a wrong singular-mode line when the
        // host reached the table upload while the stream was still busy with the prologue -- about every second 3-level solve of
        // 512x512x256 ended with hundreds of extra iterations on the last level (tools/diag_levels.py)
        const char* pe = getenv("DOTSOCP_PREP");
        if (!(pe && pe[0] == '0')) {
            for (Slab* s : c->slabs) {
                if (world == 1) rc = poisson_prepare(c->pp, 0, g.P, 0, 0, c->st);
                else if (c->tpipe) rc = poisson_prepare(c->pp, 0, 0, s->tr.tn0, s->tr.tn1, c->st);
                else rc = poisson_prepare(c->pp, s->p0, s->p1 - s->p0, 0, 0, c->st);
                if (rc) { dotsocp_destroy(c); return rc; }
            }
            cudaDeviceSynchronize();
        }
    }
    { const char* kk = getenv("DOTSOCP_KKT"); c->fuse_kkt = !(kk && strcmp(kk, "separate") == 0); }
    if (c->trace) fprintf(stderr, "[dotsocp trace] create %d x %d x %d: arrays %.1f ms, Poisson plan %.1f ms, tables + sync %.1f ms\n", nt, nx, ny,
                          ms_slabs, ms_plan, since(t_c2));
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMalloc(&c->d_lvl, (size_t)nt * KSL * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&c->d_tot, KSL * sizeof(double));
    if (e == cudaSuccess && c->tpipe) e = cudaMalloc(&c->d_line0, (size_t)nt * sizeof(double));
    if (e == cudaSuccess) e = cudaMallocHost(&c->h_tot, KSL * sizeof(double));
    if (e == cudaSuccess) e = cudaMallocHost(&c->h_elapsed, sizeof(double));
    if (e != cudaSuccess) {
        cudaGetLastError();
        dotsocp_destroy(c);
        return set_err(DOTSOCP_ECUDA, "plan / reduction buffers: %s", cudaGetErrorString(e));
    }
    *out = c;
    return DOTSOCP_OK;
}

extern "C" double dotsocp_launch_count(const dotsocp_ctx* c) { return c ? c->launches : 0.0; }

// ------------------------------------------------------------------------------------------------ slab communication
// move `count` doubles at element offset `off` of the array selected by `sel` from slab `from` to slab `to`
typedef double* (*ArrSel)(Slab*, int);
static double* sel_phi(Slab* s, int) { return s->phi; }
static double* sel_q(Slab* s, int k) { return s->q[k]; }
static double* sel_alpha(Slab* s, int) { return s->alpha; }
static double* sel_beta(Slab* s, int k) { return s->beta[k]; }
static double* sel_weight(Slab* s, int) { return s->weight; }

struct Xfer { ArrSel sel; int k; i64 off, count; int from, to; };

static int do_xfers(dotsocp_ctx* c, const std::vector<Xfer>& xs, cudaStream_t st)
{
    if (xs.empty()) return 0;
    const NcclApi& n = nccl_api();
    bool grouped = false;
    for (const Xfer& x : xs) {
        if (x.count <= 0) continue;
        Slab* a = c->local(x.from);
        Slab* b = c->local(x.to);
        if (a && b) {
            CU(cudaMemcpyAsync(x.sel(b, x.k) + x.off, x.sel(a, x.k) + x.off, (size_t)x.count * sizeof(double),
                               cudaMemcpyDeviceToDevice, st));
        } else if (a || b) {
            if (!grouped) { NC(n.GroupStart()); grouped = true; }
            if (a) NC(n.Send(x.sel(a, x.k) + x.off, (size_t)x.count, NCCL_FLOAT64, x.to, c->comm, st));
            else NC(n.Recv(x.sel(b, x.k) + x.off, (size_t)x.count, NCCL_FLOAT64, x.from, c->comm, st));
        }
    }
    if (grouped) NC(n.GroupEnd());
    return 0;
}

// ghost exchanges across every slab boundary (node level T between slab r and r+1)
enum { GH_PHI_UP = 1, GH_Q_UP = 2, GH_Q_DOWN = 4, GH_ALPHA0_DOWN = 8, GH_BETA_DOWN = 16, GH_W = 32, GH_PHI_DOWN = 64 };
static int ghosts(dotsocp_ctx* c, int what, int qk, int bk, cudaStream_t st = nullptr)
{
    if (c->world == 1) return 0;
    if (!st) st = c->st;
    const Geo& g = c->g;
    std::vector<Xfer> xs;
    for (int r = 0; r + 1 < c->world; r++) {
        const i64 T = c->part[r].tn1;   // == part[r+1].tn0
        if (what & GH_PHI_UP) xs.push_back({sel_phi, 0, T * g.P, g.P, r + 1, r});
        if (what & GH_PHI_DOWN) xs.push_back({sel_phi, 0, (T - 1) * g.P, g.P, r, r + 1});
        if (what & GH_Q_UP) {
            xs.push_back({sel_q, qk, g.L + T * g.PBX, g.PBX, r + 1, r});
            xs.push_back({sel_q, qk, g.L + g.NBX + T * g.PBY, g.PBY, r + 1, r});
        }
        if (what & GH_Q_DOWN) {
            xs.push_back({sel_q, qk, (T - 1) * g.PC, g.PC, r, r + 1});
            xs.push_back({sel_q, qk, g.L + (T - 1) * g.PBX, g.PBX, r, r + 1});
            xs.push_back({sel_q, qk, g.L + g.NBX + (T - 1) * g.PBY, g.PBY, r, r + 1});
        }
        if (what & GH_ALPHA0_DOWN) xs.push_back({sel_alpha, 0, (T - 1) * g.PC, g.PC, r, r + 1});
        if (what & GH_BETA_DOWN)
            for (int j = 0; j < 10; j++) xs.push_back({sel_beta, bk, j * g.L + (T - 1) * g.PC, g.PC, r, r + 1});
        if ((what & GH_W) && c->weighted) {
            xs.push_back({sel_weight, 0, g.L + T * g.PBX, g.PBX, r + 1, r});
            xs.push_back({sel_weight, 0, g.L + g.NBX + T * g.PBY, g.PBY, r + 1, r});
            xs.push_back({sel_weight, 0, (T - 1) * g.PC, g.PC, r, r + 1});
            xs.push_back({sel_weight, 0, g.L + (T - 1) * g.PBX, g.PBX, r, r + 1});
            xs.push_back({sel_weight, 0, g.L + g.NBX + (T - 1) * g.PBY, g.PBY, r, r + 1});
        }
    }
    return do_xfers(c, xs, st);
}

// Poisson solve: rhs -> phi.  One slab: 5 in-place passes.  Several slabs: (y,x) forward locally, transpose so that
// every slab holds all t for its chunk of (x,y) modes, t-pass, transpose back, (x,y) inverse.
static int solve_poisson(dotsocp_ctx* c, double D2)
{
    const Geo& g = c->g;
    if (c->world == 1) {
        Slab* s = c->slabs[0];
        return poisson_solve(c->pp, s->rhs, s->phi, D2, c->st, &c->launches);
    }
    const NcclApi& n = nccl_api();
    if (c->tpipe) {
        // Pipelined Thomas: (y,x) transforms of the own levels in place, forward elimination slab 0 -> W-1 and back substitution
        // W-1 -> 0 with one carry plane per boundary, (x,y) inverse transforms.  The sweeps are sequential over the slabs, so the
        // modes are cut into chunks: slab r works on chunk k while slab r+1 works on chunk k-1.  No transposes, no packed
        // buffers; bit-identical to the single-slab solve.
        int rc = 0;
        cudaStream_t st = c->st;
        for (Slab* s : c->slabs)
            if ((rc = poisson_xy(c->pp, s->rhs, s->phi, s->tr.tn0, s->tr.tn1 - s->tr.tn0, false, st, &c->launches))) return rc;
        // the singular mode (zero eigenvalue := 1, initialize_FFTkernel.m:15) needs its whole line: nt doubles, every rank
        // contributes its levels (all-reduce of disjoint rows) and solves the line redundantly
        CU(cudaMemsetAsync(c->d_line0, 0, (size_t)g.nt * sizeof(double), st));
        for (Slab* s : c->slabs) poisson_line0_gather(c->pp, s->phi, c->d_line0, s->tr.tn0, s->tr.tn1, st);
        if (c->comm) NC(n.AllReduce(c->d_line0, c->d_line0, (size_t)g.nt, NCCL_FLOAT64, NCCL_SUM, c->comm, st));
        poisson_line0_solve(c->pp, c->d_line0, D2, st);
        for (Slab* s : c->slabs) poisson_line0_scatter(c->pp, s->phi, c->d_line0, s->tr.tn0, s->tr.tn1, st);
        c->launches += 1 + 2.0 * c->slabs.size();
        const int K = (int)std::max<i64>(1, std::min<i64>(std::min(c->tchunks, 64), g.P / 1024 + 1));
        if (c->tpush) c->tepoch++;
        for (int dir = 0; dir < 2; dir++) {            // 0: forward elimination (ascending slabs), 1: back substitution (descending)
            const bool bwd = dir == 1;
            for (int k = 0; k < K; k++) {
                const i64 m0 = g.P * k / K, m1 = g.P * (k + 1) / K;
                const int ns = (int)c->slabs.size();
                for (int i = 0; i < ns; i++) {
                    Slab* s = c->slabs[bwd ? ns - 1 - i : i];
                    const int from = bwd ? s->id + 1 : s->id - 1, to = bwd ? s->id - 1 : s->id + 1;
                    const bool has_in = from >= 0 && from < c->world, has_out = to >= 0 && to < c->world;
                    if (c->tpush) {
                        // hand-off inside the kernel: wait for the neighbour's flag of this chunk, store the carry plane into the
                        // next slab's buffer, publish the epoch there (flags: [0, 64) forward chunks, [64, 128) backward chunks)
                        SlabSync sy{nullptr, c->tepoch, s->tdone + dir * 64 + k, nullptr, c->tepoch};
                        if (has_in) sy.wait_flag = s->tflags + dir * 64 + k;
                        double* peer = bwd ? s->peer_dn : s->peer_up;      // xbuf of the slab the carry goes to
                        double* out = s->carry[bwd ? 3 : 1];
                        if (has_out) {
                            out = bwd ? peer + g.P : peer;                 // its backward-in / forward-in plane
                            sy.signal_flag = reinterpret_cast<int*>(peer + 2 * g.P) + dir * 64 + k;
                        }
                        if ((rc = poisson_thomas_slab(c->pp, s->phi, s->tr.tn0, s->tr.tn1, m0, m1, D2, bwd, has_in ? s->carry[bwd ? 2 : 0] : nullptr,
                                                      out, st, &c->launches, &sy))) return rc;
                        continue;
                    }
                    double* cin = s->carry[bwd ? 2 : 0];
                    double* cout = s->carry[bwd ? 3 : 1];
                    const double* in_ptr = nullptr;
                    if (has_in) {
                        Slab* o = c->local(from);
                        if (o) in_ptr = o->carry[bwd ? 3 : 1];          // same process: read the neighbour's carry-out directly
                        else { NC(n.Recv(cin + m0, (size_t)(m1 - m0), NCCL_FLOAT64, from, c->comm, st)); in_ptr = cin; }
                    }
                    if ((rc = poisson_thomas_slab(c->pp, s->phi, s->tr.tn0, s->tr.tn1, m0, m1, D2, bwd, in_ptr, cout, st, &c->launches))) return rc;
                    if (has_out && !c->local(to)) NC(n.Send(cout + m0, (size_t)(m1 - m0), NCCL_FLOAT64, to, c->comm, st));
                }
            }
        }
        for (Slab* s : c->slabs)
            if ((rc = poisson_xy(c->pp, s->phi, s->phi, s->tr.tn0, s->tr.tn1 - s->tr.tn0, true, st, &c->launches))) return rc;
        return ghosts(c, GH_PHI_UP, 0, 0);
    }
    const bool fused_pack = poisson_can_pack(c->pp);
    // rows [r0, r1) (local level indices of the slab that owns the rows) of the transposed exchange, both directions
    auto exchange_rows = [&](bool forward, int grp, int ngrp, cudaStream_t st) -> int {
        bool grouped = false, forked = false;
        cudaEvent_t e_fork = nullptr;
        for (int a = 0; a < c->world; a++)          // a: owner of the time rows
            for (int b = 0; b < c->world; b++) {    // b: owner of the mode chunk
                Slab* sa = c->local(a);
                Slab* sb = c->local(b);
                if (!sa && !sb) continue;
                const int nlev = c->part[a].tn1 - c->part[a].tn0;
                const int r0 = (int)((i64)grp * nlev / ngrp), r1 = (int)((i64)(grp + 1) * nlev / ngrp);
                if (r1 <= r0) continue;
                const i64 ch = c->pcut[b + 1] - c->pcut[b];
                const size_t cnt = (size_t)(r1 - r0) * ch;
                double* pa = sa ? sa->tsend + (i64)nlev * c->pcut[b] + (i64)r0 * ch : nullptr;
                double* pb = sb ? sb->trecv + (i64)(c->part[a].tn0 + r0) * ch : nullptr;
                if (sa && sb) {
                    CU(cudaMemcpyAsync(forward ? pb : pa, forward ? pa : pb, cnt * sizeof(double), cudaMemcpyDeviceToDevice, st));
                } else if (c->ipc) {
                    // push model over CUDA IPC: the owner of the source writes straight into the peer's buffer (copy engines);
                    // each peer has its own stream, forked from / joined to `st` by events
                    const int peer = forward ? b : a;
                    if ((forward && sa) || (!forward && sb)) {
                        cudaStream_t cs = c->cps[peer];
                        if (!forked) { e_fork = c->comm_event(); CU(cudaEventRecord(e_fork, st)); forked = true; }
                        CU(cudaStreamWaitEvent(cs, e_fork, 0));
                        if (forward) {
                            double* dst = sa->peer_trecv[b] + (i64)(c->part[a].tn0 + r0) * ch;
                            CU(cudaMemcpyAsync(dst, pa, cnt * sizeof(double), cudaMemcpyDeviceToDevice, cs));
                        } else {
                            double* dst = sb->peer_tsend[a] + (i64)nlev * c->pcut[b] + (i64)r0 * ch;
                            CU(cudaMemcpyAsync(dst, pb, cnt * sizeof(double), cudaMemcpyDeviceToDevice, cs));
                        }
                        cudaEvent_t ej = c->comm_event();
                        CU(cudaEventRecord(ej, cs));
                        CU(cudaStreamWaitEvent(st, ej, 0));
                    }
                } else {
                    if (!grouped) { NC(n.GroupStart()); grouped = true; }
                    if (forward) {
                        if (sa) NC(n.Send(pa, cnt, NCCL_FLOAT64, b, c->comm, st));
                        else NC(n.Recv(pb, cnt, NCCL_FLOAT64, a, c->comm, st));
                    } else {
                        if (sb) NC(n.Send(pb, cnt, NCCL_FLOAT64, a, c->comm, st));
                        else NC(n.Recv(pa, cnt, NCCL_FLOAT64, b, c->comm, st));
                    }
                }
            }
        if (grouped) NC(n.GroupEnd());
        // pushes are complete everywhere once every rank has passed this point of its stream
        if (c->ipc) NC(n.AllReduce(c->barrier_buf, c->barrier_buf, 1, NCCL_FLOAT64, NCCL_SUM, c->comm, st));
        return 0;
    };
    int rc = 0;
    if (fused_pack) {
        // pipelined in groups of time levels (8 by default, DOTSOCP_NGRP; 4 -> 8 groups: 20.6 -> 20.2 ms per iteration on
        // 4 GPUs): the transforms of group i overlap the all-to-all of group i-1 (forward), the
        // all-to-all of group i+1 overlaps the inverse transforms of group i (backward); rows of a level group are contiguous
        // in the packed buffers, so no data layout changes.  st = compute stream, st2 = communication stream.
        int minlev = g.nt;
        for (auto& tr : c->part) minlev = std::min(minlev, tr.tn1 - tr.tn0);
        static const int ngrp_want = [] { const char* e = getenv("DOTSOCP_NGRP"); const int v = e ? atoi(e) : 0; return v > 0 ? v : 8; }();
        const int ngrp = std::max(1, std::min(ngrp_want, minlev));
        c->cev_used = 0;
        auto tmark = [&](int i) {
            if (!c->trace) return;
            if (c->tev.size() < 6) { c->tev.resize(6); for (auto& e : c->tev) cudaEventCreate(&e); }
            cudaEventRecord(c->tev[i], c->st);
        };
        if (c->trace && c->tev.size() == 6) {   // fold the previous solve's timings (its events have completed by now or will block briefly)
            cudaEventSynchronize(c->tev[5]);
            if (++c->tseen > 3) {   // skip the first solves (NCCL connection set-up, table creation)
                for (int i = 0; i < 5; i++) { float ms = 0; cudaEventElapsedTime(&ms, c->tev[i], c->tev[i + 1]); c->tacc[i] += ms; }
                c->tcount++;
            }
        }
        if (c->direct) {
            // the kernels do the exchange: the forward x pass stores every result in the t-solve buffer of the owner of its
            // mode chunk, the t-solve stores every level in the packed buffer of the owner of the level (peer memory over
            // NVLink, plain local memory when one process emulates the slabs); the only communication calls left are the
            // two barriers that order "everybody has written" before "anybody reads".
            auto barrier = [&]() -> int {
                if (c->comm) NC(n.AllReduce(c->barrier_buf, c->barrier_buf, 1, NCCL_FLOAT64, NCCL_SUM, c->comm, c->st));
                return 0;
            };
            tmark(0);
            for (int i = 0; i < ngrp; i++)
                for (Slab* s : c->slabs) {
                    const int nlev = s->tr.tn1 - s->tr.tn0;
                    const int r0 = (int)((i64)i * nlev / ngrp), r1 = (int)((i64)(i + 1) * nlev / ngrp);
                    if (r1 > r0)
                        if ((rc = poisson_xy(c->pp, s->rhs, s->phi, s->tr.tn0 + r0, r1 - r0, false, c->st, &c->launches, s->tsend, c->world, nlev, r0, s->d_fwd))) return rc;
                }
            tmark(1);
            if ((rc = barrier())) return rc;
            tmark(2);
            for (Slab* s : c->slabs)
                if ((rc = poisson_t_chunk(c->pp, s->trecv, s->p1 - s->p0, s->p0, D2, c->st, &c->launches, s->d_bwd, c->d_tcut, c->world))) return rc;
            tmark(3);
            if ((rc = barrier())) return rc;
            for (int i = 0; i < ngrp; i++)
                for (Slab* s : c->slabs) {
                    const int nlev = s->tr.tn1 - s->tr.tn0;
                    const int r0 = (int)((i64)i * nlev / ngrp), r1 = (int)((i64)(i + 1) * nlev / ngrp);
                    if (r1 > r0)
                        if ((rc = poisson_xy(c->pp, s->phi, s->phi, s->tr.tn0 + r0, r1 - r0, true, c->st, &c->launches, s->tsend, c->world, nlev, r0))) return rc;
                }
            tmark(4);
            rc = ghosts(c, GH_PHI_UP, 0, 0);
            tmark(5);
            return rc;
        }
        tmark(0);
        for (int i = 0; i < ngrp; i++) {
            for (Slab* s : c->slabs) {
                const int nlev = s->tr.tn1 - s->tr.tn0;
                const int r0 = (int)((i64)i * nlev / ngrp), r1 = (int)((i64)(i + 1) * nlev / ngrp);
                if (r1 > r0)
                    if ((rc = poisson_xy(c->pp, s->rhs, s->phi, s->tr.tn0 + r0, r1 - r0, false, c->st, &c->launches, s->tsend, c->world, nlev, r0))) return rc;
            }
            cudaEvent_t e = c->comm_event();
            CU(cudaEventRecord(e, c->st));
            CU(cudaStreamWaitEvent(c->st2, e, 0));
            if ((rc = exchange_rows(true, i, ngrp, c->st2))) return rc;
        }
        tmark(1);
        cudaEvent_t e1 = c->comm_event();
        CU(cudaEventRecord(e1, c->st2));
        CU(cudaStreamWaitEvent(c->st, e1, 0));
        tmark(2);
        for (Slab* s : c->slabs) if ((rc = poisson_t_chunk(c->pp, s->trecv, s->p1 - s->p0, s->p0, D2, c->st, &c->launches))) return rc;
        tmark(3);
        cudaEvent_t e2 = c->comm_event();
        CU(cudaEventRecord(e2, c->st));
        CU(cudaStreamWaitEvent(c->st2, e2, 0));
        for (int i = 0; i < ngrp; i++) {
            if ((rc = exchange_rows(false, i, ngrp, c->st2))) return rc;
            cudaEvent_t e = c->comm_event();
            CU(cudaEventRecord(e, c->st2));
            CU(cudaStreamWaitEvent(c->st, e, 0));
            for (Slab* s : c->slabs) {
                const int nlev = s->tr.tn1 - s->tr.tn0;
                const int r0 = (int)((i64)i * nlev / ngrp), r1 = (int)((i64)(i + 1) * nlev / ngrp);
                if (r1 > r0)
                    if ((rc = poisson_xy(c->pp, s->phi, s->phi, s->tr.tn0 + r0, r1 - r0, true, c->st, &c->launches, s->tsend, c->world, nlev, r0))) return rc;
            }
        }
        tmark(4);
        rc = ghosts(c, GH_PHI_UP, 0, 0);
        tmark(5);
        return rc;
    }
    for (Slab* s : c->slabs) {
        const int nlev = s->tr.tn1 - s->tr.tn0;
        if ((rc = poisson_xy(c->pp, s->rhs, s->phi, s->tr.tn0, nlev, false, c->st, &c->launches))) return rc;
        // pack: block r of the send buffer = rows of this slab x modes of slab r
        for (int r = 0; r < c->world; r++) {
            const i64 ch = c->pcut[r + 1] - c->pcut[r];
            CU(cudaMemcpy2DAsync(s->tsend + (i64)nlev * c->pcut[r], ch * sizeof(double), s->phi + s->tr.tn0 * g.P + c->pcut[r],
                                 g.P * sizeof(double), ch * sizeof(double), nlev, cudaMemcpyDeviceToDevice, c->st));
        }
    }
    if ((rc = exchange_rows(true, 0, 1, c->st))) return rc;
    for (Slab* s : c->slabs) if ((rc = poisson_t_chunk(c->pp, s->trecv, s->p1 - s->p0, s->p0, D2, c->st, &c->launches))) return rc;
    if ((rc = exchange_rows(false, 0, 1, c->st))) return rc;
    for (Slab* s : c->slabs) {
        const int nlev = s->tr.tn1 - s->tr.tn0;
        for (int r = 0; r < c->world; r++) {
            const i64 ch = c->pcut[r + 1] - c->pcut[r];
            CU(cudaMemcpy2DAsync(s->phi + s->tr.tn0 * g.P + c->pcut[r], g.P * sizeof(double), s->tsend + (i64)nlev * c->pcut[r],
                                 ch * sizeof(double), ch * sizeof(double), nlev, cudaMemcpyDeviceToDevice, c->st));
        }
        if ((rc = poisson_xy(c->pp, s->phi, s->phi, s->tr.tn0, nlev, true, c->st, &c->launches))) return rc;
    }
    return ghosts(c, GH_PHI_UP, 0, 0);
}

// ------------------------------------------------------------------------------------------------ upload / download
// Host layout: world == 1 or emulation -> GLOBAL arrays; NCCL mode -> the slab's owned part only:
//   phi, c : owned node levels ; q, alpha, weight : [q0 owned cells | bx owned levels | by owned levels] ;
//   z, beta: ncol columns of (owned cells) doubles, column-major.
struct HostMap {
    bool local;
    const Geo* g;     // the PACKED geometry (ctx->gh): host arrays always have the reference's layout
    TRange tr;
    i64 nodes(i64 t) const { return (local ? t - tr.tn0 : t) * g->P; }
    i64 q0(i64 t) const { return (local ? t - tr.tc0 : t) * g->P; }
    i64 bx(i64 t) const { return local ? (i64)(tr.tc1 - tr.tc0) * g->P + (t - tr.tn0) * g->PBX : g->L + t * g->PBX; }
    i64 by(i64 t) const
    {
        return local ? (i64)(tr.tc1 - tr.tc0) * g->P + (i64)(tr.tn1 - tr.tn0) * g->PBX + (t - tr.tn0) * g->PBY : g->L + g->NBX + t * g->PBY;
    }
    i64 col(int j, i64 t) const { return local ? (i64)j * (tr.tc1 - tr.tc0) * g->P + (t - tr.tc0) * g->P : (i64)j * g->L + t * g->P; }
};

// host <-> device transfer of one contiguous piece: staged through the pinned ring (hostcopy.h) unless the piece is small or
// DOTSOCP_HOSTCOPY=plain asks for the driver's own pageable path (A/B measurements)
static int xfer(dotsocp_ctx* c, double* dev, double* host, size_t n, bool up)
{
    static const bool plain = [] { const char* e = getenv("DOTSOCP_HOSTCOPY"); return e && strcmp(e, "plain") == 0; }();
    const size_t bytes = n * sizeof(double);
    if (!plain && bytes >= ((size_t)1 << 20)) {
        HostCopier* hc = HostCopier::get();
        if (hc->ok()) {
            const int e = up ? hc->h2d(dev, host, bytes, c->st) : hc->d2h(host, dev, bytes, c->st);
            if (e) return set_err(DOTSOCP_ECUDA, "staged %s copy failed: %s", up ? "host-to-device" : "device-to-host", cudaGetErrorString((cudaError_t)e));
            return 0;
        }
    }
    if (up) CU(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, c->st));
    else CU(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, c->st));
    return 0;
}
// `rows` rows of `w` doubles: packed on the host, `dpitch` doubles apart on the device
static int xfer_rows(dotsocp_ctx* c, double* dev, i64 dpitch, double* host, i64 w, i64 rows, bool up, double pad_value = 0.0)
{
    if (rows <= 0 || w <= 0) return 0;
    if (dpitch == w) return xfer(c, dev, host, (size_t)(w * rows), up);
    static const bool plain = [] { const char* e = getenv("DOTSOCP_HOSTCOPY"); return e && strcmp(e, "plain") == 0; }();
    const size_t wb = (size_t)w * sizeof(double), db = (size_t)dpitch * sizeof(double);
    if (wb * (size_t)rows < ((size_t)1 << 20)) {
        // small pieces (coarse multilevel grids): a 2-D copy from pageable memory is driven row by row (milliseconds per array);
        // re-pitch on the host instead and move ONE contiguous block (the pads travel with it and keep their value)
        std::vector<double> tmp((size_t)rows * dpitch, pad_value);
        if (up) {
            for (i64 r = 0; r < rows; r++) memcpy(tmp.data() + r * dpitch, host + r * w, wb);
            CU(cudaMemcpyAsync(dev, tmp.data(), tmp.size() * sizeof(double), cudaMemcpyHostToDevice, c->st));   // returns once tmp is staged
        } else {
            CU(cudaMemcpyAsync(tmp.data(), dev, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost, c->st));
            CU(cudaStreamSynchronize(c->st));
            for (i64 r = 0; r < rows; r++) memcpy(host + r * w, tmp.data() + r * dpitch, wb);
        }
        return 0;
    }
    if (!plain && wb * (size_t)rows >= ((size_t)1 << 20)) {
        HostCopier* hc = HostCopier::get();
        if (hc->ok()) {
            const int e = up ? hc->h2d_rows(dev, db, host, wb, (size_t)rows, c->st) : hc->d2h_rows(host, dev, db, wb, (size_t)rows, c->st);
            if (e) return set_err(DOTSOCP_ECUDA, "staged %s copy failed: %s", up ? "host-to-device" : "device-to-host", cudaGetErrorString((cudaError_t)e));
            return 0;
        }
    }
    if (up) CU(cudaMemcpy2DAsync(dev, db, host, wb, wb, (size_t)rows, cudaMemcpyHostToDevice, c->st));
    else CU(cudaMemcpy2DAsync(host, wb, dev, db, wb, (size_t)rows, cudaMemcpyDeviceToHost, c->st));
    return 0;
}

static int copy_stag(dotsocp_ctx* c, Slab* s, const HostMap& hm, double* dev, double* host, bool up, double pad_value = 0.0)
{
    const Geo& g = c->g;
    const TRange& tr = s->tr;
    // q0 rows of the owned cell layers, bx / by rows of the owned node levels
    struct P { i64 d, h, rows, w, dp; } parts[3] = {
        {tr.tc0 * g.PC, hm.q0(tr.tc0), (i64)(tr.tc1 - tr.tc0) * g.nx, g.ny, g.py},
        {g.L + tr.tn0 * g.PBX, hm.bx(tr.tn0), (i64)(tr.tn1 - tr.tn0) * (g.nx - 1), g.ny, g.py},
        {g.L + g.NBX + tr.tn0 * g.PBY, hm.by(tr.tn0), (i64)(tr.tn1 - tr.tn0) * g.nx, g.ny - 1, g.pyb}};
    for (auto& p : parts) {
        int rc = xfer_rows(c, dev + p.d, p.dp, host + p.h, p.w, p.rows, up, pad_value);
        if (rc) return rc;
    }
    return 0;
}

static int copy_cols(dotsocp_ctx* c, Slab* s, const HostMap& hm, double* dev10, double* host, bool up)
{
    const Geo& g = c->g;
    const TRange& tr = s->tr;
    const int ncol = c->one_d ? 6 : 10;
    // 1-D variant: 6 columns at the boundary (c0..c4 -> 0..4, c5 -> 9; columns 5..8 are structural zeros on the device)
    for (int jh = 0; jh < ncol; jh++) {
        const int j = (c->one_d && jh == 5) ? 9 : jh;
        int rc = xfer_rows(c, dev10 + j * g.L + tr.tc0 * g.PC, g.py, host + hm.col(jh, tr.tc0), g.ny, (i64)(tr.tc1 - tr.tc0) * g.nx, up);
        if (rc) return rc;
    }
    if (up && c->one_d)
        for (int j = 5; j < 9; j++)
            CU(cudaMemsetAsync(dev10 + j * g.L + s->lo_c * g.PC, 0, (size_t)(s->hi_c - s->lo_c) * g.PC * sizeof(double), c->st));
    return 0;
}

// pad entries of the weight are ones: the level transfer divides whole ranges by the weight (k_prolong_alpha)
static int fill_weight_pads(dotsocp_ctx* c, Slab* s)
{
    if (c->g.packed() || !s->weight) return 0;
    for (auto& x : s->q_all) launch_fill(s->weight + x.b, x.e - x.b, 1.0, c->st);
    CU(cudaGetLastError());
    return 0;
}

// The session's weight from level `level` of a device pyramid (weights.cu) instead of a host array: every window the slab
// backs (owned levels and ghost layers alike, so no exchange follows) is filled by a device-to-device scatter from the packed
// level array into the pitched layout.  Call it before dotsocp_upload / dotsocp_prolong, which then take weight == NULL.
extern "C" int dotsocp_set_weight(dotsocp_ctx* c, const dotsocp_weights* w, int level)
{
    if (!c || !w) return set_err(DOTSOCP_EINVAL, "NULL argument");
    if (!c->weighted) return set_err(DOTSOCP_EINVAL, "set_weight: not a weighted (WDOT2D) session");
    if (level < 0 || level >= (int)w->lv.size()) return set_err(DOTSOCP_EINVAL, "set_weight: level %d of %d", level, (int)w->lv.size());
    if (level >= w->filled) return set_err(DOTSOCP_ESTATE, "set_weight: level %d of the pyramid has not been computed", level);
    const WeightLevel& l = w->lv[level];
    const Geo& g = c->g;
    if (l.nt != g.nt || l.nx != g.nx || l.ny != g.ny)
        return set_err(DOTSOCP_EINVAL, "set_weight: level %d is %d x %d x %d, the session %d x %d x %d", level, l.nt, l.nx, l.ny, g.nt, g.nx, g.ny);
    if (w->device != c->device) return set_err(DOTSOCP_EINVAL, "set_weight: the pyramid lives on device %d, the session on %d", w->device, c->device);
    for (Slab* s : c->slabs) {
        for (auto& x : s->q_all) launch_weight_scatter(g, x.b, x.e, l.w, s->weight, c->st);
        CU(cudaGetLastError());
        c->launches += (double)s->q_all.size();
    }
    CU(cudaStreamSynchronize(c->st));
    c->weight_set = true;
    return DOTSOCP_OK;
}

extern "C" int dotsocp_upload(dotsocp_ctx* c, const double* phi, const double* q, const double* z, const double* alpha,
                              const double* beta, const double* cvec, const double* weight)
{
    if (!c || !phi || !q || !alpha || !beta || !cvec) return set_err(DOTSOCP_EINVAL, "NULL array");
    if (c->weighted && !weight && !c->weight_set)
        return set_err(DOTSOCP_EINVAL, "weighted variant needs weight (a host array here, or dotsocp_set_weight before)");
    const Geo& g = c->g;
    c->qcur = 0;
    c->bcur = 0;
    for (Slab* s : c->slabs) {
        HostMap hm{c->world > 1 && !c->emulate, &c->gh, s->tr};
        const TRange& tr = s->tr;
        // c is -rho0/ht on the first time level, +rho1/ht on the last and zero in between (initialize.m:41-44); only the
        // two planes are kept on the device.
        {
            std::atomic<long long> bad(-1);
            const int nrows = (int)(tr.tn1 - tr.tn0);
            HostCopier::get()->pool().parallel_for(nrows, [&](int k) {
                const i64 t = tr.tn0 + k;
                if (t == 0 || t == g.nt - 1 || bad.load(std::memory_order_relaxed) >= 0) return;
                const double* row = cvec + hm.nodes(t);
                for (i64 i = 0; i < g.P; i++)
                    if (row[i] != 0.0) { bad.store((long long)t); return; }
            });
            if (bad.load() >= 0) return set_err(DOTSOCP_EINVAL, "model.c has a non-zero interior entry (t=%lld): unsupported", bad.load());
        }
        int rc;
        if ((rc = xfer(c, s->phi + tr.tn0 * g.P, const_cast<double*>(phi) + hm.nodes(tr.tn0), (size_t)(tr.tn1 - tr.tn0) * g.P, true))) return rc;
        if ((rc = copy_stag(c, s, hm, s->q[0], const_cast<double*>(q), true))) return rc;
        if ((rc = copy_stag(c, s, hm, s->alpha, const_cast<double*>(alpha), true))) return rc;
        if (c->weighted && weight) {
            if ((rc = fill_weight_pads(c, s))) return rc;
            if ((rc = copy_stag(c, s, hm, s->weight, const_cast<double*>(weight), true, 1.0))) return rc;
        }
        if (tr.tn0 == 0) CU(cudaMemcpyAsync(s->c0, cvec + hm.nodes(0), g.P * sizeof(double), cudaMemcpyHostToDevice, c->st));
        if (tr.tn1 == g.nt) CU(cudaMemcpyAsync(s->c1, cvec + hm.nodes(g.nt - 1), g.P * sizeof(double), cudaMemcpyHostToDevice, c->st));
        if ((rc = copy_cols(c, s, hm, s->beta[0], const_cast<double*>(beta), true))) return rc;
        if (z && (rc = copy_cols(c, s, hm, s->beta[1], const_cast<double*>(z), true))) return rc;
    }
    c->z_materialised = true;
    c->z_absent = (z == nullptr);
    int rc = ghosts(c, GH_PHI_UP | GH_PHI_DOWN | GH_Q_UP | GH_Q_DOWN | GH_ALPHA0_DOWN | GH_BETA_DOWN | GH_W, 0, 0);
    if (rc) return rc;
    CU(cudaStreamSynchronize(c->st));
    c->uploaded = true;
    return DOTSOCP_OK;
}

// Level transfer on the device: recoverOrgVar + interpolate + jump_nextLevel + InitialScaling of
// solver_dotsocp2d.m:230-250 without the state ever leaving HBM.  `coarse` holds the output state of a finished level
// (after dotsocp_run), `fine` is a fresh session of the refined grid (2n-1 nodes per refined axis); afterwards `fine` is
// in the state dotsocp_upload would have left it in (z = 0).  Time slabs: `fine` must come from dotsocp_create_refined
// (aligned partition), every slab fills its own part from the coarse slab with the same index and the ghost planes are
// exchanged exactly as after an upload; host arrays follow the upload convention (global, or the slab's part in NCCL mode).
extern "C" int dotsocp_prolong(dotsocp_ctx* coarse, dotsocp_ctx* fine, const dotsocp_prolong_scal* ps, const double* c_first,
                               const double* c_last, const double* weight)
{
    if (!coarse || !fine || !ps) return set_err(DOTSOCP_EINVAL, "NULL argument");
    if (coarse->variant != fine->variant) return set_err(DOTSOCP_EINVAL, "prolong: variants differ");
    if (!coarse->uploaded || !coarse->z_materialised || coarse->iter_open)
        return set_err(DOTSOCP_ESTATE, "prolong: the coarse session must hold the output state of a finished run");
    const Geo &gc = coarse->g, &gf = fine->g;
    if (gf.nt != 2 * gc.nt - 1 || gf.nx != 2 * gc.nx - 1 || gf.ny != (gc.ny > 1 ? 2 * gc.ny - 1 : 1))
        return set_err(DOTSOCP_EINVAL, "prolong: the fine grid must have 2n-1 nodes per axis (%d,%d,%d) -> (%d,%d,%d)", gc.nt, gc.nx,
                       gc.ny, gf.nt, gf.nx, gf.ny);
    if (fine->weighted && !weight && !fine->weight_set)
        return set_err(DOTSOCP_EINVAL, "weighted variant needs weight (a host array here, or dotsocp_set_weight on the fine session before)");
    if (coarse->world != fine->world || coarse->emulate != fine->emulate || coarse->comm != fine->comm || coarse->my != fine->my)
        return set_err(DOTSOCP_EINVAL, "prolong: the two sessions must share rank, world and communicator (dotsocp_create_refined)");
    for (int r = 0; r < fine->world; r++)
        if (fine->part[r].tc0 != 2 * coarse->part[r].tc0 || fine->part[r].tc1 != 2 * coarse->part[r].tc1)
            return set_err(DOTSOCP_EINVAL, "prolong: slab %d of the fine session is not the refined coarse slab (dotsocp_create_refined)", r);
    fine->qcur = 0;
    fine->bcur = 0;
    int rc;
    CU(cudaStreamSynchronize(coarse->st));   // the coarse state is final
    const ProlongScal k{ps->phi_recover, ps->beta_recover, ps->grad_t, ps->grad_x, ps->grad_y,
                        ps->phi_scale, ps->q_scale, ps->alpha_scale, ps->beta_scale};
    cudaStream_t st = fine->st;
    const bool local_host = fine->world > 1 && !fine->emulate;
    // stage A (per slab): weight and c from the host, unscaled fine phi on the owned node levels
    for (Slab* sf : fine->slabs) {
        Slab* sc = coarse->local(sf->id);
        const TRange& tr = sf->tr;
        if (fine->weighted && weight) {
            HostMap hm{local_host, &fine->gh, tr};
            if ((rc = fill_weight_pads(fine, sf))) return rc;
            if ((rc = copy_stag(fine, sf, hm, sf->weight, const_cast<double*>(weight), true, 1.0))) return rc;
        }
        if (tr.tn0 == 0) {
            if (!c_first) return set_err(DOTSOCP_EINVAL, "prolong: c_first is required on the slab that owns the first time level");
            CU(cudaMemcpyAsync(sf->c0, c_first, gf.P * sizeof(double), cudaMemcpyHostToDevice, st));
        }
        if (tr.tn1 == gf.nt) {
            if (!c_last) return set_err(DOTSOCP_EINVAL, "prolong: c_last is required on the slab that owns the last time level");
            CU(cudaMemcpyAsync(sf->c1, c_last, gf.P * sizeof(double), cudaMemcpyHostToDevice, st));
        }
        launch_prolong_phi(gc, gf, k.phi_recover, sc->phi, sf->phi, tr.tn0, tr.tn1, st);
        fine->launches += 1;
    }
    // q0 = Dt phi needs the level above the slab
    if ((rc = ghosts(fine, GH_PHI_UP, 0, 0))) return rc;
    // stage B: q = A phi (unscaled phi), then phi and its ghost level scaled; beta on the owned cells AND the ghost layer below
    // (both neighbours keep that layer), alpha = (BF)^*(-beta), scalings, z = 0
    for (Slab* sf : fine->slabs) {
        Slab* sc = coarse->local(sf->id);
        const TRange& tr = sf->tr;
        const double* w = fine->weighted ? sf->weight : nullptr;
        launch_prolong_q(gf, k, sf->phi, w, sf->q[0], tr.tn0, tr.tn1, st);
        const int n_hi = std::min(tr.tn1 + 1, gf.nt);
        launch_mul_inplace(sf->phi + (i64)tr.tn0 * gf.P, (i64)(n_hi - tr.tn0) * gf.P, k.phi_scale, st);
        launch_prolong_beta(gc, gf, k.beta_recover, sc->beta[coarse->bcur], sf->beta[0], sf->lo_c, tr.tc1, st);
        launch_bfdconj(gf, 1.0, sf->beta[0], sf->alpha, st, &tr);                                   // mexBFdConj(alpha, ., 1)
        for (auto& x : sf->q_own) launch_prolong_alpha(sf->alpha + x.b, w ? w + x.b : nullptr, x.e - x.b, k.alpha_scale, st);
        for (int j = 0; j < 10; j++)
            launch_mul_inplace(sf->beta[0] + (i64)j * gf.L + (i64)sf->lo_c * gf.PC, (i64)(tr.tc1 - sf->lo_c) * gf.PC, k.beta_scale, st);
        for (auto& x : sf->b_all) CU(cudaMemsetAsync(sf->beta[1] + x.b, 0, (size_t)(x.e - x.b) * sizeof(double), st));   // z = 0 (jump_nextLevel.m:9)
        fine->launches += 5 + (double)sf->q_own.size() + 10;
    }
    CU(cudaGetLastError());
    if ((rc = ghosts(fine, GH_PHI_DOWN | GH_Q_UP | GH_Q_DOWN | GH_ALPHA0_DOWN | GH_W, 0, 0))) return rc;   // (the sGS phi-step reads phi[t-1] too)
    CU(cudaStreamSynchronize(st));
    fine->z_materialised = true;
    fine->z_absent = false;
    fine->uploaded = true;
    return DOTSOCP_OK;
}

extern "C" int dotsocp_download(dotsocp_ctx* c, double* phi, double* q, double* z, double* alpha, double* beta)
{
    if (!c) return set_err(DOTSOCP_EINVAL, "NULL ctx");
    if (!c->uploaded) return set_err(DOTSOCP_ESTATE, "download before upload");
    if (!c->z_materialised && z) return set_err(DOTSOCP_ESTATE, "z is not materialised (session still open)");
    if (c->z_absent && z) return set_err(DOTSOCP_ESTATE, "z was neither uploaded nor computed yet");
    const Geo& g = c->g;
    for (Slab* s : c->slabs) {
        HostMap hm{c->world > 1 && !c->emulate, &c->gh, s->tr};
        const TRange& tr = s->tr;
        int rc;
        if (phi && (rc = xfer(c, s->phi + tr.tn0 * g.P, phi + hm.nodes(tr.tn0), (size_t)(tr.tn1 - tr.tn0) * g.P, false))) return rc;
        if (q && (rc = copy_stag(c, s, hm, s->q[c->qcur], q, false))) return rc;
        if (alpha && (rc = copy_stag(c, s, hm, s->alpha, alpha, false))) return rc;
        if (beta && (rc = copy_cols(c, s, hm, s->beta[c->bcur], beta, false))) return rc;
        if (z && (rc = copy_cols(c, s, hm, s->beta[1 - c->bcur], z, false))) return rc;
    }
    CU(cudaStreamSynchronize(c->st));
    return DOTSOCP_OK;
}

// ------------------------------------------------------------------------------------------------ loop helpers
static const double UPDATE_RULE[11][2] = {   // solver_socp_inPALM.m:39-51
    {1.1, 1.10}, {1.2, 1.15}, {1.5, 1.20}, {2, 1.26}, {2.5, 1.28}, {3.33, 1.32},
    {5, 1.35}, {10, 1.40}, {20, 1.60}, {40, 1.80}, {50, 2.00}};

static const double SGS_UPDATE_RULE[6][2] = {   // solver_socp_sGSinPALM.m:37-44
    {1.5, 1.20}, {2, 1.26}, {2.5, 1.28}, {3.33, 1.32}, {5, 1.35}, {10, 1.40}};

static double get_factor(double xi, const double (*rule)[2], int nrule)   // adjust_lagrangianParam.m:47-59
{
    double factor = 1;
    for (int i = 0; i < nrule; i++) {
        if (xi >= rule[i][0]) factor = rule[i][1];
        else break;
    }
    return factor;
}
static void adjust_lagrangianParam(double& sigma, double xi, double& factor, const double (*rule)[2] = UPDATE_RULE,
                                   int nrule = 11)   // adjust_lagrangianParam.m:14-39
{
    const double lower = 1e-3, upper = 1e3;
    if (xi >= 1) factor = get_factor(xi, rule, nrule);
    else if (xi < 1) factor = 1 / get_factor(1 / xi, rule, nrule);
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
static bool IfAdjustSigma_sGS(double it, double last, double scale)   // solver_socp_sGSinPALM.m:431-456
{
    double passed = it - last;
    it = it / scale;
    passed = passed / scale;
    if (it < 20 && passed >= 5) return true;
    if (it < 50 && passed >= 10) return true;
    if (it < 100 && passed >= 20) return true;
    if (it < 200 && passed >= 35) return true;
    if (it < 500 && passed >= 50) return true;
    return passed >= 100;
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

static int ensure_alloc(double*& p, i64 n)
{
    if (p) return 0;
    cudaError_t e = cudaMalloc(&p, (size_t)n * sizeof(double));
    if (e != cudaSuccess) { cudaGetLastError(); return set_err(DOTSOCP_ENOMEM, "cudaMalloc(%lld doubles): %s", (long long)n, cudaGetErrorString(e)); }
    if (const char* poison = getenv("DOTSOCP_POISON"))
        if (poison[0] == '1') { launch_fill(p, n, std::numeric_limits<double>::quiet_NaN(), 0); cudaDeviceSynchronize(); }
    return 0;
}

static bool dbg_sums() { static const bool on = [] { const char* e = getenv("DOTSOCP_DEBUG_SUMS"); return e && e[0] == '1'; }(); return on; }
static void dbg_state(dotsocp_ctx* c, const char* tag, int point = 0)
{
    if (!dbg_sums() || c->world != 1) return;
    Slab* s = c->slabs[0];
    const Geo& g = c->g;
    char nm[64];
    auto one = [&](const char* a, const double* x, i64 n) { snprintf(nm, sizeof nm, "%d:%s %s", g.nx, tag, a); debug_sum(nm, x, n, c->st); };
    one("phi", s->phi, g.N); one("rhs", s->rhs, g.N); one("q[0]", s->q[0], g.Q); one("q[1]", s->q[1], g.Q);
    one("alpha", s->alpha, g.Q); one("q2", s->q2, g.Q); one("beta[0]", s->beta[0], 10 * g.L); one("beta[1]", s->beta[1], 10 * g.L);
    if (point == 5) debug_sum_flush(c->st);
}

enum ArrKind { A_PHI, A_Q, A_ALPHA, A_BETA, A_ZMAT, A_C };

struct Loop {
    dotsocp_ctx* c;
    IterScal sc;
    double sc_D2 = 1.0;
    UpdateArgs ua(Slab* s) const
    {
        UpdateArgs a;
        a.g = c->g; a.tr = s->tr; a.sc = sc; a.phi = s->phi; a.q_old = s->q[c->qcur]; a.q_new = s->q[1 - c->qcur];
        a.alpha = s->alpha; a.weight = s->weight; a.beta_in = s->beta[c->bcur]; a.beta_out = s->beta[1 - c->bcur];
        a.q2 = s->q2; a.rhs = s->rhs; a.c0 = s->c0; a.c1 = s->c1; a.kkt_t0 = s->tr.tn0;
        a.side = s->side; a.side_t0 = s->lo_c; a.side_layers = s->tr.tc1 - s->lo_c;
        a.maps = nullptr;
        if (s->maps_ok) {
            s->km.beta = s->tm_beta[c->bcur];
            for (int i = 0; i < 3; i++) {
                s->km.qn[i] = s->tm_q[1 - c->qcur][i];
                s->km.qo[i] = s->tm_q[c->qcur][i];
                s->km.al[i] = s->tm_alpha[i];
                s->km.w[i] = c->weighted ? s->tm_w[i] : s->tm_alpha[i];
            }
            s->km.toc = s->lo_c;
            s->km.ton = s->lo_n;
            a.maps = &s->km;
        }
        return a;
    }
    // q2, rhs from the current (q, alpha, beta): the z-step part of the first iteration / after any rescaling
    void prologue()
    {
        for (Slab* s : c->slabs) {
            UpdateArgs a = ua(s);
            a.q_old = nullptr;
            a.q_new = s->q[c->qcur];
            a.beta_out = nullptr;
            launch_mult(a, c->weighted, c->one_d, false, c->st);
            c->launches += 1;
        }
    }
    int step_phi() { return solve_poisson(c, sc_D2); }
    // phi-step of the sGS loops: mexsGS(phi, rhs, 0, D^2, nt, nx, ny, 1) = odd, even, odd half sweeps (sgs.cu); time slabs
    // refresh one ghost plane of phi per side after every half sweep -- the only communication of this phi-step
    int step_sgs()
    {
        static const int order[3] = {1, 0, 1};
        for (int hs = 0; hs < 3; hs++) {
            for (Slab* s : c->slabs) {
                launch_sgs_half(c->g, 0.0, sc_D2, order[hs], s->rhs, s->phi, s->tr.tn0, s->tr.tn1, c->st);
                c->launches += 1;
            }
            int rc = ghosts(c, GH_PHI_UP | GH_PHI_DOWN, 0, 0);
            if (rc) return rc;
        }
        return 0;
    }
    // fused KKT: per-slab partial buffers (allocated at the first fused check) + the scalars of the terms
    int kkt_fused(Slab* s, double sigma, double cScale, double dScale, double D, double E, KktFused* kf)
    {
        const int nlev = s->tr.tn1 - s->tr.tn0;
        int rc = ensure_alloc(s->partial_q, (i64)nlev * kkt_blocks_x(c->g) * KQ_COUNT);
        if (!rc) rc = ensure_alloc(s->partial_m, (i64)nlev * mult_tiles(c->g) * KM_COUNT);
        if (rc) return rc;
        *kf = KktFused{sigma, cScale, dScale, D, E, s->partial_q, s->partial_m, c->d_lvl};
        return 0;
    }
    int step_q(bool acc, const KktFused* kf_tmpl = nullptr)
    {
        auto fused = [&](Slab* s, KktFused* kf) -> const KktFused* {
            if (!kf_tmpl) return nullptr;
            *kf = *kf_tmpl;
            kf->partial_q = s->partial_q; kf->partial_m = s->partial_m;
            return kf;
        };
        if (c->world == 1) {
            KktFused kf;
            launch_qstep(ua(c->slabs[0]), c->weighted, acc, c->st, nullptr, nullptr, true, fused(c->slabs[0], &kf));
            c->launches += 1;
            return 0;
        }
        // time slabs: first the two node levels whose q / alpha_0 the neighbours need, then their exchange on the
        // communication stream while the interior levels are processed
        auto part = [&](Slab* s, int t0, int t1) {
            if (t1 <= t0) return;
            UpdateArgs a = ua(s);
            a.tr.tn0 = t0; a.tr.tn1 = t1;
            KktFused kf;
            launch_qstep(a, c->weighted, acc, c->st, nullptr, nullptr, true, fused(s, &kf));
            c->launches += 1;
        };
        for (Slab* s : c->slabs) {
            const int t0 = s->tr.tn0, t1 = s->tr.tn1;
            part(s, t0, t0 + 1);
            if (t1 - 1 > t0) part(s, t1 - 1, t1);
        }
        c->cev_used = 0;
        cudaEvent_t e = c->comm_event();
        CU(cudaEventRecord(e, c->st));
        CU(cudaStreamWaitEvent(c->st2, e, 0));
        int rc = ghosts(c, GH_Q_UP | GH_Q_DOWN | GH_ALPHA0_DOWN, 1 - c->qcur, 0, c->st2);
        if (rc) return rc;
        for (Slab* s : c->slabs) part(s, s->tr.tn0 + 1, s->tr.tn1 - 1);
        cudaEvent_t e2 = c->comm_event();
        CU(cudaEventRecord(e2, c->st2));
        CU(cudaStreamWaitEvent(c->st, e2, 0));
        return 0;
    }
    void step_mult(const KktFused* kf_tmpl = nullptr)
    {
        for (Slab* s : c->slabs) {
            KktFused kf;
            if (kf_tmpl) { kf = *kf_tmpl; kf.partial_q = s->partial_q; kf.partial_m = s->partial_m; }
            c->launches += launch_mult(ua(s), c->weighted, c->one_d, true, c->st, kf_tmpl ? &kf : nullptr);
        }
        c->qcur ^= 1;
        c->bcur ^= 1;
        c->z_materialised = false;
    }
    // x = (x*mul)/div on every backed window (owned + ghost copies, so the ghosts stay consistent without traffic)
    void scale(ArrKind k, double mul, double div)
    {
        const Geo& g = c->g;
        for (Slab* s : c->slabs) {
            if (k == A_C) {
                launch_scale(s->c0, g.P, mul, div, c->st);
                launch_scale(s->c1, g.P, mul, div, c->st);
                c->launches += 2;
                continue;
            }
            double* base = k == A_PHI ? s->phi : k == A_Q ? s->q[c->qcur] : k == A_ALPHA ? s->alpha
                         : k == A_BETA ? s->beta[c->bcur] : s->beta[1 - c->bcur];
            const Ranges& r = k == A_PHI ? s->n_all : (k == A_Q || k == A_ALPHA) ? s->q_all : s->b_all;
            for (auto& x : r) { launch_scale(base + x.b, x.e - x.b, mul, div, c->st); c->launches += 1; }
        }
    }
};

// Level-table reductions (kernels.h): the producers of a check fill their own rows of c->d_lvl; finish_sums() brings in the
// rows of the other ranks (all-reduce of rows with exactly one non-zero contributor: exact), adds the nt rows in a fixed order
// and returns the KSL totals on the host.  The result is bit-identical for any number of slabs / GPUs.
static int begin_sums(dotsocp_ctx* c)
{
    CU(cudaMemsetAsync(c->d_lvl, 0, (size_t)c->g.nt * KSL * sizeof(double), c->st));
    return 0;
}
static int finish_sums(dotsocp_ctx* c, const double** out)
{
    if (c->comm)
        NC(nccl_api().AllReduce(c->d_lvl, c->d_lvl, (size_t)c->g.nt * KSL, NCCL_FLOAT64, NCCL_SUM, c->comm, c->st));
    launch_levels_total(c->d_lvl, c->g.nt, c->d_tot, c->st);
    c->launches += 1;
    CU(cudaMemcpyAsync(c->h_tot, c->d_tot, KSL * sizeof(double), cudaMemcpyDeviceToHost, c->st));
    CU(cudaStreamSynchronize(c->st));
    *out = c->h_tot;
    return 0;
}

static double now_s()
{
    using namespace std::chrono;
    return duration<double>(steady_clock::now().time_since_epoch()).count();
}

// z = Pi_Q(d + BF q_old - beta_old) of the owned cells, written over beta_old
static int zstep_all(dotsocp_ctx* c, const IterScal& sc)
{
    for (Slab* s : c->slabs) {
        launch_zstep(c->g, sc, c->one_d, s->q[1 - c->qcur], s->beta[1 - c->bcur], s->beta[1 - c->bcur], c->st, &s->tr);
        c->launches += 1;
    }
    return 0;
}

// Output recovery on the device: what solver_dotsocp2d.m:268-287 does after the last level --
// recoverOrgVar, recover_RhoE, recover_q, check_massConservation -- from the resident state of a finished run.  Every field
// is produced into the (now free) Poisson rhs buffer and copied to the host array, NULL fields are skipped; host arrays
// follow the download convention (global, or the slab's own levels in a one-process-per-GPU session).
extern "C" int dotsocp_recover(dotsocp_ctx* c, const dotsocp_recover_scal* rs, const double* rho0, const double* rho1, double* rho,
                               double* Ex, double* Ey, double* q0, double* bx, double* by, double* sumRho, double* sumNegRho,
                               double* w2cost)
{
    if (!c || !rs) return set_err(DOTSOCP_EINVAL, "NULL argument");
    if (!c->uploaded || c->iter_open) return set_err(DOTSOCP_ESTATE, "recover: needs the state of a finished run");
    if (c->one_d && (Ey || by)) return set_err(DOTSOCP_EINVAL, "recover: the 1-D variant has no y fields");
    const Geo& g = c->g;
    const bool local_host = c->world > 1 && !c->emulate;
    const bool need_rho = rho || sumRho || sumNegRho || w2cost;
    double *d0 = nullptr, *d1 = nullptr;
    struct Free { double*& a; double*& b; ~Free() { cudaFree(a); cudaFree(b); } } fr{d0, d1};
    if (need_rho) {
        if (!rho0 || !rho1) return set_err(DOTSOCP_EINVAL, "recover: rho0 / rho1 are required for rho, the mass check and the cost");
        CU(cudaMalloc(&d0, g.P * sizeof(double)));
        CU(cudaMalloc(&d1, g.P * sizeof(double)));
        CU(cudaMemcpyAsync(d0, rho0, g.P * sizeof(double), cudaMemcpyHostToDevice, c->st));
        CU(cudaMemcpyAsync(d1, rho1, g.P * sizeof(double), cudaMemcpyHostToDevice, c->st));
    }
    auto args = [&](Slab* s) {
        RecoverArgs a;
        a.g = g; a.tr = s->tr; a.arec = rs->alpha_recover; a.qrec = rs->q_recover;
        a.alpha = s->alpha; a.q = s->q[c->qcur]; a.weight = c->weighted ? s->weight : nullptr;
        a.rho0 = d0; a.rho1 = d1; a.out = s->rhs; a.partial = s->partial; a.lvl = c->d_lvl;
        return a;
    };
    int rc;
    double* outs[6] = {rho, Ex, Ey, q0, bx, by};
    for (int which = 0; which < 6; which++) {
        if (!outs[which]) continue;
        const bool cells = which >= RC_Q0;
        for (Slab* s : c->slabs) {
            const TRange& tr = s->tr;
            launch_recover(args(s), which, c->weighted, c->st);
            c->launches += 1;
            const int t0 = tr.tn0, t1 = cells ? tr.tc1 : tr.tn1;
            if (t1 <= t0) continue;
            double* host = outs[which] + (local_host ? 0 : (i64)t0 * g.P);
            if ((rc = xfer(c, s->rhs + (i64)t0 * g.P, host, (size_t)(t1 - t0) * g.P, false))) return rc;
        }
        CU(cudaStreamSynchronize(c->st));   // the scratch buffer is reused by the next field
    }
    if (sumRho || sumNegRho || w2cost) {
        if ((rc = begin_sums(c))) return rc;
        for (Slab* s : c->slabs) { launch_recover_stats(args(s), c->weighted, c->one_d, c->st); c->launches += 2; }
        if (c->comm) NC(nccl_api().AllReduce(c->d_lvl, c->d_lvl, (size_t)g.nt * KSL, NCCL_FLOAT64, NCCL_SUM, c->comm, c->st));
        std::vector<double> lv((size_t)g.nt * KSL);
        CU(cudaMemcpyAsync(lv.data(), c->d_lvl, lv.size() * sizeof(double), cudaMemcpyDeviceToHost, c->st));
        CU(cudaStreamSynchronize(c->st));
        double w2 = 0.0;
        for (int t = 0; t < g.nt; t++) {
            if (sumRho) sumRho[t] = lv[(size_t)t * KSL + RS_SUMRHO] / (double)g.P;        // check_massConservation.m:20-24
            if (sumNegRho) sumNegRho[t] = lv[(size_t)t * KSL + RS_SUMNEG] / (double)g.P;
            w2 += lv[(size_t)t * KSL + RS_W2];
        }
        if (w2cost) *w2cost = w2 / (double)g.N;
    }
    CU(cudaStreamSynchronize(c->st));
    CU(cudaGetLastError());
    return DOTSOCP_OK;
}

// ------------------------------------------------------------------------------------------------ the level loops
// One function for the three reference loops; the shared parts (rescaling, KKT, sigma rule, output) are literally the
// same code in solver_socp_inPALM.m, solver_socp_PALM.m and solver_socp_accADMM.m, only the iteration body differs.
//   inPALM / ALG2 : fused kernels, z never stored (recomputed from (q_old, beta_old) where the reference reads it)
//   PALM, acc-ADMM: z is genuine state (it enters the first q-step / the extrapolation), kept in beta[1-bcur]
static int run_level(dotsocp_ctx* c, const dotsocp_level_opts& o, dotsocp_hist* hist, dotsocp_level_result* res)
{
    const Geo& g = c->g;
    const int method = o.method;
    // sGS-inPALM (solver_socp_sGSinPALM.m) is the inPALM iteration with the phi-step replaced by one symmetric Gauss-Seidel
    // sweep, its own check schedule and its own sigma voting: it shares the fused kernels and everything marked `inpalm`
    const bool sgs = method == DOTSOCP_METHOD_SGSINPALM;
    // acc-sGS-ADMM (solver_socp_accsGSADMM.m) is acc-ADMM with the same replacement of the phi-step and the same voting
    const bool accsgs = method == DOTSOCP_METHOD_ACCSGSADMM, vote = sgs || accsgs;
    const bool inpalm = method == DOTSOCP_METHOD_INPALM || sgs, palm = method == DOTSOCP_METHOD_PALM;
    const bool acc = method == DOTSOCP_METHOD_ACCADMM || accsgs;
    const bool weighted = c->weighted;
    const bool checkPD = o.checkPrimDualFeas < 0 ? !weighted : (o.checkPrimDualFeas != 0);   // :20-24 / wsocp :25-29
    // NaN = opts.time_limit absent (default 3600 s, :26-30); a non-positive value is a budget that is already spent (the
    // multilevel drivers pass time_limit - Total_Time on, solver_dotsocp2d.m:244): one iteration, one check, stop (:287-289)
    const double time_limit = std::isnan(o.time_limit) ? 3600 : o.time_limit;
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
    const bool halpern = stepAlpha == 2;   // :30
    if (c->z_absent && (!inpalm || maxit < 1))
        return set_err(DOTSOCP_ESTATE, "z was not uploaded: only inPALM with maxit >= 1 never reads the incoming z");
    if (palm && (weighted || c->one_d)) return set_err(DOTSOCP_EINVAL, "PALM exists only for socp/dot2d");
    if (acc && c->one_d) return set_err(DOTSOCP_EINVAL, "acc-ADMM does not exist for socp/dot1d");
    if (!inpalm && c->world > 1) return set_err(DOTSOCP_EINVAL, "PALM / acc-ADMM run on a single slab only (world == 1)");
    if (vote && (c->variant != DOTSOCP_VARIANT_DOT2D || !sgs_supported(g)))
        return set_err(DOTSOCP_EINVAL, "the sGS loops exist only for socp/dot2d, and mexsGS only handles nx == ny with odd node counts");
    int kacc = 0;
    // sGS-inPALM state (:76-80, :109-112)
    const int sgs_hist = 19, sgs_victory = 12;
    const double initialSigmaScale = 1.10, sigma_adjust_val_gap = 0.95, tol_sgs_blocks = 5 * tol;
    const double sigma_adjust_it_gap = fmax(1.0, pow((double)g.nt * g.nx * g.ny, 1.0 / 3.0) / 33.0);
    bool stablePhase = false, sgs_superior_yes = false;
    std::vector<double> FeasRatio;
    if (vote) FeasRatio.assign((size_t)maxit + 1, INFINITY);   // 1-based like the reference
    const int stable_after = accsgs ? 1500 : 2500;            // accsGSADMM :379 / sGSinPALM :337
    double KRs[5] = {0, 0, 0, 0, 0}, norm_Aphi_s = 0, norm_q_s = 0;   // values of the last check (the :385-402 branch re-uses them)

    Loop L;
    L.c = c;
    L.sc = make_scal(o, D, E, dScale, tau);
    L.sc_D2 = D * D;
    const i64 nB = 10 * g.L;
    Slab* S0 = c->slabs[0];
    double* zmat = nullptr;      // PALM / acc: z lives here
    double* tmpq = nullptr;      // PALM: stored A*phi
    i64 vn[5] = {g.N, nB, g.Q, g.Q, nB};
    double* cur[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    if (!inpalm) {
        if (!c->z_materialised) return set_err(DOTSOCP_ESTATE, "z is not materialised");
        zmat = S0->beta[1 - c->bcur];
    }
    if (palm) { int rc = ensure_alloc(S0->tmpq, g.Q); if (rc) return rc; tmpq = S0->tmpq; }
    if (acc) {
        for (int i = 0; i < 5; i++) {
            int rc = ensure_alloc(S0->old_[i], vn[i]); if (rc) return rc;
            rc = ensure_alloc(S0->anc_[i], vn[i]); if (rc) return rc;
        }
    }
    auto refresh_cur = [&]() { cur[0] = S0->phi; cur[1] = zmat; cur[2] = S0->q[c->qcur]; cur[3] = S0->alpha; cur[4] = S0->beta[c->bcur]; };
    auto copy_to = [&](double** dst) {
        refresh_cur();
        for (int i = 0; i < 5; i++) cudaMemcpyAsync(dst[i], cur[i], (size_t)vn[i] * sizeof(double), cudaMemcpyDeviceToDevice, c->st);
    };

    // alpha, beta, c <- ./sigma  (:102-104)
    L.scale(A_ALPHA, 1.0, sigma);
    L.scale(A_BETA, 1.0, sigma);
    L.scale(A_C, 1.0, sigma);
    if (vote) {   // phi = phi - integralL2(phi, h)   (solver_socp_sGSinPALM.m:142, solver_socp_accsGSADMM.m:165)
        int r_;
        if ((r_ = begin_sums(c))) return r_;
        for (Slab* s : c->slabs) { launch_sum_nodes(g, s->phi, s->partial, c->d_lvl, 0, s->tr.tn0, s->tr.tn1, c->st); c->launches += 2; }
        const double* v;
        if ((r_ = finish_sums(c, &v))) return r_;
        const double shift = (1.0 / (double)g.N) * v[0];
        for (Slab* s : c->slabs)
            for (auto& x : s->n_all) { launch_shift(s->phi + x.b, x.e - x.b, shift, c->st); c->launches += 1; }
    }
    dbg_state(c, "entry", 1);
    if (inpalm) L.prologue();   // z2 := d + BF q (:133) folded into the first z-step
    dbg_state(c, "prologue", 2);
    if (palm) {                 // tmp_q = A*phi ; mexBFd(z, tmp_q, ...)   (PALM :137-138)
        UpdateArgs a = L.ua(S0);
        a.q_new = S0->q[1 - c->qcur];   // scratch: only tmpq_out matters here
        launch_qstep(a, false, false, c->st, nullptr, tmpq, false);
        launch_cells_update(g, L.sc, false, 2, tmpq, zmat, nullptr, c->st);
        c->launches += 2;
    }
    if (acc) { copy_to(S0->old_); copy_to(S0->anc_); }   // :157-163

    struct EvPair {   // destroyed on every exit path
        cudaEvent_t a = nullptr, b = nullptr;
        EvPair() { cudaEventCreate(&a); cudaEventCreate(&b); }
        ~EvPair() { cudaEventDestroy(a); cudaEventDestroy(b); }
    } ev_total;
    const cudaEvent_t ev_begin = ev_total.a, ev_end = ev_total.b;
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
    int it = 0, hist_len = 0, rc = 0;
    bool z_ever = false;
    for (it = 1; it <= maxit; it++) {
        // ---------------------------------------------------------------- rescaling :138-190
        bool scaleYes = false;
        double normPhis = 0, normAlps = 0;
        auto rescale_norms = [&]() -> int {
            int r;
            if ((r = begin_sums(c))) return r;
            for (Slab* s : c->slabs) {
                KktArgs ka;
                ka.g = g; ka.tr = s->tr; ka.sc = L.sc; ka.sigma = sigma; ka.cScale = cScale; ka.dScale = dScale; ka.D = D; ka.E = E;
                ka.phi = s->phi; ka.q = s->q[c->qcur]; ka.alpha = s->alpha; ka.weight = s->weight; ka.beta = s->beta[c->bcur];
                ka.z = c->z_materialised ? s->beta[1 - c->bcur] : nullptr;
                ka.q_old = s->q[1 - c->qcur]; ka.beta_old = s->beta[1 - c->bcur];
                ka.q2b = nullptr; ka.c0 = s->c0; ka.c1 = s->c1; ka.partial = s->partial; ka.lvl = c->d_lvl;
                launch_norms(ka, c->one_d, c->st);
                c->launches += 2;
            }
            const double* v;
            if ((r = finish_sums(c, &v))) return r;
            const double normPhi = sqrt(h) * sqrt(v[NR_PHI2]), normQ = sqrt(h) * sqrt(v[NR_Q2]), normZ = sqrt(h) * sqrt(v[NR_Z2]);
            const double normAlpha = sigma * (sqrt(h) * sqrt(v[NR_ALPHA2])), normBeta = sigma * (sqrt(h) * sqrt(v[NR_BETA2]));
            normPhis = mmax({normPhi, normQ, normZ});
            normAlps = mmax({normAlpha, normBeta});
            return 0;
        };
        if (rescale >= 3 && it % checkRescaleIters == 0) {
            if ((rc = rescale_norms())) return rc;
            const double ratio = fmax(normAlps, normPhis) / fmin(normAlps, normPhis);
            if (ratio > ratioThreshold) scaleYes = true;
        }
        if ((rescale == 1 && maxFeas < 2e-2 && it >= firstScaleIter && relGap < 5e-2) ||
            (rescale == 2 && maxFeas < 5e-3 && it >= SecondScaleIter && relGap < 1e-2) || scaleYes) {
            if (!scaleYes && (rc = rescale_norms())) return rc;
            const double dScale2 = normPhis, cScale2 = normAlps;
            sigma = sigma * (cScale2 / dScale2);
            const double cs2 = cScale2 * cScale2;
            // c, alpha, beta <- x * dScale2 / cScale2^2 ; q, z <- ./dScale2 (inPALM: z is recomputed, never stored)
            L.scale(A_C, dScale2, cs2);
            norm_c = norm_c / cScale2;
            if (!weighted) norm_d = norm_d / dScale2;
            L.scale(A_ALPHA, dScale2, cs2);
            L.scale(A_BETA, dScale2, cs2);
            if (acc || sgs) L.scale(A_PHI, 1.0, dScale2);          // accADMM :207 ; sGSinPALM :184
            if (!palm) L.scale(A_Q, 1.0, dScale2);                 // :177 (absent in PALM)
            if (c->z_materialised) L.scale(A_ZMAT, 1.0, dScale2);
            if (palm) { launch_scale(tmpq, g.Q, 1.0, dScale2, c->st); c->launches += 1; }   // PALM :191
            dScale = dScale2 * dScale;
            cScale = cScale2 * cScale;
            L.sc.DF = E / dScale;                                   // scaleD
            sigmaScale = sigmaScale * (cScale2 / dScale2);
            if (inpalm) L.prologue();                               // mexBFd(z2, q, ...) refresh (:187)
            if (acc) { kacc = 0; copy_to(S0->old_); copy_to(S0->anc_); }   // accADMM :217-222
            rescale += 1;
        }

        // ---------------------------------------------------------------- iteration
        cudaEvent_t e_last;
        bool sums_begun = false;   // the level table already holds a slot of this iteration (acc-sGS residual)
        // the check schedule depends only on (it, lastSigmaIt), so an inPALM iteration knows beforehand whether a check
        // follows it and lets k_qstep / k_mult accumulate the KKT sums on the data they stream anyway (:218-267)
        bool fused_check = false;
        if (inpalm) {   // :192-216, fused order
            const bool sched = sgs ? IfAdjustSigma_sGS(it, lastSigmaIt, sigma_adjust_it_gap) : IfAdjustSigma(it, lastSigmaIt);
            // sGS: while sgs_superior_yes holds the reference also evaluates two of the residuals on non-check iterations
            // (:385-402); they come from the same fused sums
            fused_check = (c->fuse_kkt || sgs) && (checkSByS || sched || it == maxit || (sgs && sgs_superior_yes));
            KktFused kf{sigma, cScale, dScale, D, E, nullptr, nullptr, c->d_lvl};
            if (fused_check) {
                for (Slab* s : c->slabs) { KktFused tmp; if ((rc = L.kkt_fused(s, sigma, cScale, dScale, D, E, &tmp))) return rc; }
                if ((rc = begin_sums(c))) return rc;
            }
            cudaEvent_t e0 = mark();
            if ((rc = sgs ? L.step_sgs() : L.step_phi())) return rc;
            if (it == 1) dbg_state(c, "it1 phi", 3);
            if (sgs && (checkSByS || sched || it == maxit)) {
                // error of the sGS blocks (:212-216): ||A'(A phi - q + alpha) - c|| over the even nodes, q / alpha of the previous iterate
                for (Slab* s : c->slabs) {
                    launch_sgs_resid(g, L.sc, true, s->phi, s->q[c->qcur], s->alpha, s->c0, s->c1, s->partial, c->d_lvl, KS_SGS_BLOCKS,
                                     s->tr.tn0, s->tr.tn1, c->st);
                    c->launches += 2;
                }
            }
            cudaEvent_t e1 = mark();
            if ((rc = L.step_q(false, fused_check ? &kf : nullptr))) return rc;
            cudaEvent_t e2 = mark();
            if (it == 1) dbg_state(c, "it1 q", 4);
            L.step_mult(fused_check ? &kf : nullptr);
            if (it == 1) dbg_state(c, "it1 mult", 5);
            cudaEvent_t e3 = mark();
            segs.push_back({e0, e1, 0});
            segs.push_back({e1, e2, 2});
            segs.push_back({e2, e3, 3});
            e_last = e3;
            z_ever = true;
        } else if (palm) {   // solver_socp_PALM.m:196-224
            double* q = S0->q[c->qcur];
            double* beta = S0->beta[c->bcur];
            UpdateArgs a = L.ua(S0);
            a.q_new = q;
            cudaEvent_t e0 = mark();
            launch_bfdconj_sum(g, L.sc.S, zmat, beta, S0->q2, c->st);
            launch_qstep(a, false, false, c->st, tmpq, nullptr, false);              // q = (tmp_q + alpha + q2).*diagQInv
            cudaEvent_t e1 = mark();
            launch_rhs(g, L.sc, false, q, S0->alpha, nullptr, S0->c0, S0->c1, S0->rhs, c->st);
            c->launches += 3;
            if ((rc = L.step_phi())) return rc;
            cudaEvent_t e2 = mark();
            launch_zstep(g, L.sc, false, q, beta, zmat, c->st);   // mexBFd + mexProjSoc (:209-210)
            cudaEvent_t e3 = mark();
            launch_bfdconj_sum(g, L.sc.S, zmat, beta, S0->q2, c->st);
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
            double* q = S0->q[c->qcur];
            double* beta = S0->beta[c->bcur];
            UpdateArgs a = L.ua(S0);
            a.q_new = q;
            cudaEvent_t e0 = mark();
            launch_bfdconj_sum(g, L.sc.S, zmat, beta, S0->q2, c->st);
            launch_qstep(a, weighted, true, c->st);                                  // q ; alpha = (alpha + A phi) - w.*q
            cudaEvent_t e1 = mark();
            launch_rhs(g, L.sc, weighted, q, S0->alpha, S0->weight, S0->c0, S0->c1, S0->rhs, c->st);
            if ((rc = accsgs ? L.step_sgs() : L.step_phi())) return rc;
            if (accsgs && (checkSByS || IfAdjustSigma_sGS(it, lastSigmaIt, sigma_adjust_it_gap) || it == maxit)) {
                // error of the sGS blocks (accsGSADMM :262-268), on the new q and alpha
                if ((rc = begin_sums(c))) return rc;
                sums_begun = true;
                launch_sgs_resid(g, L.sc, true, S0->phi, q, S0->alpha, S0->c0, S0->c1, S0->partial, c->d_lvl, KS_SGS_BLOCKS, 0, g.nt, c->st);
                c->launches += 2;
            }
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
        const bool adjustSigmaYes = vote ? IfAdjustSigma_sGS(it, lastSigmaIt, sigma_adjust_it_gap) : IfAdjustSigma(it, lastSigmaIt);
        bool over_time = false;
        if (!c->comm) {   // (between processes the clock is only consulted at collective points, see below)
            over_time = (now_s() - clock_total) > time_limit;
            if (over_time) {   // the host runs ahead of the device: confirm against completed work
                cudaStreamSynchronize(c->st);
                over_time = (now_s() - clock_total) > time_limit;
            }
        }
        const bool check = checkSByS || adjustSigmaYes || it == maxit || over_time;
        bool stop = false;
        if (check) {
            if (!fused_check && !sums_begun && (rc = begin_sums(c))) return rc;
            for (Slab* s : c->slabs) {
                if (fused_check) {
                    KktFused kf{sigma, cScale, dScale, D, E, s->partial_q, s->partial_m, c->d_lvl};
                    launch_kkt_fused_reduce(g, s->tr, kf, c->st);
                    c->launches += 2;
                    if (sgs) {   // ||A' resi_alpha|| = ||A'(A phi - q_new)|| of :322 (all nodes)
                        launch_sgs_resid(g, L.sc, false, s->phi, s->q[c->qcur], nullptr, s->c0, s->c1, s->partial, c->d_lvl, KS_SGS_KKT,
                                         s->tr.tn0, s->tr.tn1, c->st);
                        c->launches += 2;
                    }
                } else {
                    if ((rc = ensure_qtmp(c, s))) return rc;
                    launch_bfdconj(g, L.sc.S, s->beta[c->bcur], s->qtmp, c->st, &s->tr);   // q2 = s (BF)^* beta   (:225)
                    KktArgs ka;
                    ka.g = g; ka.tr = s->tr; ka.sc = L.sc; ka.sigma = sigma; ka.cScale = cScale; ka.dScale = dScale; ka.D = D; ka.E = E;
                    ka.phi = s->phi; ka.q = s->q[c->qcur]; ka.alpha = s->alpha; ka.weight = s->weight;
                    ka.beta = s->beta[c->bcur]; ka.z = zmat; ka.q_old = s->q[1 - c->qcur]; ka.beta_old = s->beta[1 - c->bcur];
                    ka.q2b = s->qtmp; ka.c0 = s->c0; ka.c1 = s->c1; ka.partial = s->partial; ka.lvl = c->d_lvl;
                    launch_kkt_cells(ka, weighted, c->one_d, c->st);
                    launch_kkt_nodes(ka, weighted, c->st);
                    c->launches += 5;
                    if (accsgs) {   // ||A'(A phi - q)|| of accsGSADMM :360
                        launch_sgs_resid(g, L.sc, false, s->phi, s->q[c->qcur], nullptr, s->c0, s->c1, s->partial, c->d_lvl, KS_SGS_KKT,
                                         s->tr.tn0, s->tr.tn1, c->st);
                        c->launches += 2;
                    }
                }
                // row 0 carries the elapsed time seen by slab 0 so that all processes decide alike
                if (s->id == 0) {
                    *c->h_elapsed = now_s() - clock_total;
                    CU(cudaMemcpyAsync(c->d_lvl + KS_ELAPSED, c->h_elapsed, sizeof(double), cudaMemcpyHostToDevice, c->st));
                }
            }
            cudaEvent_t e4 = mark();
            segs.push_back({e_last, e4, 4});
            const double* sums;
            if ((rc = finish_sums(c, &sums))) return rc;
            flush_segs();
            const double* sc_ = sums;
            const double* sn = sums + KC_COUNT;
            const double elapsed = c->comm ? sums[KS_ELAPSED] : (now_s() - clock_total);
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
            if (vote) {   // :284 and the state the :385-402 branch re-uses
                FeasRatio[it] = mmax({KR[0], KR[1]}) / mmax({KR[2], KR[4]});
                for (int j = 0; j < 5; j++) KRs[j] = KR[j];
                norm_Aphi_s = norm_Aphi;
                norm_q_s = norm_q;
            }
            if (hist && hist_len < hist->cap) {
                if (hist->kkt) for (int j = 0; j < 7; j++) hist->kkt[(size_t)hist_len * 7 + j] = KO[j];
                if (hist->time) hist->time[hist_len] = elapsed;
                if (hist->iter) hist->iter[hist_len] = it;
                if (hist->pdGap) hist->pdGap[hist_len] = pdGap;
                if (hist->priVal) hist->priVal[hist_len] = priVal;
                if (hist->dualVal) hist->dualVal[hist_len] = dualVal;
            }
            hist_len++;
            const double stopv = checkPD ? mmax({KO[0], KO[2], KO[5], KO[6]}) : mmax({KO[0], KO[2], KO[5]});
            if (stopv < tol || elapsed > time_limit) {
                stop = true;
            } else {
                if (mmax({KR[0], KR[1], KR[2], KR[3], KR[4]}) < tol_feasOrg) use_feasOrg = 1;
                auto apply_factor = [&](double factor) {
                    L.scale(A_ALPHA, 1.0, factor);
                    L.scale(A_BETA, 1.0, factor);
                    L.scale(A_C, 1.0, factor);
                    // inPALM: q2, rhs were computed with the old alpha/beta: refresh.  The prologue reads only the
                    // current buffers, so the (q_old, beta_old) pair that defines z survives.
                    if (inpalm) L.prologue();
                    if (acc) {   // accADMM :346-358
                        launch_scale(S0->old_[3], g.Q, 1.0, factor, c->st);
                        launch_scale(S0->old_[4], nB, 1.0, factor, c->st);
                        c->launches += 2;
                        kacc = 0;
                        copy_to(S0->anc_);
                    }
                };
                if (vote) {   // solver_socp_sGSinPALM.m:321-360 ; solver_socp_accsGSADMM.m:359-412
                    const double kkt_sgs_blocks = sqrt(nrm(sums[KS_SGS_KKT]) * nrm(sums[KS_SGS_KKT]) + (dualFea1 / sigma) * (dualFea1 / sigma));
                    const double resi_sGS_blocks = nrm(sums[KS_SGS_BLOCKS]);
                    sgs_superior_yes = resi_sGS_blocks < sigma_adjust_val_gap * kkt_sgs_blocks;
                    if (adjustSigmaYes) {
                        lastSigmaIt = it;
                        const int i0 = std::max(1, it - sgs_hist);
                        double sum = 0;
                        int primWin = 0, dualWin = 0;
                        for (int i = i0; i <= it; i++) {
                            sum += FeasRatio[i];
                            primWin += FeasRatio[i] < 1;
                            dualWin += FeasRatio[i] > 1;
                        }
                        const double meanFeasRatio = sum / (double)(it - i0 + 1);
                        const bool adjust2 = sgs_superior_yes || stopv < tol_sgs_blocks || (dualWin >= sgs_victory && meanFeasRatio > 1);
                        if (adjust2) {
                            if (it > stable_after) stablePhase = true;
                            if ((primWin >= sgs_victory && meanFeasRatio < 1) || (dualWin >= sgs_victory && meanFeasRatio > 1)) {
                                double factor = 1;
                                if (stablePhase) {
                                    adjust_lagrangianParam(sigma, meanFeasRatio, factor, SGS_UPDATE_RULE, 6);
                                } else {
                                    if (meanFeasRatio < 1) factor = 1 / initialSigmaScale;
                                    else if (meanFeasRatio > 1) factor = initialSigmaScale;
                                    sigma = sigma * factor;
                                }
                                if (factor != 1) apply_factor(factor);
                            }
                        }
                    }
                } else if (adjustSigmaYes) {
                    lastSigmaIt = it;
                    double resiPri, resiDual;
                    if (use_feasOrg) { resiPri = mmax({KO[0], KO[1]}); resiDual = mmax({KO[2], KO[4]}); }
                    else { resiPri = mmax({KR[0], KR[1]}); resiDual = mmax({KR[2], KR[4]}); }
                    double factor = 1;
                    adjust_lagrangianParam(sigma, resiPri / resiDual, factor);
                    if (factor != 1) apply_factor(factor);
                }
                if (rescale > 0) {
                    maxFeas = mmax({KR[0], KR[1], KR[2], KR[3], KR[4]});
                    relGap = pdGap;
                }
            }
        }
        if (vote && !check) {
            if (fused_check || (accsgs && sgs_superior_yes)) {   // sgs_superior_yes: primal / dual feasibility of this iteration (:385-402)
                if (accsgs) {   // acc keeps z as state and has no fused sums: the stand-alone node kernel (accsGSADMM :420-425)
                    if ((rc = begin_sums(c))) return rc;
                    if ((rc = ensure_qtmp(c, S0))) return rc;
                    launch_bfdconj(g, L.sc.S, S0->beta[c->bcur], S0->qtmp, c->st, &S0->tr);
                    KktArgs ka;
                    ka.g = g; ka.tr = S0->tr; ka.sc = L.sc; ka.sigma = sigma; ka.cScale = cScale; ka.dScale = dScale; ka.D = D; ka.E = E;
                    ka.phi = S0->phi; ka.q = S0->q[c->qcur]; ka.alpha = S0->alpha; ka.weight = S0->weight;
                    ka.beta = S0->beta[c->bcur]; ka.z = zmat; ka.q_old = S0->q[1 - c->qcur]; ka.beta_old = S0->beta[1 - c->bcur];
                    ka.q2b = S0->qtmp; ka.c0 = S0->c0; ka.c1 = S0->c1; ka.partial = S0->partial; ka.lvl = c->d_lvl;
                    launch_kkt_nodes(ka, weighted, c->st);
                    c->launches += 3;
                }
                for (Slab* s : c->slabs) {
                    if (accsgs) break;
                    KktFused kf{sigma, cScale, dScale, D, E, s->partial_q, s->partial_m, c->d_lvl};
                    launch_kkt_fused_reduce(g, s->tr, kf, c->st);
                    c->launches += 2;
                }
                const double* sums;
                if ((rc = finish_sums(c, &sums))) return rc;
                flush_segs();
                const double primFea1 = sqrt(h) * sqrt(sums[KC_COUNT + KN_PRIM1]);
                const double dualFea1 = sigma * (sqrt(h) * sqrt(sums[KC_COUNT + KN_DUAL1]));
                if (use_feasOrg) {
                    const double dec = primFea1 / ((kktConst * D / dScale + norm_Aphi_s + norm_q_s) * KRs[0]);
                    KRs[0] *= dec; KRs[1] *= dec;
                    KRs[2] = dualFea1 / (kktConst / cScale + norm_c);
                } else {
                    const double dec = primFea1 / ((kktConst + norm_Aphi_s + norm_q_s) * KRs[0]);
                    KRs[0] *= dec; KRs[1] *= dec;
                    KRs[2] = dualFea1 / (kktConst + norm_c);
                }
                FeasRatio[it] = mmax({KRs[0], KRs[1]}) / mmax({KRs[2], KRs[4]});
            } else {
                FeasRatio[it] = FeasRatio[it - 1];
            }
        }
        if (stop) break;
        if (acc && halpern) {   // Halpern iteration, accADMM :371-388
            cudaEvent_t e5 = mark();
            const double c1 = 1.0 / (kacc + 2), c2 = (double)(kacc + 1) / (kacc + 2);
            kacc += 1;
            const bool anchor = kacc >= restart;
            refresh_cur();
            for (int i = 0; i < 5; i++) launch_halpern(cur[i], S0->old_[i], S0->anc_[i], vn[i], c1, c2, stepRho, anchor, c->st);
            c->launches += 5;
            if (anchor) kacc = 0;
            cudaEvent_t e6 = mark();
            segs.push_back({e5, e6, 5});
        } else if (acc) {       // general extrapolation, accADMM :389-417 (anc_ holds the previous "hat" iterate)
            cudaEvent_t e5 = mark();
            const double c1 = stepAlpha / (2 * (kacc + stepAlpha));
            const double c2 = kacc / (kacc + stepAlpha);
            const bool first = kacc == 0;
            kacc += 1;
            const bool restarted = kacc >= restart;
            refresh_cur();
            for (int i = 0; i < 5; i++)
                launch_accel3(cur[i], S0->old_[i], S0->anc_[i], vn[i], stepRho, 1 - c1, first ? c1 : c1 + c2, c2, first, !restarted, c->st);
            c->launches += 5;
            if (restarted) kacc = 0;
            cudaEvent_t e6 = mark();
            segs.push_back({e5, e6, 5});
        }
    }
    if (it > maxit) it = maxit;
    // ---------------------------------------------------------------- output :328-357
    if (!c->z_materialised && z_ever) {
        // z = Pi_Q(d + BF q_old - beta_old), written over beta_old (cell-local, safe in place)
        if ((rc = zstep_all(c, L.sc))) return rc;
        c->z_materialised = true;
        c->z_absent = false;
    }
    L.scale(A_ALPHA, sigma, 1.0);      // var.alpha = sigma*alpha
    L.scale(A_BETA, sigma, 1.0);       // var.beta  = sigma*beta
    L.scale(A_C, sigma, 1.0);          // undo the folding of model.c (the reference never writes its local copy back)
    cudaEventRecord(ev_end, c->st);
    CU(cudaStreamSynchronize(c->st));
    flush_segs();
    float total_ms = 0;
    cudaEventElapsedTime(&total_ms, ev_begin, ev_end);
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
        else if (accsgs) { res->times[0] = T[0]; res->times[1] = T[1]; res->times[2] = T[3]; res->times[3] = T[2]; res->times[4] = T[5]; res->times[5] = T[4]; res->times[6] = tot; }
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
    if (o->method < 0 || o->method > 4) return set_err(DOTSOCP_EINVAL, "unknown method %d", o->method);
    return run_level(c, *o, hist, res);
}

// dotsocp_solve_level keeps its session alive between calls (process lifetime, keyed by variant and grid): the MEX gateway is
// called once per level and per solve, and at 1024x1024x512 creating and destroying the 147 GB of device arrays costs about a
// second per call.  dotsocp_release_cached() frees it (the gateway registers it with mexAtExit); a later dotsocp_create that
// runs out of device memory releases it too and retries.  DOTSOCP_CACHE_CTX=0 restores create/destroy per call.
static dotsocp_ctx* g_cached = nullptr;
static std::mutex g_cached_mu;
static void release_cached_locked()
{
    if (g_cached) { dotsocp_destroy(g_cached); g_cached = nullptr; }
}
static void release_cached_unlocked_if_free()
{
    // called from dotsocp_create: inside dotsocp_solve_level the lock is held and g_cached is already detached (nothing to do)
    if (g_cached_mu.try_lock()) { release_cached_locked(); g_cached_mu.unlock(); }
}
extern "C" void dotsocp_release_cached(void)
{
    std::lock_guard<std::mutex> lk(g_cached_mu);
    release_cached_locked();
}

extern "C" int dotsocp_solve_level(const dotsocp_level_opts* o, double* phi, double* q, double* z, double* alpha, double* beta,
                                   const double* cvec, const double* weight, dotsocp_hist* hist, dotsocp_level_result* res)
{
    if (!o) return set_err(DOTSOCP_EINVAL, "NULL opts");
    static const bool keep = [] { const char* e = getenv("DOTSOCP_CACHE_CTX"); return !(e && e[0] == '0'); }();
    std::lock_guard<std::mutex> lk(g_cached_mu);
    int dev = 0;
    cudaGetDevice(&dev);
    dotsocp_ctx* c = g_cached;
    g_cached = nullptr;
    if (c && !(c->variant == o->variant && c->g.nt == o->nt && c->g.nx == o->nx && c->g.ny == o->ny && c->device == dev && c->world == 1)) {
        dotsocp_destroy(c);
        c = nullptr;
    }
    int rc = 0;
    if (!c) rc = dotsocp_create(&c, o->variant, o->nt, o->nx, o->ny, 0, 1, nullptr);
    if (rc) return rc;
    const double launches0 = c->launches;
    // inPALM overwrites z (solver_socp_inPALM.m:199) before it is ever read, so its incoming value need not cross PCIe
    const bool z_dead = (o->method == DOTSOCP_METHOD_INPALM || o->method == DOTSOCP_METHOD_SGSINPALM) && o->maxit >= 1;
    rc = dotsocp_upload(c, phi, q, z_dead ? nullptr : z, alpha, beta, cvec, weight);
    if (!rc) rc = dotsocp_run(c, o, hist, res);
    if (!rc) rc = dotsocp_download(c, phi, q, z, alpha, beta);
    if (!rc && res) res->gpu_launches -= launches0;
    if (rc || !keep) dotsocp_destroy(c);   // a failed call leaves no state behind
    else g_cached = c;
    return rc;
}

// ------------------------------------------------------------------------------------------------ benchmark session
extern "C" int dotsocp_iter_begin(dotsocp_ctx* c, const dotsocp_level_opts* o)
{
    if (!c || !o) return set_err(DOTSOCP_EINVAL, "NULL ctx/opts");
    if (!c->uploaded) return set_err(DOTSOCP_ESTATE, "iter_begin before upload");
    c->sc = make_scal(*o, o->D, o->E, o->dScale, o->tau);
    c->sigma_fold = o->sigma;
    c->D2 = o->D * o->D;
    c->it_cScale = o->cScale; c->it_dScale = o->dScale; c->it_D = o->D; c->it_E = o->E;
    Loop L; L.c = c; L.sc = c->sc; L.sc_D2 = c->D2;
    L.scale(A_ALPHA, 1.0, o->sigma);
    L.scale(A_BETA, 1.0, o->sigma);
    L.scale(A_C, 1.0, o->sigma);
    L.prologue();
    CU(cudaStreamSynchronize(c->st));
    c->iter_open = true;
    return DOTSOCP_OK;
}

extern "C" int dotsocp_iterate(dotsocp_ctx* c, int n_iters, int with_kkt_every, float* elapsed_ms, float* ms_by_kernel)
{
    if (!c || !c->iter_open) return set_err(DOTSOCP_ESTATE, "iterate without iter_begin");
    Loop L; L.c = c; L.sc = c->sc; L.sc_D2 = c->D2;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    std::vector<cudaEvent_t> ev;
    if (ms_by_kernel) {
        ev.resize((size_t)n_iters * 4);
        for (auto& e : ev) cudaEventCreate(&e);
    }
    int rc = 0;
    cudaEventRecord(a, c->st);
    for (int i = 0; i < n_iters && !rc; i++) {
        // with_kkt_every = k > 0: every k-th iteration is a check iteration (KKT sums fused into the update kernels, the
        // level-table reduction, the all-reduce between ranks and the read-back of the totals, like run_level's checks)
        const bool chk = with_kkt_every > 0 && (i + 1) % with_kkt_every == 0;
        KktFused kf{c->sigma_fold, c->it_cScale, c->it_dScale, c->it_D, c->it_E, nullptr, nullptr, c->d_lvl};
        if (chk) {
            for (Slab* s : c->slabs) { KktFused tmp; if ((rc = L.kkt_fused(s, kf.sigma, kf.cScale, kf.dScale, kf.D, kf.E, &tmp))) break; }
            if (!rc) rc = begin_sums(c);
            if (rc) break;
        }
        if (ms_by_kernel) cudaEventRecord(ev[4 * i + 0], c->st);
        rc = L.step_phi();
        if (ms_by_kernel) cudaEventRecord(ev[4 * i + 1], c->st);
        if (!rc) rc = L.step_q(false, chk ? &kf : nullptr);
        if (ms_by_kernel) cudaEventRecord(ev[4 * i + 2], c->st);
        L.step_mult(chk ? &kf : nullptr);
        if (ms_by_kernel) cudaEventRecord(ev[4 * i + 3], c->st);
        if (chk && !rc) {
            for (Slab* s : c->slabs) {
                KktFused k2 = kf;
                k2.partial_q = s->partial_q; k2.partial_m = s->partial_m;
                launch_kkt_fused_reduce(c->g, s->tr, k2, c->st);
                c->launches += 2;
            }
            const double* sums;
            rc = finish_sums(c, &sums);
        }
    }
    cudaEventRecord(b, c->st);
    if (rc) return rc;
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
    Loop L; L.c = c; L.sc = c->sc; L.sc_D2 = c->D2;
    if (!c->z_materialised) {
        int rc = zstep_all(c, c->sc);
        if (rc) return rc;
        c->z_materialised = true;
        c->z_absent = false;
    }
    L.scale(A_ALPHA, c->sigma_fold, 1.0);
    L.scale(A_BETA, c->sigma_fold, 1.0);
    L.scale(A_C, c->sigma_fold, 1.0);
    CU(cudaStreamSynchronize(c->st));
    c->iter_open = false;
    return DOTSOCP_OK;
}
