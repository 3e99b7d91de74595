// weights.cu -- level weights of the weighted variant generated, restricted and kept on the device.
//
// Reference: examples/wdot2d/gene_weight_circle.m:6-27 and get_weight_by_barrier.m:12-33 build the finest weight as
// [ones(L) ; repmat(weightX, nt) ; repmat(weightY, nt)] from two (x,y) planes; socp/wdot2d/utils/downSample_q.m:4-19 and
// downSample_barrier.m:4-24 restrict it level by level through sparse kron products of 1-D prolongation matrices
// (column-normalised and transposed); solver_wdotsocp2d.m:312-316 needs mean(log10(weight + 1e-10)) of every level.
// At 1025 x 1025 x 513 the weight is Q = 1.6e9 doubles (13 GB) per level chain: here only the two planes (or, for a general
// weight, one upload of the finest level) cross PCIe, the time replication, the restriction chain and the log-mean run as
// streaming kernels, and a session takes its level from the pyramid by a device-to-device scatter into its pitched layout
// (dotsocp_set_weight, solver.cu).  Every rank of a one-process-per-GPU run builds the (full-grid) pyramid on its own GPU --
// no communication -- and scatters only the windows its slab backs.
//
// The restriction is the separable form of the reference's kron products, evaluated in the order of the host mirror
// (dotsocp_b200/driver.py: t, then x, then y; `nearest` axes average a pair, `linear` axes use (.5,1,.5)/2 inside and
// (1,.5)/1.5 at the two ends) with separately rounded IEEE operations, so the arithmetic pyramid is bit-identical to the host
// mirror; the geometric one (barriers: exp of the restricted log) agrees to the rounding of log / exp.
#include "../../include/dotsocp.h"
#include "errs.h"
#include "kernels.h"
#include "reduce.cuh"

#include <new>
#include <vector>

using namespace dsocp;

namespace {

template <class F>
__device__ __forceinline__ double restrict_linear(int c, int nC, F f)    // downSample_q.m:25-31 normalised by :10-12
{
    const int nR = 2 * nC - 1;
    if (c == 0) return __ddiv_rn(__dadd_rn(f(0), __dmul_rn(0.5, f(1))), 1.5);
    if (c == nC - 1) return __ddiv_rn(__dadd_rn(f(nR - 1), __dmul_rn(0.5, f(nR - 2))), 1.5);
    return __ddiv_rn(__dadd_rn(f(2 * c), __dmul_rn(0.5, __dadd_rn(f(2 * c - 1), f(2 * c + 1)))), 2.0);
}
template <class F>
__device__ __forceinline__ double restrict_nearest(int c, F f)           // downSample_q.m:33-39: pair average
{
    return __ddiv_rn(__dadd_rn(f(2 * c), f(2 * c + 1)), 2.0);
}

// One thread per entry of the COARSE part.  PART 0 / 1 / 2 = q0 / bx / by part: that axis (t / x / y) is cell-centred
// ("nearest", 2 nC fine entries), the other two are node-centred ("linear", 2 nC - 1 fine entries).
// (TC, XC, YC) = coarse part dimensions, (XF, YF) = fine part dimensions of the two fast axes.
template <int PART, bool GEO>
__global__ void __launch_bounds__(256) k_restrict(int TC, int XC, int YC, int XF, int YF, const double* __restrict__ fine,
                                                  double* __restrict__ coarse)
{
    const i64 n = (i64)TC * XC * YC;
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int yc = (int)(i % YC);
    const i64 r = i / YC;
    const int xc = (int)(r % XC), tc = (int)(r / XC);
    auto at = [&](int t, int x, int y) {
        const double v = fine[((i64)t * XF + x) * YF + y];
        return GEO ? log(v) : v;
    };
    auto rt = [&](int x, int y) {
        auto f = [&](int t) { return at(t, x, y); };
        return PART == 0 ? restrict_nearest(tc, f) : restrict_linear(tc, TC, f);
    };
    auto rx = [&](int y) {
        auto f = [&](int x) { return rt(x, y); };
        return PART == 1 ? restrict_nearest(xc, f) : restrict_linear(xc, XC, f);
    };
    const double v = PART == 2 ? restrict_nearest(yc, rx) : restrict_linear(yc, YC, rx);
    coarse[i] = GEO ? exp(v) : v;
}

// finest level from the two planes: [1 ... 1 | WX on every node level | WY on every node level]
__global__ void __launch_bounds__(256) k_weight_from_planes(i64 L, i64 PX, i64 PY, int nt, const double* __restrict__ wx,
                                                            const double* __restrict__ wy, double* __restrict__ w)
{
    const i64 Q = L + (i64)nt * (PX + PY);
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < Q; i += stride) {
        double v = 1.0;
        if (i >= L) {
            const i64 j = i - L;
            v = j < (i64)nt * PX ? wx[j % PX] : wy[(j - (i64)nt * PX) % PY];
        }
        w[i] = v;
    }
}

// stage 1 of sum(log10(w + 1e-10)): fixed grid, fixed order
constexpr int LM_BLOCKS = 1024, LM_THREADS = 256;
__global__ void __launch_bounds__(LM_THREADS) k_logsum(i64 Q, const double* __restrict__ w, double* __restrict__ partial)
{
    double s[1] = {0.0};
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < Q; i += stride) s[0] += log10(w[i] + 1e-10);
    block_reduce_store<1, LM_THREADS>(s, partial, (i64)blockIdx.x);
}
__global__ void __launch_bounds__(32) k_logsum_total(const double* __restrict__ partial, int n, double* __restrict__ out)
{
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < n; i++) s += partial[i];
        out[0] = s;
    }
}

// One thread per element of the window [b, e) of a session's (pitched) staggered array: the value of the packed level array,
// or 1 in the pad entries (the level transfer divides whole ranges by the weight)
__global__ void __launch_bounds__(256) k_weight_scatter(Geo g, i64 b, i64 e, const double* __restrict__ packed, double* __restrict__ dst)
{
    const i64 i = b + (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= e) return;
    const i64 Lp = (i64)(g.nt - 1) * g.P, NBXp = (i64)g.nt * (g.nx - 1) * g.ny;
    double v = 1.0;
    if (i < g.L) {
        const i64 t = i / g.PC, p = i - t * g.PC;
        const int x = (int)(p / g.py), y = (int)(p - (i64)x * g.py);
        if (y < g.ny) v = packed[t * g.P + (i64)x * g.ny + y];
    } else if (i < g.L + g.NBX) {
        const i64 j = i - g.L, t = j / g.PBX, p = j - t * g.PBX;
        const int x = (int)(p / g.py), y = (int)(p - (i64)x * g.py);
        if (y < g.ny) v = packed[Lp + t * (i64)(g.nx - 1) * g.ny + (i64)x * g.ny + y];
    } else {
        const i64 j = i - g.L - g.NBX, t = j / g.PBY, p = j - t * g.PBY;
        const int x = (int)(p / g.pyb), y = (int)(p - (i64)x * g.pyb);
        if (y < g.ny - 1) v = packed[Lp + NBXp + t * (i64)g.nx * (g.ny - 1) + (i64)x * (g.ny - 1) + y];
    }
    dst[i] = v;
}

inline unsigned nblk(i64 n) { return (unsigned)((n + 255) / 256); }

}  // namespace

namespace dsocp {
void launch_weight_scatter(const Geo& g, i64 b, i64 e, const double* packed, double* dst, cudaStream_t st)
{
    if (e > b) k_weight_scatter<<<nblk(e - b), 256, 0, st>>>(g, b, e, packed, dst);
}
}  // namespace dsocp

dotsocp_weights::~dotsocp_weights()
{
    for (auto& l : lv) cudaFree(l.w);
    cudaFree(scratch);
}

static WeightLevel make_level(int nt, int nx, int ny)
{
    WeightLevel l;
    l.nt = nt; l.nx = nx; l.ny = ny;
    l.L = (i64)(nt - 1) * nx * ny;
    l.NBX = (i64)nt * (nx - 1) * ny;
    l.NBY = (i64)nt * nx * (ny - 1);
    l.Q = l.L + l.NBX + l.NBY;
    l.w = nullptr;
    return l;
}

extern "C" int dotsocp_weights_create(dotsocp_weights** out, int nt, int nx, int ny, int levels)
{
    if (!out) return set_err(DOTSOCP_EINVAL, "NULL argument");
    *out = nullptr;
    int rc = require_device();
    if (rc) return rc;
    if (levels < 1 || nt < 2 || nx < 2 || ny < 2) return set_err(DOTSOCP_EINVAL, "weights: levels >= 1 and a 2-D grid expected");
    dotsocp_weights* w = new (std::nothrow) dotsocp_weights;
    if (!w) return set_err(DOTSOCP_ENOMEM, "host allocation failed");
    for (int l = 0; l < levels; l++) {
        if (l > 0) {
            // downSample_q.m:6 halves every axis: (n+1)/2 nodes, which needs odd node counts >= 3
            if ((nt & 1) == 0 || (nx & 1) == 0 || (ny & 1) == 0 || nt < 3 || nx < 3 || ny < 3) {
                delete w;
                return set_err(DOTSOCP_EINVAL, "weights: level %d of a %d x %d x %d grid cannot be halved (odd node counts >= 3 needed)", l - 1, nt, nx, ny);
            }
            nt = (nt + 1) / 2; nx = (nx + 1) / 2; ny = (ny + 1) / 2;
        }
        w->lv.push_back(make_level(nt, nx, ny));
    }
    cudaGetDevice(&w->device);
    for (auto& l : w->lv) {
        cudaError_t e = cudaMalloc(&l.w, (size_t)l.Q * sizeof(double));
        if (e != cudaSuccess) {
            cudaGetLastError();
            const long long q_ = l.Q;
            delete w;
            return set_err(DOTSOCP_ENOMEM, "weights: %lld doubles: %s", q_, cudaGetErrorString(e));
        }
    }
    if (cudaMalloc(&w->scratch, (LM_BLOCKS + 1) * sizeof(double)) != cudaSuccess) {
        cudaGetLastError();
        delete w;
        return set_err(DOTSOCP_ENOMEM, "weights: scratch");
    }
    *out = w;
    return DOTSOCP_OK;
}

extern "C" void dotsocp_weights_destroy(dotsocp_weights* w) { delete w; }

extern "C" int dotsocp_weights_set(dotsocp_weights* w, const double* weight)
{
    if (!w || !weight) return set_err(DOTSOCP_EINVAL, "NULL argument");
    CU(cudaSetDevice(w->device));
    CU(cudaMemcpy(w->lv[0].w, weight, (size_t)w->lv[0].Q * sizeof(double), cudaMemcpyHostToDevice));
    w->filled = 1;
    return DOTSOCP_OK;
}

extern "C" int dotsocp_weights_set_planes(dotsocp_weights* w, const double* weightX, const double* weightY)
{
    if (!w || !weightX || !weightY) return set_err(DOTSOCP_EINVAL, "NULL argument");
    CU(cudaSetDevice(w->device));
    const WeightLevel& l = w->lv[0];
    const i64 PX = (i64)(l.nx - 1) * l.ny, PY = (i64)l.nx * (l.ny - 1);
    double* planes = nullptr;
    CU(cudaMalloc(&planes, (size_t)(PX + PY) * sizeof(double)));
    cudaError_t e = cudaMemcpy(planes, weightX, (size_t)PX * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(planes + PX, weightY, (size_t)PY * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        k_weight_from_planes<<<148 * 8, 256>>>(l.L, PX, PY, l.nt, planes, planes + PX, l.w);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaFree(planes);
    CU(e);
    w->launches += 1;
    w->filled = 1;
    return DOTSOCP_OK;
}

template <bool GEO>
static void restrict_level(const WeightLevel& f, const WeightLevel& c)
{
    // q0 part: (nt-1, nx, ny) ; bx part: (nt, nx-1, ny) ; by part: (nt, nx, ny-1)
    k_restrict<0, GEO><<<nblk(c.L), 256>>>(c.nt - 1, c.nx, c.ny, f.nx, f.ny, f.w, c.w);
    k_restrict<1, GEO><<<nblk(c.NBX), 256>>>(c.nt, c.nx - 1, c.ny, f.nx - 1, f.ny, f.w + f.L, c.w + c.L);
    k_restrict<2, GEO><<<nblk(c.NBY), 256>>>(c.nt, c.nx, c.ny - 1, f.nx, f.ny - 1, f.w + f.L + f.NBX, c.w + c.L + c.NBX);
}

extern "C" int dotsocp_weights_restrict(dotsocp_weights* w, int geometric)
{
    if (!w) return set_err(DOTSOCP_EINVAL, "NULL argument");
    if (w->filled < 1) return set_err(DOTSOCP_ESTATE, "weights: the finest level has not been set");
    CU(cudaSetDevice(w->device));
    for (size_t l = 1; l < w->lv.size(); l++) {
        if (geometric) restrict_level<true>(w->lv[l - 1], w->lv[l]);
        else restrict_level<false>(w->lv[l - 1], w->lv[l]);
        CU(cudaGetLastError());
        w->launches += 3;
    }
    CU(cudaDeviceSynchronize());
    w->filled = (int)w->lv.size();
    return DOTSOCP_OK;
}

static int check_level(const dotsocp_weights* w, int level)
{
    if (!w) return set_err(DOTSOCP_EINVAL, "NULL argument");
    if (level < 0 || level >= (int)w->lv.size()) return set_err(DOTSOCP_EINVAL, "weights: level %d of %d", level, (int)w->lv.size());
    if (level >= w->filled) return set_err(DOTSOCP_ESTATE, "weights: level %d has not been computed (set the finest level, then restrict)", level);
    return DOTSOCP_OK;
}

extern "C" int dotsocp_weights_get(const dotsocp_weights* w, int level, double* weight)
{
    int rc = check_level(w, level);
    if (rc) return rc;
    if (!weight) return set_err(DOTSOCP_EINVAL, "NULL argument");
    CU(cudaSetDevice(w->device));
    CU(cudaMemcpy(weight, w->lv[level].w, (size_t)w->lv[level].Q * sizeof(double), cudaMemcpyDeviceToHost));
    return DOTSOCP_OK;
}

extern "C" int dotsocp_weights_dims(const dotsocp_weights* w, int level, int* nt, int* nx, int* ny)
{
    if (!w) return set_err(DOTSOCP_EINVAL, "NULL argument");
    if (level < 0 || level >= (int)w->lv.size()) return set_err(DOTSOCP_EINVAL, "weights: level %d of %d", level, (int)w->lv.size());
    if (nt) *nt = w->lv[level].nt;
    if (nx) *nx = w->lv[level].nx;
    if (ny) *ny = w->lv[level].ny;
    return DOTSOCP_OK;
}

extern "C" int dotsocp_weights_log10_mean(dotsocp_weights* w, int level, double* mean)
{
    int rc = check_level(w, level);
    if (rc) return rc;
    if (!mean) return set_err(DOTSOCP_EINVAL, "NULL argument");
    CU(cudaSetDevice(w->device));
    const WeightLevel& l = w->lv[level];
    k_logsum<<<LM_BLOCKS, LM_THREADS>>>(l.Q, l.w, w->scratch);
    k_logsum_total<<<1, 32>>>(w->scratch, LM_BLOCKS, w->scratch + LM_BLOCKS);
    CU(cudaGetLastError());
    w->launches += 2;
    double s = 0.0;
    CU(cudaMemcpy(&s, w->scratch + LM_BLOCKS, sizeof(double), cudaMemcpyDeviceToHost));
    *mean = s / (double)l.Q;
    return DOTSOCP_OK;
}

extern "C" double dotsocp_weights_launch_count(const dotsocp_weights* w) { return w ? w->launches : 0.0; }
