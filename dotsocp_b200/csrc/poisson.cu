// poisson.cu -- Neumann Poisson solve  phi = IDCT3( DCT3(rhs) ./ (D^2 kernel) )  as batched per-axis transforms (sm_100a).
//
// Reference: oper_poisson3dim.m:4, initialize_FFTkernel.m:6-15, mirt_dctn.m:64-96, mirt_idctn.m:59-95 (orthonormal
// DCT-II per axis, implemented there as permute -> complex FFT -> twiddle).
//
// The grid lengths are n = 2^k + 1 (65, 129, 257, 513, 1025; 257 is prime), so no power-of-two FFT applies directly
// and a dense n x n contraction costs 2n flop per 16 B moved (compute-bound beyond n ~ 64 even on the FP64 tensor
// pipe).  Each 1-D DCT-II is therefore computed as
//     Makhoul permutation -> length-n complex DFT of TWO real lines packed as re/im
//                         -> Bluestein chirp-z: n-point DFT == circular convolution of length M,
//                            M = 2(n-1) = 2^(k+1)   (valid because the chirp is even, see dct_plan_create)
//                         -> two power-of-two FFTs of length M in shared memory (radix-8/4, digit-reversed middle)
// which is O(log n) flop per element for every n.  The t axis is done as ONE kernel: forward DCT, division by the
// eigenvalue table, inverse DCT, so the solve is 5 passes over N doubles (y, x, [t,/,t^-1], x^-1, y^-1).
// Lengths <= 32 (coarse multilevel grids) use a dense matrix kernel.
#include "kernels.h"
#include "errs.h"
#include "fft16.cuh"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace dsocp {

struct DctPlan {
    int n;
    int dense;         // 1: dense matrix path
    int log2m, M;
    double2* w;        // chirp  w[j] = exp(-i pi j^2 / n), j < n
    double2* bhat;     // FFT_M(b_circ)/M in the digit-reversed order produced by the forward passes
    double2* bhat16;   // same spectrum in the order of the register radix-16 passes (fft16.cuh), M >= 256 only
    double2* tw;       // exp(-2 pi i j / M), j < M
    double2* pw;       // c_k exp(-i pi k / (2n)), k < n   (forward post-twiddle incl. orthonormal scale)
    double2* ipw;      // exp(+i pi k / (2n)) / c_k / n     (inverse pre-twiddle incl. 1/n of the IDFT)
    double* cmat;      // dense: C[k*n + j] = c_k cos(pi (2j+1) k / (2n))
};

// ------------------------------------------------------------------------------------------------ host tables
static const long double PIl = 3.141592653589793238462643383279502884L;

static void radices_for(int log2m, int* r, int* nr)
{
    int n8 = 0, n4 = 0;
    switch (log2m % 3) {
        case 0: n8 = log2m / 3; break;
        case 1: n8 = (log2m - 4) / 3; n4 = 2; break;
        default: n8 = (log2m - 2) / 3; n4 = 1; break;
    }
    int k = 0;
    for (int i = 0; i < n8; i++) r[k++] = 8;
    for (int i = 0; i < n4; i++) r[k++] = 4;
    *nr = k;
}

static void host_fft(std::vector<long double>& re, std::vector<long double>& im)
{
    const size_t n = re.size();
    for (size_t i = 1, j = 0; i < n; i++) {
        size_t bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) { std::swap(re[i], re[j]); std::swap(im[i], im[j]); }
    }
    for (size_t len = 2; len <= n; len <<= 1) {
        for (size_t i = 0; i < n; i += len)
            for (size_t k = 0; k < len / 2; k++) {
                const long double ang = -2 * PIl * (long double)k / (long double)len;
                const long double wr = cosl(ang), wi = sinl(ang);
                const size_t a = i + k, b = i + k + len / 2;
                const long double xr = re[b] * wr - im[b] * wi, xi = re[b] * wi + im[b] * wr;
                re[b] = re[a] - xr; im[b] = im[a] - xi;
                re[a] += xr; im[a] += xi;
            }
    }
}

template <typename T>
static T* to_device(const std::vector<T>& v)
{
    T* d = nullptr;
    if (cudaMalloc(&d, v.size() * sizeof(T)) != cudaSuccess) return nullptr;
    // a copy from pageable memory runs on the legacy stream, which is NOT ordered against the sessions' non-blocking streams (and
    // may return before its last DMA has landed): wait for it here, the tables are read by kernels on other streams
    cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
    cudaStreamSynchronize(0);
    return d;
}

static DctPlan* dct_plan_create(int n)
{
    DctPlan* p = new DctPlan();
    memset(p, 0, sizeof(*p));
    p->n = n;
    const long double c0 = sqrtl(1.0L / n), c1 = sqrtl(2.0L / n);
    if (n <= 32) {
        p->dense = 1;
        std::vector<double> C((size_t)n * n);
        for (int k = 0; k < n; k++)
            for (int j = 0; j < n; j++) {
                // reduce the angle pi (2j+1) k / (2n) exactly modulo 2 pi: (2j+1) k mod 4n
                const long long a = ((long long)(2 * j + 1) * k) % (4LL * n);
                C[(size_t)k * n + j] = (double)((k == 0 ? c0 : c1) * cosl(PIl * (long double)a / (2.0L * n)));
            }
        p->cmat = to_device(C);
        return p;
    }
    // Bluestein length: the chirp b[m] = exp(+i pi m^2 / n) is even in m, so a circular convolution of length
    // M >= 2n-2 already reproduces the linear one (lags +(n-1) and -(n-1) share a slot and need the same value).
    int log2m = 1;
    while ((1 << log2m) < 2 * n - 2) log2m++;
    p->log2m = log2m;
    p->M = 1 << log2m;
    const int M = p->M;
    std::vector<double2> w(n), pw(n), ipw(n), tw(M), bh(M);
    std::vector<long double> wr(n), wi(n);
    for (int j = 0; j < n; j++) {
        const long long a = ((long long)j * j) % (2LL * n);   // j^2 mod 2n (exact)
        wr[j] = cosl(PIl * (long double)a / n);
        wi[j] = -sinl(PIl * (long double)a / n);
        w[j] = make_double2((double)wr[j], (double)wi[j]);
        const long double ck = (j == 0) ? c0 : c1;
        const long double ang = PIl * (long double)j / (2.0L * n);
        pw[j] = make_double2((double)(ck * cosl(ang)), (double)(-ck * sinl(ang)));
        ipw[j] = make_double2((double)(cosl(ang) / ck / n), (double)(sinl(ang) / ck / n));
    }
    for (int j = 0; j < M; j++) {
        const long double ang = -2 * PIl * (long double)j / M;
        tw[j] = make_double2((double)cosl(ang), (double)sinl(ang));
    }
    std::vector<long double> br(M, 0.0L), bi(M, 0.0L);
    for (int s = 0; s < n; s++) {
        br[s] = wr[s]; bi[s] = -wi[s];                       // b = conj(w)
        if (s > 0) { br[M - s] = wr[s]; bi[M - s] = -wi[s]; }
    }
    host_fft(br, bi);
    int rad[8], nr;
    radices_for(log2m, rad, &nr);
    for (int pos = 0; pos < M; pos++) {
        // position -> frequency of the mixed-radix DIF passes (see fft_pass): pos = sum m_i * M/(r_1..r_i),
        // freq = m_1 + r_1 (m_2 + r_2 (...))
        int rem = pos, span = M, f = 0, mult = 1;
        for (int i = 0; i < nr; i++) {
            span /= rad[i];
            const int m = rem / span;
            rem -= m * span;
            f += m * mult;
            mult *= rad[i];
        }
        bh[pos] = make_double2((double)(br[f] / M), (double)(bi[f] / M));
    }
    p->w = to_device(w);
    p->pw = to_device(pw);
    p->ipw = to_device(ipw);
    p->tw = to_device(tw);
    p->bhat = to_device(bh);
    if (log2m >= 8 && log2m <= 12) {
        std::vector<double2> b16(M);
        const int TP = M / 16, Q2 = M / 256;
        for (int pos = 0; pos < M; pos++) {
            const int m1 = pos / TP, r = pos % TP, m2 = r / Q2, m3 = r % Q2;
            const int f = m1 + 16 * m2 + 256 * m3;
            // thread j multiplies its registers e = 0..15 (positions 16j + e): stored as [e][j] so that a warp reads one
            // contiguous 512-byte run per e instead of 32 different cache lines
            b16[(size_t)(pos % 16) * TP + pos / 16] = make_double2((double)(br[f] / M), (double)(bi[f] / M));
        }
        p->bhat16 = to_device(b16);
    }
    return p;
}

static void dct_plan_destroy(DctPlan* p)
{
    if (!p) return;
    cudaFree(p->w); cudaFree(p->bhat); cudaFree(p->bhat16); cudaFree(p->tw); cudaFree(p->pw); cudaFree(p->ipw); cudaFree(p->cmat);
    delete p;
}

// ------------------------------------------------------------------------------------------------ device FFT
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cmul(double2 a, double2 b)
{
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cmulc(double2 a, double2 b)   // a * conj(b)
{
    return make_double2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
// multiply by -i (SIGN<0) or +i (SIGN>0)
template <int SIGN>
__device__ __forceinline__ double2 cmuli(double2 a)
{
    return SIGN < 0 ? make_double2(a.y, -a.x) : make_double2(-a.y, a.x);
}

template <int SIGN>
__device__ __forceinline__ void dft4(double2& a0, double2& a1, double2& a2, double2& a3)
{
    const double2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = cmuli<SIGN>(csub(a1, a3));
    a0 = cadd(t0, t2);
    a1 = cadd(t1, t3);
    a2 = csub(t0, t2);
    a3 = csub(t1, t3);
}

template <int SIGN>
__device__ __forceinline__ void dft8(double2 (&v)[8])
{
    double2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6];
    double2 o0 = v[1], o1 = v[3], o2 = v[5], o3 = v[7];
    dft4<SIGN>(e0, e1, e2, e3);
    dft4<SIGN>(o0, o1, o2, o3);
    const double h = 0.70710678118654752440;
    // w8^1 = (1 -+ i)/sqrt2, w8^2 = -+i, w8^3 = (-1 -+ i)/sqrt2
    const double2 t1 = SIGN < 0 ? make_double2((o1.x + o1.y) * h, (o1.y - o1.x) * h)
                                : make_double2((o1.x - o1.y) * h, (o1.y + o1.x) * h);
    const double2 t2 = cmuli<SIGN>(o2);
    const double2 t3 = SIGN < 0 ? make_double2((o3.y - o3.x) * h, (-o3.x - o3.y) * h)
                                : make_double2((-o3.x - o3.y) * h, (o3.x - o3.y) * h);
    v[0] = cadd(e0, o0); v[4] = csub(e0, o0);
    v[1] = cadd(e1, t1); v[5] = csub(e1, t1);
    v[2] = cadd(e2, t2); v[6] = csub(e2, t2);
    v[3] = cadd(e3, t3); v[7] = csub(e3, t3);
}

#define PADI(i) ((i) + ((i) >> 3))

// One radix-R pass over an M-point array in shared memory, TP = M/8 cooperating threads.
// forward (INV=false): decimation in frequency, butterfly then twiddle  -> digit-reversed spectrum after all passes
// inverse (INV=true) : the transposed graph, conjugate twiddle then butterfly -> natural order from digit-reversed input
template <int R, bool INV>
__device__ __forceinline__ void fft_pass(double2* __restrict__ s, int ltid, int TP, int M, int len,
                                         const double2* __restrict__ tw)
{
    const int q = len / R;
    const int nb = M / R;
    const int tstep = M / len;
    for (int jb = ltid; jb < nb; jb += TP) {
        const int b = jb / q, k = jb - b * q;
        const int base = b * len + k;
        if (R == 8) {
            double2 v[8];
#pragma unroll
            for (int m = 0; m < 8; m++) v[m] = s[PADI(base + m * q)];
            if (!INV) {
                dft8<-1>(v);
#pragma unroll
                for (int f = 1; f < 8; f++) v[f] = cmul(v[f], __ldg(&tw[f * k * tstep]));
            } else {
#pragma unroll
                for (int f = 1; f < 8; f++) v[f] = cmulc(v[f], __ldg(&tw[f * k * tstep]));
                dft8<1>(v);
            }
#pragma unroll
            for (int m = 0; m < 8; m++) s[PADI(base + m * q)] = v[m];
        } else {
            double2 a0 = s[PADI(base)], a1 = s[PADI(base + q)], a2 = s[PADI(base + 2 * q)], a3 = s[PADI(base + 3 * q)];
            if (!INV) {
                dft4<-1>(a0, a1, a2, a3);
                a1 = cmul(a1, __ldg(&tw[k * tstep]));
                a2 = cmul(a2, __ldg(&tw[2 * k * tstep]));
                a3 = cmul(a3, __ldg(&tw[3 * k * tstep]));
            } else {
                a1 = cmulc(a1, __ldg(&tw[k * tstep]));
                a2 = cmulc(a2, __ldg(&tw[2 * k * tstep]));
                a3 = cmulc(a3, __ldg(&tw[3 * k * tstep]));
                dft4<1>(a0, a1, a2, a3);
            }
            s[PADI(base)] = a0; s[PADI(base + q)] = a1; s[PADI(base + 2 * q)] = a2; s[PADI(base + 3 * q)] = a3;
        }
    }
    __syncthreads();
}

template <int LOG2M>
struct Radix {
    static constexpr int N8 = (LOG2M % 3 == 0) ? LOG2M / 3 : (LOG2M % 3 == 1) ? (LOG2M - 4) / 3 : (LOG2M - 2) / 3;
    static constexpr int N4 = (LOG2M % 3 == 0) ? 0 : (LOG2M % 3 == 1) ? 2 : 1;
};

// a <- circular convolution core: FFT_M (DIF), pointwise * bhat, IFFT_M (DIT).  Leaves the result in natural order.
template <int LOG2M>
__device__ __forceinline__ void conv_core(double2* __restrict__ s, int ltid, const double2* __restrict__ tw,
                                          const double2* __restrict__ bhat)
{
    constexpr int M = 1 << LOG2M, TP = M / 8;
    int len = M;
#pragma unroll
    for (int i = 0; i < Radix<LOG2M>::N8; i++) { fft_pass<8, false>(s, ltid, TP, M, len, tw); len >>= 3; }
#pragma unroll
    for (int i = 0; i < Radix<LOG2M>::N4; i++) { fft_pass<4, false>(s, ltid, TP, M, len, tw); len >>= 2; }
    for (int i = ltid; i < M; i += TP) s[PADI(i)] = cmul(s[PADI(i)], __ldg(&bhat[i]));
    __syncthreads();
    len = 1;
#pragma unroll
    for (int i = 0; i < Radix<LOG2M>::N4; i++) { len <<= 2; fft_pass<4, true>(s, ltid, TP, M, len, tw); }
#pragma unroll
    for (int i = 0; i < Radix<LOG2M>::N8; i++) { len <<= 3; fft_pass<8, true>(s, ltid, TP, M, len, tw); }
}

// ------------------------------------------------------------------------------------------------ batched DCT kernel
struct LineGeom {
    int n;             // transform length
    i64 estride;       // distance between consecutive elements of a line
    i64 gstride;       // distance between the first elements of adjacent lines of a group
    i64 inner;         // lines per outer index
    i64 ostride;       // distance between outer indices
    int contiguous;    // 1: lines are contiguous (estride == 1): flat staging index runs along the line
    // time-slab transposes fused into the x passes: element (t_loc = blockIdx.y, p = x*ny + y) lives in the packed
    // all-to-all buffer at  nlev*pcut[r] + t_loc*(pcut[r+1]-pcut[r]) + (p - pcut[r]),  pcut[r] = min(P, r*ceil(P/world))
    int rm_world;      // 0: no remap
    int rm_nlev;
    int rm_t0;         // first local level of this launch (the slab is processed in groups of levels)
    i64 rm_P;
    int rm_ny;
    // direct push (forward x pass only): rm_tab[r] = address of row 0 of this slab's rows inside the t-solve buffer of the
    // owner of chunk r (peer memory mapped through CUDA IPC, or a local buffer) -- the transform stores its result there
    double* const* rm_tab;
    unsigned n_inv;    // floor(2^32 / n) + 1: f / n == __umulhi(f, n_inv) for f < 2^32 / n (set by the launcher)
};

__device__ __forceinline__ i64 remap_index(const LineGeom& lg, int t_loc, i64 p)
{
    // chunks of C = ceil(P/world) modes (the last one shorter): one 32-bit division per element
    const unsigned C = (unsigned)((lg.rm_P + lg.rm_world - 1) / lg.rm_world);
    const unsigned r = (unsigned)p / C;
    const i64 c0 = (i64)r * C;
    const i64 ch = (lg.rm_P - c0) < (i64)C ? (lg.rm_P - c0) : (i64)C;
    return (i64)lg.rm_nlev * c0 + (i64)t_loc * ch + (p - c0);
}

struct ScaleArgs {     // MODE 2 (t axis): divide the spectrum by D2*((lamY[y] + lamX[x]) + lamT[k]), zero -> 1
    const double* lam_t;
    const double* lam_x;
    const double* lam_y;
    int ny;
    double D2;
    i64 line_offset;   // global index of the first line of the array (t-pass on a chunk of (x,y) modes)
};

// Makhoul index: real element i of the line -> position in v (x[2j] -> v[j], x[2j+1] -> v[n-1-j])
__device__ __forceinline__ int makhoul(int i, int n) { return (i & 1) ? (n - 1 - (i >> 1)) : (i >> 1); }

// MODE 0: forward DCT-II, 1: inverse, 2: forward -> ./kernel -> inverse
template <int LOG2M, int PAIRS, int MODE>
__global__ void __launch_bounds__((1 << LOG2M) / 8 * PAIRS)
k_dct_bluestein(LineGeom lg, const double* ain, double* aout, const double2* __restrict__ w, const double2* __restrict__ bhat,
                const double2* __restrict__ tw, const double2* __restrict__ pw, const double2* __restrict__ ipw,
                ScaleArgs sa)
{
    constexpr int M = 1 << LOG2M, TP = M / 8, G = 2 * PAIRS, NTHR = TP * PAIRS, PADLEN = M + M / 8;
    extern __shared__ double2 smem[];
    const int n = lg.n;
    const int tid = threadIdx.x;
    const int pair = tid / TP, ltid = tid - pair * TP;
    double2* __restrict__ s = smem + (size_t)pair * PADLEN;
    const i64 line0 = (i64)blockIdx.x * G;
    const i64 gbase = (i64)blockIdx.y * lg.ostride + line0 * lg.gstride;
    const int nlines = (int)((lg.inner - line0) < (i64)G ? (lg.inner - line0) : (i64)G);

    // ---- stage in: global -> shared -------------------------------------------------------------------------
    for (int i = tid; i < PAIRS * PADLEN; i += NTHR) smem[i] = make_double2(0.0, 0.0);
    __syncthreads();
    {
        const int total = G * n;
        for (int f = tid; f < total; f += NTHR) {
            int g, j;
            if (lg.contiguous) { g = f / n; j = f - g * n; } else { j = f / G; g = f - j * G; }
            if (g < nlines) {
                const double val = ain[gbase + (i64)g * lg.gstride + (i64)j * lg.estride];
                double* dst = reinterpret_cast<double*>(smem + (size_t)(g >> 1) * PADLEN +
                                                        PADI(MODE == 1 ? j : makhoul(j, n)));
                dst[g & 1] = val;
            }
        }
    }
    __syncthreads();

    if (MODE != 1) {
        // ---- forward: a[m] = (v1 + i v2)[m] * w[m]; V = w .* conv(a, b) ----------------------------------------
        for (int m = ltid; m < n; m += TP) s[PADI(m)] = cmul(s[PADI(m)], __ldg(&w[m]));
        __syncthreads();
        conv_core<LOG2M>(s, ltid, tw, bhat);
        for (int k = ltid; k < n; k += TP) s[PADI(k)] = cmul(s[PADI(k)], __ldg(&w[k]));
        __syncthreads();
        // separate the two real lines: V1 = (V[k] + conj V[n-k])/2, V2 = (V[k] - conj V[n-k])/(2i); X = Re(pw V)
        constexpr int KPT = 5;   // n <= M/2 + 1 = 4 TP + 1  => at most 5 k per thread
        double2 r[KPT];
#pragma unroll
        for (int cnt = 0; cnt < KPT; cnt++) {
            const int k = ltid + cnt * TP;
            if (k >= n) continue;
            const double2 vk = s[PADI(k)];
            const double2 vn = s[PADI(k == 0 ? 0 : n - k)];
            const double2 v1 = make_double2(0.5 * (vk.x + vn.x), 0.5 * (vk.y - vn.y));
            const double2 dd = make_double2(vk.x - vn.x, vk.y + vn.y);
            const double2 v2 = make_double2(0.5 * dd.y, -0.5 * dd.x);
            const double2 p = __ldg(&pw[k]);
            double x1 = p.x * v1.x - p.y * v1.y;
            double x2 = p.x * v2.x - p.y * v2.y;
            if (MODE == 2) {
                // spectral division, same operation order as D^2 * ((CY + CX) + CT) and rhs ./ kernel
                const i64 l1 = sa.line_offset + line0 + 2 * pair, l2 = l1 + 1;
                {
                    const int xx = (int)(l1 / sa.ny), yy = (int)(l1 - (i64)xx * sa.ny);
                    double kv = (l1 - sa.line_offset < lg.inner) ? __dadd_rn(__dadd_rn(sa.lam_y[yy], sa.lam_x[xx]), sa.lam_t[k]) : 1.0;
                    if (kv == 0.0) kv = 1.0;
                    x1 = x1 / __dmul_rn(sa.D2, kv);
                }
                {
                    const int xx = (int)(l2 / sa.ny), yy = (int)(l2 - (i64)xx * sa.ny);
                    double kv = (l2 - sa.line_offset < lg.inner) ? __dadd_rn(__dadd_rn(sa.lam_y[yy], sa.lam_x[xx]), sa.lam_t[k]) : 1.0;
                    if (kv == 0.0) kv = 1.0;
                    x2 = x2 / __dmul_rn(sa.D2, kv);
                }
            }
            r[cnt] = make_double2(x1, x2);
        }
        __syncthreads();
#pragma unroll
        for (int cnt = 0; cnt < KPT; cnt++) {
            const int k = ltid + cnt * TP;
            if (k < n) s[PADI(k)] = r[cnt];
        }
        if (MODE == 2)
            for (int m = n + ltid; m < M; m += TP) s[PADI(m)] = make_double2(0.0, 0.0);
        __syncthreads();
    }
    if (MODE != 0) {
        // ---- inverse: s[k] = (X1[k], X2[k]) natural order.  V[k] = g_k [(X1[k]+X2[n-k]) + i (X2[k]-X1[n-k])], k>=1;
        //      v = IDFT(V) = conj(DFT(conj V))/n through the same forward core. -------------------------------------
        constexpr int KPT = 5;
        double2 r[KPT];
#pragma unroll
        for (int cnt = 0; cnt < KPT; cnt++) {
            const int k = ltid + cnt * TP;
            if (k >= n) continue;
            const double2 xk = s[PADI(k)];
            double2 V;
            const double2 gk = __ldg(&ipw[k]);
            if (k == 0) {
                V = make_double2(xk.x * gk.x, xk.y * gk.x);      // ipw[0] is real: 1/(c0 n)
            } else {
                const double2 xn = s[PADI(n - k)];
                V = cmul(gk, make_double2(xk.x + xn.y, xk.y - xn.x));
            }
            // conj(V) * w[k]
            r[cnt] = cmul(make_double2(V.x, -V.y), __ldg(&w[k]));
        }
        __syncthreads();
#pragma unroll
        for (int cnt = 0; cnt < KPT; cnt++) {
            const int k = ltid + cnt * TP;
            if (k < n) s[PADI(k)] = r[cnt];
        }
        __syncthreads();
        conv_core<LOG2M>(s, ltid, tw, bhat);
        // v[j] = conj(w[j] c[j]) ; (1/n folded into ipw)
        for (int j = ltid; j < n; j += TP) {
            const double2 d = cmul(s[PADI(j)], __ldg(&w[j]));
            s[PADI(j)] = make_double2(d.x, -d.y);
        }
        __syncthreads();
    }

    // ---- stage out: shared -> global ------------------------------------------------------------------------------
    {
        const int total = G * n;
        for (int f = tid; f < total; f += NTHR) {
            int g, j;
            if (lg.contiguous) { g = f / n; j = f - g * n; } else { j = f / G; g = f - j * G; }
            if (g < nlines) {
                const double* src = reinterpret_cast<const double*>(smem + (size_t)(g >> 1) * PADLEN +
                                                                    PADI(MODE == 0 ? j : makhoul(j, n)));
                aout[gbase + (i64)g * lg.gstride + (i64)j * lg.estride] = src[g & 1];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ v2: register radix-16
// Same mathematics as k_dct_bluestein, but the two M-point FFTs of the chirp-z convolution run on TP = M/16 threads with
// 16 complex values per thread in registers (fft16.cuh): 4 shared-memory exchanges per convolution instead of ~14 passes,
// twiddles by a multiplication ladder instead of table loads, the chirp and the chirp spectrum applied in registers.
// barrier among the TP threads that share one line pair (named barrier 1 + pair; whole warps only)
template <int TP>
__device__ __forceinline__ void pair_sync(int pair)
{
    if (TP <= 32) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(pair + 1), "r"(TP) : "memory");
}

template <int LOG2M>
__device__ __forceinline__ void conv16(double2 (&v)[16], int j, int pair, double2* __restrict__ s,
                                       const double2* __restrict__ tw, const double2* __restrict__ bhat16)
{
    typedef F16<LOG2M> F;
#define __syncthreads() pair_sync<F::TP>(pair)
    F::fwd1(v, j, tw);
    F::store1(s, j, v);
    __syncthreads();
    F::load2(s, j, v);
    F::fwd2(v, j, tw);
    F::store2(s, j, v);
    __syncthreads();
    F::load3(s, j, v);
    F::fwd3(v);
#pragma unroll
    for (int e = 0; e < 16; e++) v[e] = c_mul(v[e], __ldg(&bhat16[e * F::TP + j]));
    F::inv3(v);
    F::store3(s, j, v);
    __syncthreads();
    F::load2(s, j, v);
    F::inv2(v, j, tw);
    F::store2(s, j, v);
    __syncthreads();
    F::load1(s, j, v);
    F::inv1(v, j, tw);
#undef __syncthreads
}

template <int LOG2M, int PAIRS, int MODE, bool CONTIG>
__global__ void __launch_bounds__((1 << LOG2M) / 16 * PAIRS, (LOG2M >= 12 ? 1 : 2))
k_dct_blu16(LineGeom lg, const double* ain, double* aout, const double2* __restrict__ w, const double2* __restrict__ bhat16,
            const double2* __restrict__ tw, const double2* __restrict__ pw, const double2* __restrict__ ipw, ScaleArgs sa)
{
    typedef F16<LOG2M> F;
    constexpr int M = F::M, TP = F::TP, G = 2 * PAIRS, NTHR = TP * PAIRS, PADLEN = M + M / 16 + 1;
    constexpr int NM = 9;    // n <= M/2 + 1  =>  positions j + m*TP < n only for m <= 8
    extern __shared__ double2 smem[];
    const int n = lg.n;
    const int tid = threadIdx.x;
    const int pair = tid / TP, j = tid - pair * TP;
    double2* __restrict__ s = smem + (size_t)pair * PADLEN;
    const i64 line0 = (i64)blockIdx.x * G;
    const i64 gbase = (i64)blockIdx.y * lg.ostride + line0 * lg.gstride;
    const int nlines = (int)((lg.inner - line0) < (i64)G ? (lg.inner - line0) : (i64)G);

    // ---- stage in ---------------------------------------------------------------------------------------------------
    // element f of the CTA's G x n block: CONTIG (lines contiguous, gstride == n): f = g*n + jj is also the global offset;
    // otherwise (lines strided, G adjacent lines contiguous, gstride == 1): f = jj*G + g.  All offsets relative to the
    // CTA's first element fit 32 bits (N < 2^31 is checked by the launcher).
    const double* __restrict__ in_b = ain + gbase;
    double* __restrict__ out_b = aout + gbase;
    const unsigned es = (unsigned)lg.estride;
    const unsigned total = (unsigned)(G * n);
    {
        // all global loads of a thread are issued before the first shared store (G*n/NTHR <= 2*(M/2+1)*PAIRS/NTHR = 16.x)
        constexpr int NLD = (G * (M / 2 + 1) + NTHR - 1) / NTHR;
        double vals[NLD];
#pragma unroll
        for (int u = 0; u < NLD; u++) {
            const unsigned f = tid + u * NTHR;
            unsigned g, jj;
            if (CONTIG) { g = __umulhi(f, lg.n_inv); jj = f - g * n; } else { jj = f / G; g = f % G; }
            if (f < total && (int)g < nlines) {
                if (MODE == 1 && lg.rm_world)   // x-inverse of a slab: read straight from the packed all-to-all buffer
                    vals[u] = ain[remap_index(lg, lg.rm_t0 + blockIdx.y, (i64)jj * lg.rm_ny + (line0 + g))];
                else
                    vals[u] = in_b[CONTIG ? f : jj * es + g];
            } else {
                vals[u] = 0.0;
            }
        }
#pragma unroll
        for (int u = 0; u < NLD; u++) {
            const unsigned f = tid + u * NTHR;
            if (f < total) {
                unsigned g, jj;
                if (CONTIG) { g = __umulhi(f, lg.n_inv); jj = f - g * n; } else { jj = f / G; g = f % G; }
                double* dst = reinterpret_cast<double*>(smem + (g >> 1) * PADLEN + PAD16(MODE == 1 ? (int)jj : makhoul((int)jj, n)));
                dst[g & 1] = vals[u];
            }
        }
    }
    __syncthreads();
    double2 v[16];
    if (MODE != 1) {
        // a[m] = (v1 + i v2)[m] * w[m], zero beyond n
#pragma unroll
        for (int m = 0; m < 16; m++) {
            const int pos = j + m * TP;
            v[m] = (m < NM && pos < n) ? c_mul(s[PAD16(pos)], __ldg(&w[pos])) : make_double2(0.0, 0.0);
        }
        conv16<LOG2M>(v, j, pair, s, tw, bhat16);
        pair_sync<TP>(pair);
        // V[k] = w[k] c[k], natural positions
#pragma unroll
        for (int m = 0; m < NM; m++) {
            const int pos = j + m * TP;
            if (pos < n) { v[m] = c_mul(v[m], __ldg(&w[pos])); s[PAD16(pos)] = v[m]; }
        }
        pair_sync<TP>(pair);
        double2 pr[NM];
#pragma unroll
        for (int m = 0; m < NM; m++) {
            const int pos = j + m * TP;
            if (pos < n) pr[m] = s[PAD16(pos == 0 ? 0 : n - pos)];
        }
        pair_sync<TP>(pair);
#pragma unroll
        for (int m = 0; m < NM; m++) {
            const int k = j + m * TP;
            if (k < n) {
            const double2 vk = v[m], vn = pr[m];
            const double2 v1 = make_double2(0.5 * (vk.x + vn.x), 0.5 * (vk.y - vn.y));
            const double2 dd = make_double2(vk.x - vn.x, vk.y + vn.y);
            const double2 v2 = make_double2(0.5 * dd.y, -0.5 * dd.x);
            const double2 p = __ldg(&pw[k]);
            double x1 = p.x * v1.x - p.y * v1.y;
            double x2 = p.x * v2.x - p.y * v2.y;
            s[PAD16(k)] = make_double2(x1, x2);
            }
        }
        if (MODE == 2) {
            // spectral division rhs ./ (D^2 * ((CY + CX) + CT)) on the thread's own entries; kept out of the unrolled
            // register-resident section (FP64 division needs many temporaries)
            const i64 l1 = sa.line_offset + line0 + 2 * pair, l2 = l1 + 1;
            const int xx1 = (int)(l1 / sa.ny), yy1 = (int)(l1 - (i64)xx1 * sa.ny);
            const int xx2 = (int)(l2 / sa.ny), yy2 = (int)(l2 - (i64)xx2 * sa.ny);
            const bool ok1 = l1 - sa.line_offset < lg.inner, ok2 = l2 - sa.line_offset < lg.inner;
            const double lxy1 = ok1 ? __dadd_rn(sa.lam_y[yy1], sa.lam_x[xx1]) : 1.0;
            const double lxy2 = ok2 ? __dadd_rn(sa.lam_y[yy2], sa.lam_x[xx2]) : 1.0;
#pragma unroll 1
            for (int m = 0; m < NM; m++) {
                const int k = j + m * TP;
                if (k < n) {
                    double2 x = s[PAD16(k)];
                    const double lt = sa.lam_t[k];
                    double k1 = ok1 ? __dadd_rn(lxy1, lt) : 1.0, k2 = ok2 ? __dadd_rn(lxy2, lt) : 1.0;
                    if (k1 == 0.0) k1 = 1.0;
                    if (k2 == 0.0) k2 = 1.0;
                    x.x = x.x / __dmul_rn(sa.D2, k1);
                    x.y = x.y / __dmul_rn(sa.D2, k2);
                    s[PAD16(k)] = x;
                }
            }
        }
        pair_sync<TP>(pair);
    }
    if (MODE != 0) {
        // s[k] = (X1[k], X2[k]);  V[k] = g_k [(X1[k]+X2[n-k]) + i (X2[k]-X1[n-k])];  input of the core = conj(V) w
#pragma unroll
        for (int m = 0; m < 16; m++) {
            const int k = j + m * TP;
            if (m < NM && k < n) {
                const double2 xk = s[PAD16(k)];
                const double2 gk = __ldg(&ipw[k]);
                double2 V;
                if (k == 0) {
                    V = make_double2(xk.x * gk.x, xk.y * gk.x);
                } else {
                    const double2 xn = s[PAD16(n - k)];
                    V = c_mul(gk, make_double2(xk.x + xn.y, xk.y - xn.x));
                }
                v[m] = c_mul(c_conj(V), __ldg(&w[k]));
            } else {
                v[m] = make_double2(0.0, 0.0);
            }
        }
        pair_sync<TP>(pair);
        {
            // MODE 2 runs the convolution core twice: hide the table pointers from CSE, otherwise the compiler keeps the
            // 46 twiddle / spectrum values of the first call alive in local memory (1 KB of spills per thread)
            const double2* tw_b = tw;
            const double2* bh_b = bhat16;
            if (MODE == 2) { asm volatile("" : "+l"(tw_b)); asm volatile("" : "+l"(bh_b)); }
            conv16<LOG2M>(v, j, pair, s, tw_b, bh_b);
        }
        pair_sync<TP>(pair);
#pragma unroll
        for (int m = 0; m < NM; m++) {
            const int pos = j + m * TP;
            if (pos < n) {
                const double2 d = c_mul(v[m], __ldg(&w[pos]));
                s[PAD16(pos)] = make_double2(d.x, -d.y);
            }
        }
        pair_sync<TP>(pair);
    }
    __syncthreads();
    // ---- stage out --------------------------------------------------------------------------------------------------
    {
        for (unsigned f = tid; f < total; f += NTHR) {
            unsigned g, jj;
            if (CONTIG) { g = __umulhi(f, lg.n_inv); jj = f - g * n; } else { jj = f / G; g = f % G; }
            if ((int)g < nlines) {
                const double* src = reinterpret_cast<const double*>(smem + (g >> 1) * PADLEN + PAD16(MODE == 0 ? (int)jj : makhoul((int)jj, n)));
                if (MODE == 0 && lg.rm_world) {   // x-forward of a slab: write straight into the packed all-to-all buffer
                    const i64 pp = (i64)jj * lg.rm_ny + (line0 + g);
                    if (lg.rm_tab) {              // ... or straight into the t-solve buffer of the chunk's owner
                        const unsigned C = (unsigned)((lg.rm_P + lg.rm_world - 1) / lg.rm_world);
                        const unsigned r = (unsigned)pp / C;
                        const i64 c0 = (i64)r * C;
                        const i64 ch = (lg.rm_P - c0) < (i64)C ? (lg.rm_P - c0) : (i64)C;
                        lg.rm_tab[r][(i64)(lg.rm_t0 + blockIdx.y) * ch + (pp - c0)] = src[g & 1];
                    } else
                        aout[remap_index(lg, lg.rm_t0 + blockIdx.y, pp)] = src[g & 1];
                } else
                    out_b[CONTIG ? f : jj * es + g] = src[g & 1];
            }
        }
    }
}

template <int LOG2M, int PAIRS>
static int launch_blu16(const DctPlan* p, const LineGeom& lg, i64 outer, const double* ain, double* a, int mode,
                        const ScaleArgs& sa, cudaStream_t st)
{
    constexpr int M = 1 << LOG2M, TP = M / 16, G = 2 * PAIRS, NTHR = TP * PAIRS;
    const size_t smem = (size_t)PAIRS * (M + M / 16 + 1) * sizeof(double2);
    dim3 grid((unsigned)((lg.inner + G - 1) / G), (unsigned)outer);
    static bool attr_set[3][2] = {{false, false}, {false, false}, {false, false}};
    // the kernel's two line layouts (see its stage-in comment)
    const bool contig = lg.contiguous != 0;
    if ((contig ? (lg.estride != 1 || lg.gstride != lg.n) : (lg.gstride != 1)) ||
        (i64)lg.n * lg.estride + (i64)G * lg.gstride >= ((i64)1 << 31)) {
        return dsocp_set_err(-1, "unsupported line geometry for the register-FFT DCT kernel (n = %d, element stride %lld, line stride "
                             "%lld): offsets inside one CTA must stay below 2^31", lg.n, (long long)lg.estride, (long long)lg.gstride);
    }
    LineGeom lgn = lg;
    lgn.n_inv = (unsigned)((((unsigned long long)1 << 32) / (unsigned)lg.n) + 1);
#define BLU(MODE, CT)                                                                                              \
    {                                                                                                              \
        if (!attr_set[MODE][CT]) {                                                                                 \
            cudaFuncSetAttribute(k_dct_blu16<LOG2M, PAIRS, MODE, CT>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                 (int)smem);                                                                       \
            attr_set[MODE][CT] = true;                                                                             \
        }                                                                                                          \
        k_dct_blu16<LOG2M, PAIRS, MODE, CT><<<grid, NTHR, smem, st>>>(lgn, ain, a, p->w, p->bhat16, p->tw, p->pw, p->ipw, sa); \
    }
    if (contig) { if (mode == 0) BLU(0, true) else if (mode == 1) BLU(1, true) else BLU(2, true) }
    else { if (mode == 0) BLU(0, false) else if (mode == 1) BLU(1, false) else BLU(2, false) }
#undef BLU
    return 0;
}

// ------------------------------------------------------------------------------------------------ dense small-n kernel
// One CTA = G lines staged in shared memory; thread (g,k) computes one output.  MODE as above.
template <int MODE>
__global__ void __launch_bounds__(256) k_dct_dense(LineGeom lg, int G, const double* ain, double* aout,
                                                   const double* __restrict__ cmat, ScaleArgs sa)
{
    extern __shared__ double sd[];   // [2][G][n]
    const int n = lg.n;
    double* buf0 = sd;
    double* buf1 = sd + (size_t)G * n;
    const i64 line0 = (i64)blockIdx.x * G;
    const i64 gbase = (i64)blockIdx.y * lg.ostride + line0 * lg.gstride;
    const int nlines = (int)((lg.inner - line0) < (i64)G ? (lg.inner - line0) : (i64)G);
    const int total = G * n;
    for (int f = threadIdx.x; f < total; f += blockDim.x) {
        int g, j;
        if (lg.contiguous) { g = f / n; j = f - g * n; } else { j = f / G; g = f - j * G; }
        buf0[g * n + j] = (g < nlines) ? ain[gbase + (i64)g * lg.gstride + (i64)j * lg.estride] : 0.0;
    }
    __syncthreads();
    if (MODE != 1) {
        for (int f = threadIdx.x; f < total; f += blockDim.x) {
            const int g = f / n, k = f - g * n;
            double acc = 0.0;
            for (int j = 0; j < n; j++) acc += __ldg(&cmat[k * n + j]) * buf0[g * n + j];
            if (MODE == 2) {
                const i64 l = sa.line_offset + line0 + g;
                const int xx = (int)(l / sa.ny), yy = (int)(l - (i64)xx * sa.ny);
                double kv = (l - sa.line_offset < lg.inner) ? __dadd_rn(__dadd_rn(sa.lam_y[yy], sa.lam_x[xx]), sa.lam_t[k]) : 1.0;
                if (kv == 0.0) kv = 1.0;
                acc = acc / __dmul_rn(sa.D2, kv);
            }
            buf1[g * n + k] = acc;
        }
        __syncthreads();
    }
    if (MODE != 0) {
        double* src = (MODE == 1) ? buf0 : buf1;
        double* dst = (MODE == 1) ? buf1 : buf0;
        for (int f = threadIdx.x; f < total; f += blockDim.x) {
            const int g = f / n, j = f - g * n;
            double acc = 0.0;
            for (int k = 0; k < n; k++) acc += __ldg(&cmat[k * n + j]) * src[g * n + k];
            dst[g * n + j] = acc;
        }
        __syncthreads();
    }
    const double* res = (MODE == 2) ? buf0 : buf1;
    for (int f = threadIdx.x; f < total; f += blockDim.x) {
        int g, j;
        if (lg.contiguous) { g = f / n; j = f - g * n; } else { j = f / G; g = f - j * G; }
        if (g < nlines) aout[gbase + (i64)g * lg.gstride + (i64)j * lg.estride] = res[g * n + j];
    }
}

// ------------------------------------------------------------------------------------------------ launch
template <int LOG2M, int PAIRS>
static void launch_blu(const DctPlan* p, const LineGeom& lg, i64 outer, const double* ain, double* a, int mode,
                       const ScaleArgs& sa, cudaStream_t st)
{
    constexpr int M = 1 << LOG2M, TP = M / 8, G = 2 * PAIRS, NTHR = TP * PAIRS;
    const size_t smem = (size_t)PAIRS * (M + M / 8) * sizeof(double2);
    dim3 grid((unsigned)((lg.inner + G - 1) / G), (unsigned)outer);
    static bool attr_set[3] = {false, false, false};
#define BLU(MODE)                                                                                                     \
    {                                                                                                                 \
        if (!attr_set[MODE]) {                                                                                        \
            cudaFuncSetAttribute(k_dct_bluestein<LOG2M, PAIRS, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                 (int)smem);                                                                          \
            attr_set[MODE] = true;                                                                                    \
        }                                                                                                             \
        k_dct_bluestein<LOG2M, PAIRS, MODE><<<grid, NTHR, smem, st>>>(lg, ain, a, p->w, p->bhat, p->tw, p->pw, p->ipw, sa); \
    }
    if (mode == 0) BLU(0) else if (mode == 1) BLU(1) else BLU(2)
#undef BLU
}

static int launch_dct_axis(const DctPlan* p, LineGeom lg, i64 outer, const double* ain, double* a, int mode,
                           const ScaleArgs& sa, cudaStream_t st)
{
    if (p->dense) {
        const int n = p->n;
        int G = 256 / n;
        if (G < 1) G = 1;
        if (!lg.contiguous && G < 4) G = 4;
        if (G > 64) G = 64;
        const size_t smem = (size_t)2 * G * n * sizeof(double);
        dim3 grid((unsigned)((lg.inner + G - 1) / G), (unsigned)outer);
        if (mode == 0) k_dct_dense<0><<<grid, 256, smem, st>>>(lg, G, ain, a, p->cmat, sa);
        else if (mode == 1) k_dct_dense<1><<<grid, 256, smem, st>>>(lg, G, ain, a, p->cmat, sa);
        else k_dct_dense<2><<<grid, 256, smem, st>>>(lg, G, ain, a, p->cmat, sa);
        return 0;
    }
    switch (p->log2m) {
        case 6: launch_blu<6, 32>(p, lg, outer, ain, a, mode, sa, st); return 0;
        case 7: launch_blu<7, 16>(p, lg, outer, ain, a, mode, sa, st); return 0;
        case 8: return launch_blu16<8, 16>(p, lg, outer, ain, a, mode, sa, st);
        case 9: return launch_blu16<9, 8>(p, lg, outer, ain, a, mode, sa, st);
        case 10: return launch_blu16<10, 4>(p, lg, outer, ain, a, mode, sa, st);
        case 11: return launch_blu16<11, 2>(p, lg, outer, ain, a, mode, sa, st);
        case 12: return launch_blu16<12, 2>(p, lg, outer, ain, a, mode, sa, st);
        default: return dsocp_set_err(-1, "unsupported transform length %d (Bluestein length 2^%d): supported up to 2^12", p->n, p->log2m);
    }
}

// ------------------------------------------------------------------------------------------------ t-solve without transforms
// After the (y,x) transforms every (kx,ky) mode is an independent tridiagonal system along t:
//     D^2 ( (nt-1)^2 T + (CY[ky]+CX[kx]) I ) phi = rhs ,   T = tridiag(-1, 2, -1) with 1 in the two corners,
// exactly the operator whose eigenvalues the reference divides by (initialize_FFTkernel.m:6-15).  It is solved by the
// Thomas algorithm with the pivot reciprocals g_t(mode) tabulated once per plan (they depend only on the grid), which
// replaces two of the six O(n log n) transforms of the solve by 6N doubles of streaming traffic.  The singular mode
// kx = ky = 0 keeps the reference's convention (zero eigenvalue := 1) through a dense DCT of that single line.
__global__ void __launch_bounds__(256) k_thomas_table(int nt, int ny, i64 lines, i64 stride, i64 p0, const double* __restrict__ lam_x,
                                                      const double* __restrict__ lam_y, double* __restrict__ gtab,
                                                      int* __restrict__ t_fix)
{
    const i64 l = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    if (l >= lines) return;
    const i64 p = p0 + l;
    const int kx = (int)(p / ny), ky = (int)(p - (i64)kx * ny);
    const double ct = (double)(nt - 1) * (double)(nt - 1);
    const double nu = (lam_y[ky] + lam_x[kx]) / ct;
    double g = 1.0 / (1.0 + nu);
    gtab[l] = g;
    int fix = nt;   // g_t == g_{t-1} => every later interior level repeats it (same expression, same inputs)
    for (int t = 1; t < nt; t++) {
        const double diag = (t == nt - 1 ? 1.0 : 2.0) + nu;
        const double gn = 1.0 / (diag - g);
        if (fix == nt && t < nt - 1 && gn == g) fix = t;
        g = gn;
        gtab[(i64)t * stride + l] = g;
    }
    t_fix[l] = fix;
}

// PUSH: the solution of time level t is stored in the buffer of the slab that owns the level (obase[owner] = first owned
// row of this mode chunk there, tcut = first level of every slab) -- peer memory over NVLink -- instead of in place
template <bool PUSH>
__global__ void __launch_bounds__(256) k_thomas(int nt, i64 lines, i64 stride, i64 p0, double inv_scale, const double* __restrict__ gtab,
                                                const int* __restrict__ t_fix, double* __restrict__ a,
                                                double* const* __restrict__ obase, const int* __restrict__ tcut, int world)
{
    const i64 l = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    if (l >= lines) return;
    // table reads only for t < tf and for the last level; in between g_t is the fixed point gs (bit-identical to the table)
    const int tf = t_fix[l];
    const double gs = tf < nt ? gtab[(i64)tf * stride + l] : 0.0;
    auto gt = [&](int t) { return (t < tf || t == nt - 1) ? gtab[(i64)t * stride + l] : gs; };
    int ow = world - 1, tlo = 0;
    double* ob = nullptr;
    if (PUSH) { tlo = tcut[ow]; ob = obase[ow] + l; }
    auto put = [&](int t, double v) {
        if (PUSH) {
            while (t < tlo) { ow--; tlo = tcut[ow]; ob = obase[ow] + l; }
            ob[(i64)(t - tlo) * lines] = v;
        } else
            a[(i64)t * stride + l] = v;
    };
    if (p0 + l == 0) {      // mode (0,0) is handled by k_tline0 (in place); only forward its result
        if (PUSH)
            for (int t = nt - 1; t >= 0; t--) put(t, a[(i64)t * stride]);
        return;
    }
    double d = 0.0;
    int t = 0;
    // forward elimination, 4 time levels per trip so that the loads of a trip are in flight together
    for (; t + 4 <= nt; t += 4) {
        double r[4], g[4];
#pragma unroll
        for (int u = 0; u < 4; u++) { r[u] = a[(i64)(t + u) * stride + l]; g[u] = gt(t + u); }
#pragma unroll
        for (int u = 0; u < 4; u++) { d = __dmul_rn(__fma_rn(r[u], inv_scale, d), g[u]); a[(i64)(t + u) * stride + l] = d; }
    }
    for (; t < nt; t++) { d = __dmul_rn(__fma_rn(a[(i64)t * stride + l], inv_scale, d), gt(t)); a[(i64)t * stride + l] = d; }
    // back substitution
    double x = d;
    if (PUSH) put(nt - 1, x);
    t = nt - 2;
    for (; t - 3 >= 0; t -= 4) {
        double dd[4], g[4];
#pragma unroll
        for (int u = 0; u < 4; u++) { dd[u] = a[(i64)(t - u) * stride + l]; g[u] = gt(t - u); }
#pragma unroll
        for (int u = 0; u < 4; u++) { x = __fma_rn(g[u], x, dd[u]); put(t - u, x); }
    }
    for (; t >= 0; t--) { x = __fma_rn(gt(t), x, a[(i64)t * stride + l]); put(t, x); }
}

// ---- the same solve with the time axis cut into slabs ("pipelined Thomas") -----------------------------------------------
// A slab holds levels [t0, t1) of ALL modes in the natural layout a[t*P + mode] (the output of the (y,x) transforms, in place).
// The forward elimination needs d of level t0-1 from the slab below, the back substitution x of level t1 from the slab above:
// one plane of P doubles per slab boundary and direction instead of two all-to-all transposes of the whole array.  Every mode
// runs through exactly the operations of k_thomas in the same order, so the result is bit-identical to the single-slab solve.
// The sweeps are sequential across slabs; cutting the modes into chunks (solver.cu) lets slab r work on chunk c while slab
// r+1 works on chunk c-1.  Pivot table: rows t0..t1-1 only (gtab[(t-t0)*P + mode]); t_fix as for k_thomas.
__global__ void __launch_bounds__(256) k_thomas_table_slab(int nt, int ny, i64 P, int t0, int t1, const double* __restrict__ lam_x,
                                                           const double* __restrict__ lam_y, double* __restrict__ gtab,
                                                           int* __restrict__ t_fix)
{
    const i64 p = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    if (p >= P) return;
    const int kx = (int)(p / ny), ky = (int)(p - (i64)kx * ny);
    const double ct = (double)(nt - 1) * (double)(nt - 1);
    const double nu = (lam_y[ky] + lam_x[kx]) / ct;
    double g = 1.0 / (1.0 + nu);
    if (t0 == 0) gtab[p] = g;
    int fix = nt;
    for (int t = 1; t < nt; t++) {
        const double diag = (t == nt - 1 ? 1.0 : 2.0) + nu;
        const double gn = 1.0 / (diag - g);
        if (fix == nt && t < nt - 1 && gn == g) fix = t;
        g = gn;
        if (t >= t0 && t < t1) gtab[(i64)(t - t0) * P + p] = g;
    }
    t_fix[p] = fix;
}

// value of g_t for a mode: table below t_fix and on the last level, else the fixed point (gs = g at level tf, which a slab may
// not hold: it is recomputed by iterating the recurrence to the fixed point -- the same expression, hence the same bits)
struct SlabG {
    const double* tab;
    i64 P, p;
    int t0, nt, tf;
    double gs;
    __device__ __forceinline__ double at(int t) const { return (t < tf || t == nt - 1) ? tab[(i64)(t - t0) * P + p] : gs; }
};
__device__ __forceinline__ double thomas_fixed_point(int nt, int tf, double nu)
{
    double g = 1.0 / (1.0 + nu);
    for (int t = 1; t <= tf && t < nt; t++) g = 1.0 / (2.0 + nu - g);
    return g;
}

// Hand-off between the GPUs of neighbouring slabs without NCCL (SlabSync, kernels.h): the producer's kernel stores its carry
// plane straight into the consumer's buffer (peer memory over NVLink) and, once all its CTAs have done so, publishes the solve's
// epoch in the consumer's flag; the consumer's kernel -- already launched -- spins on that flag before it reads the plane.
__device__ __forceinline__ void slab_wait(const SlabSync& sy)
{
    if (sy.wait_flag != nullptr) {
        if (threadIdx.x == 0) {
            int v;
            const long long t_begin = clock64();
            do {
                asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(sy.wait_flag) : "memory");
                if (v < sy.wait_value && clock64() - t_begin > 8000000000LL) {    // ~4 s: never hang the GPU on a lost signal
                    if (blockIdx.x == 0) printf("[dotsocp] slab hand-off timed out: flag %p holds %d, waiting for %d\n", (const void*)sy.wait_flag, v, sy.wait_value);
                    break;
                }
            } while (v < sy.wait_value);
        }
        __syncthreads();
    }
}
__device__ __forceinline__ void slab_signal(const SlabSync& sy)
{
    if (sy.signal_flag != nullptr) {
        __threadfence_system();              // this thread's carry stores are visible system-wide ...
        __syncthreads();                     // ... for every thread of the CTA
        if (threadIdx.x == 0) {
            if (atomicAdd(sy.done, 1) == (int)gridDim.x - 1) {       // last CTA of the chunk
                *sy.done = 0;
                __threadfence_system();
                asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(sy.signal_flag), "r"(sy.signal_value) : "memory");
            }
        }
    }
}

__global__ void __launch_bounds__(256) k_thomas_fwd_slab(int nt, int ny, i64 P, i64 m0, i64 m1, int t0, int t1, double inv_scale,
                                                         const double* __restrict__ lam_x, const double* __restrict__ lam_y,
                                                         const double* __restrict__ gtab, const int* __restrict__ t_fix,
                                                         double* __restrict__ a, const double* carry_in,
                                                         double* __restrict__ carry_out, SlabSync sy)
{
    slab_wait(sy);
    const i64 p = m0 + blockIdx.x * (i64)blockDim.x + threadIdx.x;
    // (no early return: slab_signal holds a CTA barrier, which every thread must reach from the same place)
    if (p < m1 && p != 0) {        // mode (0,0) is solved separately (k_tline0)
        const int kx = (int)(p / ny), ky = (int)(p - (i64)kx * ny);
        const double ct = (double)(nt - 1) * (double)(nt - 1);
        const double nu = (lam_y[ky] + lam_x[kx]) / ct;
        SlabG G{gtab, P, p, t0, nt, t_fix[p], 0.0};
        if (G.tf < nt) G.gs = thomas_fixed_point(nt, G.tf, nu);
        double d = t0 > 0 ? __ldcg(carry_in + p) : 0.0;      // (written by the neighbour's GPU: not through the read-only path)
        int t = t0;
        for (; t + 4 <= t1; t += 4) {
            double r[4], g[4];
#pragma unroll
            for (int u = 0; u < 4; u++) { r[u] = a[(i64)(t + u) * P + p]; g[u] = G.at(t + u); }
#pragma unroll
            for (int u = 0; u < 4; u++) { d = __dmul_rn(__fma_rn(r[u], inv_scale, d), g[u]); a[(i64)(t + u) * P + p] = d; }
        }
        for (; t < t1; t++) { d = __dmul_rn(__fma_rn(a[(i64)t * P + p], inv_scale, d), G.at(t)); a[(i64)t * P + p] = d; }
        if (t1 < nt) carry_out[p] = d;
    }
    slab_signal(sy);
}

__global__ void __launch_bounds__(256) k_thomas_bwd_slab(int nt, int ny, i64 P, i64 m0, i64 m1, int t0, int t1,
                                                         const double* __restrict__ lam_x, const double* __restrict__ lam_y,
                                                         const double* __restrict__ gtab, const int* __restrict__ t_fix,
                                                         double* __restrict__ a, const double* carry_in,
                                                         double* __restrict__ carry_out, SlabSync sy)
{
    slab_wait(sy);
    const i64 p = m0 + blockIdx.x * (i64)blockDim.x + threadIdx.x;
    if (p < m1 && p != 0) {
        const int kx = (int)(p / ny), ky = (int)(p - (i64)kx * ny);
        const double ct = (double)(nt - 1) * (double)(nt - 1);
        const double nu = (lam_y[ky] + lam_x[kx]) / ct;
        SlabG G{gtab, P, p, t0, nt, t_fix[p], 0.0};
        if (G.tf < nt) G.gs = thomas_fixed_point(nt, G.tf, nu);
        double x;
        int t;
        if (t1 == nt) { x = a[(i64)(nt - 1) * P + p]; t = nt - 2; }     // last level: x = d
        else { x = __ldcg(carry_in + p); t = t1 - 1; }
        for (; t - 3 >= t0; t -= 4) {
            double dd[4], g[4];
#pragma unroll
            for (int u = 0; u < 4; u++) { dd[u] = a[(i64)(t - u) * P + p]; g[u] = G.at(t - u); }
#pragma unroll
            for (int u = 0; u < 4; u++) { x = __fma_rn(g[u], x, dd[u]); a[(i64)(t - u) * P + p] = x; }
        }
        for (; t >= t0; t--) { x = __fma_rn(G.at(t), x, a[(i64)t * P + p]); a[(i64)t * P + p] = x; }
        if (t0 > 0) carry_out[p] = x;       // x of my first level, for the slab below
    }
    slab_signal(sy);
}

// the singular mode: its nt values live one per level on the owning slabs; line[t] <-> a[t*P]
__global__ void k_line0_gather(int t0, int t1, i64 P, const double* __restrict__ a, double* __restrict__ line)
{
    const int t = t0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (t < t1) line[t] = a[(i64)t * P];
}
__global__ void k_line0_scatter(int t0, int t1, i64 P, double* __restrict__ a, const double* __restrict__ line)
{
    const int t = t0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (t < t1) a[(i64)t * P] = line[t];
}

// mode (0,0): phi = IDCT_t( DCT_t(r) ./ (D2 * lam_t) ), lam_t[0] := 1, dense nt x nt transform by one CTA
__global__ void __launch_bounds__(1024) k_tline0(int nt, i64 stride, double D2, const double* __restrict__ lam_t,
                                                const double* __restrict__ cmat, double* __restrict__ a)
{
    extern __shared__ double sl[];   // [2][nt]
    double* r = sl;
    double* X = sl + nt;
    for (int t = threadIdx.x; t < nt; t += blockDim.x) r[t] = a[(i64)t * stride];
    __syncthreads();
    // forward: one warp per coefficient row (coalesced reads of the row, shuffle reduction)
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int k = wid; k < nt; k += nw) {
        double acc = 0.0;
        for (int t = lane; t < nt; t += 32) acc += cmat[(i64)k * nt + t] * r[t];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
        if (lane == 0) {
            double kv = lam_t[k];
            if (kv == 0.0) kv = 1.0;
            X[k] = acc / (D2 * kv);
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < nt; t += blockDim.x) {
        double acc = 0.0;
        for (int k = 0; k < nt; k++) acc += cmat[(i64)k * nt + t] * X[k];
        a[(i64)t * stride] = acc;
    }
}

static double* dense_dct_matrix(int n)
{
    std::vector<double> C((size_t)n * n);
    const long double c0 = sqrtl(1.0L / n), c1 = sqrtl(2.0L / n);
    for (int k = 0; k < n; k++)
        for (int j = 0; j < n; j++) {
            const long long a = ((long long)(2 * j + 1) * k) % (4LL * n);
            C[(size_t)k * n + j] = (double)((k == 0 ? c0 : c1) * cosl(PIl * (long double)a / (2.0L * n)));
        }
    return to_device(C);
}

static double* lam_table(int n)
{
    std::vector<double> v(n);
    // (2*(n-1)^2) * (1 - cos(pi*k/n))   initialize_FFTkernel.m:6-8 (double arithmetic as in MATLAB)
    for (int k = 0; k < n; k++) v[k] = (2.0 * (double)(n - 1) * (double)(n - 1)) * (1.0 - cos(M_PI * (double)k / (double)n));
    return to_device(v);
}

PoissonPlan* poisson_plan_create(int nt, int nx, int ny)
{
    PoissonPlan* p = new PoissonPlan();
    p->g = make_geo(nt, nx, ny);
    p->py = dct_plan_create(ny);
    p->px = dct_plan_create(nx);
    p->pt = dct_plan_create(nt);
    p->lam_t = lam_table(nt);
    p->lam_x = lam_table(nx);
    p->lam_y = lam_table(ny);
    p->cmat_t = nullptr;
    const char* e = getenv("DOTSOCP_TSOLVE");
    p->use_thomas = !(e && strcmp(e, "dct") == 0);
    return p;
}

void poisson_plan_destroy(PoissonPlan* p)
{
    if (!p) return;
    for (auto& e : p->gtabs) { cudaFree(e.tab); cudaFree(e.t_fix); }
    cudaFree(p->cmat_t);
    dct_plan_destroy(p->py); dct_plan_destroy(p->px); dct_plan_destroy(p->pt);
    cudaFree(p->lam_t); cudaFree(p->lam_x); cudaFree(p->lam_y);
    delete p;
}

// t-solve on a [nt][lines] array (line stride = lines) holding the modes p0 .. p0+lines-1
static int t_solve(PoissonPlan* p, double* buf, i64 lines, i64 p0, double D2, cudaStream_t st, double* launches,
                   double* const* push_tab = nullptr, const int* tcut = nullptr, int world = 1)
{
    const Geo& g = p->g;
    if (!p->use_thomas || g.nt < 3) {
        ScaleArgs sa{p->lam_t, p->lam_x, p->lam_y, g.ny, D2, p0};
        LineGeom gt{g.nt, lines, 1, lines, 0, 0, 0, 0, 0, 0, 0};
        if (launches) *launches += 1;
        return launch_dct_axis(p->pt, gt, 1, buf, buf, 2, sa, st);
    }
    // one table per mode range (a process that emulates several slabs keeps one per slab)
    double* gtab = nullptr;
    int* t_fix = nullptr;
    for (auto& e : p->gtabs)
        if (e.p0 == p0 && e.lines == lines) { gtab = e.tab; t_fix = e.t_fix; }
    if (!gtab) {
        cudaMalloc(&gtab, (size_t)g.nt * lines * sizeof(double));
        cudaMalloc(&t_fix, (size_t)lines * sizeof(int));
        k_thomas_table<<<(unsigned)((lines + 255) / 256), 256, 0, st>>>(g.nt, g.ny, lines, lines, p0, p->lam_x, p->lam_y, gtab, t_fix);
        p->gtabs.push_back({p0, lines, gtab, t_fix});
        if (!p->cmat_t) p->cmat_t = dense_dct_matrix(g.nt);
        if (launches) *launches += 1;
    }
    const double ct = (double)(g.nt - 1) * (double)(g.nt - 1);
    if (p0 == 0) {
        k_tline0<<<1, 1024, (size_t)2 * g.nt * sizeof(double), st>>>(g.nt, lines, D2, p->lam_t, p->cmat_t, buf);
        if (launches) *launches += 1;
    }
    if (push_tab)
        k_thomas<true><<<(unsigned)((lines + 255) / 256), 256, 0, st>>>(g.nt, lines, lines, p0, 1.0 / (D2 * ct), gtab, t_fix, buf, push_tab, tcut, world);
    else
        k_thomas<false><<<(unsigned)((lines + 255) / 256), 256, 0, st>>>(g.nt, lines, lines, p0, 1.0 / (D2 * ct), gtab, t_fix, buf, nullptr, nullptr, 1);
    if (launches) *launches += 1;
    return 0;
}

// y lines are contiguous; they are grouped PER TIME LEVEL (outer index = level) although the whole array is one contiguous run of
// nt*nx lines: the kernels transform two real lines as one complex sequence, and the rounding of a line depends on its partner,
// so the pairing must not depend on where a time slab starts -- otherwise 1, 2, 4 and 8 GPUs would not give the same bits.
static LineGeom geom_y(const Geo& g) { return LineGeom{g.ny, 1, (i64)g.ny, (i64)g.nx, g.P, 1, 0, 0, 0, 0, 0}; }
static LineGeom geom_x(const Geo& g)
{
    if (g.ny == 1) return LineGeom{g.nx, 1, (i64)g.nx, (i64)g.nt, 0, 1, 0, 0, 0, 0, 0};   // 1-D variant: x lines are contiguous
    return LineGeom{g.nx, (i64)g.ny, 1, (i64)g.ny, g.P, 0, 0, 0, 0, 0, 0};
}
static LineGeom geom_t(const Geo& g) { return LineGeom{g.nt, g.P, 1, g.P, 0, 0, 0, 0, 0, 0, 0}; }

int poisson_solve(PoissonPlan* p, const double* rhs, double* a, double D2, cudaStream_t st, double* launches)
{
    // first pass reads rhs and writes a (out of place), the rest works in place on a
    const Geo& g = p->g;
    ScaleArgs sa{p->lam_t, p->lam_x, p->lam_y, g.ny, D2, 0};
    const i64 xo = (g.ny == 1) ? 1 : g.nt;
    const double* src = rhs;
    int rc = 0;
    if (g.ny > 1) { rc = launch_dct_axis(p->py, geom_y(g), g.nt, src, a, 0, sa, st); src = a; if (launches) *launches += 1; }
    if (!rc) rc = launch_dct_axis(p->px, geom_x(g), xo, src, a, 0, sa, st);
    if (!rc) rc = t_solve(p, a, g.P, 0, D2, st, launches);
    if (!rc) rc = launch_dct_axis(p->px, geom_x(g), xo, a, a, 1, sa, st);
    if (launches) *launches += 2;
    if (!rc && g.ny > 1) { rc = launch_dct_axis(p->py, geom_y(g), g.nt, a, a, 1, sa, st); if (launches) *launches += 1; }
    return rc;
}

// ---- pieces of the solve for a time slab (node levels [tn0, tn0+nlev) of the global array) --------------------------------
int poisson_xy(PoissonPlan* p, const double* src, double* a, int tn0, int nlev, bool inverse, cudaStream_t st, double* launches,
               double* packed, int world, int slab_nlev, int slab_t0, double* const* push_tab)
{
    // (tn0, nlev) may be a group of levels of a slab that owns slab_nlev levels starting slab_t0 levels before tn0
    // packed != NULL (and the x length uses the register-FFT kernel): the forward x pass writes, and the inverse x pass
    // reads, the packed all-to-all buffer directly instead of the slab rows of `a`
    const Geo& g = p->g;
    ScaleArgs sa{p->lam_t, p->lam_x, p->lam_y, g.ny, 1.0, 0};
    const i64 off = (i64)tn0 * g.P;
    const LineGeom gy = geom_y(g);   // grouped per level: the same line pairs as the single-slab solve
    LineGeom gx = (g.ny == 1) ? LineGeom{g.nx, 1, (i64)g.nx, (i64)nlev, 0, 1, 0, 0, 0, 0, 0} : LineGeom{g.nx, (i64)g.ny, 1, (i64)g.ny, g.P, 0, 0, 0, 0, 0, 0};
    if (packed) { gx.rm_world = world; gx.rm_nlev = slab_nlev > 0 ? slab_nlev : nlev; gx.rm_t0 = slab_t0; gx.rm_P = g.P; gx.rm_ny = g.ny; gx.rm_tab = inverse ? nullptr : push_tab; }
    const i64 xo = (g.ny == 1) ? 1 : nlev;
    int rc = 0;
    if (!inverse) {
        const double* s0 = src + off;
        if (g.ny > 1) { rc = launch_dct_axis(p->py, gy, nlev, s0, a + off, 0, sa, st); s0 = a + off; if (launches) *launches += 1; }
        if (!rc) rc = launch_dct_axis(p->px, gx, xo, s0, packed ? packed : a + off, 0, sa, st);
        if (launches) *launches += 1;
    } else {
        rc = launch_dct_axis(p->px, gx, xo, packed ? packed : a + off, a + off, 1, sa, st);
        if (launches) *launches += 1;
        if (!rc && g.ny > 1) { rc = launch_dct_axis(p->py, gy, nlev, a + off, a + off, 1, sa, st); if (launches) *launches += 1; }
    }
    return rc;
}
bool poisson_can_pack(const PoissonPlan* p) { return p->g.ny > 1 && !p->px->dense && p->px->log2m >= 8 && p->px->log2m <= 12; }
// t-pass (DCT_t, ./kernel, IDCT_t) on a [nt][chunk] array holding the (x,y) modes p0 .. p0+chunk-1
int poisson_t_chunk(PoissonPlan* p, double* buf, i64 chunk, i64 p0, double D2, cudaStream_t st, double* launches,
                    double* const* push_tab, const int* tcut, int world)
{
    // push_tab != NULL needs the Thomas solve (the transform-based t pass works in place only)
    return t_solve(p, buf, chunk, p0, D2, st, launches, p->use_thomas && p->g.nt >= 3 ? push_tab : nullptr, tcut, world);
}

// ---- pipelined slab Thomas (see the kernels): pieces called by solver.cu ----------------------------------------------------
bool poisson_slab_thomas_ok(const PoissonPlan* p) { return p->use_thomas && p->g.nt >= 3; }
static PoissonPlan::GTab* slab_table(PoissonPlan* p, int t0, int t1, cudaStream_t st, double* launches)
{
    const Geo& g = p->g;
    for (auto& e : p->gtabs)
        if (e.p0 == -(i64)(t0 + 1) && e.lines == (i64)(t1 - t0)) return &e;     // slab tables are keyed by (-(t0+1), nlev)
    double* gtab = nullptr;
    int* t_fix = nullptr;
    if (cudaMalloc(&gtab, (size_t)(t1 - t0) * g.P * sizeof(double)) != cudaSuccess || cudaMalloc(&t_fix, (size_t)g.P * sizeof(int)) != cudaSuccess) {
        cudaGetLastError();
        cudaFree(gtab);
        return nullptr;
    }
    k_thomas_table_slab<<<(unsigned)((g.P + 255) / 256), 256, 0, st>>>(g.nt, g.ny, g.P, t0, t1, p->lam_x, p->lam_y, gtab, t_fix);
    if (launches) *launches += 1;
    if (!p->cmat_t) p->cmat_t = dense_dct_matrix(g.nt);
    p->gtabs.push_back({-(i64)(t0 + 1), (i64)(t1 - t0), gtab, t_fix});
    return &p->gtabs.back();
}
// Build everything the first t-solve of a session would otherwise build lazily in the middle of the iteration stream: the
// pivot table of the mode range (lines > 0: t_solve's [nt][lines] table for modes p0 ..; lines == 0: the slab table of levels
// [t0, t1)) and the dense DCT matrix of the singular mode.  Called once per session right after the plan is created.
int poisson_prepare(PoissonPlan* p, i64 p0, i64 lines, int t0, int t1, cudaStream_t st)
{
    const Geo& g = p->g;
    if (!p->use_thomas || g.nt < 3) return 0;
    if (lines > 0) {
        bool have = false;
        for (auto& e : p->gtabs) have = have || (e.p0 == p0 && e.lines == lines);
        if (!have) {
            double* gtab = nullptr;
            int* t_fix = nullptr;
            if (cudaMalloc(&gtab, (size_t)g.nt * lines * sizeof(double)) != cudaSuccess || cudaMalloc(&t_fix, (size_t)lines * sizeof(int)) != cudaSuccess) {
                cudaGetLastError();
                cudaFree(gtab);
                return dsocp_set_err(-4, "pivot table of the Thomas solve: out of device memory");
            }
            k_thomas_table<<<(unsigned)((lines + 255) / 256), 256, 0, st>>>(g.nt, g.ny, lines, lines, p0, p->lam_x, p->lam_y, gtab, t_fix);
            p->gtabs.push_back({p0, lines, gtab, t_fix});
        }
    } else if (!slab_table(p, t0, t1, st, nullptr)) {
        return dsocp_set_err(-4, "pivot table of the slab Thomas solve: out of device memory");
    }
    if (!p->cmat_t) p->cmat_t = dense_dct_matrix(g.nt);
    return 0;
}

// forward elimination / back substitution of modes [m0, m1) on levels [t0, t1) of the natural-layout array `a`
int poisson_thomas_slab(PoissonPlan* p, double* a, int t0, int t1, i64 m0, i64 m1, double D2, bool backward, const double* carry_in,
                        double* carry_out, cudaStream_t st, double* launches, const SlabSync* sync)
{
    const SlabSync sy = sync ? *sync : SlabSync{nullptr, 0, nullptr, nullptr, 0};
    const Geo& g = p->g;
    if (m1 <= m0) return 0;
    PoissonPlan::GTab* tb = slab_table(p, t0, t1, st, launches);
    if (!tb) return dsocp_set_err(-4, "pivot table of the slab Thomas solve: out of device memory");
    const double ct = (double)(g.nt - 1) * (double)(g.nt - 1);
    const unsigned nb = (unsigned)((m1 - m0 + 255) / 256);
    if (!backward)
        k_thomas_fwd_slab<<<nb, 256, 0, st>>>(g.nt, g.ny, g.P, m0, m1, t0, t1, 1.0 / (D2 * ct), p->lam_x, p->lam_y, tb->tab, tb->t_fix, a,
                                              carry_in, carry_out, sy);
    else
        k_thomas_bwd_slab<<<nb, 256, 0, st>>>(g.nt, g.ny, g.P, m0, m1, t0, t1, p->lam_x, p->lam_y, tb->tab, tb->t_fix, a, carry_in, carry_out, sy);
    if (launches) *launches += 1;
    return 0;
}
// the singular mode (kx = ky = 0): gather its values of levels [t0, t1) into line[nt] / solve the complete line / scatter back
void poisson_line0_gather(PoissonPlan* p, const double* a, double* line, int t0, int t1, cudaStream_t st)
{
    if (t1 > t0) k_line0_gather<<<(unsigned)((t1 - t0 + 127) / 128), 128, 0, st>>>(t0, t1, p->g.P, a, line);
}
void poisson_line0_solve(PoissonPlan* p, double* line, double D2, cudaStream_t st)
{
    if (!p->cmat_t) p->cmat_t = dense_dct_matrix(p->g.nt);
    k_tline0<<<1, 1024, (size_t)2 * p->g.nt * sizeof(double), st>>>(p->g.nt, 1, D2, p->lam_t, p->cmat_t, line);
}
void poisson_line0_scatter(PoissonPlan* p, double* a, const double* line, int t0, int t1, cudaStream_t st)
{
    if (t1 > t0) k_line0_scatter<<<(unsigned)((t1 - t0 + 127) / 128), 128, 0, st>>>(t0, t1, p->g.P, a, line);
}

int poisson_dctn(PoissonPlan* p, double* a, bool inverse, cudaStream_t st, double* launches)
{
    const Geo& g = p->g;
    ScaleArgs sa{p->lam_t, p->lam_x, p->lam_y, g.ny, 1.0, 0};
    const int mode = inverse ? 1 : 0;
    int rc = 0;
    if (g.ny > 1) rc = launch_dct_axis(p->py, geom_y(g), g.nt, a, a, mode, sa, st);
    if (!rc && g.nx > 1) rc = launch_dct_axis(p->px, geom_x(g), (g.ny == 1) ? 1 : g.nt, a, a, mode, sa, st);
    if (!rc && g.nt > 1) rc = launch_dct_axis(p->pt, geom_t(g), 1, a, a, mode, sa, st);
    if (launches) *launches += 3;
    return rc;
}

}  // namespace dsocp
