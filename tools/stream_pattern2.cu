// stream_pattern2.cu -- why does a 21-read / 13-write stream mix reach only ~4.5 TB/s on B200, and which layout fixes it?
// All variants move the same (21 + 13) * N doubles; no arithmetic beyond one add chain.
//   soa      : 34 separate arrays, flat grid-stride (the reference point of stream_pattern.cu)
//   soa_ro   : 34 arrays, all read (no stores)          -> is it the read/write mix?
//   soa_v2   : as soa with 16-byte accesses             -> is it the access width?
//   soa_u4   : as soa, 4 elements per thread in flight  -> is it memory-level parallelism?
//   blk10    : 10 of the reads and 10 of the writes interleaved as [block of 32 cells][10][32] (one stream each): 11+1 R, 3+1 W
//   blkall   : all reads in one [block][21][32] array, all writes in one [block][13][32] array: 1 R, 1 W
//   blk10_t  : blk10 layout with k_mult's march (8x32 tiles over (x, yblock), loop over t)
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/bin/stream_pattern2 tools/stream_pattern2.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

struct Ptrs { const double* r[34]; double* w[16]; };
constexpr int NR = 21, NW = 13;

__global__ void __launch_bounds__(256) k_soa(Ptrs p, long long n)
{
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += gridDim.x * 256LL) {
        double s = 0;
#pragma unroll
        for (int a = 0; a < NR; a++) s += p.r[a][i];
#pragma unroll
        for (int a = 0; a < NW; a++) p.w[a][i] = s + a;
    }
}
__global__ void __launch_bounds__(256) k_soa_ro(Ptrs p, long long n, double* sink)
{
    double tot = 0;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += gridDim.x * 256LL) {
        double s = 0;
#pragma unroll
        for (int a = 0; a < NR + NW; a++) s += p.r[a][i];
        tot += s;
    }
    if (tot == 1.2345) *sink = tot;
}
__global__ void __launch_bounds__(256) k_soa_v2(Ptrs p, long long n2)
{
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n2; i += gridDim.x * 256LL) {
        double2 s = make_double2(0, 0);
#pragma unroll
        for (int a = 0; a < NR; a++) { const double2 v = reinterpret_cast<const double2*>(p.r[a])[i]; s.x += v.x; s.y += v.y; }
#pragma unroll
        for (int a = 0; a < NW; a++) reinterpret_cast<double2*>(p.w[a])[i] = make_double2(s.x + a, s.y + a);
    }
}
__global__ void __launch_bounds__(256) k_soa_u4(Ptrs p, long long n)
{
    const long long stride = gridDim.x * 256LL;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += 4 * stride) {
        double s[4] = {0, 0, 0, 0};
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int a = 0; a < NR; a++)
                if (i + u * stride < n) s[u] += p.r[a][i + u * stride];
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int a = 0; a < NW; a++)
                if (i + u * stride < n) p.w[a][i + u * stride] = s[u] + a;
    }
}
// blocked: KB of the reads / writes live in one array laid out [cell / 32][KB][32]
template <int KB>
__global__ void __launch_bounds__(256) k_blk(Ptrs p, const double* __restrict__ rb, double* __restrict__ wb, long long n)
{
    constexpr int KW = KB < NW ? KB : NW;    // blocked writes
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += gridDim.x * 256LL) {
        const long long blk = i >> 5;
        const int lane = (int)(i & 31);
        double s = 0;
#pragma unroll
        for (int a = 0; a < KB; a++) s += rb[(blk * KB + a) * 32 + lane];
#pragma unroll
        for (int a = 0; a < NR - KB; a++) s += p.r[a][i];
#pragma unroll
        for (int a = 0; a < KW; a++) wb[(blk * KW + a) * 32 + lane] = s + a;
#pragma unroll
        for (int a = 0; a < NW - KW; a++) p.w[a][i] = s + a;
    }
}
// k_mult's march with the blocked layout: tile = 8 x-rows x 32 y, loop over t; nyb = blocks of 32 along y
template <int KB, bool BLOCKED>
__global__ void __launch_bounds__(256, 2) k_march(Ptrs p, const double* __restrict__ rb, double* __restrict__ wb, int nt, int nx, int nyp)
{
    constexpr int KW = KB < NW ? KB : NW;
    const int ly = threadIdx.x, lx = threadIdx.y;
    const int x = blockIdx.y * 8 + lx, yb = blockIdx.x;
    if (x >= nx) return;
    const long long P = (long long)nx * nyp;
    const long long base = (long long)x * nyp + yb * 32 + ly;
    for (int t = 0; t < nt; t++) {
        const long long i = t * P + base;
        const long long blk = i >> 5;
        double s = 0;
        if (BLOCKED) {
#pragma unroll
            for (int a = 0; a < KB; a++) s += rb[(blk * KB + a) * 32 + ly];
#pragma unroll
            for (int a = 0; a < NR - KB; a++) s += p.r[a][i];
#pragma unroll
            for (int a = 0; a < KW; a++) wb[(blk * KW + a) * 32 + ly] = s + a;
#pragma unroll
            for (int a = 0; a < NW - KW; a++) p.w[a][i] = s + a;
        } else {
#pragma unroll
            for (int a = 0; a < NR; a++) s += p.r[a][i];
#pragma unroll
            for (int a = 0; a < NW; a++) p.w[a][i] = s + a;
        }
    }
}

int main(int argc, char** argv)
{
    const int nt = argc > 1 ? atoi(argv[1]) : 256, nx = argc > 2 ? atoi(argv[2]) : 513, nyp = argc > 3 ? atoi(argv[3]) : 544;  // y padded to 32
    const long long N = (long long)nt * nx * nyp;
    printf("grid %d x %d x %d, %.2f GB per array, %d reads + %d writes per cell\n", nt, nx, nyp, N * 8 / 1e9, NR, NW);
    Ptrs p;
    for (int a = 0; a < NR + NW; a++) { double* d; cudaMalloc(&d, N * 8); cudaMemset(d, 0, N * 8); p.r[a] = d; }
    for (int a = 0; a < NW; a++) p.w[a] = const_cast<double*>(p.r[NR + a]);
    double *rb, *wb, *sink;
    cudaMalloc(&rb, N * 8 * NR); cudaMemset(rb, 0, N * 8 * NR);
    cudaMalloc(&wb, N * 8 * NW);
    cudaMalloc(&sink, 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const double gb = (double)(NR + NW) * N * 8 / 1e9;
    const int G = 148 * 16;
    auto timeit = [&](const char* name, auto&& launch) {
        float best = 1e30f;
        for (int rep = 0; rep < 5; rep++) {
            cudaEventRecord(e0);
            launch();
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0 && ms < best) best = ms;
        }
        cudaError_t e = cudaGetLastError();
        printf("%-10s %8.3f ms  %7.1f GB/s %s\n", name, best, gb / (best * 1e-3), e == cudaSuccess ? "" : cudaGetErrorString(e));
    };
    timeit("soa", [&] { k_soa<<<G, 256>>>(p, N); });
    timeit("soa_ro", [&] { k_soa_ro<<<G, 256>>>(p, N, sink); });
    timeit("soa_v2", [&] { k_soa_v2<<<G, 256>>>(p, N / 2); });
    timeit("soa_u4", [&] { k_soa_u4<<<G, 256>>>(p, N); });
    timeit("soa_g8", [&] { k_soa<<<148 * 8, 256>>>(p, N); });
    timeit("soa_g4", [&] { k_soa<<<148 * 4, 256>>>(p, N); });
    timeit("blk10", [&] { k_blk<10><<<G, 256>>>(p, rb, wb, N); });
    timeit("blk21", [&] { k_blk<21><<<G, 256>>>(p, rb, wb, N); });
    dim3 grid(nyp / 32, (nx + 7) / 8), block(32, 8);
    timeit("soa_t", [&] { k_march<10, false><<<grid, block>>>(p, rb, wb, nt, nx, nyp); });
    timeit("blk10_t", [&] { k_march<10, true><<<grid, block>>>(p, rb, wb, nt, nx, nyp); });
    timeit("blk21_t", [&] { k_march<21, true><<<grid, block>>>(p, rb, wb, nt, nx, nyp); });
    return 0;
}
