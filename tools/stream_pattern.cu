// stream_pattern.cu -- what bandwidth does k_mult's ACCESS PATTERN reach with no arithmetic at all?
// NR read streams + NW write streams over (nt, nx, ny) arrays (y fastest), 8 x 32 tiles of 256 threads:
//   mode t : tile over (x, y), march along t  (k_mult today: consecutive steps of a stream are nx*ny*8 bytes apart)
//   mode x : tile over (t, y), march along x  (consecutive steps are ny*8 bytes apart: same 2 MB page for many steps)
//   mode f : flat grid-stride copy of the same arrays (the reference point)
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/bin/stream_pattern tools/stream_pattern.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

struct Ptrs { const double* r[24]; double* w[16]; };

template <int NR, int NW, int MODE>
__global__ void __launch_bounds__(256, 2) k(Ptrs p, int nt, int nx, int ny)
{
    const int ly = threadIdx.x, l2 = threadIdx.y;
    const long long P = (long long)nx * ny;
    if (MODE == 0) {          // march t
        const int x = blockIdx.y * 8 + l2, y = blockIdx.x * 32 + ly;
        if (x >= nx || y >= ny) return;
        const long long base = (long long)x * ny + y;
        for (int t = 0; t < nt; t++) {
            const long long i = t * P + base;
            double s = 0;
#pragma unroll
            for (int a = 0; a < NR; a++) s += p.r[a][i];
#pragma unroll
            for (int a = 0; a < NW; a++) p.w[a][i] = s + a;
        }
    } else {                  // march x
        const int t = blockIdx.y * 8 + l2, y = blockIdx.x * 32 + ly;
        if (t >= nt || y >= ny) return;
        const long long base = (long long)t * P + y;
        for (int x = 0; x < nx; x++) {
            const long long i = base + (long long)x * ny;
            double s = 0;
#pragma unroll
            for (int a = 0; a < NR; a++) s += p.r[a][i];
#pragma unroll
            for (int a = 0; a < NW; a++) p.w[a][i] = s + a;
        }
    }
}
template <int NR, int NW>
__global__ void __launch_bounds__(256) kflat(Ptrs p, long long n)
{
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += gridDim.x * 256LL) {
        double s = 0;
#pragma unroll
        for (int a = 0; a < NR; a++) s += p.r[a][i];
#pragma unroll
        for (int a = 0; a < NW; a++) p.w[a][i] = s + a;
    }
}

template <int NR, int NW>
void run(int nt, int nx, int ny)
{
    const long long N = (long long)nt * nx * ny;
    Ptrs p;
    std::vector<void*> all;
    for (int a = 0; a < NR; a++) { double* d; cudaMalloc(&d, N * 8); cudaMemset(d, 0, N * 8); p.r[a] = d; all.push_back(d); }
    for (int a = 0; a < NW; a++) { double* d; cudaMalloc(&d, N * 8); p.w[a] = d; all.push_back(d); }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const double gb = (double)(NR + NW) * N * 8 / 1e9;
    for (int mode = 0; mode < 3; mode++) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; rep++) {
            cudaEventRecord(e0);
            if (mode == 0) k<NR, NW, 0><<<dim3((ny + 31) / 32, (nx + 7) / 8), dim3(32, 8)>>>(p, nt, nx, ny);
            else if (mode == 1) k<NR, NW, 1><<<dim3((ny + 31) / 32, (nt + 7) / 8), dim3(32, 8)>>>(p, nt, nx, ny);
            else kflat<NR, NW><<<148 * 16, 256>>>(p, N);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0 && ms < best) best = ms;
        }
        printf("NR=%2d NW=%2d %s: %8.3f ms  %7.1f GB/s\n", NR, NW, mode == 0 ? "march-t" : mode == 1 ? "march-x" : "flat   ", best, gb / (best * 1e-3));
    }
    for (void* d : all) cudaFree(d);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("CUDA error %s\n", cudaGetErrorString(e));
}

int main(int argc, char** argv)
{
    const int nt = argc > 1 ? atoi(argv[1]) : 257, nx = argc > 2 ? atoi(argv[2]) : 513, ny = argc > 3 ? atoi(argv[3]) : 513;
    printf("grid %d x %d x %d (nt, nx, ny), %.2f GB per array\n", nt, nx, ny, (double)nt * nx * ny * 8 / 1e9);
    run<1, 1>(nt, nx, ny);
    run<4, 2>(nt, nx, ny);
    run<10, 10>(nt, nx, ny);
    run<21, 13>(nt, nx, ny);
    return 0;
}
