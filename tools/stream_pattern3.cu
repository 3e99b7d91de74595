// stream_pattern3.cu -- which misalignment costs k_mult's march its bandwidth?  (follow-up of stream_pattern2.cu: the same
// 21-read / 13-write march over (x, y) tiles reaches 6.36 TB/s with 256-byte-aligned rows and 4.25 TB/s with ny = 513.)
// One kernel, run-time geometry: arrays (nt, nx, pitch); CTA = 8 rows x 32 lanes, tile origin y0 = blockIdx.x * TS;
// reads at column y0 + lane + roff, writes at y0 + lane + woff for lanes < OW.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/bin/stream_pattern3 tools/stream_pattern3.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int NR = 21, NW = 13;
struct Ptrs { const double* r[NR]; double* w[NW]; };

__global__ void __launch_bounds__(256, 2) k(Ptrs p, int nt, int nx, int ny, int pitch, int TS, int OW, int roff, int woff)
{
    const int ly = threadIdx.x, lx = threadIdx.y;
    const int x = blockIdx.y * 8 + lx, y = blockIdx.x * TS + ly;
    if (x >= nx || y >= ny) return;
    const long long P = (long long)nx * pitch;
    const long long base = (long long)x * pitch + y;
    const bool st = ly < OW;
    for (int t = 0; t < nt; t++) {
        const long long i = t * P + base;
        double s = 0;
#pragma unroll
        for (int a = 0; a < NR; a++) s += p.r[a][i + roff];
        if (st) {
#pragma unroll
            for (int a = 0; a < NW; a++) p.w[a][i + woff] = s + a;
        }
    }
}

int main()
{
    const int nt = 256, nx = 513, ny = 513, maxpitch = 544;
    const long long NA = (long long)nt * nx * maxpitch + 64;
    Ptrs p;
    for (int a = 0; a < NR; a++) { double* d; cudaMalloc(&d, NA * 8); cudaMemset(d, 0, NA * 8); p.r[a] = d; }
    for (int a = 0; a < NW; a++) { double* d; cudaMalloc(&d, NA * 8); p.w[a] = d; }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto run = [&](const char* name, int pitch, int TS, int OW, int roff, int woff) {
        dim3 grid((ny + TS - 1) / TS, (nx + 7) / 8), block(32, 8);
        float best = 1e30f;
        for (int rep = 0; rep < 5; rep++) {
            cudaEventRecord(e0);
            k<<<grid, block>>>(p, nt, nx, ny, pitch, TS, OW, roff, woff);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0 && ms < best) best = ms;
        }
        // useful bytes: every cell read NR times; written NW times where owned
        const double cells = (double)nt * nx * ny;
        const double gb = cells * 8 * (NR + NW * (double)OW / TS) / 1e9;
        cudaError_t e = cudaGetLastError();
        printf("%-44s pitch %3d TS %2d OW %2d roff %d woff %d : %7.3f ms  %7.1f GB/s %s\n", name, pitch, TS, OW, roff, woff, best,
               gb / (best * 1e-3), e == cudaSuccess ? "" : cudaGetErrorString(e));
    };
    run("aligned rows, aligned tiles", 544, 32, 32, 0, 0);
    run("ny = 513 rows, 32-wide tiles (v1)", 513, 32, 32, 0, 0);
    run("ny = 513 rows, 31-wide tiles (k_mult today)", 513, 31, 31, 0, 0);
    run("aligned rows, 31-wide tiles", 544, 31, 31, 0, 0);
    run("aligned rows, tiles 32 / owned 31 stores", 544, 32, 31, 0, 0);
    run("aligned, reads shifted by 1", 544, 32, 32, 1, 0);
    run("aligned, writes shifted by 1", 544, 32, 32, 0, 1);
    run("aligned, reads and writes shifted by 1", 544, 32, 32, 1, 1);
    run("aligned, writes shifted by 4 (32 B)", 544, 32, 32, 0, 4);
    run("aligned, writes shifted by 16 (128 B)", 544, 32, 32, 0, 16);
    run("pitch 516 (32-byte rows)", 516, 32, 32, 0, 0);
    run("pitch 528 (128-byte rows)", 528, 32, 32, 0, 0);
    run("pitch 520 (64-byte rows)", 520, 32, 32, 0, 0);
    run("pitch 516, 28-wide tiles (32 B-aligned origins)", 516, 28, 28, 0, 0);
    return 0;
}
