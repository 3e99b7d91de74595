// reduce.cuh -- stage 1 of the fixed-order reductions (kernels.h): every CTA leaves K partial sums.
#pragma once
#include "common.cuh"

namespace dsocp {

template <int K, int NT>
__device__ __forceinline__ void block_reduce_store(double (&s)[K], double* __restrict__ partial, i64 block)
{
    __shared__ double red[K][NT / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; k++) {
        double v = s[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) red[k][wid] = v;
    }
    __syncthreads();
    if (threadIdx.x < K) {
        double v = 0.0;
        for (int i = 0; i < NT / 32; i++) v += red[threadIdx.x][i];
        partial[block * K + threadIdx.x] = v;
    }
}
template <int K, int NT>
__device__ __forceinline__ void block_reduce_store(double (&s)[K], double* __restrict__ partial)
{
    block_reduce_store<K, NT>(s, partial, (i64)blockIdx.y * gridDim.x + blockIdx.x);
}

}  // namespace dsocp
