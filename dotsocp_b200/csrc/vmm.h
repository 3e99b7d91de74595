// vmm.h -- sparse device arrays (CUDA virtual memory management) and a lazily loaded NCCL.
//
// Time-slab partition: every array keeps its GLOBAL index space (so the kernels index exactly as on one GPU) but only
// the windows a slab touches -- its own time levels plus one ghost level on each side -- are backed by physical
// memory.  The driver entry points are fetched with cudaGetDriverEntryPoint and NCCL with dlopen, so the shared
// library has no load-time dependency on libcuda/libnccl (it must load on a machine without a GPU).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include <utility>
#include <vector>

namespace dsocp {

// A device array of `count` doubles of which only the listed element ranges are backed.
class SparseArray {
public:
    SparseArray() = default;
    ~SparseArray() { release(); }
    SparseArray(const SparseArray&) = delete;
    SparseArray& operator=(const SparseArray&) = delete;
    // windows: [begin, end) element ranges; dense=true backs everything with one plain cudaMalloc.
    // returns cudaSuccess or an error; err_text gets a description.
    int create(long long count, const std::vector<std::pair<long long, long long>>& windows, bool dense, int device,
               const char** err_text);
    void release();
    double* ptr() const { return reinterpret_cast<double*>(base_); }
    size_t backed_bytes() const { return backed_; }

private:
    CUdeviceptr base_ = 0;
    size_t va_size_ = 0;
    size_t backed_ = 0;
    bool dense_ = false;
    struct Chunk { CUmemGenericAllocationHandle h; size_t off, size; };
    std::vector<Chunk> chunks_;
};

// FP64 tensor map (TMA descriptor) of a `rank`-dimensional view: dims[0] is the contiguous dimension, strides_bytes[i] the
// distance between consecutive indices of dimension i+1 (rank-1 entries, multiples of 16), box[] the tile one copy moves.
// Out-of-range parts of a box are filled with zeros.  Returns 0, or -1 when the driver lacks cuTensorMapEncodeTiled / rejects it.
int make_tensor_map_f64(CUtensorMap* out, const void* base, int rank, const unsigned long long* dims,
                        const unsigned long long* strides_bytes, const unsigned* box);

// ---- NCCL, loaded at run time --------------------------------------------------------------------------------------
struct NcclId { char internal[128]; };
struct NcclApi {
    bool ok = false;
    const char* why = "";
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, NcclId /* ncclUniqueId by value */, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
const NcclApi& nccl_api();
enum { NCCL_INT8 = 0, NCCL_FLOAT64 = 8, NCCL_SUM = 0 };

}  // namespace dsocp
