// vmm.cu -- see vmm.h
#include "vmm.h"

#include <dlfcn.h>

#include <algorithm>
#include <cstdio>
#include <cstring>

namespace dsocp {

namespace {
struct DrvApi {
    bool ok = false;
    CUresult (*MemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*MemAddressFree)(CUdeviceptr, size_t) = nullptr;
    CUresult (*MemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
    CUresult (*MemRelease)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*MemUnmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
    CUresult (*MemGetAllocationGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
};

template <typename F>
bool entry(const char* name, F& fn)
{
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &qr) != cudaSuccess || p == nullptr) {
        cudaGetLastError();
        return false;
    }
    fn = reinterpret_cast<F>(p);
    return true;
}

const DrvApi& drv()
{
    static DrvApi a = [] {
        DrvApi d;
        d.ok = entry("cuMemAddressReserve", d.MemAddressReserve) && entry("cuMemAddressFree", d.MemAddressFree) &&
               entry("cuMemCreate", d.MemCreate) && entry("cuMemRelease", d.MemRelease) && entry("cuMemMap", d.MemMap) &&
               entry("cuMemUnmap", d.MemUnmap) && entry("cuMemSetAccess", d.MemSetAccess) &&
               entry("cuMemGetAllocationGranularity", d.MemGetAllocationGranularity);
        return d;
    }();
    return a;
}
}  // namespace

int SparseArray::create(long long count, const std::vector<std::pair<long long, long long>>& windows, bool dense, int device,
                        const char** err_text)
{
    release();
    static const char* e_none = "";
    if (err_text) *err_text = e_none;
    if (count <= 0) count = 1;
    if (dense) {
        void* p = nullptr;
        cudaError_t e = cudaMalloc(&p, (size_t)count * sizeof(double));
        if (e != cudaSuccess) {
            cudaGetLastError();
            if (err_text) *err_text = cudaGetErrorString(e);
            return (int)e;
        }
        base_ = (CUdeviceptr)p;
        dense_ = true;
        backed_ = (size_t)count * sizeof(double);
        return 0;
    }
    const DrvApi& d = drv();
    if (!d.ok) {
        if (err_text) *err_text = "CUDA virtual memory management entry points are unavailable";
        return -1;
    }
    cudaFree(0);   // make sure the primary context exists
    CUmemAllocationProp prop;
    memset(&prop, 0, sizeof prop);
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = device;
    size_t gran = 0;
    if (d.MemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM) != CUDA_SUCCESS || gran == 0) {
        if (err_text) *err_text = "cuMemGetAllocationGranularity failed";
        return -1;
    }
    const size_t bytes = (size_t)count * sizeof(double);
    va_size_ = (bytes + gran - 1) / gran * gran;
    if (d.MemAddressReserve(&base_, va_size_, gran, 0, 0) != CUDA_SUCCESS) {
        base_ = 0;
        if (err_text) *err_text = "cuMemAddressReserve failed";
        return -1;
    }
    // union of the windows in units of pages
    std::vector<std::pair<size_t, size_t>> pg;
    for (auto& w : windows) {
        long long b = std::max(0LL, w.first), e = std::min(count, w.second);
        if (e <= b) continue;
        pg.emplace_back((size_t)b * sizeof(double) / gran, ((size_t)e * sizeof(double) + gran - 1) / gran);
    }
    std::sort(pg.begin(), pg.end());
    std::vector<std::pair<size_t, size_t>> runs;
    for (auto& p : pg) {
        if (!runs.empty() && p.first <= runs.back().second) runs.back().second = std::max(runs.back().second, p.second);
        else runs.push_back(p);
    }
    CUmemAccessDesc acc;
    memset(&acc, 0, sizeof acc);
    acc.location = prop.location;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    for (auto& r : runs) {
        Chunk c;
        c.off = r.first * gran;
        c.size = (r.second - r.first) * gran;
        if (d.MemCreate(&c.h, c.size, &prop, 0) != CUDA_SUCCESS) {
            if (err_text) *err_text = "cuMemCreate failed (out of device memory?)";
            release();
            return 2;   // cudaErrorMemoryAllocation
        }
        if (d.MemMap(base_ + c.off, c.size, 0, c.h, 0) != CUDA_SUCCESS) {
            d.MemRelease(c.h);
            if (err_text) *err_text = "cuMemMap failed";
            release();
            return -1;
        }
        chunks_.push_back(c);
        if (d.MemSetAccess(base_ + c.off, c.size, &acc, 1) != CUDA_SUCCESS) {
            if (err_text) *err_text = "cuMemSetAccess failed";
            release();
            return -1;
        }
        backed_ += c.size;
    }
    return 0;
}

void SparseArray::release()
{
    if (!base_) return;
    if (dense_) {
        cudaFree((void*)base_);
    } else {
        const DrvApi& d = drv();
        for (auto& c : chunks_) {
            d.MemUnmap(base_ + c.off, c.size);
            d.MemRelease(c.h);
        }
        d.MemAddressFree(base_, va_size_);
    }
    chunks_.clear();
    base_ = 0;
    va_size_ = backed_ = 0;
    dense_ = false;
}

int make_tensor_map_f64(CUtensorMap* out, const void* base, int rank, const unsigned long long* dims,
                        const unsigned long long* strides_bytes, const unsigned* box)
{
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn fn = [] { EncodeFn f = nullptr; return entry("cuTensorMapEncodeTiled", f) ? f : (EncodeFn) nullptr; }();
    if (!fn || rank < 1 || rank > 5) return -1;
    cuuint64_t d[5], st[4];
    cuuint32_t b[5], es[5];
    for (int i = 0; i < rank; i++) { d[i] = dims[i]; b[i] = box[i]; es[i] = 1; }
    for (int i = 0; i + 1 < rank; i++) st[i] = strides_bytes[i];
    const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, (cuuint32_t)rank, const_cast<void*>(base), d, st, b, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -1;
}

const NcclApi& nccl_api()
{
    static NcclApi api = [] {
        NcclApi a;
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) {
            a.why = "libnccl.so.2 could not be loaded";
            return a;
        }
#define SYM(field, name)                                             \
    *(void**)(&a.field) = dlsym(h, name);                            \
    if (!a.field) {                                                  \
        a.why = "libnccl is missing " name;                         \
        return a;                                                    \
    }
        SYM(GetUniqueId, "ncclGetUniqueId")
        SYM(CommInitRank, "ncclCommInitRank")
        SYM(CommDestroy, "ncclCommDestroy")
        SYM(Send, "ncclSend")
        SYM(Recv, "ncclRecv")
        SYM(AllReduce, "ncclAllReduce")
        SYM(AllGather, "ncclAllGather")
        SYM(GroupStart, "ncclGroupStart")
        SYM(GroupEnd, "ncclGroupEnd")
        SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
        a.ok = true;
        return a;
    }();
    return api;
}

}  // namespace dsocp
