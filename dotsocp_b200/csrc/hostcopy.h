// hostcopy.h -- pageable host memory <-> HBM at PCIe speed.
//
// The reference-facing entry points take ordinary (pageable) host arrays: MATLAB / numpy own them.  cudaMemcpy from
// pageable memory is staged by the driver through one thread and reaches a small fraction of the PCIe rate, which made
// the transfers -- not the iterations -- the largest part of a dotsocp_solve_level() call.
// HostCopier stages through a ring of pinned buffers itself: a small pool of host threads copies pageable <-> pinned
// while the copy engine moves the previous chunk, so the transfer runs at min(host memcpy, PCIe) bandwidth.
#pragma once
#include <cuda_runtime.h>

#include <condition_variable>
#include <cstddef>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace dsocp {

class WorkerPool {
public:
    explicit WorkerPool(int nthreads);
    ~WorkerPool();
    int size() const { return (int)threads_.size() + 1; }
    // fn(i) for i in [0, n): the calling thread takes part; returns when all are done
    void parallel_for(int n, const std::function<void(int)>& fn);

private:
    void worker();
    std::vector<std::thread> threads_;
    std::mutex mu_, submit_mu_;   // submit_mu_ is held for the whole of a parallel_for (serialises concurrent callers)
    std::condition_variable cv_, done_cv_;
    const std::function<void(int)>* fn_ = nullptr;
    int n_ = 0, next_ = 0, pending_ = 0;
    unsigned long gen_ = 0;
    bool stop_ = false;
};

// memcpy with non-temporal stores (x86-64; plain memcpy elsewhere)
void host_copy_streaming(char* dst, const char* src, size_t n);

class HostCopier {
public:
    // per-device instance of the calling thread's current device (created on first use; pinned ring + thread pool are
    // reused by every session on that device)
    static HostCopier* get();
    // both return a cudaError_t-compatible code (0 = ok); `st` orders the transfer against the caller's other work and is
    // synchronised before d2h returns (h2d returns once the last chunk has been queued and its staging buffer may be reused
    // only after the copy, which later calls check by event).
    int h2d(void* dev, const void* host, size_t bytes, cudaStream_t st);
    int d2h(void* host, const void* dev, size_t bytes, cudaStream_t st);
    // `rows` rows of `width` bytes: packed on the host, `dpitch` bytes apart on the device (the pitched device layout of the
    // staggered arrays, common.cuh); same staging, the copy engine does the strided side (cudaMemcpy2DAsync)
    int h2d_rows(void* dev, size_t dpitch, const void* host, size_t width, size_t rows, cudaStream_t st);
    int d2h_rows(void* host, const void* dev, size_t dpitch, size_t width, size_t rows, cudaStream_t st);
    WorkerPool& pool() { return *pool_; }
    bool ok() const { return ok_; }
    ~HostCopier();

private:
    HostCopier();
    void par_memcpy(char* dst, const char* src, size_t bytes);
    int h2d_impl(char* dev, size_t dpitch, const char* host, size_t width, size_t rows, cudaStream_t st);
    int d2h_impl(char* host, const char* dev, size_t dpitch, size_t width, size_t rows, cudaStream_t st);
    static constexpr int NBUF = 4;
    size_t chunk_ = 0;
    char* pinned_[NBUF] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_[NBUF] = {nullptr, nullptr, nullptr, nullptr};
    bool busy_[NBUF] = {false, false, false, false};
    int cursor_ = 0;
    WorkerPool* pool_ = nullptr;
    std::mutex mu_;
    bool ok_ = false;
};

}  // namespace dsocp
