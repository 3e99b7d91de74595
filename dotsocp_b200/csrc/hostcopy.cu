// hostcopy.cu -- see hostcopy.h
#include "hostcopy.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <deque>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

namespace dsocp {

// ------------------------------------------------------------------------------------------------ worker pool
WorkerPool::WorkerPool(int nthreads)
{
    for (int i = 1; i < nthreads; i++) threads_.emplace_back([this] { worker(); });
}

WorkerPool::~WorkerPool()
{
    {
        std::lock_guard<std::mutex> lk(mu_);
        stop_ = true;
    }
    cv_.notify_all();
    for (auto& t : threads_) t.join();
}

void WorkerPool::worker()
{
    unsigned long seen = 0;
    std::unique_lock<std::mutex> lk(mu_);
    for (;;) {
        cv_.wait(lk, [&] { return stop_ || (gen_ != seen && fn_ != nullptr); });
        if (stop_) return;
        seen = gen_;
        while (fn_ && next_ < n_) {
            const int i = next_++;
            const std::function<void(int)>* fn = fn_;
            lk.unlock();
            (*fn)(i);
            lk.lock();
            if (--pending_ == 0) done_cv_.notify_all();
        }
    }
}

void WorkerPool::parallel_for(int n, const std::function<void(int)>& fn)
{
    if (n <= 0) return;
    if (n == 1 || threads_.empty()) {
        for (int i = 0; i < n; i++) fn(i);
        return;
    }
    // one job at a time: a second caller (another host thread driving its own session on this device) waits here
    // instead of overwriting the shared job state
    std::lock_guard<std::mutex> job(submit_mu_);
    std::unique_lock<std::mutex> lk(mu_);
    fn_ = &fn;
    n_ = n;
    next_ = 0;
    pending_ = n;
    gen_++;
    cv_.notify_all();
    while (next_ < n_) {
        const int i = next_++;
        lk.unlock();
        fn(i);
        lk.lock();
        --pending_;
    }
    done_cv_.wait(lk, [&] { return pending_ == 0; });
    fn_ = nullptr;
}

// ------------------------------------------------------------------------------------------------ staged copies
static int env_int(const char* name, int dflt)
{
    const char* e = getenv(name);
    if (!e || !*e) return dflt;
    const int v = atoi(e);
    return v > 0 ? v : dflt;
}

HostCopier::HostCopier()
{
    // threads: the host cores divided among the ranks that share the node (torchrun exports LOCAL_WORLD_SIZE)
    const int hw = (int)std::max(1u, std::thread::hardware_concurrency());
    const int lws = env_int("LOCAL_WORLD_SIZE", 1);
    const int nthr = env_int("DOTSOCP_COPY_THREADS", std::min(16, std::max(2, hw / lws)));
    chunk_ = (size_t)env_int("DOTSOCP_COPY_CHUNK_MB", 32) << 20;
    pool_ = new WorkerPool(nthr);
    ok_ = true;
    for (int b = 0; b < NBUF; b++) {
        if (cudaHostAlloc((void**)&pinned_[b], chunk_, cudaHostAllocDefault) != cudaSuccess ||
            cudaEventCreateWithFlags(&ev_[b], cudaEventDisableTiming) != cudaSuccess) {
            ok_ = false;
            cudaGetLastError();
            break;
        }
    }
}

HostCopier::~HostCopier()
{
    for (int b = 0; b < NBUF; b++) {
        if (pinned_[b]) cudaFreeHost(pinned_[b]);
        if (ev_[b]) cudaEventDestroy(ev_[b]);
    }
    delete pool_;
}

HostCopier* HostCopier::get()
{
    // one instance per device (its events belong to that device's context); deliberately never destroyed: at process
    // exit the CUDA context may already be gone
    static std::mutex mu;
    static HostCopier* inst[64] = {nullptr};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { cudaGetLastError(); dev = 0; }
    std::lock_guard<std::mutex> lk(mu);
    if (!inst[dev]) inst[dev] = new HostCopier();
    return inst[dev];
}

// memcpy with non-temporal stores: the destination (a pinned staging buffer about to be read by the copy engine, or the
// caller's array that nobody reads before the transfer is over) should not be pulled into the cache first -- an ordinary
// store reads every destination line before overwriting it, which costs a third of the host memory traffic.
void host_copy_streaming(char* dst, const char* src, size_t n)
{
#if defined(__SSE2__)
    if (n < 4096) { memcpy(dst, src, n); return; }
    const size_t head = (16 - ((size_t)dst & 15)) & 15;
    if (head) { memcpy(dst, src, head); dst += head; src += head; n -= head; }
    size_t i = 0;
    for (; i + 64 <= n; i += 64) {
        const __m128i a = _mm_loadu_si128((const __m128i*)(src + i)), b = _mm_loadu_si128((const __m128i*)(src + i + 16));
        const __m128i c = _mm_loadu_si128((const __m128i*)(src + i + 32)), d = _mm_loadu_si128((const __m128i*)(src + i + 48));
        _mm_stream_si128((__m128i*)(dst + i), a);
        _mm_stream_si128((__m128i*)(dst + i + 16), b);
        _mm_stream_si128((__m128i*)(dst + i + 32), c);
        _mm_stream_si128((__m128i*)(dst + i + 48), d);
    }
    _mm_sfence();
    if (i < n) memcpy(dst + i, src + i, n - i);
#else
    memcpy(dst, src, n);
#endif
}

void HostCopier::par_memcpy(char* dst, const char* src, size_t bytes)
{
    static const bool nt = [] { const char* e = getenv("DOTSOCP_COPY_NT"); return !(e && e[0] == '0'); }();   // 0: plain memcpy
    const int T = pool_->size();
    const size_t piece = ((bytes + T - 1) / T + 4095) & ~(size_t)4095;
    const int n = (int)((bytes + piece - 1) / piece);
    pool_->parallel_for(n, [&](int i) {
        const size_t o = (size_t)i * piece;
        if (nt) host_copy_streaming(dst + o, src + o, std::min(piece, bytes - o));
        else memcpy(dst + o, src + o, std::min(piece, bytes - o));
    });
}

// One implementation for flat and row-pitched transfers: the host side is always contiguous (rows * width bytes), cut into
// chunks of whole rows; the device side is either the same bytes (dpitch == width: one 1-D copy per chunk) or rows dpitch
// bytes apart (one 2-D copy per chunk).
int HostCopier::h2d_impl(char* d, size_t dpitch, const char* h, size_t width, size_t rows, cudaStream_t st)
{
    std::lock_guard<std::mutex> lk(mu_);
    const bool flat = dpitch == width;
    if (!flat && width > chunk_) return (int)cudaErrorInvalidValue;
    const size_t rpc = flat ? rows : chunk_ / width;                 // rows per chunk (flat: bytes are cut freely below)
    const size_t bytes = width * rows;
    for (size_t off = 0, r0 = 0; off < bytes;) {
        const size_t len = flat ? std::min(chunk_, bytes - off) : std::min(rpc, rows - r0) * width;
        const int b = cursor_;
        cursor_ = (cursor_ + 1) % NBUF;
        cudaError_t e;
        if (busy_[b] && (e = cudaEventSynchronize(ev_[b])) != cudaSuccess) return (int)e;
        par_memcpy(pinned_[b], h + off, len);
        if (flat) e = cudaMemcpyAsync(d + off, pinned_[b], len, cudaMemcpyHostToDevice, st);
        else e = cudaMemcpy2DAsync(d + r0 * dpitch, dpitch, pinned_[b], width, width, len / width, cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) return (int)e;
        if ((e = cudaEventRecord(ev_[b], st)) != cudaSuccess) return (int)e;
        busy_[b] = true;
        off += len;
        r0 += len / width;
    }
    return 0;
}

int HostCopier::d2h_impl(char* h, const char* d, size_t dpitch, size_t width, size_t rows, cudaStream_t st)
{
    std::lock_guard<std::mutex> lk(mu_);
    const bool flat = dpitch == width;
    if (!flat && width > chunk_) return (int)cudaErrorInvalidValue;
    const size_t rpc = flat ? rows : chunk_ / width;
    const size_t bytes = width * rows;
    struct Item { int b; size_t off, len; };
    std::deque<Item> fifo;
    auto drain_one = [&]() -> int {
        const Item it = fifo.front();
        fifo.pop_front();
        cudaError_t e = cudaEventSynchronize(ev_[it.b]);
        if (e != cudaSuccess) return (int)e;
        par_memcpy(h + it.off, pinned_[it.b], it.len);
        busy_[it.b] = false;
        return 0;
    };
    for (size_t off = 0, r0 = 0; off < bytes;) {
        const size_t len = flat ? std::min(chunk_, bytes - off) : std::min(rpc, rows - r0) * width;
        const int b = cursor_;
        cursor_ = (cursor_ + 1) % NBUF;
        cudaError_t e;
        if (busy_[b]) {   // still owned by an earlier h2d (its event), or by this call's pipeline (then it is the oldest item)
            if (!fifo.empty() && fifo.front().b == b) { int rc = drain_one(); if (rc) return rc; }
            else if ((e = cudaEventSynchronize(ev_[b])) != cudaSuccess) return (int)e;
        }
        if (flat) e = cudaMemcpyAsync(pinned_[b], d + off, len, cudaMemcpyDeviceToHost, st);
        else e = cudaMemcpy2DAsync(pinned_[b], width, d + r0 * dpitch, dpitch, width, len / width, cudaMemcpyDeviceToHost, st);
        if (e != cudaSuccess) return (int)e;
        if ((e = cudaEventRecord(ev_[b], st)) != cudaSuccess) return (int)e;
        busy_[b] = true;
        fifo.push_back({b, off, len});
        off += len;
        r0 += len / width;
        if ((int)fifo.size() >= NBUF - 1) { int rc = drain_one(); if (rc) return rc; }
    }
    while (!fifo.empty()) { int rc = drain_one(); if (rc) return rc; }
    return 0;
}

int HostCopier::h2d(void* dev, const void* host, size_t bytes, cudaStream_t st)
{
    return h2d_impl((char*)dev, bytes, (const char*)host, bytes, bytes ? 1 : 0, st);
}
int HostCopier::d2h(void* host, const void* dev, size_t bytes, cudaStream_t st)
{
    return d2h_impl((char*)host, (const char*)dev, bytes, bytes, bytes ? 1 : 0, st);
}
int HostCopier::h2d_rows(void* dev, size_t dpitch, const void* host, size_t width, size_t rows, cudaStream_t st)
{
    return h2d_impl((char*)dev, dpitch, (const char*)host, width, rows, st);
}
int HostCopier::d2h_rows(void* host, const void* dev, size_t dpitch, size_t width, size_t rows, cudaStream_t st)
{
    return d2h_impl((char*)host, (const char*)dev, dpitch, width, rows, st);
}

}  // namespace dsocp
