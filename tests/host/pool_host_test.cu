// Host check of the worker pool behind the staged host<->device copies (dotsocp_b200/csrc/hostcopy.cu): every index is
// visited exactly once, for many back-to-back jobs of varying size (generation / wake-up races), including n < threads.
#include <atomic>
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <vector>

#include "hostcopy.h"

using namespace dsocp;

int main()
{
    for (int threads : {1, 2, 5, 8}) {
        WorkerPool pool(threads);
        if (pool.size() != threads) { printf("size %d != %d\n", pool.size(), threads); return 1; }
        for (int job = 0; job < 3000; job++) {
            const int n = job % 37;
            std::vector<std::atomic<int>> hits(n > 0 ? n : 1);
            for (auto& h : hits) h.store(0);
            long long sum = 0;
            std::atomic<long long> asum(0);
            pool.parallel_for(n, [&](int i) { hits[i].fetch_add(1); asum.fetch_add(i + 1); });
            for (int i = 0; i < n; i++) {
                if (hits[i].load() != 1) { printf("threads %d job %d: index %d visited %d times\n", threads, job, i, hits[i].load()); return 1; }
                sum += i + 1;
            }
            if (asum.load() != sum) { printf("threads %d job %d: sum mismatch\n", threads, job); return 1; }
        }
    }
    // streaming copy: every size / alignment combination around the vector and head/tail boundaries, and a large block
    {
        std::vector<char> src(1 << 22), dst(1 << 22), ref(1 << 22);
        for (size_t i = 0; i < src.size(); i++) src[i] = (char)(i * 131 + 7);
        const size_t sizes[] = {0, 1, 15, 16, 17, 63, 64, 65, 4095, 4096, 4097, 4111, 8191, 100003, (1 << 21) + 5};
        for (size_t n : sizes)
            for (int da = 0; da < 17; da += 4)
                for (int sa = 0; sa < 17; sa += 5) {
                    std::fill(dst.begin(), dst.end(), (char)0x55);
                    ref = dst;
                    memcpy(ref.data() + 64 + da, src.data() + 32 + sa, n);
                    host_copy_streaming(dst.data() + 64 + da, src.data() + 32 + sa, n);
                    if (dst != ref) { printf("streaming copy differs: n=%zu da=%d sa=%d\n", n, da, sa); return 1; }
                }
    }
    printf("POOL_HOST_OK\n");
    return 0;
}
