// Host check of the worker pool behind the staged host<->device copies (dotsocp_b200/csrc/hostcopy.cu): every index is
// visited exactly once, for many back-to-back jobs of varying size (generation / wake-up races), including n < threads.
#include <atomic>
#include <cstdio>
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
    printf("POOL_HOST_OK\n");
    return 0;
}
