// Work partitioning of one batched cycle over the ranks (GPUs) of a job -- host only, no CUDA.
//
// Level 2 of SURVEY.md section 8(e): inside a cycle the Step-1 problems ((fit, module) pairs,
// independent: R/scregclust.R:1332-1388) are spread over the ranks, and Step 2a/2b (NNLS columns,
// residuals, votes: independent per target gene when prior_weight == 0, R/scregclust.R:1533-1615,
// src/allocation.cpp:39-56) is split into contiguous target-gene slices.  Every rank evaluates the
// same plan from the same inputs, so no message is needed to agree on it.
#pragma once
#include <algorithm>
#include <cstdint>
#include <numeric>
#include <vector>

namespace scrc {

// Longest-processing-time-first assignment of `n` problems with the given costs to `world` ranks:
// problems in order of decreasing cost (ties: lower index first) go to the currently least loaded
// rank (ties: lower rank).  Deterministic.  owner[i] in [0, world).
inline void shard_assign(const double* cost, int n, int world, int* owner) {
    if (world <= 1) { std::fill(owner, owner + n, 0); return; }
    std::vector<int> order(n);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cost[a] > cost[b]; });
    std::vector<double> load(world, 0.0);
    for (int i : order) {
        int best = 0;
        for (int r = 1; r < world; ++r) if (load[r] < load[best]) best = r;
        owner[i] = best;
        load[best] += cost[i];
    }
}

// Contiguous slice [begin, end) of `n` items for `rank`, in units of `align` items (the vote kernel
// works on 32-target tiles); the slices of all ranks tile [0, n) exactly, earlier ranks take the
// larger shares.
inline void shard_slice(int64_t n, int64_t align, int rank, int world, int64_t* begin, int64_t* end) {
    if (world <= 1) { *begin = 0; *end = n; return; }
    const int64_t units = (n + align - 1) / align;
    const int64_t base = units / world, extra = units % world;
    const int64_t u0 = base * rank + std::min<int64_t>(rank, extra);
    const int64_t u1 = u0 + base + (rank < extra ? 1 : 0);
    *begin = std::min(u0 * align, n);
    *end = std::min(u1 * align, n);
}

}  // namespace scrc
