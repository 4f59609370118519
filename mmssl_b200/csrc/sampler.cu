// GPU triple sampler -- SURVEY section 8f "next" row 1.  Semantics of the reference's host sampler
// Data.sample (utility/load_data.py:153-191):
//   users : `batch` DISTINCT users drawn uniformly from the users that have >= 1 training item
//           (rd.sample without replacement, :154-155)
//   pos   : one uniform draw from the user's training items            (:160-171)
//   neg   : uniform item id, rejected while it is in the user's row    (:173-180)
// One CTA (batch <= 1024).  Distinctness without atomics races: rounds of "draw, atomicMin-claim,
// check" -- the winner of a claim is the smallest thread id, so the result depends only on
// (seed, step), never on scheduling; the claim table cleans itself up.  Counter-based RNG
// (splitmix64 of (seed, step, thread, draw)), so the kernel is replayable inside a CUDA graph with
// the step number read from device memory.
#include "common.cuh"
#include "../../include/mmssl_b200.h"

namespace mmssl {

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__device__ __forceinline__ uint32_t rnd32(uint64_t seed, uint32_t step, uint32_t tid, uint32_t draw) {
    return (uint32_t)(splitmix64(splitmix64(seed ^ ((uint64_t)step << 32 | tid)) + draw) >> 32);
}
// unbiased enough for sampling: 32-bit multiply-shift range reduction
__device__ __forceinline__ uint32_t below(uint32_t r, uint32_t n) { return (uint32_t)(((uint64_t)r * n) >> 32); }

__global__ void __launch_bounds__(1024) sample_triples_kernel(const int64_t* __restrict__ indptr,
                                                              const int64_t* __restrict__ indices,
                                                              const int64_t* __restrict__ exist, int64_t n_exist,
                                                              int64_t n_items, int batch, uint64_t seed,
                                                              const int32_t* __restrict__ step_dev, int32_t step_host,
                                                              int32_t* __restrict__ claim, int64_t* __restrict__ users,
                                                              int64_t* __restrict__ pos, int64_t* __restrict__ neg) {
    __shared__ int pending;
    const int t = threadIdx.x;
    const uint32_t step = (uint32_t)(step_dev ? *step_dev : step_host);
    const bool with_replacement = batch > n_exist;
    uint32_t draw = 0;
    int64_t slot = -1;                      // index into `exist`
    bool done = t >= batch;
    if (with_replacement && !done) { slot = below(rnd32(seed, step, t, draw++), (uint32_t)n_exist); done = true; }
    // ---- distinct users: rounds of claim-by-minimum ----
    // claim[x]: INT_MAX = free, -1 = taken in an earlier round, otherwise the smallest contender id.
    for (int round = 0; round < 64; ++round) {
        if (t == 0) pending = 0;
        __syncthreads();
        int64_t cand = -1;
        if (!done) {
            cand = below(rnd32(seed, step, t, draw++), (uint32_t)n_exist);
            atomicMin(&claim[cand], t);
        }
        __syncthreads();
        bool won = false;
        if (!done) {
            won = (claim[cand] == t);          // the smallest contender of a free slot wins it
            if (!won) atomicAdd(&pending, 1);
        }
        __syncthreads();
        const int left = pending;              // read before thread 0 may reset it for the next round
        if (won) { slot = cand; done = true; claim[cand] = -1; }
        __syncthreads();
        if (left == 0) break;
    }
    if (t < batch && !with_replacement && slot >= 0) claim[slot] = 0x7fffffff;   // leave the table clean
    if (t >= batch) return;
    if (slot < 0) slot = below(rnd32(seed, step, t, draw++), (uint32_t)n_exist);  // (never in practice: 64 rounds)
    const int64_t u = exist[slot];
    const int64_t b = indptr[u], e = indptr[u + 1];
    const uint32_t deg = (uint32_t)(e - b);
    const int64_t p = indices[b + below(rnd32(seed, step, t, draw++), deg)];
    int64_t ng = 0;
    for (int tries = 0; tries < 4096; ++tries) {
        ng = below(rnd32(seed, step, t, draw++), (uint32_t)n_items);
        int64_t lo = b, hi = e;             // binary search in the (sorted) row
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (indices[mid] < ng) lo = mid + 1; else hi = mid;
        }
        if (!(lo < e && indices[lo] == ng)) break;
    }
    users[t] = u; pos[t] = p; neg[t] = ng;
}

__global__ void fill_claim_kernel(int32_t* claim, int64_t n) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) claim[i] = 0x7fffffff;
}

}  // namespace mmssl

using namespace mmssl;

extern "C" int mmssl_sampler_init(int32_t* claim, int64_t n_exist, void* stream_) {
    if (n_exist == 0) return 0;
    fill_claim_kernel<<<(unsigned)((n_exist + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(claim, n_exist);
    MMSSL_LAUNCH_OK();
    return 0;
}

extern "C" int mmssl_sample_triples(const int64_t* indptr, const int64_t* indices, const int64_t* exist, int64_t n_exist,
                                    int64_t n_items, int batch, uint64_t seed, const int32_t* step_dev, int32_t step_host,
                                    int32_t* claim, int64_t* users, int64_t* pos, int64_t* neg, void* stream_) {
    MMSSL_REQUIRE(batch >= 1 && batch <= 1024, "one sampler launch draws at most 1024 triples");
    MMSSL_REQUIRE(n_exist >= 1 && n_exist < (1ll << 31) && n_items >= 1 && n_items < (1ll << 31), "bad sizes");
    sample_triples_kernel<<<1, 1024, 0, (cudaStream_t)stream_>>>(indptr, indices, exist, n_exist, n_items, batch, seed, step_dev,
                                                                  step_host, claim, users, pos, neg);
    MMSSL_LAUNCH_OK();
    return 0;
}
