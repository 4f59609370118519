// Modality-graph bookkeeping of the reference's training step (main.py:372-405) on the device -- SURVEY section 8f
// "next" row 2.  Per step the reference either
//   * takes torch.topk(G_*_u_sim, k = int(n_items * m_topk_rate)) of every batch row, moves it to the host and appends
//     python lists  x += users.repeat(1, k).view(-1),  y += ids.view(-1)      (main.py:397-402), or
//   * turns the collected lists into scipy CSR (duplicates summed), normalises rows by (rowsum + 1e-8)^-1/2 and
//     converts both the matrix and its (separately normalised) transpose to torch sparse tensors (main.py:379-391).
// Here the pairs never leave the GPU:
//   topk_rows_kernel      one CTA per row, k rounds of "largest key below the previous pick"; keys are (orderable fp32
//                         score, ~column) packed in 64 bits and unique, so the result is deterministic and needs no
//                         exclusion list; equal scores keep the lower column first.  k is tiny in the reference's
//                         configurations (0 at Baby, 1 at Sports with the default rate 1e-4), rows are L2 resident.
//   pair_append_kernel    x[j] = users[j % B], y[j] = ids[j]  -- the reference's pairing quirk, kept: the x list tiles the
//                         whole user vector k times while the y list is row-major, so pair j joins user j % B with an item
//                         of row j / k.
//   degree_count / degree_scale   per-edge values (deg(row) + 1e-8)^-1/2 for the matrix and for its transpose (evaluated in
//                         double like numpy, rounded to fp32 like sparse_mx_to_torch_sparse_tensor); duplicates stay
//                         separate COO entries -- the CSR builder keeps them and the SpMM sums them.
#include "common.cuh"
#include "../../include/mmssl_b200.h"

namespace mmssl {

__device__ __forceinline__ uint64_t topk_key(float s, uint32_t col) {
    uint32_t b = __float_as_uint(s + 0.0f);                     // -0 -> +0 so that equal scores tie
    b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
    return ((uint64_t)b << 32) | (uint64_t)(0xFFFFFFFFu - col);
}

__global__ void __launch_bounds__(256) topk_rows_kernel(const float* __restrict__ x, int64_t ldx, int64_t w, int k,
                                                        int64_t* __restrict__ ids) {
    __shared__ uint64_t sh[8];
    __shared__ uint64_t prev_s;
    const float* row = x + (int64_t)blockIdx.x * ldx;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint64_t prev = ~0ull;
    for (int r = 0; r < k; ++r) {
        uint64_t best = 0ull;                                    // below every real key (see eval.cu)
        for (int64_t c = tid; c < w; c += 256) {
            const float v = row[c];
            const uint64_t key = (v == v) ? topk_key(v, (uint32_t)c) : (uint64_t)(0xFFFFFFFFu - (uint32_t)c);   // NaN ranks last
            if (key < prev && key > best) best = key;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const uint64_t other = __shfl_xor_sync(0xffffffffu, best, o);
            best = other > best ? other : best;
        }
        if (lane == 0) sh[warp] = best;
        __syncthreads();
        if (tid == 0) {
            uint64_t b = sh[0];
#pragma unroll
            for (int i = 1; i < 8; ++i) b = sh[i] > b ? sh[i] : b;
            prev_s = b;
            ids[(int64_t)blockIdx.x * k + r] = (int64_t)(0xFFFFFFFFu - (uint32_t)b);
        }
        __syncthreads();
        prev = prev_s;
    }
}

__global__ void __launch_bounds__(256) pair_append_kernel(const int64_t* __restrict__ users, int64_t batch, const int64_t* __restrict__ ids,
                                                          int64_t total, int64_t* __restrict__ x, int64_t* __restrict__ y) {
    const int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (j >= total) return;
    x[j] = users[j % batch];
    y[j] = ids[j];
}

__global__ void __launch_bounds__(256) degree_count_kernel(const int64_t* __restrict__ idx, int64_t n, int32_t* __restrict__ deg) {
    const int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (j < n) atomicAdd(&deg[idx[j]], 1);
}
__global__ void __launch_bounds__(256) degree_scale_kernel(const int64_t* __restrict__ idx, int64_t n, const int32_t* __restrict__ deg,
                                                           float* __restrict__ vals) {
    const int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (j < n) vals[j] = (float)pow((double)deg[idx[j]] + 1e-8, -0.5);
}

}  // namespace mmssl

using namespace mmssl;

extern "C" int mmssl_topk_rows(const float* x, int64_t ldx, int64_t rows, int64_t w, int k, int64_t* ids, void* stream_) {
    MMSSL_REQUIRE(rows >= 0 && w >= 0 && w < (1ll << 31) && k >= 0 && k <= w, "bad sizes (k must not exceed the row width)");
    if (rows == 0 || k == 0) return 0;
    topk_rows_kernel<<<(unsigned)rows, 256, 0, (cudaStream_t)stream_>>>(x, ldx, w, k, ids);
    MMSSL_LAUNCH_OK();
    return 0;
}

extern "C" int mmssl_pair_append(const int64_t* users, int64_t batch, const int64_t* ids, int k, int64_t* x, int64_t* y, void* stream_) {
    MMSSL_REQUIRE(batch >= 0 && k >= 0, "bad sizes");
    const int64_t total = batch * k;
    if (total == 0) return 0;
    pair_append_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(users, batch, ids, total, x, y);
    MMSSL_LAUNCH_OK();
    return 0;
}

extern "C" int mmssl_degree_values(const int64_t* idx, int64_t n, int64_t n_rows, int32_t* deg_scratch, float* vals, void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    MMSSL_REQUIRE(n >= 0 && n_rows >= 0, "bad sizes");
    if (n == 0) return 0;
    MMSSL_CUDA(cudaMemsetAsync(deg_scratch, 0, sizeof(int32_t) * (size_t)n_rows, st));
    const unsigned grid = (unsigned)((n + 255) / 256);
    degree_count_kernel<<<grid, 256, 0, st>>>(idx, n, deg_scratch);
    MMSSL_LAUNCH_OK();
    degree_scale_kernel<<<grid, 256, 0, st>>>(idx, n, deg_scratch, vals);
    MMSSL_LAUNCH_OK();
    return 0;
}
