// Row-sharded hot step (mmssl_b200/rowshard_step.py): the two kernels that connect the batch -- global
// user / item ids (main.py:368-370, :411-412 index the full tables with them) -- to a rank's row block [lo, hi).
//   gather_owned      out[j] = table[idx[j] - lo] if lo <= idx[j] < hi else 0     -> summed over ranks by one all-reduce
//                     this gives every rank the batch rows of the full table
//   scatter_add_owned table[idx[j] - lo] += src[j] for the owned j only           -> the loss kernels' gradient rows
//                     return to the rank that owns the row; duplicates (the same item drawn twice) accumulate atomically
// One thread per float4 of a row (d % 4 == 0, rows 16-byte aligned like every table of the library).
#include "common.cuh"
#include "../../include/mmssl_b200.h"

namespace mmssl {

__global__ void __launch_bounds__(256) gather_owned_kernel(const float* __restrict__ table, int64_t ld, const int64_t* __restrict__ idx,
                                                           int64_t lo, int64_t hi, int64_t n, int d4, float* __restrict__ out, int64_t ldo) {
    const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (t >= n * d4) return;
    const int64_t j = t / d4;
    const int c = (int)(t - j * d4) * 4;
    const int64_t r = idx[j];
    float4 v = f4zero();
    if (r >= lo && r < hi) v = ld4(table + (r - lo) * ld + c);
    st4(out + j * ldo + c, v);
}

__global__ void __launch_bounds__(256) scatter_add_owned_kernel(float* __restrict__ table, int64_t ld, const int64_t* __restrict__ idx,
                                                                int64_t lo, int64_t hi, int64_t n, int d4, const float* __restrict__ src,
                                                                int64_t lds) {
    const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (t >= n * d4) return;
    const int64_t j = t / d4;
    const int c = (int)(t - j * d4) * 4;
    const int64_t r = idx[j];
    if (r < lo || r >= hi) return;
    atomicAdd(reinterpret_cast<float4*>(table + (r - lo) * ld + c), ld4(src + j * lds + c));     // 128-bit reduction (sm_90+)
}

}  // namespace mmssl

using namespace mmssl;

extern "C" int mmssl_gather_owned(const float* table, int64_t ld, const int64_t* idx, int64_t lo, int64_t hi, int64_t n, int d, float* out,
                                  int64_t ldo, void* stream_) {
    MMSSL_REQUIRE(d % 4 == 0 && ld % 4 == 0 && ldo % 4 == 0 && aligned16(table) && aligned16(out), "alignment");
    MMSSL_REQUIRE(lo >= 0 && hi >= lo && n >= 0, "bad range");
    if (n == 0) return 0;
    gather_owned_kernel<<<(unsigned)((n * (d / 4) + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(table, ld, idx, lo, hi, n, d / 4, out, ldo);
    MMSSL_LAUNCH_OK();
    return 0;
}

extern "C" int mmssl_scatter_add_owned(float* table, int64_t ld, const int64_t* idx, int64_t lo, int64_t hi, int64_t n, int d,
                                       const float* src, int64_t lds, void* stream_) {
    MMSSL_REQUIRE(d % 4 == 0 && ld % 4 == 0 && lds % 4 == 0 && aligned16(table) && aligned16(src), "alignment");
    MMSSL_REQUIRE(lo >= 0 && hi >= lo && n >= 0, "bad range");
    if (n == 0) return 0;
    scatter_add_owned_kernel<<<(unsigned)((n * (d / 4) + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(table, ld, idx, lo, hi, n, d / 4, src, lds);
    MMSSL_LAUNCH_OK();
    return 0;
}
