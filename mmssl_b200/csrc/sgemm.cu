// CUDA-core fp32 GEMM (64x64x16 tiles, 4x4 per thread).  Not the projection's production path
// (that is the tcgen05 kernel in proj_tc.cu); used for the small d x d mixing GEMMs of the
// "attention" closed form (Models.py:139-169), their gradients, and as the
// independent verification path for the tensor-core projection.
#include "common.cuh"
#include "../../include/mmssl_b200.h"

namespace mmssl {

constexpr int BM = 64, BN = 64, BK = 16;

template <bool TA, bool TB>
__global__ void __launch_bounds__(256) sgemm_kernel(int64_t M, int64_t N, int64_t K, float alpha,
                                                    const float* __restrict__ A, int64_t lda,
                                                    const float* __restrict__ B, int64_t ldb, float beta,
                                                    float* __restrict__ C, int64_t ldc, int64_t k_per_split) {
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];
    const int t = threadIdx.x;
    const int64_t m0 = blockIdx.x * (int64_t)BM, n0 = blockIdx.y * (int64_t)BN;
    const int64_t kbeg = blockIdx.z * k_per_split;
    const int64_t kend = min(K, kbeg + k_per_split);
    const int ty = t / 16, tx = t % 16;
    float acc[4][4] = {};
    for (int64_t k0 = kbeg; k0 < kend; k0 += BK) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int mm, kk;
            if (!TA) { kk = t % 16; mm = t / 16 + 16 * i; } else { mm = t % 64; kk = t / 64 + 4 * i; }
            const int64_t gm = m0 + mm, gk = k0 + kk;
            float v = 0.f;
            if (gm < M && gk < kend) v = TA ? A[gk * lda + gm] : A[gm * lda + gk];
            As[kk][mm] = v;
            int nn, kb;
            if (!TB) { nn = t % 64; kb = t / 64 + 4 * i; } else { kb = t % 16; nn = t / 16 + 16 * i; }
            const int64_t gn = n0 + nn, gkb = k0 + kb;
            float w = 0.f;
            if (gn < N && gkb < kend) w = TB ? B[gn * ldb + gkb] : B[gkb * ldb + gn];
            Bs[kb][nn] = w;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w};
            const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t gm = m0 + ty * 4 + i;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t gn = n0 + tx * 4 + j;
            if (gn >= N) continue;
            float* cp = C + gm * ldc + gn;
            if (gridDim.z > 1) atomicAdd(cp, alpha * acc[i][j]);
            else *cp = alpha * acc[i][j] + (beta != 0.f ? beta * (*cp) : 0.f);
        }
    }
}

}  // namespace mmssl

using namespace mmssl;

extern "C" int mmssl_sgemm(int trans_a, int trans_b, int64_t m, int64_t n, int64_t k, float alpha, const float* a,
                           int64_t lda, const float* b, int64_t ldb, float beta, float* c, int64_t ldc, int split_k,
                           void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    MMSSL_REQUIRE(m >= 0 && n >= 0 && k >= 0 && split_k >= 1, "bad sizes");
    if (m == 0 || n == 0) return 0;
    dim3 grid((unsigned)((m + BM - 1) / BM), (unsigned)((n + BN - 1) / BN), (unsigned)split_k);
    MMSSL_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "grid too large");
    int64_t kps = (k + split_k - 1) / split_k;
    kps = (kps + BK - 1) / BK * BK;
    if (kps == 0) kps = BK;
#define LAUNCH(TA, TB) sgemm_kernel<TA, TB><<<grid, 256, 0, st>>>(m, n, k, alpha, a, lda, b, ldb, beta, c, ldc, kps)
    if (!trans_a && !trans_b) LAUNCH(false, false);
    else if (!trans_a && trans_b) LAUNCH(false, true);
    else if (trans_a && !trans_b) LAUNCH(true, false);
    else LAUNCH(true, true);
#undef LAUNCH
    MMSSL_LAUNCH_OK();
    return 0;
}
