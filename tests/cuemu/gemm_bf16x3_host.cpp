// Host stand-in for the tcgen05 GEMM entry points of csrc/proj_tc.cu, for the cuemu library ONLY (tests).
// The tensor-core kernel itself cannot be emulated (TMA, TMEM, tcgen05 PTX) and has its own GPU test
// (tests/test_gpu_ops.py::test_gemm_bf16x3_tensor_core).  What this file provides is the CONTRACT of the entry points
// as include/mmssl_b200.h states it -- bf16 hi/lo K-major operands, lo*lo dropped, fp32 accumulation, the same split-K
// slicing and partial[s][m][n] layout -- so that everything around the GEMM (operand splits, leading dimensions,
// epilogues, the callers' shapes) is exercised on the CPU.
#include <stdint.h>
#include <string.h>

#include <vector>

#include "common.cuh"
#include "../../include/mmssl_b200.h"

namespace {
constexpr int kBlockM = 128, kBlockK = 64, kMaxSplit = 64;
inline float bf(uint16_t b) {
    const uint32_t u = (uint32_t)b << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}
int choose_split(int64_t m, int64_t n, int64_t k) {
    const int64_t total_kb = (k + kBlockK - 1) / kBlockK, m_tiles = (m + kBlockM - 1) / kBlockM;
    const int64_t slots = (int64_t)mmssl::kNumSMs * (n == 64 ? 2 : 1);
    int64_t split = slots / m_tiles;
    if (split > kMaxSplit) split = kMaxSplit;
    if (split > total_kb) split = total_kb;
    if (split < 1) split = 1;
    const int64_t per = (total_kb + split - 1) / split;
    return (int)((total_kb + per - 1) / per);
}
}  // namespace

extern "C" int64_t mmssl_gemm_bf16x3_workspace_floats(int64_t m, int64_t n, int64_t k, int* split_k_out) {
    const int split = choose_split(m, n, k);
    if (split_k_out) *split_k_out = split;
    return (int64_t)split * m * n;
}

extern "C" int mmssl_gemm_bf16x3(const uint16_t* a_hi, const uint16_t* a_lo, int64_t lda, const uint16_t* b_hi, const uint16_t* b_lo,
                                 int64_t ldb, int64_t m, int64_t n, int64_t k, int split_k, float* partial, void*) {
    MMSSL_REQUIRE(n == 64 || n == 128 || n == 256, "n (embedding width) must be 64, 128 or 256");
    MMSSL_REQUIRE(m >= 1 && k >= 1, "bad m / k");
    MMSSL_REQUIRE(lda % 8 == 0 && ldb % 8 == 0 && lda >= k && ldb >= k, "lda/ldb must be >= k and multiples of 8 (16-byte TMA strides)");
    MMSSL_REQUIRE(mmssl::aligned16(a_hi) && mmssl::aligned16(a_lo) && mmssl::aligned16(b_hi) && mmssl::aligned16(b_lo) &&
                      mmssl::aligned16(partial), "alignment");
    const int total_kb = (int)((k + kBlockK - 1) / kBlockK);
    MMSSL_REQUIRE(split_k >= 1 && split_k <= total_kb, "split_k out of range");
    const int per = (total_kb + split_k - 1) / split_k;
    MMSSL_REQUIRE((int64_t)per * (split_k - 1) < total_kb, "split_k leaves an empty K slice (use mmssl_gemm_bf16x3_workspace_floats)");
    for (int s = 0; s < split_k; ++s) {
        const int64_t k0 = (int64_t)s * per * kBlockK, k1 = (k0 + (int64_t)per * kBlockK < k) ? k0 + (int64_t)per * kBlockK : k;
        for (int64_t i = 0; i < m; ++i)
            for (int64_t j = 0; j < n; ++j) {
                float acc = 0.f;
                for (int64_t kk = k0; kk < k1; ++kk) {          // TMA zero-fills beyond k; the operands' padding is never read here
                    const float ah = bf(a_hi[i * lda + kk]), al = bf(a_lo[i * lda + kk]);
                    const float bh = bf(b_hi[j * ldb + kk]), bl = bf(b_lo[j * ldb + kk]);
                    acc += ah * bh;
                    acc += ah * bl;
                    acc += al * bh;
                }
                partial[((int64_t)s * m + i) * n + j] = acc;
            }
    }
    return 0;
}

// Contract of csrc/gemm_wide.cu:mmssl_gemm_bf16x3_wide (general n, no split-K, alpha / accumulate epilogue).
extern "C" int mmssl_gemm_bf16x3_wide(const uint16_t* a_hi, const uint16_t* a_lo, int64_t lda, const uint16_t* b_hi, const uint16_t* b_lo,
                                      int64_t ldb, int64_t m, int64_t n, int64_t k, float alpha, int accumulate, float* c, int64_t ldc,
                                      void*) {
    MMSSL_REQUIRE(m >= 1 && n >= 1 && k >= 1, "bad m / n / k");
    MMSSL_REQUIRE(lda % 8 == 0 && ldb % 8 == 0 && lda >= k && ldb >= k, "lda/ldb must be >= k and multiples of 8 (16-byte TMA strides)");
    MMSSL_REQUIRE(mmssl::aligned16(a_hi) && mmssl::aligned16(a_lo) && mmssl::aligned16(b_hi) && mmssl::aligned16(b_lo), "operand alignment");
    MMSSL_REQUIRE(c != nullptr && ldc >= n, "bad output");
    // the tensor maps cover [rows][ld]: the padding columns k..ld are READ by the kernel, so they must be zero
    for (int64_t i = 0; i < m; ++i)
        for (int64_t kk = k; kk < lda; ++kk) MMSSL_REQUIRE(a_hi[i * lda + kk] == 0 && a_lo[i * lda + kk] == 0, "A padding is not zero");
    for (int64_t j = 0; j < n; ++j)
        for (int64_t kk = k; kk < ldb; ++kk) MMSSL_REQUIRE(b_hi[j * ldb + kk] == 0 && b_lo[j * ldb + kk] == 0, "B padding is not zero");
    for (int64_t i = 0; i < m; ++i)
        for (int64_t j = 0; j < n; ++j) {
            float acc = 0.f;
            for (int64_t kk = 0; kk < k; ++kk) {
                const float ah = bf(a_hi[i * lda + kk]), al = bf(a_lo[i * lda + kk]);
                const float bh = bf(b_hi[j * ldb + kk]), bl = bf(b_lo[j * ldb + kk]);
                acc += ah * bh;
                acc += ah * bl;
                acc += al * bh;
            }
            const float r = alpha * acc;
            c[i * ldc + j] = accumulate ? c[i * ldc + j] + r : r;
        }
    return 0;
}
