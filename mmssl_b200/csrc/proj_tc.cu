// Projection GEMM on the 5th-gen tensor cores (tcgen05) fed by TMA, accumulators in TMEM.
//
//   partial[s][m][n] = sum_{k in slice s} (A_hi + A_lo)[m][k] * (B_hi + B_lo)[n][k]      (lo*lo dropped)
//
// replaces nn.Linear image_trans / text_trans forward (X = F W^T, Models.py:173-174) and its weight
// gradient (dW^T = F^T dX, autograd of the same lines).  fp32 contract (1e-4 rel) is met with a
// bf16 hi/lo operand split: three kind::f16 MMAs per product, fp32 accumulation in TMEM.  The
// feature matrix is constant (Models.py:46-47), so its split (and its transposed split for the
// weight gradient) is made once; both GEMMs then run the SAME K-major kernel.
//
// Shape of the work: HBM-bound (AI = 3*2*N/4 flop per byte of A at N = d), M is small (items) or
// medium (feature dim) -> split-K so that >= 2 waves of CTAs pull from HBM; partials are reduced in
// a fixed order by the epilogue kernels in proj_common.cu (deterministic, no float atomics).
//
// Kernel anatomy (one CTA = one 128 x N output tile of one K slice, 192 threads):
//   warp 0  : TMA producer   (cp.async.bulk.tensor.2d, SWIZZLE_128B, mbarrier complete_tx)
//   warp 1  : TMEM alloc + single-thread tcgen05.mma issuer, tcgen05.commit frees smem stages
//   warps 2-5: epilogue       (tcgen05.ld 32x32b -> registers -> 128-bit global stores)
#include "tc_common.cuh"
#include "../../include/mmssl_b200.h"

namespace mmssl {

template <int N>
struct GemmCfg {
    static constexpr int kTileBBytes = N * kBlockK * 2;
    static constexpr int kStageBytes = 2 * kTileABytes + 2 * kTileBBytes;
    // N = 64: two CTAs are co-resident per SM (2 stages each = the same 192 KB of loads in flight),
    // so one CTA's prologue / epilogue overlaps the other's main loop.
    static constexpr int kCtasPerSm = (N == 64) ? 2 : 1;
    static constexpr int kStages = (N == 64) ? 2 : (N == 128 ? 3 : 2);
    static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
    static constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) |
                                       ((uint32_t)(kBlockM >> 4) << 24);   // F32 acc, BF16 x BF16, K-major both
};

template <int N>
__global__ void __launch_bounds__(kThreads, GemmCfg<N>::kCtasPerSm)
gemm_bf16x3_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                   const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
                   float* __restrict__ partial, int M, int total_kb, int kb_per_split) {
    using Cfg = GemmCfg<N>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
    uint64_t* empty_bar = full_bar + Cfg::kStages;
    uint64_t* accum_bar = empty_bar + Cfg::kStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tile = blockIdx.x, split = blockIdx.y;
    const int kb0 = split * kb_per_split;
    const int nkb = max(0, min(total_kb, kb0 + kb_per_split) - kb0);

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_a_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_a_lo) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_b_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_b_lo) : "memory");
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < Cfg::kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
            mbar_init(accum_bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(N) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int i = 0; i < nkb; ++i) {
                const int s = i % Cfg::kStages;
                const uint32_t ph = (uint32_t)(i / Cfg::kStages) & 1u;
                mbar_wait(&empty_bar[s], ph ^ 1u);
                uint8_t* st = smem + s * Cfg::kStageBytes;
                mbar_expect_tx(&full_bar[s], Cfg::kStageBytes);
                const int kx = (kb0 + i) * kBlockK;
                tma_load_2d(&tm_a_hi, &full_bar[s], st, kx, m_tile * kBlockM, kEvictFirst);
                tma_load_2d(&tm_a_lo, &full_bar[s], st + kTileABytes, kx, m_tile * kBlockM, kEvictFirst);
                tma_load_2d(&tm_b_hi, &full_bar[s], st + 2 * kTileABytes, kx, 0, kEvictLast);
                tma_load_2d(&tm_b_lo, &full_bar[s], st + 2 * kTileABytes + Cfg::kTileBBytes, kx, 0, kEvictLast);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            for (int i = 0; i < nkb; ++i) {
                const int s = i % Cfg::kStages;
                const uint32_t ph = (uint32_t)(i / Cfg::kStages) & 1u;
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                const uint32_t a_hi = smem_u32(smem + s * Cfg::kStageBytes);
                const uint32_t a_lo = a_hi + kTileABytes;
                const uint32_t b_hi = a_hi + 2 * kTileABytes;
                const uint32_t b_lo = b_hi + Cfg::kTileBBytes;
#pragma unroll
                for (int k = 0; k < kBlockK / 16; ++k) {
                    const uint32_t off = k * 32;   // 16 bf16 = 32 bytes inside the 128-byte swizzle span
                    const uint64_t dah = make_sw128_kmajor_desc(a_hi + off), dal = make_sw128_kmajor_desc(a_lo + off);
                    const uint64_t dbh = make_sw128_kmajor_desc(b_hi + off), dbl = make_sw128_kmajor_desc(b_lo + off);
                    umma_bf16(tmem_base, dah, dbh, Cfg::kIdesc, (i > 0 || k > 0) ? 1u : 0u);
                    umma_bf16(tmem_base, dah, dbl, Cfg::kIdesc, 1u);
                    umma_bf16(tmem_base, dal, dbh, Cfg::kIdesc, 1u);
                }
                umma_commit(&empty_bar[s]);   // frees the smem stage once the MMAs have read it
            }
            umma_commit(accum_bar);           // accumulator complete
        }
    } else {
        const int q = warp & 3;               // TMEM lane quarter this warp may access
        const int64_t row = (int64_t)m_tile * kBlockM + q * 32 + lane;
        float* out = partial + ((int64_t)split * M + row) * N;
        if (nkb > 0) {
            mbar_wait(accum_bar, 0);
            tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < N; c += 32) {
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
                if (row < M) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        st4(out + c + j, make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                     __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])));
                }
            }
        } else if (row < M) {
            for (int c = 0; c < N; c += 4) st4(out + c, f4zero());
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(N) : "memory");
    }
}

static int choose_split(int64_t m, int64_t n, int64_t k) {
    // HBM-bound streaming GEMM: what matters is that all SMs pull from HBM for the whole kernel.
    // Pick the split-K that fills ONE wave of co-resident CTAs as completely as possible
    // (e.g. 56 M-tiles x 5 slices = 280 of 296 slots) instead of leaving a nearly empty tail wave.
    const int64_t total_kb = (k + kBlockK - 1) / kBlockK;
    const int64_t m_tiles = (m + kBlockM - 1) / kBlockM;
    const int64_t slots = (int64_t)kNumSMs * (n == 64 ? 2 : 1);
    int64_t split = slots / m_tiles;
    // accuracy: the tensor core's fp32 accumulate is not round-to-nearest and its error grows linearly with the number of MMAs
    // chained into one TMEM accumulator (gemm_wide.cu; measured round 2: the weight gradient at 200k items, 347 k-blocks per
    // slice, was 1e-4 off).  A slice is therefore at most kMaxChainKb k-blocks long; the slices are summed in fp32 by the epilogue.
    const int64_t need = (total_kb + kMaxChainKb - 1) / kMaxChainKb;
    if (split < need) split = need;
    if (split > kMaxSplit) split = kMaxSplit;
    if (split > total_kb) split = total_kb;
    if (split < 1) split = 1;
    // make every slice non-empty
    const int64_t per = (total_kb + split - 1) / split;
    split = (total_kb + per - 1) / per;
    return (int)split;
}

template <int N>
static int launch_gemm(const CUtensorMap& ah, const CUtensorMap& al, const CUtensorMap& bh, const CUtensorMap& bl,
                       float* partial, int64_t m, int64_t k, int split_k, cudaStream_t st) {
    using Cfg = GemmCfg<N>;
    static bool attr_done = false;
    if (!attr_done) {
        MMSSL_CUDA(cudaFuncSetAttribute(gemm_bf16x3_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
        attr_done = true;
    }
    const int total_kb = (int)((k + kBlockK - 1) / kBlockK);
    const int per = (total_kb + split_k - 1) / split_k;
    dim3 grid((unsigned)((m + kBlockM - 1) / kBlockM), (unsigned)split_k);
    gemm_bf16x3_kernel<N><<<grid, kThreads, Cfg::kSmemBytes, st>>>(ah, al, bh, bl, partial, (int)m, total_kb, per);
    MMSSL_LAUNCH_OK();
    return 0;
}

}  // namespace mmssl

using namespace mmssl;

extern "C" int64_t mmssl_gemm_bf16x3_workspace_floats(int64_t m, int64_t n, int64_t k, int* split_k_out) {
    const int split = choose_split(m, n, k);
    if (split_k_out) *split_k_out = split;
    return (int64_t)split * m * n;
}

extern "C" int mmssl_gemm_bf16x3(const uint16_t* a_hi, const uint16_t* a_lo, int64_t lda, const uint16_t* b_hi,
                                 const uint16_t* b_lo, int64_t ldb, int64_t m, int64_t n, int64_t k, int split_k,
                                 float* partial, void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    MMSSL_REQUIRE(n == 64 || n == 128 || n == 256, "n (embedding width) must be 64, 128 or 256");
    MMSSL_REQUIRE(m >= 1 && k >= 1 && m < (1ll << 31) && k < (1ll << 31), "bad m / k");
    MMSSL_REQUIRE(lda % 8 == 0 && ldb % 8 == 0 && lda >= k && ldb >= k, "lda/ldb must be >= k and multiples of 8 (16-byte TMA strides)");
    MMSSL_REQUIRE(aligned16(a_hi) && aligned16(a_lo) && aligned16(b_hi) && aligned16(b_lo) && aligned16(partial), "alignment");
    const int total_kb = (int)((k + kBlockK - 1) / kBlockK);
    MMSSL_REQUIRE(split_k >= 1 && split_k <= total_kb, "split_k out of range");
    {
        const int per = (total_kb + split_k - 1) / split_k;
        MMSSL_REQUIRE((int64_t)per * (split_k - 1) < total_kb, "split_k leaves an empty K slice (use mmssl_gemm_bf16x3_workspace_floats)");
    }
    CUtensorMap ah, al, bh, bl;
    if (int rc = make_map(&ah, a_hi, m, lda, kBlockM)) return rc;
    if (int rc = make_map(&al, a_lo, m, lda, kBlockM)) return rc;
    if (int rc = make_map(&bh, b_hi, n, ldb, (int)n)) return rc;
    if (int rc = make_map(&bl, b_lo, n, ldb, (int)n)) return rc;
    if (n == 64) return launch_gemm<64>(ah, al, bh, bl, partial, m, k, split_k, st);
    if (n == 128) return launch_gemm<128>(ah, al, bh, bl, partial, m, k, split_k, st);
    return launch_gemm<256>(ah, al, bh, bl, partial, m, k, split_k, st);
}
