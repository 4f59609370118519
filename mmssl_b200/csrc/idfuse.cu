// Fused id-fusion kernels (reference Models.py:139-169 "attention" closed form + :188-197):
//   forward : m = coef*(Ya [+ Yb]) ; z = m * Wsum ; out = E + rate * z / max(|z|, eps)
//   backward: dz = rate * d(normalize)(g) ; dY = coef * dz * Wsum^T (+ external grads) ;
//             dWsum += m^T dz  (per-block partial tiles, reduced in a fixed order)
// with Wsum = sum_h Wcat[h*d:(h+1)*d, :]   and dWcat[h] = dWsum for every head.
// The d x d matrix lives in shared memory; a lane group owns a row, m[k] is broadcast by shuffles.
// Replaces 8 CUDA-core GEMM launches of the first version with 2 + 3 launches.
#include "common.cuh"
#include "../../include/mmssl_b200.h"

namespace mmssl {

constexpr float kNormEps = 1e-12f;

__global__ void wsum_kernel(const float* __restrict__ wcat, int d, int heads, float* __restrict__ wsum) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d * d) return;
    float s = 0.f;
    for (int h = 0; h < heads; ++h) s += wcat[(int64_t)h * d * d + i];
    wsum[i] = s;
}

// z4 = sum_k m[k] * W[k][lane*4 .. +3], m distributed as one float4 per lane of the group
template <int G>
__device__ __forceinline__ float4 row_times_matrix(const float4 m4, const float* __restrict__ Wsm, int lane, unsigned mask) {
    constexpr int D = 4 * G;
    float4 z = f4zero();
#pragma unroll 4
    for (int kk = 0; kk < G; ++kk) {
        const float mx = __shfl_sync(mask, m4.x, kk, G), my = __shfl_sync(mask, m4.y, kk, G);
        const float mz = __shfl_sync(mask, m4.z, kk, G), mw = __shfl_sync(mask, m4.w, kk, G);
        const float* wr = Wsm + (kk * 4) * D + lane * 4;
        fma4(z, mx, *reinterpret_cast<const float4*>(wr));
        fma4(z, my, *reinterpret_cast<const float4*>(wr + D));
        fma4(z, mz, *reinterpret_cast<const float4*>(wr + 2 * D));
        fma4(z, mw, *reinterpret_cast<const float4*>(wr + 3 * D));
    }
    return z;
}

template <int G>
__global__ void __launch_bounds__(256) id_fuse2_fwd_kernel(const float* __restrict__ ya, int64_t lda,
                                                           const float* __restrict__ yb, int64_t ldb, float coef,
                                                           const float* __restrict__ wsum, const float* __restrict__ e,
                                                           int64_t lde, int64_t n, float rate, float* __restrict__ out,
                                                           int64_t ldo, float* __restrict__ zn, float* __restrict__ nrm) {
    constexpr int D = 4 * G;
    extern __shared__ __align__(16) float Wsm[];
    for (int i = threadIdx.x; i < D * D / 4; i += blockDim.x)
        reinterpret_cast<float4*>(Wsm)[i] = __ldg(reinterpret_cast<const float4*>(wsum) + i);
    __syncthreads();
    const unsigned mask = group_mask<G>();
    const int lane = threadIdx.x & (G - 1);
    const int gpb = blockDim.x / G;
    for (int64_t row = blockIdx.x * (int64_t)gpb + threadIdx.x / G; row < n; row += (int64_t)gridDim.x * gpb) {
        float4 m4 = ld4(ya + row * lda + lane * 4);
        if (yb != nullptr) m4 = add4(m4, ld4(yb + row * ldb + lane * 4));
        m4 = scale4(m4, coef);
        float4 z = row_times_matrix<G>(m4, Wsm, lane, mask);
        const float nr = sqrtf(group_sum<G>(dot4(z, z), mask));
        const float inv = 1.f / fmaxf(nr, kNormEps);
        z = scale4(z, inv);
        float4 o = ld4(e + row * lde + lane * 4);
        fma4(o, rate, z);
        st4(out + row * ldo + lane * 4, o);
        st4(zn + row * D + lane * 4, z);
        if (lane == 0) nrm[row] = nr;
    }
}

// Backward.  Tiles of TR rows: phase 1 (per lane group): dz, dY ; phase 2 (whole block): acc += M^T DZ.
template <int G>
__global__ void __launch_bounds__(256) id_fuse2_bwd_kernel(const float* __restrict__ g, int64_t ldg,
                                                           const float* __restrict__ zn, const float* __restrict__ nrm,
                                                           const float* __restrict__ ya, int64_t lda,
                                                           const float* __restrict__ yb, int64_t ldb, float coef,
                                                           const float* __restrict__ wsum, int64_t n, float rate,
                                                           const float* __restrict__ ext_a, int64_t ldea,
                                                           const float* __restrict__ ext_b, int64_t ldeb,
                                                           float* __restrict__ out_a, int64_t ldoa,
                                                           float* __restrict__ out_b, int64_t ldob,
                                                           float* __restrict__ dw_part) {
    constexpr int D = 4 * G;
    constexpr int TR = 64;                 // rows per tile
    constexpr int T = D / 16;              // per-thread edge of the D x D accumulator tile (256 threads = 16 x 16)
    extern __shared__ __align__(16) float sm[];
    float* Wt = sm;                        // Wt[c][k] = Wsum[k][c]
    float* Ms = sm + D * D;                // [TR][D]
    float* Zs = Ms + TR * D;               // [TR][D]
    for (int i = threadIdx.x; i < D * D; i += blockDim.x) {
        const int k = i / D, c = i - k * D;
        Wt[c * D + k] = __ldg(wsum + i);
    }
    __syncthreads();
    const unsigned mask = group_mask<G>();
    const int lane = threadIdx.x & (G - 1);
    const int gib = threadIdx.x / G, gpb = blockDim.x / G;
    const int ty = threadIdx.x / 16, tx = threadIdx.x % 16;
    float acc[T][T];
#pragma unroll
    for (int i = 0; i < T; ++i)
#pragma unroll
        for (int j = 0; j < T; ++j) acc[i][j] = 0.f;

    for (int64_t tile0 = blockIdx.x * (int64_t)TR; tile0 < n; tile0 += (int64_t)gridDim.x * TR) {
        for (int rl = gib; rl < TR; rl += gpb) {
            const int64_t row = tile0 + rl;
            float4 m4 = f4zero(), dz = f4zero();
            if (row < n) {
                const float4 gv = ld4(g + row * ldg + lane * 4);
                const float4 nv = ld4(zn + row * D + lane * 4);
                const float nr = nrm[row];
                if (nr > kNormEps) {
                    const float dot = group_sum<G>(dot4(nv, gv), mask);
                    const float k = rate / nr;
                    dz = make_float4(k * (gv.x - nv.x * dot), k * (gv.y - nv.y * dot), k * (gv.z - nv.z * dot), k * (gv.w - nv.w * dot));
                } else {
                    dz = scale4(gv, rate / kNormEps);
                }
                m4 = ld4(ya + row * lda + lane * 4);
                if (yb != nullptr) m4 = add4(m4, ld4(yb + row * ldb + lane * 4));
                m4 = scale4(m4, coef);
                // dY = coef * dz * Wsum^T  (+ external gradients)
                float4 dy = scale4(row_times_matrix<G>(dz, Wt, lane, mask), coef);
                if (out_b == nullptr) {
                    if (ext_a) dy = add4(dy, ld4(ext_a + row * ldea + lane * 4));
                    if (ext_b) dy = add4(dy, ld4(ext_b + row * ldeb + lane * 4));
                    st4(out_a + row * ldoa + lane * 4, dy);
                } else {
                    float4 da = dy, db = dy;
                    if (ext_a) da = add4(da, ld4(ext_a + row * ldea + lane * 4));
                    if (ext_b) db = add4(db, ld4(ext_b + row * ldeb + lane * 4));
                    st4(out_a + row * ldoa + lane * 4, da);
                    st4(out_b + row * ldob + lane * 4, db);
                }
            }
            st4(Ms + rl * D + lane * 4, m4);
            st4(Zs + rl * D + lane * 4, dz);
        }
        __syncthreads();
#pragma unroll 4
        for (int r = 0; r < TR; ++r) {
            float mv[T], zv[T];
#pragma unroll
            for (int i = 0; i < T; i += 4) {
                const float4 a = *reinterpret_cast<const float4*>(Ms + r * D + ty * T + i);
                mv[i] = a.x; mv[i + 1] = a.y; mv[i + 2] = a.z; mv[i + 3] = a.w;
                const float4 b = *reinterpret_cast<const float4*>(Zs + r * D + tx * T + i);
                zv[i] = b.x; zv[i + 1] = b.y; zv[i + 2] = b.z; zv[i + 3] = b.w;
            }
#pragma unroll
            for (int i = 0; i < T; ++i)
#pragma unroll
                for (int j = 0; j < T; ++j) acc[i][j] = fmaf(mv[i], zv[j], acc[i][j]);
        }
        __syncthreads();
    }
    float* part = dw_part + (int64_t)blockIdx.x * D * D;
#pragma unroll
    for (int i = 0; i < T; ++i)
#pragma unroll
        for (int j = 0; j < T; j += 4)
            st4(part + (ty * T + i) * D + tx * T + j, make_float4(acc[i][j], acc[i][j + 1], acc[i][j + 2], acc[i][j + 3]));
}

// dWcat[h*d*d + i] = sum over the partial tiles of both sides (fixed order), for every head h
__global__ void dwcat_reduce_kernel(const float* __restrict__ part_u, int nu, const float* __restrict__ part_i, int ni,
                                    int d, int heads, float* __restrict__ dwcat) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d * d) return;
    const int64_t dd = (int64_t)d * d;
    float s = 0.f;
    for (int pass = 0; pass < 2; ++pass) {
        const float* p = pass ? part_i : part_u;
        const int nb = pass ? ni : nu;
        int b = 0;
        for (; b + 4 <= nb; b += 4) {
            const float v0 = p[(b + 0) * dd + i], v1 = p[(b + 1) * dd + i], v2 = p[(b + 2) * dd + i], v3 = p[(b + 3) * dd + i];
            s = (((s + v0) + v1) + v2) + v3;
        }
        for (; b < nb; ++b) s += p[b * dd + i];
    }
    for (int h = 0; h < heads; ++h) dwcat[h * dd + i] = s;
}

static int fuse_blocks(int64_t n) {
    int64_t b = (n + 63) / 64;
    if (b > 2 * kNumSMs) b = 2 * kNumSMs;
    return (int)(b < 1 ? 1 : b);
}

}  // namespace mmssl

using namespace mmssl;

extern "C" int mmssl_wsum(const float* wcat, int d, int heads, float* wsum, void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    wsum_kernel<<<(d * d + 255) / 256, 256, 0, st>>>(wcat, d, heads, wsum);
    MMSSL_LAUNCH_OK();
    return 0;
}

extern "C" int mmssl_id_fuse2_blocks(int64_t n) { return fuse_blocks(n); }

extern "C" int mmssl_id_fuse2_fwd(const float* ya, int64_t lda, const float* yb, int64_t ldb, float coef, const float* wsum,
                                  const float* e, int64_t lde, int64_t n, int d, float rate, float* out, int64_t ldo,
                                  float* zn, float* nrm, void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    MMSSL_REQUIRE(d == 64 || d == 128, "fused id-fusion supports d = 64 or 128 (d = 256 uses the GEMM path)");
    MMSSL_REQUIRE(aligned16(ya) && lda % 4 == 0 && (yb == nullptr || (aligned16(yb) && ldb % 4 == 0)) && aligned16(e) &&
                      lde % 4 == 0 && aligned16(out) && ldo % 4 == 0 && aligned16(zn) && aligned16(wsum), "alignment");
    if (n == 0) return 0;
    const int blocks = fuse_blocks(n);
    const size_t smem = (size_t)d * d * 4;
    if (d == 64) {
        id_fuse2_fwd_kernel<16><<<blocks, 256, smem, st>>>(ya, lda, yb, ldb, coef, wsum, e, lde, n, rate, out, ldo, zn, nrm);
    } else {
        static bool attr = false;
        if (!attr) { MMSSL_CUDA(cudaFuncSetAttribute(id_fuse2_fwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = true; }
        id_fuse2_fwd_kernel<32><<<blocks, 256, smem, st>>>(ya, lda, yb, ldb, coef, wsum, e, lde, n, rate, out, ldo, zn, nrm);
    }
    MMSSL_LAUNCH_OK();
    return 0;
}

extern "C" int mmssl_id_fuse2_bwd(const float* g, int64_t ldg, const float* zn, const float* nrm, const float* ya,
                                  int64_t lda, const float* yb, int64_t ldb, float coef, const float* wsum, int64_t n, int d,
                                  float rate, const float* ext_a, int64_t ldea, const float* ext_b, int64_t ldeb,
                                  float* out_a, int64_t ldoa, float* out_b, int64_t ldob, float* dw_part, void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    MMSSL_REQUIRE(d == 64 || d == 128, "fused id-fusion supports d = 64 or 128 (d = 256 uses the GEMM path)");
    MMSSL_REQUIRE(aligned16(g) && ldg % 4 == 0 && aligned16(zn) && aligned16(ya) && lda % 4 == 0 && aligned16(out_a) &&
                      ldoa % 4 == 0 && aligned16(dw_part), "alignment");
    MMSSL_REQUIRE((yb == nullptr || (aligned16(yb) && ldb % 4 == 0)) && (ext_a == nullptr || (aligned16(ext_a) && ldea % 4 == 0)) &&
                      (ext_b == nullptr || (aligned16(ext_b) && ldeb % 4 == 0)) && (out_b == nullptr || (aligned16(out_b) && ldob % 4 == 0)), "alignment");
    if (n == 0) return 0;
    const int blocks = fuse_blocks(n);
    const size_t smem = ((size_t)d * d + 2 * 64 * (size_t)d) * 4;
    if (d == 64) {
        static bool attr = false;
        if (!attr) { MMSSL_CUDA(cudaFuncSetAttribute(id_fuse2_bwd_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = true; }
        id_fuse2_bwd_kernel<16><<<blocks, 256, smem, st>>>(g, ldg, zn, nrm, ya, lda, yb, ldb, coef, wsum, n, rate, ext_a, ldea,
                                                         ext_b, ldeb, out_a, ldoa, out_b, ldob, dw_part);
    } else {
        static bool attr = false;
        if (!attr) { MMSSL_CUDA(cudaFuncSetAttribute(id_fuse2_bwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = true; }
        id_fuse2_bwd_kernel<32><<<blocks, 256, smem, st>>>(g, ldg, zn, nrm, ya, lda, yb, ldb, coef, wsum, n, rate, ext_a, ldea,
                                                         ext_b, ldeb, out_a, ldoa, out_b, ldob, dw_part);
    }
    MMSSL_LAUNCH_OK();
    return 0;
}

extern "C" int mmssl_dwcat_reduce(const float* part_u, int nu, const float* part_i, int ni, int d, int heads, float* dwcat,
                                  void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    dwcat_reduce_kernel<<<(d * d + 255) / 256, 256, 0, st>>>(part_u, nu, part_i, ni, d, heads, dwcat);
    MMSSL_LAUNCH_OK();
    return 0;
}
