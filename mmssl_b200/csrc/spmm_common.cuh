// Shared between the SpMM variants (spmm.cu: LDG gather; spmm_hot.cu: TMA-staged hot rows).
#pragma once
#include "common.cuh"
#include "../../include/mmssl_b200.h"

namespace mmssl {

constexpr int kMaxRhs = MMSSL_SPMM_MAX_RHS;

struct SpmmParams {
    const int32_t* rowptr;
    const int32_t* colidx;
    const float* vals;
    const int4* items;
    int64_t n_items;
    const int4* split_table;
    int32_t* counters;
    float* partials;
    const float* x[kMaxRhs];  int64_t ldx[kMaxRhs];
    float* y[kMaxRhs];        int64_t ldy[kMaxRhs];
    const float* c[kMaxRhs];  int64_t ldc[kMaxRhs];   // optional addend (alpha * C[row])
    const float* ys[kMaxRhs]; int64_t ldys[kMaxRhs];  // saved softmax output (softmax-backward epilogue)
    float* s[kMaxRhs];        int64_t lds[kMaxRhs];   // optional running sum
    const float* sb[kMaxRhs]; int64_t ldsb[kMaxRhs];  // s_mode 2: S = SB[row] + out
    float alpha;
    int epilogue;   // MMSSL_EPI_*
    int s_mode;     // 0 none, 1: S += out, 2: S = SB + out
    int has_c;
    int y_mode[kMaxRhs];            // 0 local, 1 multicast (multimem.st), 2 local + peers
    int n_peers[kMaxRhs];
    float* y_peers[kMaxRhs][8];
    unsigned long long x_policy;    // L2 policy word of the gathered rows (hinted variants of the LDG kernel)
};

// G lanes per group, C float4 chunks per lane per rhs (d = 4*G*C), R right-hand sides.
// UNR neighbour gathers are issued back to back before the first FMA consumes one, and the next
// chunk of (col, val) pairs is prefetched while the current one is processed, so a row walk costs
// about one memory round trip per UNR non-zeros instead of one per load.
//
// Persistent: the grid is one resident wave; every lane group strides over the work items.  The
// next item descriptor and its first (col, val) chunk are fetched while the current row is being
// processed, so the item -> indices -> gather dependency chain is paid once per group, not per row.
template <int G, int C, int R>
__device__ __forceinline__ void spmm_epilogue(const SpmmParams& p, float4 (&acc)[R][C], int row, int lane, unsigned gmask) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int64_t col0 = lane * 4;
        if (p.has_c && p.c[r] != nullptr) {
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const float4 cv = ld4(p.c[r] + (int64_t)row * p.ldc[r] + col0 + c * (4 * G));   // may alias Y
                fma4(acc[r][c], p.alpha, cv);
            }
        }
        if (p.epilogue == MMSSL_EPI_SOFTMAX) {
            float m = -INFINITY;
#pragma unroll
            for (int c = 0; c < C; ++c) m = fmaxf(m, max4(acc[r][c]));
            m = group_max<G>(m, gmask);
            float sum = 0.f;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                acc[r][c].x = __expf(acc[r][c].x - m); acc[r][c].y = __expf(acc[r][c].y - m);
                acc[r][c].z = __expf(acc[r][c].z - m); acc[r][c].w = __expf(acc[r][c].w - m);
                sum += (acc[r][c].x + acc[r][c].y) + (acc[r][c].z + acc[r][c].w);
            }
            sum = group_sum<G>(sum, gmask);
            const float inv = 1.f / sum;
#pragma unroll
            for (int c = 0; c < C; ++c) acc[r][c] = scale4(acc[r][c], inv);
        } else if (p.epilogue == MMSSL_EPI_SOFTMAX_BWD) {
            float4 yv[C];
            float dotp = 0.f;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                yv[c] = ldg4(p.ys[r] + (int64_t)row * p.ldys[r] + col0 + c * (4 * G));
                dotp += dot4(acc[r][c], yv[c]);
            }
            dotp = group_sum<G>(dotp, gmask);
#pragma unroll
            for (int c = 0; c < C; ++c) {
                acc[r][c].x = yv[c].x * (acc[r][c].x - dotp); acc[r][c].y = yv[c].y * (acc[r][c].y - dotp);
                acc[r][c].z = yv[c].z * (acc[r][c].z - dotp); acc[r][c].w = yv[c].w * (acc[r][c].w - dotp);
            }
        }
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const int64_t off = (int64_t)row * p.ldy[r] + col0 + c * (4 * G);
            if (p.y_mode[r] == 1) {           // NVSwitch multicast: the store is replicated into every GPU's table
                asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p.y[r] + off), "f"(acc[r][c].x),
                             "f"(acc[r][c].y), "f"(acc[r][c].z), "f"(acc[r][c].w) : "memory");
            } else {
                st4(p.y[r] + off, acc[r][c]);
                if (p.y_mode[r] == 2)         // peer-mapped tables over NVLink
                    for (int q = 0; q < p.n_peers[r]; ++q) st4(p.y_peers[r][q] + off, acc[r][c]);
            }
        }
        if (p.s_mode != 0 && p.s[r] != nullptr) {
#pragma unroll
            for (int c = 0; c < C; ++c) {
                float* sp = p.s[r] + (int64_t)row * p.lds[r] + col0 + c * (4 * G);
                const float4 prev = (p.s_mode == 1) ? ld4(sp)
                                                    : ldg4(p.sb[r] + (int64_t)row * p.ldsb[r] + col0 + c * (4 * G));
                st4(sp, add4(prev, acc[r][c]));
            }
        }
    }
}

// Fills the kernel parameter block from the C-ABI descriptors (validation included).
int fill_spmm_params(SpmmParams& p, const mmssl_csr_t* a, int d, int nrhs, const mmssl_spmm_rhs_t* rhs, int epilogue,
                     float alpha, int s_mode, float* partials, int64_t partials_floats);

// spmm_hot.cu
int launch_spmm_hot(const SpmmParams& p, int d, int nrhs, const int32_t* colidx_hot, const int32_t* hot_ids, int n_hot,
                    cudaStream_t stream);

}  // namespace mmssl
