// SpMM, bulk-copy gather pipeline:  Y_r = epilogue( A * X_r ),  r < nrhs  -- same operator contract as spmm.cu
// (reference: MMSSL.mm / torch.sparse.mm, Models.py:69-73, and torch.mm(sparse, dense), Models.py:203-208).
//
// Why a second kernel.  The LDG kernel (spmm.cu) keeps the neighbour rows it has in flight in REGISTERS: 8 float4 per lane,
// and a row walk of n non-zeros costs n/8 dependent memory round trips.  On the small graphs of the reference (Baby, Sports:
// everything L2-resident, 8-16 MB of compulsory traffic = 1-2 us at HBM speed) the kernel is a chain of round trips, not a
// bandwidth problem.  Here the neighbour rows travel as TMA bulk copies (cp.async.bulk global -> shared, completion counted in
// bytes on an mbarrier): no register holds them, ONE lane issues a whole 256 B .. 2 KB row, and a warp has up to 64 of them in
// flight after a single instruction per lane.
//
// Work decomposition (plan: graph.cu, mmssl_spmm_bulk_plan): the non-zeros are cut into BUCKETS of 32 consecutive positions --
// one position per lane.  One warp owns one bucket:
//   round trip 1   the bucket descriptor (32 B) and, coalesced, the 32 (col, val) pairs of the bucket + the 32 after it
//   round trip 2   all neighbour-row copies of the bucket's stream, the work items (row, begin, end) that start in the bucket,
//                  the split-row table entries they name and the row-indexed epilogue operands (alpha*C[row], saved softmax
//                  output, running-sum base: rows of a bucket are consecutive, they are staged by bulk copies as well)
//   then           accumulate from shared memory (every lane owns d/32 columns of every right-hand side), epilogue, store.
// Rows of up to 32 non-zeros are never cut: one that straddles the end of its bucket drags up to 31 positions of the next
// one along (the "overhang"; its (col, val) pairs were fetched in round trip 1).  Longer rows are cut AT bucket boundaries;
// their segments publish partial sums and the last one to arrive adds them in bucket order (deterministic), as in spmm.cu;
// rows of more than 32 segments accumulate with vector reductions into a zeroed slot instead.
//
// Shared memory per warp: ring of NST stages x 16 slots x (nrhs * d * 4) bytes + 2 x 8 rows per staged epilogue operand.
// No tensor cores: the contraction is a sparse gather.
//
// Two copy engines, same pipeline (template parameter TMA):
//   TMA    one cp.async.bulk per neighbour row, issued by the lane that owns the position.  MEASURED (B200, round 2,
//          profiles/r02_probe_bulk_tma.txt): the TMA unit retires about one bulk copy per 44 cycles per SM whatever its size --
//          6.5 G copies/s chip-wide -- so 256-512 B rows reach 6-12 B/cycle/SM, a quarter of what L2 can deliver:
//          3.06 ms at 1M x 200k (LDG kernel 1.04 ms), 20 us at Baby (LDG 9.4 us).  Kept for 1-2 KB slots and as the measured
//          record of why the per-neighbour TMA ring north_star sketches is not the product path at d <= 128.
//   LDGSTS the whole warp issues 16-byte cp.async.cg copies (512 B per instruction: one d=128 row, two d=64 rows), completion
//          through the same per-stage mbarriers (cp.async.mbarrier.arrive.noinc, 32 arrivals).  Still no register holds a row in
//          flight, and the issue rate is the LSU's (8 cycles per 512 B), above what L2 delivers.
#include "spmm_common.cuh"

namespace mmssl {

constexpr int kBk = 32;        // positions per bucket (one per lane)
constexpr int kSL = 16;        // ring slots per stage (one mbarrier per stage)
constexpr int kEG = 8;         // rows per staged epilogue-operand group (two groups: double buffer)

struct BulkParams {
    SpmmParams p;
    const int4* buckets;       // [n_buckets][2]: {item0, n_items, row0, n_rows}, {nz0, nz_end, 0, 0}
    int64_t n_buckets;
    int64_t nnz;
    int tasks_per_warp;
    int n_ops;                 // staged epilogue operands: 0, 1 or 2   (A = alpha*C, B = ysaved | S | SB)
    int warp_bytes;            // shared memory per warp
};

__device__ __forceinline__ uint32_t bsm(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bk_mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bsm(bar)), "r"(count)); }
__device__ __forceinline__ void bk_cp16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(bsm(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void bk_cp_arrive(uint64_t* bar) {      // this lane's earlier cp.async copies arrive on `bar` when they land
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bsm(bar)) : "memory");
}
__device__ __forceinline__ void bk_expect(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bsm(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bk_copy(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(bsm(dst)), "l"(src), "r"(bytes), "r"(bsm(bar)) : "memory");
}
__device__ __forceinline__ void bk_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = bsm(bar);
    uint32_t done = 0, spins = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done) break;
        if (++spins > (1u << 24)) __trap();   // never hang the GPU on a protocol bug
    }
}

// Lane layout of a d-wide row: V = d/32 floats per lane in chunks of CW (vector width) floats.
template <int V>
struct Lay {
    static constexpr int CW = V >= 4 ? 4 : 2;
    static constexpr int NCH = V / CW;
    __device__ static __forceinline__ int off(int lane, int ch) { return ch * (32 * CW) + lane * CW; }
};
template <int CW> __device__ __forceinline__ void ldv(float (&dst)[CW], const float* src);
template <> __device__ __forceinline__ void ldv<2>(float (&dst)[2], const float* src) {
    const float2 v = *reinterpret_cast<const float2*>(src); dst[0] = v.x; dst[1] = v.y;
}
template <> __device__ __forceinline__ void ldv<4>(float (&dst)[4], const float* src) {
    const float4 v = *reinterpret_cast<const float4*>(src); dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
}
template <int CW> __device__ __forceinline__ void ldv_cg(float (&dst)[CW], const float* src);
template <> __device__ __forceinline__ void ldv_cg<2>(float (&dst)[2], const float* src) {
    const float2 v = __ldcg(reinterpret_cast<const float2*>(src)); dst[0] = v.x; dst[1] = v.y;
}
template <> __device__ __forceinline__ void ldv_cg<4>(float (&dst)[4], const float* src) {
    const float4 v = __ldcg(reinterpret_cast<const float4*>(src)); dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
}
template <int CW> __device__ __forceinline__ void stv(float* dst, const float (&v)[CW]);
template <> __device__ __forceinline__ void stv<2>(float* dst, const float (&v)[2]) { *reinterpret_cast<float2*>(dst) = make_float2(v[0], v[1]); }
template <> __device__ __forceinline__ void stv<4>(float* dst, const float (&v)[4]) { *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]); }
template <int CW> __device__ __forceinline__ void stv_cg(float* dst, const float (&v)[CW]);
template <> __device__ __forceinline__ void stv_cg<2>(float* dst, const float (&v)[2]) { __stcg(reinterpret_cast<float2*>(dst), make_float2(v[0], v[1])); }
template <> __device__ __forceinline__ void stv_cg<4>(float* dst, const float (&v)[4]) { __stcg(reinterpret_cast<float4*>(dst), make_float4(v[0], v[1], v[2], v[3])); }
template <int CW> __device__ __forceinline__ void redv(float* dst, const float (&v)[CW]);
template <> __device__ __forceinline__ void redv<2>(float* dst, const float (&v)[2]) { atomicAdd(reinterpret_cast<float2*>(dst), make_float2(v[0], v[1])); }
template <> __device__ __forceinline__ void redv<4>(float* dst, const float (&v)[4]) { atomicAdd(reinterpret_cast<float4*>(dst), make_float4(v[0], v[1], v[2], v[3])); }
template <int CW> __device__ __forceinline__ void mcstv(float* dst, const float (&v)[CW]);
template <> __device__ __forceinline__ void mcstv<2>(float* dst, const float (&v)[2]) {
    asm volatile("multimem.st.relaxed.sys.global.v2.f32 [%0], {%1, %2};" ::"l"(dst), "f"(v[0]), "f"(v[1]) : "memory");
}
template <> __device__ __forceinline__ void mcstv<4>(float* dst, const float (&v)[4]) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
}

// Warp-cooperative copy of one D-float row (16 bytes per lane per instruction): rows of 64 floats take a half warp, so two rows
// travel per instruction (`half` selects which one a lane serves); 128 floats one instruction; 256 floats two.
template <int D>
__device__ __forceinline__ void bk_row_ldgsts(float* dst, const float* src, int lane, bool on) {
    if (D == 64) {
        if (on) bk_cp16(dst + (lane & 15) * 4, src + (lane & 15) * 4);
    } else {
#pragma unroll
        for (int o = 0; o < D; o += 128)
            if (on) bk_cp16(dst + o + lane * 4, src + o + lane * 4);
    }
}

// V floats per lane (d = 32 V), R right-hand sides, NST ring stages of kSL slots, TMA: copy engine (see the file header).
template <int V, int R, int NST, bool TMA>
__global__ void __launch_bounds__(256) spmm_bulk_kernel(const BulkParams bp) {
    using L = Lay<V>;
    constexpr int CW = L::CW, NCH = L::NCH;
    constexpr int D = 32 * V;
    constexpr int RD = R * D;                       // floats per ring slot / staged operand row
    constexpr uint32_t ROWB = RD * 4;
    const SpmmParams& p = bp.p;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* wbase = smem_raw + (size_t)warp * bp.warp_bytes;
    float* ring = reinterpret_cast<float*>(wbase);                                   // [NST][kSL][RD]
    float* stg = ring + NST * kSL * RD;                                              // [n_ops][2][kEG][RD]
    uint64_t* bars = reinterpret_cast<uint64_t*>(stg + (size_t)bp.n_ops * 2 * kEG * RD);   // [NST] ring + [2] operand groups
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < NST + 2; ++s) bk_mbar_init(&bars[s], TMA ? 1u : 32u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    pdl_wait();
    uint32_t phase = 0;                             // bit s: parity the next wait on barrier s uses

    const int64_t gwarp = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    const int64_t t_begin = gwarp * bp.tasks_per_warp;
    const int64_t t_end = min(bp.n_buckets, t_begin + bp.tasks_per_warp);
    if (t_begin >= t_end) return;

    // operand sources: A = alpha * C[row];  B = saved softmax output | running sum | running-sum base
    const bool has_a = p.has_c != 0;
    const bool has_b = bp.n_ops > (has_a ? 1 : 0);

    // ---- round trip 1 of the first task
    int4 bk0 = __ldg(&bp.buckets[2 * t_begin]), bk1 = __ldg(&bp.buckets[2 * t_begin + 1]);
    int c0 = 0, c1 = 0;
    float v0 = 0.f, v1 = 0.f;
    {
        const int64_t q0 = t_begin * kBk + lane;
        if (q0 < bp.nnz) { c0 = __ldg(p.colidx + q0); v0 = __ldg(p.vals + q0); }
        if (q0 + kBk < bp.nnz) { c1 = __ldg(p.colidx + q0 + kBk); v1 = __ldg(p.vals + q0 + kBk); }
    }

    for (int64_t t = t_begin; t < t_end; ++t) {
        const int64_t base = t * kBk;
        const int item0 = bk0.x, n_it = bk0.y, row0 = bk0.z, n_rows = bk0.w;
        const int nz0 = bk1.x, nz_end = bk1.y;
        const int my_c0 = c0, my_c1 = c1;
        const float my_v0 = v0, my_v1 = v1;
        // prefetch round trip 1 of the next task (hidden behind this one's copies)
        if (t + 1 < t_end) {
            bk0 = __ldg(&bp.buckets[2 * (t + 1)]); bk1 = __ldg(&bp.buckets[2 * (t + 1) + 1]);
            const int64_t q0 = (t + 1) * kBk + lane;
            c0 = 0; v0 = 0.f; c1 = 0; v1 = 0.f;
            if (q0 < bp.nnz) { c0 = __ldg(p.colidx + q0); v0 = __ldg(p.vals + q0); }
            if (q0 + kBk < bp.nnz) { c1 = __ldg(p.colidx + q0 + kBk); v1 = __ldg(p.vals + q0 + kBk); }
        }
        if (n_it == 0) continue;                    // nothing starts in this bucket (interior of a straddling short row)

        // ---- issue: neighbour-row copies.  stream = positions [nz0, nz_end), q = position - base in [0, 64)
        const int q_lo = (int)(nz0 - base), q_hi = (int)(nz_end - base);
        const int n_chunks = (q_hi + kSL - 1) / kSL;          // chunks 0 .. n_chunks-1 (leading ones may be empty)
        int issued = 0;                                       // chunks [0, issued) are armed or empty
        auto issue_chunk = [&](int ch) {
            const int lo = max(q_lo, ch * kSL), hi = min(q_hi, ch * kSL + kSL);
            if (hi <= lo) return;
            uint64_t* bar = &bars[ch % NST];
            const int creg = (ch * kSL < kBk) ? my_c0 : my_c1;       // a chunk never straddles the two position registers
            float* stage = ring + (size_t)((ch % NST) * kSL) * RD;
            if (TMA) {
                const bool mine = ((lane >> 4) == (ch & 1));           // chunk ch lives in lanes 16 (ch & 1) .. +15
                if (lane == ((ch * kSL) & 31)) bk_expect(bar, (uint32_t)(hi - lo) * ROWB);
                __syncwarp();
                const int qq = ch * kSL + (lane & (kSL - 1));          // this lane's position within [0, 64)
                if (mine && qq >= lo && qq < hi) {
                    float* dst = stage + (size_t)(qq & (kSL - 1)) * RD;
#pragma unroll
                    for (int r = 0; r < R; ++r) bk_copy(dst + r * D, p.x[r] + (int64_t)creg * p.ldx[r], D * 4, bar);
                }
            } else {
                constexpr int RPI = (D == 64) ? 2 : 1;                 // rows per instruction
                for (int qq = lo; qq < hi; qq += RPI) {
                    const int myq = qq + (RPI == 2 ? (lane >> 4) : 0);
                    const int col = __shfl_sync(0xffffffffu, creg, myq & 31);
                    float* dst = stage + (size_t)(myq & (kSL - 1)) * RD;
#pragma unroll
                    for (int r = 0; r < R; ++r) bk_row_ldgsts<D>(dst + r * D, p.x[r] + (int64_t)col * p.ldx[r], lane, myq < hi);
                }
                bk_cp_arrive(bar);
            }
        };
        __syncwarp();                                         // every lane is done with the ring contents of the previous task
        for (; issued < n_chunks && issued < NST; ++issued) issue_chunk(issued);

        // ---- issue: row-indexed epilogue operands of rows [row0, row0 + n_rows), groups of kEG rows, two buffers
        int groups_issued = 0, groups_waited = 0;
        const int n_groups = bp.n_ops > 0 ? (n_rows + kEG - 1) / kEG : 0;
        auto operand_src = [&](int o, int r, int64_t row) -> const float* {   // o: 0 = A (alpha * C) when present, else B
            if (o == 0 && has_a) return p.c[r] + row * p.ldc[r];
            return (p.epilogue == MMSSL_EPI_SOFTMAX_BWD) ? p.ys[r] + row * p.ldys[r]
                   : (p.s_mode == 2)                     ? p.sb[r] + row * p.ldsb[r]
                                                         : p.s[r] + row * p.lds[r];
        };
        auto issue_group = [&](int g) {
            uint64_t* bar = &bars[NST + (g & 1)];
            const int r_lo = g * kEG, r_n = min(kEG, n_rows - r_lo);
            if (TMA) {
                if (lane == 0) bk_expect(bar, (uint32_t)(r_n * bp.n_ops) * ROWB);
                __syncwarp();
                if (lane < r_n) {
                    const int64_t row = row0 + r_lo + lane;
                    for (int o = 0; o < bp.n_ops; ++o) {
                        float* dst = stg + (size_t)((o * 2 + (g & 1)) * kEG + lane) * RD;
#pragma unroll
                        for (int r = 0; r < R; ++r) bk_copy(dst + r * D, operand_src(o, r, row), D * 4, bar);
                    }
                }
            } else {
                constexpr int RPI = (D == 64) ? 2 : 1;
                for (int o = 0; o < bp.n_ops; ++o)
                    for (int i = 0; i < r_n; i += RPI) {
                        const int mi = i + (RPI == 2 ? (lane >> 4) : 0);
                        const int64_t row = row0 + r_lo + min(mi, r_n - 1);
                        float* dst = stg + (size_t)((o * 2 + (g & 1)) * kEG + mi) * RD;
#pragma unroll
                        for (int r = 0; r < R; ++r) bk_row_ldgsts<D>(dst + r * D, operand_src(o, r, row), lane, mi < r_n);
                    }
                bk_cp_arrive(bar);
            }
        };
        for (; groups_issued < n_groups && groups_issued < 2; ++groups_issued) issue_group(groups_issued);

        // ---- consume
        int arrived = 0;                                      // chunks [0, arrived) have landed (or are empty)
        for (int ib = 0; ib < n_it; ib += 32) {
            int4 item = make_int4(-1, 0, 0, -1);
            if (ib + lane < n_it) item = __ldg(&p.items[item0 + ib + lane]);
            int4 st = make_int4(0, 0, 0, 0);
            if (item.w >= 0) st = __ldg(&p.split_table[item.w]);     // {first partial slot, #segments, first bucket, heavy}
            const int nb = min(32, n_it - ib);
            for (int k = 0; k < nb; ++k) {
                const int row = __shfl_sync(0xffffffffu, item.x, k);
                const int ib_ = __shfl_sync(0xffffffffu, item.y, k);
                const int ie_ = __shfl_sync(0xffffffffu, item.z, k);
                const int sp = __shfl_sync(0xffffffffu, item.w, k);
                float acc[R][V];
#pragma unroll
                for (int r = 0; r < R; ++r)
#pragma unroll
                    for (int e = 0; e < V; ++e) acc[r][e] = 0.f;
                for (int pos = ib_; pos < ie_; ++pos) {
                    const int q = (int)(pos - base);
                    const int ch = q / kSL;
                    if (ch >= arrived) {
                        for (; arrived <= ch; ++arrived) {
                            if (arrived >= issued) {          // ring smaller than the stream: refill behind the consumer
                                __syncwarp();
                                for (; issued < n_chunks && issued < arrived + NST; ++issued) issue_chunk(issued);
                            }
                            const int lo = max(q_lo, arrived * kSL), hi = min(q_hi, arrived * kSL + kSL);
                            if (hi > lo) {
                                const int s = arrived % NST;
                                bk_wait(&bars[s], (phase >> s) & 1u);
                                phase ^= 1u << s;
                            }
                        }
                    }
                    const float v = (q < kBk) ? __shfl_sync(0xffffffffu, my_v0, q) : __shfl_sync(0xffffffffu, my_v1, q - kBk);
                    const float* slot = ring + (size_t)((ch % NST) * kSL + (q & (kSL - 1))) * RD;
#pragma unroll
                    for (int r = 0; r < R; ++r)
#pragma unroll
                        for (int c = 0; c < NCH; ++c) {
                            float x[CW];
                            ldv<CW>(x, slot + r * D + L::off(lane, c));
#pragma unroll
                            for (int e = 0; e < CW; ++e) acc[r][c * CW + e] = fmaf(v, x[e], acc[r][c * CW + e]);
                        }
                }

                // ---- staged operands of this row: make sure its group has landed, keep the next one coming
                const int rs = row - row0;
                const int g = rs / kEG;
                if (bp.n_ops > 0) {
                    for (; groups_waited <= g; ++groups_waited) {
                        const int s = NST + (groups_waited & 1);
                        bk_wait(&bars[s], (phase >> s) & 1u);
                        phase ^= 1u << s;
                        // buffer (groups_waited - 1) & 1 == (groups_waited + 1) & 1 has been consumed: refill it
                        if (groups_issued < n_groups && groups_issued <= groups_waited + 1) {
                            __syncwarp();
                            issue_group(groups_issued);
                            ++groups_issued;
                        }
                    }
                }

                // ---- split rows: publish the partial sum, the last segment to arrive reduces in bucket order
                bool fin = true;
                if (sp >= 0) {
                    const int st_x = __shfl_sync(0xffffffffu, st.x, k), st_y = __shfl_sync(0xffffffffu, st.y, k);
                    const int st_z = __shfl_sync(0xffffffffu, st.z, k), st_w = __shfl_sync(0xffffffffu, st.w, k);
                    if (st_w != 0) {
                        float* slot = p.partials + (int64_t)st_x * RD;
#pragma unroll
                        for (int r = 0; r < R; ++r)
#pragma unroll
                            for (int c = 0; c < NCH; ++c) {
                                float x[CW];
#pragma unroll
                                for (int e = 0; e < CW; ++e) x[e] = acc[r][c * CW + e];
                                redv<CW>(slot + r * D + L::off(lane, c), x);
                            }
                    } else {
                        const int seg = (int)(t - st_z);
                        float* part = p.partials + ((int64_t)st_x + seg) * RD;
#pragma unroll
                        for (int r = 0; r < R; ++r)
#pragma unroll
                            for (int c = 0; c < NCH; ++c) {
                                float x[CW];
#pragma unroll
                                for (int e = 0; e < CW; ++e) x[e] = acc[r][c * CW + e];
                                stv<CW>(part + r * D + L::off(lane, c), x);
                            }
                    }
                    __threadfence();
                    __syncwarp();
                    int old = 0;
                    if (lane == 0) old = atomicAdd(p.counters + sp, 1);
                    old = __shfl_sync(0xffffffffu, old, 0);
                    fin = (old == st_y - 1);
                    if (fin) {
                        __threadfence();
                        if (lane == 0) p.counters[sp] = 0;    // self-cleaning for the next launch
                        if (st_w != 0) {
                            float* slot = p.partials + (int64_t)st_x * RD;
#pragma unroll
                            for (int r = 0; r < R; ++r)
#pragma unroll
                                for (int c = 0; c < NCH; ++c) {
                                    float x[CW], z[CW];
                                    ldv_cg<CW>(x, slot + r * D + L::off(lane, c));
#pragma unroll
                                    for (int e = 0; e < CW; ++e) { acc[r][c * CW + e] = x[e]; z[e] = 0.f; }
                                    stv_cg<CW>(slot + r * D + L::off(lane, c), z);      // leave the slot clean for the next launch
                                }
                        } else {
#pragma unroll
                            for (int r = 0; r < R; ++r)
#pragma unroll
                                for (int e = 0; e < V; ++e) acc[r][e] = 0.f;
                            constexpr int PB = (8 / (R * NCH)) >= 1 ? (8 / (R * NCH)) : 1;   // partial rows fetched per round trip
                            for (int s0 = 0; s0 < st_y; s0 += PB) {
                                float pv[PB][R][V];
#pragma unroll
                                for (int qq = 0; qq < PB; ++qq) {
                                    const bool on = (s0 + qq) < st_y;
                                    const float* ps = p.partials + ((int64_t)st_x + s0 + qq) * RD;
#pragma unroll
                                    for (int r = 0; r < R; ++r)
#pragma unroll
                                        for (int c = 0; c < NCH; ++c) {
                                            float x[CW];
                                            if (on) ldv_cg<CW>(x, ps + r * D + L::off(lane, c));
#pragma unroll
                                            for (int e = 0; e < CW; ++e) pv[qq][r][c * CW + e] = on ? x[e] : 0.f;
                                        }
                                }
#pragma unroll
                                for (int qq = 0; qq < PB; ++qq)      // fixed (bucket) order -> deterministic sum
#pragma unroll
                                    for (int r = 0; r < R; ++r)
#pragma unroll
                                        for (int e = 0; e < V; ++e) acc[r][e] += pv[qq][r][e];
                            }
                        }
                    }
                }
                if (!fin) continue;

                // ---- epilogue
                const float* sa = stg + (size_t)((0 * 2 + (g & 1)) * kEG + (rs & (kEG - 1))) * RD;                    // operand A rows
                const float* sb_ = stg + (size_t)(((has_a ? 1 : 0) * 2 + (g & 1)) * kEG + (rs & (kEG - 1))) * RD;     // operand B rows
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if (has_a && p.c[r] != nullptr) {
#pragma unroll
                        for (int c = 0; c < NCH; ++c) {
                            float x[CW];
                            ldv<CW>(x, sa + r * D + L::off(lane, c));
#pragma unroll
                            for (int e = 0; e < CW; ++e) acc[r][c * CW + e] = fmaf(p.alpha, x[e], acc[r][c * CW + e]);
                        }
                    }
                    if (p.epilogue == MMSSL_EPI_SOFTMAX) {
                        float m = -INFINITY;
#pragma unroll
                        for (int e = 0; e < V; ++e) m = fmaxf(m, acc[r][e]);
                        m = group_max<32>(m, 0xffffffffu);
                        float sum = 0.f;
#pragma unroll
                        for (int e = 0; e < V; ++e) { acc[r][e] = __expf(acc[r][e] - m); sum += acc[r][e]; }
                        sum = group_sum<32>(sum, 0xffffffffu);
                        const float inv = 1.f / sum;
#pragma unroll
                        for (int e = 0; e < V; ++e) acc[r][e] *= inv;
                    } else if (p.epilogue == MMSSL_EPI_SOFTMAX_BWD) {
                        float yv[V];
                        float dotp = 0.f;
#pragma unroll
                        for (int c = 0; c < NCH; ++c) {
                            float x[CW];
                            ldv<CW>(x, sb_ + r * D + L::off(lane, c));
#pragma unroll
                            for (int e = 0; e < CW; ++e) { yv[c * CW + e] = x[e]; dotp = fmaf(acc[r][c * CW + e], x[e], dotp); }
                        }
                        dotp = group_sum<32>(dotp, 0xffffffffu);
#pragma unroll
                        for (int e = 0; e < V; ++e) acc[r][e] = yv[e] * (acc[r][e] - dotp);
                    }
#pragma unroll
                    for (int c = 0; c < NCH; ++c) {
                        float x[CW];
#pragma unroll
                        for (int e = 0; e < CW; ++e) x[e] = acc[r][c * CW + e];
                        const int64_t off = (int64_t)row * p.ldy[r] + L::off(lane, c);
                        if (p.y_mode[r] == 1) {       // NVSwitch multicast: the store is replicated into every GPU's table
                            mcstv<CW>(p.y[r] + off, x);
                        } else {
                            stv<CW>(p.y[r] + off, x);
                            if (p.y_mode[r] == 2)     // peer-mapped tables over NVLink
                                for (int qq = 0; qq < p.n_peers[r]; ++qq) stv<CW>(p.y_peers[r][qq] + off, x);
                        }
                    }
                    if (p.s_mode != 0 && p.s[r] != nullptr) {
#pragma unroll
                        for (int c = 0; c < NCH; ++c) {
                            float x[CW], o[CW];
                            ldv<CW>(x, sb_ + r * D + L::off(lane, c));
#pragma unroll
                            for (int e = 0; e < CW; ++e) o[e] = x[e] + acc[r][c * CW + e];
                            stv<CW>(p.s[r] + (int64_t)row * p.lds[r] + L::off(lane, c), o);
                        }
                    }
                }
            }
        }
    }
}

template <int V, int R, int NST, bool TMA>
static int launch_bulk(const BulkParams& bp_in, cudaStream_t stream, int wpb, int tpw) {
    constexpr int D = 32 * V, RD = R * D;
    BulkParams bp = bp_in;
    const int warp_bytes = ((NST * kSL + bp.n_ops * 2 * kEG) * RD * 4 + (NST + 2) * 8 + 127) & ~127;
    bp.warp_bytes = warp_bytes;
    bp.tasks_per_warp = tpw;
    while (wpb > 1 && warp_bytes * wpb > 227 * 1024) --wpb;       // wide slots: fewer warps per block
    const int smem = warp_bytes * wpb;
    if (smem > 227 * 1024) return fail("mmssl_spmm_bulk_f32", "one warp's ring does not fit shared memory (use 2 ring stages)");
    static int attr_smem = 0;
    if (smem > attr_smem) {
        MMSSL_CUDA(cudaFuncSetAttribute(spmm_bulk_kernel<V, R, NST, TMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_smem = smem;
    }
    const int64_t warps = (bp.n_buckets + tpw - 1) / tpw;
    const int64_t blocks = (warps + wpb - 1) / wpb;
    if (blocks == 0) return 0;
    if (blocks > 0x7fffffffll) return fail("mmssl_spmm_bulk_f32", "grid too large");
    MMSSL_CUDA_LAUNCH((spmm_bulk_kernel<V, R, NST, TMA>), dim3((unsigned)blocks), dim3(32 * wpb), (size_t)smem, stream, bp);
    MMSSL_LAUNCH_OK();
    return 0;
}

}  // namespace mmssl

using namespace mmssl;

// variant: bits 0-3 ring stages (0 = automatic: 4 when a slot is <= 512 B, else 2), bits 4-7 warps per block (0 = 4),
// bits 8-15 buckets per warp (0 = automatic: 1 under 2M edges, 4 above), bit 16: TMA bulk copies instead of LDGSTS.
extern "C" int mmssl_spmm_bulk_f32(const mmssl_csr_t* a, const int32_t* buckets8, int64_t n_buckets, int d, int nrhs,
                                   const mmssl_spmm_rhs_t* rhs, int epilogue, float alpha, int s_mode, float* partials,
                                   int64_t partials_floats, int variant, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MMSSL_REQUIRE(buckets8 != nullptr && n_buckets >= 0, "missing bucket table (mmssl_spmm_bulk_plan)");
    MMSSL_REQUIRE(nrhs >= 1 && nrhs <= 2, "the bulk-copy SpMM takes 1 or 2 right-hand sides");
    BulkParams bp;
    if (int rc = fill_spmm_params(bp.p, a, d, nrhs, rhs, epilogue, alpha, s_mode, partials, partials_floats)) return rc;
    bp.buckets = reinterpret_cast<const int4*>(buckets8);
    bp.n_buckets = n_buckets;
    bp.nnz = a->nnz;
    const bool has_b = (epilogue == MMSSL_EPI_SOFTMAX_BWD) || (s_mode != 0 && bp.p.s[0] != nullptr);
    MMSSL_REQUIRE(!(epilogue == MMSSL_EPI_SOFTMAX_BWD && s_mode != 0), "softmax-backward epilogue and running sum together: use mmssl_spmm_csr_f32");
    for (int r = 0; r < nrhs; ++r) {
        MMSSL_REQUIRE((bp.p.c[r] != nullptr) == (bp.p.c[0] != nullptr), "C must be given for all right-hand sides or none");
        MMSSL_REQUIRE((bp.p.s[r] != nullptr) == (bp.p.s[0] != nullptr), "S must be given for all right-hand sides or none");
    }
    bp.n_ops = (bp.p.has_c ? 1 : 0) + (has_b ? 1 : 0);
    const int slot_bytes = nrhs * d * 4;
    int nst = variant & 15, wpb = (variant >> 4) & 15, tpw = (variant >> 8) & 255;
    const bool tma = (variant >> 16) & 1;
    if (nst == 0) nst = slot_bytes <= 512 ? 4 : 2;
    if (wpb == 0) wpb = 4;
    if (tpw == 0) tpw = a->nnz >= (1ll << 21) ? 4 : 1;
    MMSSL_REQUIRE(nst == 2 || nst == 4, "ring stages must be 2 or 4");
    MMSSL_REQUIRE(wpb >= 1 && wpb <= 8, "warps per block must be 1..8");
#define MMSSL_BULK_CASE2(V, RR, NN) return tma ? launch_bulk<V, RR, NN, true>(bp, stream, wpb, tpw) : launch_bulk<V, RR, NN, false>(bp, stream, wpb, tpw);
#define MMSSL_BULK_CASE(V)                                                       \
    if (nrhs == 1) { if (nst == 4) { MMSSL_BULK_CASE2(V, 1, 4) } MMSSL_BULK_CASE2(V, 1, 2) } \
    if (nst == 4) { MMSSL_BULK_CASE2(V, 2, 4) } MMSSL_BULK_CASE2(V, 2, 2)
    if (d == 64) { MMSSL_BULK_CASE(2) }
    if (d == 128) { MMSSL_BULK_CASE(4) }
    MMSSL_BULK_CASE(8)
#undef MMSSL_BULK_CASE
}
