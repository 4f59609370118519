// cuemu: functional model of the inline PTX the library uses.  TEST INFRASTRUCTURE ONLY (see include/cuemu.h).
//
// Scope = exactly what csrc/tc_common.cuh, proj_tc.cu and gemm_wide.cu issue:
//   mbarrier.init / arrive / arrive.expect_tx / try_wait.parity         phase + pending-arrival + transaction-byte counters
//   cp.async.bulk (1-D) ... mbarrier::complete_tx::bytes       contiguous copy global -> shared (spmm_hot.cu)
//   cp.async.bulk.tensor.2d ... mbarrier::complete_tx::bytes    box copy global -> shared through the tensor map: out-of-bounds
//                                                               elements read as zero, SWIZZLE_128B (address bits [4:6] ^= [7:9])
//   tcgen05.alloc / dealloc / relinquish_alloc_permit           a 128-lane x 512-column fp32 TMEM per block
//   tcgen05.mma.cta_group::1.kind::f16                          D[128 x N] (+)= A[128 x 16] * B[N x 16]^T, bf16 operands read from
//                                                               shared memory through K-major SWIZZLE_128B descriptors
//   tcgen05.commit ... mbarrier::arrive::one                    (the model executes an MMA when it is issued, so commit = arrive)
//   tcgen05.ld.sync.aligned.32x32b.x32                          lane t of warp w reads TMEM lane 32*(w%4)+t, 32 columns
//   multimem.ld_reduce.add.v4.f32 / multimem.st.v4.f32         over multicast groups registered by the test (cuemu_mc_register):
//                                                               several ranks' replicas in one process
//   fences, prefetch.tensormap, tcgen05.wait::ld, griddepcontrol.wait   no-ops
// What the model asserts on the way: 1024-byte aligned swizzled tiles, transaction bytes that add up, barrier phases, a warp
// touching only its TMEM lane quarter, descriptor fields the kernels are supposed to encode.  It is calibrated by running
// proj_tc.cu -- which is parity-green on real B200s -- through it (tests/test_emu_tensor_core.py): a kernel that passes here
// and shares tc_common.cuh's descriptor / swizzle code with it differs from hardware-proven code only in its own logic.
// It says nothing about timing, asynchrony bugs that need real concurrency, or PTX outside this list (fails the launch).
#include <map>
#include <string>
#include <vector>

#include "cuda.h"

namespace cuemu {

namespace {
struct TensorMap {
    uint32_t magic;
    int elem_bytes, swizzle;
    const uint8_t* base;
    uint64_t dims[2];          // elements: [0] = contiguous dimension
    uint64_t stride1;          // bytes between rows
    uint32_t box[2];
};
static_assert(sizeof(TensorMap) <= sizeof(CUtensorMap), "tensor map record does not fit");
constexpr uint32_t kMagic = 0x7E45AB01u;

struct MBar {
    bool init = false;
    int expected = 0, pending = 0;
    int64_t tx = 0;
    unsigned phase = 0;
};
std::map<uint32_t, MBar> g_bars;

// NVSwitch multicast objects, for tests that play several ranks in one process: a fake "multicast address range" stands for n
// replicas (one buffer per rank).  multimem.st writes every replica, multimem.ld_reduce.add sums them in rank order.
struct McGroup { uint64_t base, bytes; std::vector<uint8_t*> replicas; };
std::vector<McGroup> g_mc;
const McGroup& mc_find(uint64_t addr, size_t bytes, uint64_t* off) {
    for (const McGroup& g : g_mc)
        if (addr >= g.base && addr + bytes <= g.base + g.bytes) { *off = addr - g.base; return g; }
    fail("multimem: the address is not inside a registered multicast group");
    static McGroup none{};
    *off = 0;
    return none;
}
std::vector<uint32_t> g_tmem;              // [128 lanes][512 columns]
int g_tmem_cols = 0;

uint8_t* smem_at(uint64_t addr, size_t bytes) {
    if (addr + bytes > dyn_smem_bytes()) fail("shared-memory address outside the block's dynamic shared memory");
    return static_cast<uint8_t*>(dyn_smem()) + addr;
}
inline uint64_t swz128(uint64_t a) { return a ^ (((a >> 7) & 7u) << 4); }
inline float bf16(uint16_t b) {
    const uint32_t u = (uint32_t)b << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}
MBar& bar_at(uint64_t addr, bool must_exist = true) {
    if (addr % 8) fail("mbarrier address is not 8-byte aligned");
    MBar& b = g_bars[(uint32_t)addr];
    if (must_exist && !b.init) fail("mbarrier used before mbarrier.init");
    return b;
}
void bar_check(MBar& b) {
    if (b.tx < 0) fail("mbarrier: more bytes completed than expected (complete_tx without a matching expect_tx)");
    if (b.pending == 0 && b.tx == 0) {
        b.phase ^= 1u;
        b.pending = b.expected;
        note_progress();
    }
}
void bar_arrive(MBar& b) {
    if (b.pending <= 0) fail("mbarrier: more arrivals than the count it was initialised with");
    --b.pending;
    note_progress();
    bar_check(b);
}
void out32(void** outs, const int* sizes, int i, uint32_t v) {
    if (sizes[i] != 4) fail("ptx model: a 32-bit result is written to an operand of another size");
    memcpy(outs[i], &v, 4);
}
}  // namespace

CUresult encode_tiled(CUtensorMap* map, CUtensorMapDataType dt, cuuint32_t rank, void* base, const cuuint64_t* dims, const cuuint64_t* strides,
                      const cuuint32_t* box, const cuuint32_t* estr, CUtensorMapInterleave il, CUtensorMapSwizzle sw, CUtensorMapL2promotion,
                      CUtensorMapFloatOOBfill) {
    // the driver's documented requirements for the case the library uses
    if (rank != 2 || dt != CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 || il != CU_TENSOR_MAP_INTERLEAVE_NONE) return CUDA_ERROR_INVALID_VALUE;
    if (((uintptr_t)base & 15u) || (strides[0] & 15u) || strides[0] < dims[0] * 2) return CUDA_ERROR_INVALID_VALUE;
    if (box[0] == 0 || box[1] == 0 || box[0] > 256 || box[1] > 256 || estr[0] != 1 || estr[1] != 1) return CUDA_ERROR_INVALID_VALUE;
    if (sw == CU_TENSOR_MAP_SWIZZLE_128B && box[0] * 2 > 128) return CUDA_ERROR_INVALID_VALUE;      // inner box <= swizzle span
    if (sw != CU_TENSOR_MAP_SWIZZLE_128B && sw != CU_TENSOR_MAP_SWIZZLE_NONE) return CUDA_ERROR_INVALID_VALUE;
    if (dims[0] == 0 || dims[1] == 0) return CUDA_ERROR_INVALID_VALUE;
    TensorMap t{};
    t.magic = kMagic; t.elem_bytes = 2; t.swizzle = (int)sw; t.base = static_cast<const uint8_t*>(base);
    t.dims[0] = dims[0]; t.dims[1] = dims[1]; t.stride1 = strides[0]; t.box[0] = box[0]; t.box[1] = box[1];
    memset(map, 0, sizeof(*map));
    memcpy(map, &t, sizeof(t));
    return CUDA_SUCCESS;
}

void ptx_block_reset() {
    g_bars.clear();
    g_tmem.assign(128 * 512, 0x7FC00000u);      // NaN until written
    g_tmem_cols = 0;
}

void ptx_op(const char* text, void** outs, const int* out_sizes, int n_out, const uint64_t* in, int n_in) {
    auto has = [&](const char* s) { return strstr(text, s) != nullptr; };
    if (has("griddepcontrol") || has("tcgen05.fence") || has("fence.mbarrier_init") || has("prefetch.tensormap") ||
        has("tcgen05.wait::ld") || has("tcgen05.relinquish_alloc_permit"))
        return;
    if (has("ex2.approx")) {                                  // MUFU.EX2: the model returns the correctly rounded value
        float x;
        const uint32_t xb = (uint32_t)in[0];
        memcpy(&x, &xb, 4);
        const float y = exp2f(x);
        uint32_t yb;
        memcpy(&yb, &y, 4);
        out32(outs, out_sizes, 0, yb);
        return;
    }
    if (has("mbarrier.init")) {
        MBar& b = bar_at(in[0], false);
        b = MBar();
        b.init = true;
        int count = 0;
        if (n_in >= 2) count = (int)in[1];
        else if (const char* c = strrchr(text, ',')) count = atoi(c + 1);          // immediate count in the template
        b.expected = b.pending = count;
        if (b.expected <= 0) fail("mbarrier.init with a non-positive count");
        return;
    }
    if (has("mbarrier.arrive.expect_tx")) {
        MBar& b = bar_at(in[0]);
        b.tx += (int64_t)in[1];
        bar_arrive(b);
        return;
    }
    if (has("mbarrier.arrive.shared::cta.b64 _")) {       // plain arrival (consumer release)
        bar_arrive(bar_at(in[0]));
        return;
    }
    if (has("mbarrier.try_wait.parity")) {
        MBar& b = bar_at(in[0]);
        const unsigned parity = (unsigned)in[1] & 1u;
        const bool done = b.phase != parity;     // the phase with this parity has completed
        out32(outs, out_sizes, 0, done ? 1u : 0u);
        if (!done) yield_blocked();
        return;
    }
    if (has("cp.async.cg.shared.global")) {                 // Ampere-style 16-byte asynchronous copy: executed at once in this model
        if ((in[0] % 16) || (in[1] % 16)) fail("cp.async.cg 16: addresses must be multiples of 16 bytes");
        memcpy(smem_at(in[0], 16), reinterpret_cast<const void*>((uintptr_t)in[1]), 16);
        return;
    }
    if (has("cp.async.commit_group") || has("cp.async.wait_group")) return;      // copies execute at once in this model
    if (has("cp.async.mbarrier.arrive.noinc")) {            // the lane's earlier copies have landed (model: they always have)
        bar_arrive(bar_at(in[0]));
        return;
    }
    if (has("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes")) {      // 1-D bulk copy: dst, src, bytes, barrier
        const uint64_t dst = in[0], bytes = in[2];
        if ((dst % 16) || (in[1] % 16) || (bytes % 16)) fail("cp.async.bulk: addresses and size must be multiples of 16 bytes");
        memcpy(smem_at(dst, bytes), reinterpret_cast<const void*>((uintptr_t)in[1]), bytes);
        MBar& b = bar_at(in[3]);
        b.tx -= (int64_t)bytes;
        note_progress();
        bar_check(b);
        return;
    }
    if (has("cp.async.bulk.tensor.2d")) {
        TensorMap t;
        memcpy(&t, reinterpret_cast<const void*>((uintptr_t)in[1]), sizeof(t));
        if (t.magic != kMagic) fail("cp.async.bulk.tensor: the tensor map was not made by cuTensorMapEncodeTiled");
        const uint64_t dst = in[0];
        const int64_t x = (int32_t)in[3], y = (int32_t)in[4];
        const uint32_t row_bytes = t.box[0] * t.elem_bytes;
        if (t.swizzle == CU_TENSOR_MAP_SWIZZLE_128B && (dst % 1024)) fail("TMA destination of a SWIZZLE_128B tile is not 1024-byte aligned");
        if (dst % 128) fail("TMA destination is not 128-byte aligned");
        const size_t total = (size_t)t.box[1] * row_bytes;
        uint8_t* sm = smem_at(dst, total);
        (void)sm;
        for (uint32_t r = 0; r < t.box[1]; ++r)
            for (uint32_t c = 0; c < t.box[0]; ++c) {
                const int64_t gx = x + c, gy = y + r;
                uint16_t v = 0;                                    // out of bounds reads as zero
                if (gx >= 0 && gy >= 0 && (uint64_t)gx < t.dims[0] && (uint64_t)gy < t.dims[1])
                    memcpy(&v, t.base + (uint64_t)gy * t.stride1 + (uint64_t)gx * 2, 2);
                uint64_t a = dst + (uint64_t)r * row_bytes + (uint64_t)c * 2;
                if (t.swizzle == CU_TENSOR_MAP_SWIZZLE_128B) a = swz128(a);
                memcpy(smem_at(a, 2), &v, 2);
            }
        MBar& b = bar_at(in[2]);
        b.tx -= (int64_t)total;
        note_progress();
        bar_check(b);
        return;
    }
    if (has("tcgen05.alloc")) {
        if (cur->lane != 0) return;               // .sync.aligned: one allocation per warp, every lane executes the instruction
        const int ncols = (int)in[1];
        if (ncols < 32 || ncols > 512 || (ncols & (ncols - 1))) fail("tcgen05.alloc: column count must be a power of two in [32, 512]");
        if (g_tmem_cols + ncols > 512) fail("tcgen05.alloc: TMEM exhausted");
        const uint32_t base = (uint32_t)g_tmem_cols;              // lane 0, first free column
        g_tmem_cols += ncols;
        memcpy(smem_at(in[0], 4), &base, 4);
        return;
    }
    if (has("tcgen05.dealloc")) {
        if (cur->lane != 0) return;
        g_tmem_cols -= (int)in[1];
        if (g_tmem_cols < 0) fail("tcgen05.dealloc of more columns than allocated");
        return;
    }
    if (has("tcgen05.mma")) {
        if (!has("kind::f16") || !has("cta_group::1")) fail("tcgen05.mma: only cta_group::1.kind::f16 is modelled");
        const uint32_t tm = (uint32_t)in[0], idesc = (uint32_t)in[3];
        const uint64_t da = in[1], db = in[2];
        const bool acc = in[4] != 0;
        const int N = (int)((idesc >> 17) & 0x3Fu) << 3, M = (int)((idesc >> 24) & 0x1Fu) << 4;
        // instruction descriptor fields the kernels are meant to set: D = F32 (bits 4-5 = 1), A = B = BF16 (bits 7-9, 10-12 = 1),
        // no negate / transpose (K-major both), dense
        if (((idesc >> 4) & 3u) != 1u || ((idesc >> 7) & 7u) != 1u || ((idesc >> 10) & 7u) != 1u) fail("tcgen05.mma: unexpected operand formats in the instruction descriptor");
        if ((idesc >> 13) & 0xFu) fail("tcgen05.mma: negate / transpose bits set (the kernels use K-major operands)");
        if (M != 128 || N < 16 || N > 256 || (N % 16)) fail("tcgen05.mma: M must be 128 and N a multiple of 16 up to 256");
        const int lane0 = (int)(tm >> 16), col0 = (int)(tm & 0xFFFFu);
        if (lane0 != 0 || col0 + N > g_tmem_cols) fail("tcgen05.mma: accumulator outside the allocated TMEM columns");
        auto dec = [&](uint64_t d, uint64_t& start, uint64_t& sbo) {
            start = (d & 0x3FFFu) << 4;
            sbo = ((d >> 32) & 0x3FFFu) << 4;
            if ((d >> 61) != 2u) fail("tcgen05.mma: shared-memory descriptor is not SWIZZLE_128B");
            if (((d >> 46) & 3u) != 1u) fail("tcgen05.mma: shared-memory descriptor version field is not 1 (sm_100)");
            if (sbo != 1024) fail("tcgen05.mma: stride-byte-offset of a K-major SWIZZLE_128B tile must be 1024 (8 rows x 128 B)");
            if ((start & ~(uint64_t)127) % 1024) fail("tcgen05.mma: swizzled operand tile is not 1024-byte aligned");
        };
        uint64_t sa, sboa, sb, sbob;
        dec(da, sa, sboa);
        dec(db, sb, sbob);
        auto elem = [&](uint64_t start, uint64_t sbo, int row, int k) {
            const uint64_t a = swz128(start + (uint64_t)(row / 8) * sbo + (uint64_t)(row % 8) * 128 + (uint64_t)k * 2);
            uint16_t v;
            memcpy(&v, smem_at(a, 2), 2);
            return bf16(v);
        };
        float av[128][16];
        for (int i = 0; i < M; ++i)
            for (int k = 0; k < 16; ++k) av[i][k] = elem(sa, sboa, i, k);
        for (int j = 0; j < N; ++j) {
            float bv[16];
            for (int k = 0; k < 16; ++k) bv[k] = elem(sb, sbob, j, k);
            for (int i = 0; i < M; ++i) {
                uint32_t& cell = g_tmem[(size_t)i * 512 + col0 + j];
                float d;
                memcpy(&d, &cell, 4);
                if (!acc) d = 0.f;
                for (int k = 0; k < 16; ++k) d += av[i][k] * bv[k];
                memcpy(&cell, &d, 4);
            }
        }
        note_progress();
        return;
    }
    if (has("tcgen05.commit")) {
        bar_arrive(bar_at(in[0]));               // every MMA issued so far has already executed in this model
        return;
    }
    if (has("tcgen05.ld") && has("32x32b.x32")) {
        if (n_out != 32) fail("tcgen05.ld.32x32b.x32 needs 32 destination registers");
        const uint32_t ta = (uint32_t)in[0];
        const int lane_base = (int)(ta >> 16), col0 = (int)(ta & 0xFFFFu);
        if (lane_base != (cur->warp % 4) * 32) fail("tcgen05.ld: a warp may only access the TMEM lane quarter 32 * (warp % 4)");
        if (col0 + 32 > g_tmem_cols) fail("tcgen05.ld beyond the allocated TMEM columns");
        for (int j = 0; j < 32; ++j) out32(outs, out_sizes, j, g_tmem[(size_t)(lane_base + cur->lane) * 512 + col0 + j]);
        return;
    }
    if (has("multimem.ld_reduce") && has("add.v4.f32")) {
        uint64_t off;
        const McGroup& g = mc_find(in[0], 16, &off);
        if (in[0] % 16) fail("multimem.ld_reduce.v4: misaligned address");
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (uint8_t* r : g.replicas) {
            float v[4];
            memcpy(v, r + off, 16);
            for (int k = 0; k < 4; ++k) acc[k] += v[k];
        }
        for (int k = 0; k < 4; ++k) {
            uint32_t u;
            memcpy(&u, &acc[k], 4);
            out32(outs, out_sizes, k, u);
        }
        return;
    }
    if (has("multimem.st") && has("v4.f32")) {
        uint64_t off;
        const McGroup& g = mc_find(in[0], 16, &off);
        if (in[0] % 16) fail("multimem.st.v4: misaligned address");
        uint32_t v[4] = {(uint32_t)in[1], (uint32_t)in[2], (uint32_t)in[3], (uint32_t)in[4]};
        for (uint8_t* r : g.replicas) memcpy(r + off, v, 16);
        return;
    }
    (void)n_in;
    std::string m = std::string("inline PTX is not emulated: ") + text;
    fail(m.c_str());
}

}  // namespace cuemu

extern "C" void cuemu_mc_register(uint64_t base, uint64_t bytes, int n, void** replicas) {
    cuemu::McGroup g;
    g.base = base; g.bytes = bytes;
    for (int i = 0; i < n; ++i) g.replicas.push_back(static_cast<uint8_t*>(replicas[i]));
    cuemu::g_mc.push_back(g);
}
extern "C" void cuemu_mc_clear() { cuemu::g_mc.clear(); }
