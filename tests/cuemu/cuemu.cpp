// cuemu scheduler: fibers, block / warp barriers, deadlock detection.  See include/cuemu.h.  TEST INFRASTRUCTURE ONLY.
#include <sys/mman.h>

#include <algorithm>
#include <string>
#include <vector>

#include "cuemu.h"

#if !defined(__x86_64__)
#error "cuemu's context switch is written for x86-64"
#endif

// Save the callee-saved registers on the current stack, publish the stack pointer, adopt the other stack.
extern "C" void cuemu_switch(void** save_sp, void* load_sp);
__asm__(
    ".text\n"
    ".globl cuemu_switch\n"
    ".type cuemu_switch,@function\n"
    "cuemu_switch:\n"
    "    pushq %rbp\n"
    "    pushq %rbx\n"
    "    pushq %r12\n"
    "    pushq %r13\n"
    "    pushq %r14\n"
    "    pushq %r15\n"
    "    movq %rsp, (%rdi)\n"
    "    movq %rsi, %rsp\n"
    "    popq %r15\n"
    "    popq %r14\n"
    "    popq %r13\n"
    "    popq %r12\n"
    "    popq %rbx\n"
    "    popq %rbp\n"
    "    ret\n"
    ".size cuemu_switch,.-cuemu_switch\n");

uint3 threadIdx = {0, 0, 0}, blockIdx = {0, 0, 0};
dim3 blockDim, gridDim;

namespace cuemu {

Fiber* cur = nullptr;
int g_error = 0;

namespace {
constexpr size_t kStack = 256 * 1024;

struct Bar {
    int arrived = 0;
    unsigned gen = 0;
    unsigned mask = 0;
    int acc_or = 0, acc_and = 1, acc_cnt = 0;
    int res[3] = {0, 0, 0};
};
struct Warp {
    uint64_t slot[32];
    unsigned exist = 0, exited = 0;
    // One barrier per distinct membermask, like the hardware: lanes 0-15 may run shuffles under mask 0x0000ffff while
    // lanes 16-31 already wait in a full-mask shuffle further down the program.
    static constexpr int kMaxMasks = 96;
    int n_bars = 0;
    unsigned bar_mask[kMaxMasks];
    Bar bar[kMaxMasks];
};

std::string g_err_text;
void* g_sched_sp = nullptr;
std::vector<Fiber> g_fibers;
std::vector<Warp> g_warps;
Bar g_block_bar;
int g_alive = 0;
uint64_t g_events = 0;           // anything that is progress: arrival, release, exit
void (*g_thunk)(void*) = nullptr;
void* g_thunk_arg = nullptr;
unsigned char* g_smem = nullptr;
size_t g_smem_bytes = 0;
bool g_abandon = false;
const char* g_kernel_name = "";
char* g_stack_slab = nullptr;
size_t g_stack_slab_fibers = 0;

void yield() {
    Fiber* f = cur;
    cuemu_switch(&f->sp, g_sched_sp);
    __asm__ volatile("" ::: "memory");
}

[[noreturn]] void park_forever() {
    for (;;) yield();
}

void trampoline() {
    Fiber* f = cur;
    g_thunk(g_thunk_arg);
    f = cur;
    f->done = true;
    g_warps[f->warp].exited |= 1u << f->lane;
    --g_alive;
    ++g_events;
    park_forever();
}

void init_fiber(Fiber& f, char* stack_top) {
    uintptr_t top = reinterpret_cast<uintptr_t>(stack_top) & ~uintptr_t(15);
    void** sp = reinterpret_cast<void**>(top - 64);
    // [r15 r14 r13 r12 rbx rbp ret pad]
    for (int i = 0; i < 6; ++i) sp[i] = nullptr;
    sp[6] = reinterpret_cast<void*>(&trampoline);
    sp[7] = nullptr;
    f.sp = sp;
}

std::vector<int> fiber_order(int n) {
    std::vector<int> o(n);
    for (int i = 0; i < n; ++i) o[i] = i;
    const char* e = getenv("CUEMU_ORDER");
    if (e && strcmp(e, "rev") == 0) std::reverse(o.begin(), o.end());
    else if (e && strncmp(e, "shuffle", 7) == 0) {
        uint64_t s = 0x9E3779B97F4A7C15ull;
        if (e[7] == ':') s ^= strtoull(e + 8, nullptr, 10) * 0xBF58476D1CE4E5B9ull;
        for (int i = n - 1; i > 0; --i) {
            s ^= s << 13; s ^= s >> 7; s ^= s << 17;
            std::swap(o[i], o[(int)(s % (uint64_t)(i + 1))]);
        }
    }
    return o;
}
}  // namespace

const char* error_text() { return g_err_text.c_str(); }

void fail(const char* what) {
    if (g_error == 0) {
        char buf[640];
        snprintf(buf, sizeof(buf), "cuemu: %s [kernel %s, block (%u,%u,%u)", what, g_kernel_name, blockIdx.x, blockIdx.y, blockIdx.z);
        g_err_text = buf;
        if (cur) {
            snprintf(buf, sizeof(buf), ", thread (%u,%u,%u)", cur->tid.x, cur->tid.y, cur->tid.z);
            g_err_text += buf;
        }
        g_err_text += "]";
        g_error = cudaErrorLaunchFailure;
        fprintf(stderr, "%s\n", g_err_text.c_str());
    }
    g_abandon = true;
    if (cur) park_forever();
}

void* dyn_smem() { return g_smem; }
size_t dyn_smem_bytes() { return g_smem_bytes; }
void yield_blocked() {
    yield();
    if (g_abandon) park_forever();
}
void note_progress() { ++g_events; }

static void wait_bar(Bar& b, unsigned my_gen) {
    while (b.gen == my_gen) {
        yield();
        if (g_abandon) park_forever();
    }
}

void syncthreads() { (void)syncthreads_red(0, 0); }

int syncthreads_red(int pred, int op) {
    Bar& b = g_block_bar;
    const unsigned my = b.gen;
    b.acc_or |= pred != 0;
    b.acc_and &= pred != 0;
    b.acc_cnt += pred != 0;
    ++b.arrived;
    ++g_events;
    // released by the scheduler when arrived == live threads (a thread that exits may complete the barrier)
    wait_bar(b, my);
    return b.res[op];
}

static Bar& warp_bar(unsigned mask, Warp** wout) {
    Fiber* f = cur;
    Warp& w = g_warps[f->warp];
    if (!((mask >> f->lane) & 1u)) fail("warp primitive called by a lane that is not in its own mask");
    *wout = &w;
    for (int i = 0; i < w.n_bars; ++i)
        if (w.bar_mask[i] == mask) return w.bar[i];
    if (w.n_bars == Warp::kMaxMasks) fail("too many distinct warp masks in one block");
    w.bar_mask[w.n_bars] = mask;
    w.bar[w.n_bars] = Bar();
    return w.bar[w.n_bars++];
}

static void warp_arrive_wait(unsigned mask) {
    Warp* w;
    Bar& b = warp_bar(mask, &w);
    const unsigned my = b.gen;
    ++b.arrived;
    ++g_events;
    for (;;) {
        const unsigned need = mask & w->exist & ~w->exited;
        if (b.gen != my) return;
        if (b.arrived >= __builtin_popcount(need)) {
            b.arrived = 0;
            ++b.gen;
            ++g_events;
            return;
        }
        yield();
        if (g_abandon) park_forever();
    }
}

void syncwarp(unsigned mask) { warp_arrive_wait(mask); }

void exchange_begin(unsigned mask, uint64_t bits) {
    g_warps[cur->warp].slot[cur->lane] = bits;
    warp_arrive_wait(mask);
}
uint64_t exchange_peek(int lane, bool* valid) {
    Warp& w = g_warps[cur->warp];
    *valid = lane >= 0 && lane < 32 && ((w.exist & ~w.exited) >> lane) & 1u;
    return *valid ? w.slot[lane] : 0;
}
void exchange_end(unsigned mask) { warp_arrive_wait(mask); }

void run_grid(dim3 grid, dim3 block, size_t smem, void (*thunk)(void*), void* arg, const char* name) {
    const size_t nthreads = (size_t)block.x * block.y * block.z;
    g_kernel_name = name;
    cur = nullptr;
    if (g_error != 0) return;                        // sticky error: like CUDA, nothing runs after a failed launch
    if (nthreads == 0 || nthreads > 1024 || grid.x == 0 || grid.y == 0 || grid.z == 0 || grid.y > 65535 || grid.z > 65535 ||
        smem > 227 * 1024) {
        g_abandon = false;
        char buf[160];
        snprintf(buf, sizeof(buf), "invalid launch configuration grid (%u,%u,%u) block (%u,%u,%u) smem %zu", grid.x, grid.y, grid.z,
                 block.x, block.y, block.z, smem);
        blockIdx = {0, 0, 0};
        fail(buf);
        g_abandon = false;
        return;
    }
    if (g_stack_slab_fibers < nthreads) {
        if (g_stack_slab) munmap(g_stack_slab, g_stack_slab_fibers * kStack);
        g_stack_slab = static_cast<char*>(mmap(nullptr, nthreads * kStack, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0));
        g_stack_slab_fibers = nthreads;
    }
    if (g_smem_bytes < smem + 128) {
        free(g_smem);
        g_smem_bytes = smem + 128;
        g_smem = static_cast<unsigned char*>(aligned_alloc(1024, (g_smem_bytes + 1023) / 1024 * 1024));
    }
    g_thunk = thunk;
    g_thunk_arg = arg;
    blockDim = block;
    gridDim = grid;
    const std::vector<int> order = fiber_order((int)nthreads);
    const int nwarps = (int)((nthreads + 31) / 32);

    // Blocks run one after the other; CUEMU_BLOCK_ORDER=rev runs them from the last to the first, so that code which relies
    // on an arrival order between blocks (last-arriver reductions, self-resetting counters) is exercised in both directions.
    const char* bo = getenv("CUEMU_BLOCK_ORDER");
    const bool brev = bo != nullptr && strcmp(bo, "rev") == 0;
    const uint64_t nblocks = (uint64_t)grid.x * grid.y * grid.z;
    for (uint64_t bi = 0; bi < nblocks; ++bi) {
            {
                const uint64_t lin = brev ? nblocks - 1 - bi : bi;
                const unsigned bx = (unsigned)(lin % grid.x), by = (unsigned)((lin / grid.x) % grid.y), bz = (unsigned)(lin / ((uint64_t)grid.x * grid.y));
                {
                blockIdx = {bx, by, bz};
                g_fibers.assign(nthreads, Fiber());
                g_warps.assign(nwarps, Warp());
                g_block_bar = Bar();
                memset(g_smem, 0xFF, g_smem_bytes);
                ptx_block_reset();
                for (size_t t = 0; t < nthreads; ++t) {
                    Fiber& f = g_fibers[t];
                    f.lin = (int)t;
                    f.tid = {(unsigned)(t % block.x), (unsigned)((t / block.x) % block.y), (unsigned)(t / ((size_t)block.x * block.y))};
                    f.warp = (int)(t / 32);
                    f.lane = (int)(t % 32);
                    f.done = false;
                    g_warps[f.warp].exist |= 1u << f.lane;
                    init_fiber(f, g_stack_slab + (t + 1) * kStack);
                }
                g_alive = (int)nthreads;
                g_abandon = false;
                while (g_alive > 0 && !g_abandon) {
                    const uint64_t before = g_events;
                    for (int idx : order) {
                        Fiber& f = g_fibers[idx];
                        if (f.done) continue;
                        cur = &f;
                        threadIdx = f.tid;
                        cuemu_switch(&g_sched_sp, f.sp);
                        __asm__ volatile("" ::: "memory");
                        cur = nullptr;
                        if (g_abandon) break;
                        Bar& b = g_block_bar;
                        if (b.arrived > 0 && b.arrived >= g_alive) {      // block barrier complete (exited threads do not count)
                            b.res[0] = b.acc_or;
                            b.res[1] = b.acc_and;
                            b.res[2] = b.acc_cnt;
                            b.acc_or = 0; b.acc_and = 1; b.acc_cnt = 0;
                            b.arrived = 0;
                            ++b.gen;
                            ++g_events;
                        }
                    }
                    if (!g_abandon && g_alive > 0 && g_events == before) {
                        int at_block = g_block_bar.arrived;
                        char buf[200];
                        snprintf(buf, sizeof(buf), "deadlock: %d live threads, %d of them at __syncthreads, the rest at warp primitives that cannot complete",
                                 g_alive, at_block);
                        fail(buf);
                        g_abandon = true;
                    }
                }
                if (g_abandon) return;
                }
            }
    }
}

}  // namespace cuemu

extern "C" void cuemu_clear_error() { cuemu::g_error = 0; }
