// TEST INFRASTRUCTURE ONLY - cooperative-fiber execution of CUDA kernels on the CPU (see shim/cuda_runtime.h and
// tests/emu/README.md).  One CTA at a time; each CUDA thread is a fiber; __syncthreads, warp collectives and
// mbarrier waits yield to a round-robin scheduler whose visiting order can be shuffled per round
// (FDFD_EMU_SHUFFLE=seed) so that warps run ahead of / behind each other as they may on hardware.  Not a performance
// model and not a memory-model checker: it finds indexing, phase/parity, buffer-reuse and missing-wait mistakes.
#include <cuda_runtime.h>
#include <sys/mman.h>
#if !defined(__x86_64__)
#include <ucontext.h>
#endif

#include <fcntl.h>
#include <unistd.h>

#include <cstdarg>
#include <map>
#include <random>
#include <string>
#include <vector>

#if defined(__x86_64__)
// Minimal context switch (callee-saved registers + stack pointer); swapcontext costs a sigprocmask system call.
extern "C" void emu_switch(void **save_sp, void *load_sp);
asm(R"(
.text
.globl emu_switch
.type emu_switch,@function
emu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size emu_switch,.-emu_switch
)");
#endif

namespace emu {

thread_local ThreadCtx *cur = nullptr;

namespace {

constexpr size_t STACK_BYTES = 256 * 1024;
constexpr size_t SMEM_GUARD = 4096;
constexpr uint64_t NAN_PATTERN = 0x7ff8dead0badbeefull;   // quiet NaN: unwritten memory that reaches an output shows

[[noreturn]] void die(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    std::fprintf(stderr, "emu: ");
    std::vfprintf(stderr, fmt, ap);
    std::fprintf(stderr, "\n");
    va_end(ap);
    std::abort();
}

void fill_nan(void *p, size_t bytes) {
    uint64_t *q = static_cast<uint64_t *>(p);
    for (size_t i = 0; i < bytes / 8; ++i) q[i] = NAN_PATTERN;
}

struct BulkOp {
    void *dst;
    const void *src;   // nullptr: zero fill (out-of-range part of a tensor-map box)
    uint32_t bytes;
    uint64_t *bar;   // g2s only
};

struct Fiber {
#if defined(__x86_64__)
    void *sp = nullptr;
#else
    ucontext_t ctx;
#endif
    void *stack = nullptr;
    bool done = false;
    const volatile uint64_t *wait_word = nullptr;   // blocked while *wait_word == wait_val (the scheduler skips it)
    uint64_t wait_val = 0;
    ThreadCtx tc;
    int linear = 0;
    std::vector<BulkOp> open_stores, committed_stores;   // lazy mode: shared->global copies not yet performed
};

struct Warp {
    int arrived = 0;
    uint64_t gen = 0;
    uint64_t slot[32];
};

struct MBar {            // lives in the 8 bytes of the kernel's mbarrier word
    int32_t tx;
    int16_t pending;
    uint8_t phase;
    uint8_t count;
};
static_assert(sizeof(MBar) == 8, "mbarrier state must fit the 64-bit word");

struct Block {
    std::vector<Fiber> fibers;
    std::vector<Warp> warps;
    int nthreads = 0;
    unsigned char *smem_alloc = nullptr, *smem = nullptr;
    size_t smem_bytes = 0;
    int bar_arrived = 0, bar_or = 0, bar_or_result = 0;
    uint64_t bar_gen = 0;
    std::vector<BulkOp> pending_loads;   // lazy mode: global->shared copies not yet performed
    uint64_t progress_count = 0;
    const std::function<void()> *body = nullptr;
#if defined(__x86_64__)
    void *sched_sp = nullptr;
#else
    ucontext_t sched;
#endif
    Fiber *running = nullptr;
};

thread_local Block *blk = nullptr;
std::vector<void *> stack_pool;

bool env_lazy() {
    static const bool v = [] { const char *e = getenv("FDFD_EMU_ASYNC"); return e && std::strcmp(e, "lazy") == 0; }();
    return v;
}
long env_shuffle() {
    static const long v = [] { const char *e = getenv("FDFD_EMU_SHUFFLE"); return e ? atol(e) : 0L; }();
    return v;
}

void *get_stack() {
    if (!stack_pool.empty()) { void *s = stack_pool.back(); stack_pool.pop_back(); return s; }
    void *s = mmap(nullptr, STACK_BYTES, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (s == MAP_FAILED) die("mmap of a fiber stack failed");
    return s;
}

void mbar_try_complete(MBar *m) {
    if (m->pending == 0 && m->tx == 0) {
        m->phase ^= 1;
        m->pending = m->count;
        progress();
    }
}

void perform_load(const BulkOp &op) {
    if (op.src) std::memcpy(op.dst, op.src, op.bytes);
    else std::memset(op.dst, 0, op.bytes);
    MBar *m = reinterpret_cast<MBar *>(op.bar);
    m->tx -= (int32_t)op.bytes;
    mbar_try_complete(m);
}

void check_smem_range(const void *p, uint32_t bytes, const char *what) {
    const unsigned char *q = static_cast<const unsigned char *>(p);
    if (q < blk->smem || q + bytes > blk->smem + blk->smem_bytes)
        die("%s: shared-memory range [%td, %td) outside the %zu bytes of dynamic shared memory", what, q - blk->smem,
            q - blk->smem + bytes, blk->smem_bytes);
}

void check_bulk_args(const void *smem_p, const void *glob_p, uint32_t bytes, const char *what) {
    if (bytes == 0 || bytes % 16 != 0) die("%s: size %u is not a positive multiple of 16", what, bytes);
    if ((uintptr_t)smem_p % 16 != 0) die("%s: shared address not 16-byte aligned", what);
    if ((uintptr_t)glob_p % 16 != 0) die("%s: global address not 16-byte aligned", what);
    check_smem_range(smem_p, bytes, what);
}

void flush_stores(std::vector<BulkOp> &v) {
    for (const BulkOp &op : v) std::memcpy(op.dst, op.src, op.bytes);
    if (!v.empty()) progress();
    v.clear();
}

void to_scheduler(Fiber *f) {
#if defined(__x86_64__)
    emu_switch(&f->sp, blk->sched_sp);
#else
    swapcontext(&f->ctx, &blk->sched);
#endif
}
void to_fiber(Fiber *f) {
#if defined(__x86_64__)
    emu_switch(&blk->sched_sp, f->sp);
#else
    swapcontext(&blk->sched, &f->ctx);
#endif
}

void fiber_entry() {
    Fiber *f = blk->running;
    (*blk->body)();
    // kernel end: every bulk store of this thread completes
    flush_stores(f->open_stores);
    flush_stores(f->committed_stores);
    f->done = true;
    progress();
    to_scheduler(f);
    die("a finished fiber was resumed");
}

void init_fiber(Fiber &f, Block &b) {
#if defined(__x86_64__)
    // stack as emu_switch expects it: six callee-saved registers, then the address `ret` jumps to; the entry function
    // must see rsp % 16 == 8, as after a call
    uintptr_t top = ((uintptr_t)f.stack + STACK_BYTES) & ~(uintptr_t)15;
    void **ret_slot = reinterpret_cast<void **>(top - 16);
    *ret_slot = reinterpret_cast<void *>(&fiber_entry);
    void **sp = ret_slot - 6;
    for (int i = 0; i < 6; ++i) sp[i] = nullptr;
    f.sp = sp;
    (void)b;
#else
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp = f.stack;
    f.ctx.uc_stack.ss_size = STACK_BYTES;
    f.ctx.uc_link = &b.sched;
    makecontext(&f.ctx, fiber_entry, 0);
#endif
}

void run_block(dim3 grid, dim3 block, uint3 bid, size_t smem_bytes, const std::function<void()> &body) {
    Block b;
    blk = &b;
    b.body = &body;
    b.nthreads = (int)(block.x * block.y * block.z);
    if (b.nthreads % 32 != 0 && b.nthreads != 1) die("block size %d is not a multiple of 32", b.nthreads);
    b.smem_bytes = smem_bytes;
    if (smem_bytes) {
        b.smem_alloc = static_cast<unsigned char *>(aligned_alloc(1024, ((smem_bytes + 2 * SMEM_GUARD + 1023) / 1024) * 1024));
        fill_nan(b.smem_alloc, ((smem_bytes + 2 * SMEM_GUARD + 1023) / 1024) * 1024);
        b.smem = b.smem_alloc + SMEM_GUARD;
    }
    b.fibers.resize(b.nthreads);
    b.warps.resize((b.nthreads + 31) / 32);
    for (int t = 0; t < b.nthreads; ++t) {
        Fiber &f = b.fibers[t];
        f.linear = t;
        f.tc.tid = uint3{(unsigned)t % block.x, ((unsigned)t / block.x) % block.y, (unsigned)t / (block.x * block.y)};
        f.tc.bid = bid;
        f.tc.bdim = block;
        f.tc.gdim = grid;
        f.stack = get_stack();
        init_fiber(f, b);
    }
    std::mt19937_64 rng((uint64_t)env_shuffle() * 0x9e3779b97f4a7c15ull + bid.x + 7919u * bid.y + 104729u * bid.z);
    std::vector<int> warp_order(b.warps.size());
    for (size_t w = 0; w < warp_order.size(); ++w) warp_order[w] = (int)w;
    int live = b.nthreads;
    uint64_t stuck_rounds = 0;
    while (live > 0) {
        const uint64_t before = b.progress_count;
        if (env_shuffle()) std::shuffle(warp_order.begin(), warp_order.end(), rng);
        for (int w : warp_order) {
            // a shuffled round also lets a warp take several turns in a row, so that it can run ahead
            const int turns = env_shuffle() ? 1 + (int)(rng() % 3) : 1;
            for (int turn = 0; turn < turns; ++turn)
                for (int l = 0; l < 32 && w * 32 + l < b.nthreads; ++l) {
                    Fiber &f = b.fibers[w * 32 + l];
                    if (f.done) continue;
                    if (f.wait_word) {
                        if (*f.wait_word == f.wait_val) continue;   // still blocked: no switch needed
                        f.wait_word = nullptr;
                    }
                    b.running = &f;
                    cur = &f.tc;
                    to_fiber(&f);
                    if (f.done) --live;
                }
        }
        if (b.progress_count != before) { stuck_rounds = 0; continue; }
        // nobody moved: the copy engine "finishes" its oldest outstanding load (lazy mode), else this is a dead-lock
        if (!b.pending_loads.empty()) {
            const BulkOp op = b.pending_loads.front();
            b.pending_loads.erase(b.pending_loads.begin());
            perform_load(op);
            progress();
            continue;
        }
        if (++stuck_rounds > 3)
            die("dead-lock in block (%u,%u,%u): %d threads alive, none can proceed (barrier %d/%d arrived)", bid.x, bid.y,
                bid.z, live, b.bar_arrived, b.nthreads);
    }
    if (!b.pending_loads.empty()) die("block (%u,%u,%u) exited with %zu bulk loads in flight", bid.x, bid.y, bid.z, b.pending_loads.size());
    for (Fiber &f : b.fibers) stack_pool.push_back(f.stack);
    if (b.smem_alloc) {
        // guard regions must still hold the fill pattern
        const uint64_t *g0 = reinterpret_cast<const uint64_t *>(b.smem_alloc);
        const uint64_t *g1 = reinterpret_cast<const uint64_t *>(b.smem + ((smem_bytes + 7) / 8) * 8);
        for (size_t i = 0; i < SMEM_GUARD / 8; ++i)
            if (g0[i] != NAN_PATTERN) die("block (%u,%u,%u) wrote below its dynamic shared memory", bid.x, bid.y, bid.z);
        for (size_t i = 0; i < (SMEM_GUARD - 8) / 8; ++i)
            if (g1[i] != NAN_PATTERN) die("block (%u,%u,%u) wrote beyond its dynamic shared memory", bid.x, bid.y, bid.z);
        free(b.smem_alloc);
    }
    blk = nullptr;
    cur = nullptr;
}

Fiber *self() { return blk->running; }

}  // namespace

void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()> &body) {
    if (blk) die("nested kernel launch");
    if (grid.x == 0 || grid.y == 0 || grid.z == 0) return;   // CUDA reports an error; callers never do this
    if (smem_bytes > 227 * 1024) die("launch asks for %zu bytes of shared memory (> 227 KB)", smem_bytes);
    if (block.x * block.y * block.z > 1024) die("more than 1024 threads per block");
    for (unsigned z = 0; z < grid.z; ++z)
        for (unsigned y = 0; y < grid.y; ++y)
            for (unsigned x = 0; x < grid.x; ++x) run_block(grid, block, uint3{x, y, z}, smem_bytes, body);
}

unsigned char *dyn_smem() { return blk->smem; }
bool lazy_async() { return env_lazy(); }
void progress() { if (blk) ++blk->progress_count; }

void yield() { to_scheduler(self()); }

// block the calling thread until *word != val
static void wait_while_equal(const volatile uint64_t *word, uint64_t val) {
    Fiber *f = self();
    while (*word == val) {
        f->wait_word = word;
        f->wait_val = val;
        to_scheduler(f);
    }
    f->wait_word = nullptr;
}

void syncthreads() { (void)syncthreads_or(0); }

int syncthreads_or(int v) {
    Block *b = blk;
    b->bar_or |= (v != 0);
    if (++b->bar_arrived == b->nthreads) {
        b->bar_arrived = 0;
        b->bar_or_result = b->bar_or;
        b->bar_or = 0;
        ++b->bar_gen;
        progress();
        return b->bar_or_result;
    }
    wait_while_equal(&b->bar_gen, b->bar_gen);
    return b->bar_or_result;
}

static void warp_barrier(Warp &w, int nl) {
    if (++w.arrived == nl) {
        w.arrived = 0;
        ++w.gen;
        progress();
        return;
    }
    wait_while_equal(&w.gen, w.gen);
}

int lane_id() { return self()->linear & 31; }

void syncwarp() {
    Block *b = blk;
    const int wid = self()->linear >> 5;
    warp_barrier(b->warps[wid], std::min(32, b->nthreads - wid * 32));
}

uint64_t shfl_xor_bits(uint64_t v, int lane_mask) {
    Block *b = blk;
    const int lin = self()->linear, wid = lin >> 5, lane = lin & 31;
    const int nl = std::min(32, b->nthreads - wid * 32);
    Warp &w = b->warps[wid];
    w.slot[lane] = v;
    warp_barrier(w, nl);
    const int src = lane ^ lane_mask;
    const uint64_t r = src < nl ? w.slot[src] : v;
    warp_barrier(w, nl);
    return r;
}

uint64_t shfl_idx_bits(uint64_t v, int src) {
    Block *b = blk;
    const int lin = self()->linear, wid = lin >> 5, lane = lin & 31;
    const int nl = std::min(32, b->nthreads - wid * 32);
    Warp &w = b->warps[wid];
    w.slot[lane] = v;
    warp_barrier(w, nl);
    const uint64_t r = (src >= 0 && src < nl) ? w.slot[src] : v;
    warp_barrier(w, nl);
    return r;
}

// ---- mbarrier / bulk copies --------------------------------------------------------------------------------------
void mbar_init(uint64_t *bar, uint32_t count) {
    check_smem_range(bar, 8, "mbarrier.init");
    MBar *m = reinterpret_cast<MBar *>(bar);
    m->tx = 0;
    m->pending = (int16_t)count;
    m->phase = 0;
    m->count = (uint8_t)count;
}
void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    MBar *m = reinterpret_cast<MBar *>(bar);
    if (m->pending <= 0) die("mbarrier.arrive: more arrivals than the barrier was initialised for");
    m->tx += (int32_t)bytes;
    m->pending -= 1;
    mbar_try_complete(m);
}
void mbar_wait(uint64_t *bar, uint32_t parity) {
    const MBar *m = reinterpret_cast<const MBar *>(bar);
    while (m->phase == (parity & 1u)) yield();   // the phase with this parity has not completed yet
    progress();
}
void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    check_bulk_args(dst, src, bytes, "cp.async.bulk global->shared");
    const BulkOp op{dst, src, bytes, bar};
    if (env_lazy()) blk->pending_loads.push_back(op);
    else perform_load(op);
}
// tensor-map box load: row by row, the in-range run copied, the rest zero-filled; every piece counts on the mbarrier
void tma_load_3d(void *dst, const unsigned long long *map, int c0, int c1, int c2, uint64_t *bar) {
    const unsigned char *base = reinterpret_cast<const unsigned char *>(map[0]);
    const long d0 = (long)map[1], d1 = (long)map[2], d2 = (long)map[3];
    const long s1 = (long)map[4], s2 = (long)map[5];
    const long b0 = (long)map[6], b1 = (long)map[7], b2 = (long)map[8];
    if (!base || b0 <= 0) die("tensor-map load through an unencoded map");
    if (((uintptr_t)dst & 127) != 0) die("tensor-map load: shared destination not 128-byte aligned");
    if (((long)c0 * 8) % 16 != 0) die("tensor-map load: box starts at element %d - not on a 16-byte boundary of the global array", c0);
    if ((b0 * 8) % 16 != 0) die("tensor-map load: inner box extent %ld bytes is not a multiple of 16", b0 * 8);
    check_smem_range(dst, (uint32_t)(b0 * b1 * b2 * 8), "cp.async.bulk.tensor global->shared");
    unsigned char *out = static_cast<unsigned char *>(dst);
    auto push = [&](void *d, const void *s, long n) {
        if (n <= 0) return;
        const BulkOp op{d, s, (uint32_t)n, bar};
        if (env_lazy()) blk->pending_loads.push_back(op);
        else perform_load(op);
    };
    for (long z = 0; z < b2; ++z)
        for (long y = 0; y < b1; ++y) {
            unsigned char *row = out + ((z * b1 + y) * b0) * 8;
            const long gz = c2 + z, gy = c1 + y;
            if (gz < 0 || gz >= d2 || gy < 0 || gy >= d1) { push(row, nullptr, b0 * 8); continue; }
            const long lo = std::max<long>(c0, 0), hi = std::min<long>(c0 + b0, d0);
            if (hi <= lo) { push(row, nullptr, b0 * 8); continue; }
            push(row, nullptr, (lo - c0) * 8);
            push(row + (lo - c0) * 8, base + gz * s2 + gy * s1 + lo * 8, (hi - lo) * 8);
            push(row + (hi - c0) * 8, nullptr, (c0 + b0 - hi) * 8);
        }
}
// tensor-map box store: the in-range part of every row, as one bulk-group member per piece
void tma_store_3d(const unsigned long long *map, int c0, int c1, int c2, const void *src) {
    unsigned char *base = reinterpret_cast<unsigned char *>(map[0]);
    const long d0 = (long)map[1], d1 = (long)map[2], d2 = (long)map[3];
    const long s1 = (long)map[4], s2 = (long)map[5];
    const long b0 = (long)map[6], b1 = (long)map[7], b2 = (long)map[8];
    if (!base || b0 <= 0) die("tensor-map store through an unencoded map");
    if (((uintptr_t)src & 127) != 0) die("tensor-map store: shared source not 128-byte aligned");
    if (((long)c0 * 8) % 16 != 0) die("tensor-map store: box starts at element %d - not on a 16-byte boundary of the global array", c0);
    check_smem_range(src, (uint32_t)(b0 * b1 * b2 * 8), "cp.async.bulk.tensor shared->global");
    const unsigned char *in = static_cast<const unsigned char *>(src);
    for (long z = 0; z < b2; ++z)
        for (long y = 0; y < b1; ++y) {
            const long gz = c2 + z, gy = c1 + y;
            if (gz < 0 || gz >= d2 || gy < 0 || gy >= d1) continue;
            const long lo = std::max<long>(c0, 0), hi = std::min<long>(c0 + b0, d0);
            if (hi <= lo) continue;
            void *d = base + gz * s2 + gy * s1 + lo * 8;
            const void *s = in + ((z * b1 + y) * b0 + (lo - c0)) * 8;
            if (env_lazy()) self()->open_stores.push_back(BulkOp{d, s, (uint32_t)((hi - lo) * 8), nullptr});
            else std::memcpy(d, s, (hi - lo) * 8);
        }
}
void bulk_s2g(void *dst, const void *src, uint32_t bytes) {
    check_bulk_args(src, dst, bytes, "cp.async.bulk shared->global");
    if (env_lazy()) self()->open_stores.push_back(BulkOp{dst, src, bytes, nullptr});
    else std::memcpy(dst, src, bytes);
}
void bulk_commit() {
    Fiber *f = self();
    f->committed_stores.insert(f->committed_stores.end(), f->open_stores.begin(), f->open_stores.end());
    f->open_stores.clear();
}
void bulk_wait_read0() { flush_stores(self()->committed_stores); }
void bulk_wait0() { flush_stores(self()->committed_stores); }

}  // namespace emu

// ---- runtime API -------------------------------------------------------------------------------------------------
const char *cudaGetErrorString(cudaError_t e) {
    switch (e) {
        case cudaSuccess: return "no error";
        case cudaErrorInvalidValue: return "invalid argument";
        case cudaErrorMemoryAllocation: return "out of memory";
        case cudaErrorInvalidConfiguration: return "invalid configuration argument";
        case cudaErrorNotSupported: return "operation not supported (emulation)";
        default: return "unknown error";
    }
}
cudaError_t cudaGetLastError() { return cudaSuccess; }
// "Device" memory.  With FDFD_EMU_IPC=1 (multi-rank runs, one process per rank) every allocation is a named POSIX
// shared-memory object so that cudaIpcGetMemHandle / cudaIpcOpenMemHandle can map it into a neighbour's process.
namespace {
struct DevAlloc { size_t bytes; std::string shm_name; };
std::map<void *, DevAlloc> dev_allocs;
bool env_ipc() {
    static const bool v = getenv("FDFD_EMU_IPC") != nullptr;
    return v;
}
struct IpcPayload { char name[48]; uint64_t offset; uint64_t bytes; };
static_assert(sizeof(IpcPayload) <= sizeof(cudaIpcMemHandle_t), "payload must fit the handle");
void unlink_all_shm() {
    for (auto &kv : dev_allocs)
        if (!kv.second.shm_name.empty()) shm_unlink(kv.second.shm_name.c_str());
}
}  // namespace

cudaError_t cudaMalloc(void **p, size_t bytes) {
    const size_t n = ((bytes + 255) / 256) * 256 + 256;
    if (!env_ipc()) {
        *p = aligned_alloc(256, n);
        if (!*p) return cudaErrorMemoryAllocation;
        emu::fill_nan(*p, n);
        dev_allocs[*p] = DevAlloc{n, ""};
        return cudaSuccess;
    }
    static int counter = 0;
    static bool registered = false;
    if (!registered) { atexit(unlink_all_shm); registered = true; }
    const std::string name = "/fdfd_emu_" + std::to_string((long)getpid()) + "_" + std::to_string(counter++);
    const int fd = shm_open(name.c_str(), O_CREAT | O_EXCL | O_RDWR, 0600);
    if (fd < 0 || ftruncate(fd, (off_t)n) != 0) return cudaErrorMemoryAllocation;
    void *q = mmap(nullptr, n, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (q == MAP_FAILED) return cudaErrorMemoryAllocation;
    emu::fill_nan(q, n);
    dev_allocs[q] = DevAlloc{n, name};
    *p = q;
    return cudaSuccess;
}
cudaError_t cudaFree(void *p) {
    if (!p) return cudaSuccess;
    auto it = dev_allocs.find(p);
    if (it == dev_allocs.end()) { std::fprintf(stderr, "emu: cudaFree of an unknown pointer\n"); std::abort(); }
    if (it->second.shm_name.empty()) free(p);
    else {
        munmap(p, it->second.bytes);
        shm_unlink(it->second.shm_name.c_str());
    }
    dev_allocs.erase(it);
    return cudaSuccess;
}
cudaError_t cudaMallocHost(void **p, size_t bytes) {
    *p = aligned_alloc(256, ((bytes + 255) / 256) * 256 + 256);
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
cudaError_t cudaMemcpy(void *dst, const void *src, size_t n, cudaMemcpyKind) { std::memmove(dst, src, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t n, cudaMemcpyKind, cudaStream_t) { std::memmove(dst, src, n); return cudaSuccess; }
cudaError_t cudaMemset(void *p, int v, size_t n) { std::memset(p, v, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t) { std::memset(p, v, n); return cudaSuccess; }
cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
cudaError_t cudaDeviceGetAttribute(int *v, cudaDeviceAttr, int) {   // FDFD_EMU_SMS: persistent grids of a few CTAs
    const char *e = getenv("FDFD_EMU_SMS");
    *v = e ? atoi(e) : 148;
    return cudaSuccess;
}
cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
cudaError_t cudaDeviceGetStreamPriorityRange(int *lo, int *hi) { *lo = 0; *hi = 0; return cudaSuccess; }
static cudaError_t new_stream(cudaStream_t *s) { *s = reinterpret_cast<cudaStream_t>(malloc(8)); return cudaSuccess; }
cudaError_t cudaStreamCreate(cudaStream_t *s) { return new_stream(s); }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { return new_stream(s); }
cudaError_t cudaStreamCreateWithPriority(cudaStream_t *s, unsigned, int) { return new_stream(s); }
cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaStreamBeginCapture(cudaStream_t, int) { return cudaErrorNotSupported; }
cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t *g) { *g = nullptr; return cudaErrorNotSupported; }
cudaError_t cudaGraphInstantiate(cudaGraphExec_t *e, cudaGraph_t, unsigned long long) { *e = nullptr; return cudaErrorNotSupported; }
cudaError_t cudaGraphLaunch(cudaGraphExec_t, cudaStream_t) { return cudaErrorNotSupported; }
cudaError_t cudaGraphDestroy(cudaGraph_t) { return cudaSuccess; }
cudaError_t cudaGraphExecDestroy(cudaGraphExec_t) { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = reinterpret_cast<cudaEvent_t>(malloc(8)); return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 1.0f; return cudaSuccess; }
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *p) {
    auto it = dev_allocs.upper_bound(p);
    if (it == dev_allocs.begin()) return cudaErrorInvalidValue;
    --it;
    const char *base = static_cast<const char *>(it->first);
    if (static_cast<const char *>(p) >= base + it->second.bytes || it->second.shm_name.empty()) return cudaErrorNotSupported;
    IpcPayload pl{};
    std::strncpy(pl.name, it->second.shm_name.c_str(), sizeof(pl.name) - 1);
    pl.offset = (uint64_t)(static_cast<const char *>(p) - base);
    pl.bytes = it->second.bytes;
    std::memset(h, 0, sizeof(*h));
    std::memcpy(h, &pl, sizeof(pl));
    return cudaSuccess;
}
namespace { std::map<void *, std::pair<void *, size_t>> ipc_maps; }   // returned pointer -> (mapping base, bytes)
cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned) {
    IpcPayload pl;
    std::memcpy(&pl, &h, sizeof(pl));
    const int fd = shm_open(pl.name, O_RDWR, 0600);
    if (fd < 0) return cudaErrorInvalidValue;
    void *q = mmap(nullptr, pl.bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (q == MAP_FAILED) return cudaErrorMemoryAllocation;
    *p = static_cast<char *>(q) + pl.offset;
    ipc_maps[*p] = {q, (size_t)pl.bytes};
    return cudaSuccess;
}
cudaError_t cudaIpcCloseMemHandle(void *p) {
    auto it = ipc_maps.find(p);
    if (it == ipc_maps.end()) return cudaErrorInvalidValue;
    munmap(it->second.first, it->second.second);
    ipc_maps.erase(it);
    return cudaSuccess;
}
