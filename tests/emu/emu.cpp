// emu.cpp -- fiber scheduler of the SIMT emulator and a stand-in for the handful of CUDA runtime calls the library
// makes (device memory = host memory, one in-order "stream").  TEST INFRASTRUCTURE ONLY, see emu.h.
#include "emu.h"

#include <fcntl.h>
#include <sched.h>
#include <stdio.h>
#include <sys/mman.h>
#include <time.h>
#include <unistd.h>

#include <map>
#include <string>
#include <vector>

void emu_yield_cpu();

namespace emu {

ThreadCtx *cur = nullptr;
int block_or_flag = 0;
unsigned char *dyn_smem_ptr = nullptr;

// ---- context switch ---------------------------------------------------------------------------------------------------
#if defined(__x86_64__)
extern "C" void emu_switch(void **from_sp, void *to_sp);
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
#else
#error "the SIMT emulator's context switch is written for x86-64"
#endif

enum State { READY, WAIT_BLOCK, WAIT_WARP, DONE };

struct Fiber {
    void *sp = nullptr;
    State state = DONE;
    ThreadCtx tc;
};

struct Warp {
    unsigned arrived = 0, alive = 0, mask = 0;
    Op op = OP_SYNC;
    int width = 32;
    uint64_t val[32];
    int param[32];
    uint64_t res[32];
};

static const size_t STACK_BYTES = 256 * 1024;
static std::vector<Fiber> fibers;
static std::vector<Warp> warps;
static unsigned char *stack_pool = nullptr;
static size_t stack_pool_threads = 0;
static void *sched_sp = nullptr;
static const std::function<void()> *body_fn = nullptr;
static int barrier_waiting = 0, alive_threads = 0;
static Fiber *running = nullptr;

// OSPH_EMU_ORDER=reverse | shuffle[:seed] -- the order in which the scheduler resumes the ready threads of a CTA and runs the
// CTAs of a launch.  A kernel without races (inside a CTA: every shared-memory hand-over behind a barrier or a warp
// collective; between CTAs: no dependence on the order of the grid) gives the same bits under every order, so running the
// parity tests under two more orders flushes out what the ascending default hides (a reader that happens to run after
// its writer).  Default: ascending.
static int g_order_mode = -1;       // -1: read OSPH_EMU_ORDER on first use
static uint64_t g_order_state = 0;
static int order_mode()
{
    if (g_order_mode < 0) {
        const char *e = getenv("OSPH_EMU_ORDER");
        g_order_mode = !e ? 0 : (!strncmp(e, "reverse", 7) ? 1 : (!strncmp(e, "shuffle", 7) ? 2 : 0));
        const char *c = e ? strchr(e, ':') : nullptr;
        g_order_state = 0x9e3779b97f4a7c15ull ^ (c ? strtoull(c + 1, nullptr, 0) : 1);
    }
    return g_order_mode;
}
static uint64_t order_rng()
{
    uint64_t &s = g_order_state;
    if (!s) s = 0x9e3779b97f4a7c15ull;
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    return s;
}
static void fill_order(std::vector<int> &o, size_t n)
{
    o.resize(n);
    for (size_t k = 0; k < n; k++) o[k] = (int)(order_mode() == 1 ? n - 1 - k : k);
    if (order_mode() == 2) for (size_t k = n; k > 1; k--) std::swap(o[k - 1], o[order_rng() % k]);
}

static void yield_to_scheduler()
{
    Fiber *f = running;
    emu_switch(&f->sp, sched_sp);
}

static void fiber_entry()
{
    (*body_fn)();
    running->state = DONE;
    yield_to_scheduler();
    abort();      // a finished fiber is never resumed
}

static void prepare_fiber(Fiber &f, int slot)
{
    unsigned char *top = stack_pool + (size_t)(slot + 1) * STACK_BYTES;
    void **sp = reinterpret_cast<void **>(top);
    *--sp = nullptr;                                      // fake return address of fiber_entry: keeps rsp = 8 mod 16
    *--sp = reinterpret_cast<void *>(&fiber_entry);       // popped by emu_switch's ret
    for (int k = 0; k < 6; k++) *--sp = nullptr;          // rbp rbx r12 r13 r14 r15
    f.sp = sp;
    f.state = READY;
}

}  // namespace emu

// __nanosleep inside a spin loop (the mailbox kernels of slab_p2p.cu wait for a peer PROCESS): give the CPU away and let the
// other fibers of the CTA run, as the other warps of a CTA do on the GPU while one of them polls.  Without the fiber switch
// a schedule other than the ascending default can deadlock ranks against each other (rank A polls for B before the fiber
// that publishes to C has run, B polls for C, C for A).
long long emu_clock64()
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ((long long)ts.tv_sec * 1000000000ll + ts.tv_nsec) * 2;
}

void emu_yield_cpu()
{
    sched_yield();
    if (emu::running) { emu::running->state = emu::READY; emu::yield_to_scheduler(); }
}

namespace emu {

void block_barrier()
{
    running->state = WAIT_BLOCK;
    barrier_waiting++;
    yield_to_scheduler();
}

uint64_t warp_collective(Op op, unsigned mask, uint64_t value, int param, int width)
{
    Fiber *f = running;
    Warp &w = warps[f->tc.warp];
    const int lane = f->tc.lane;
    if (w.arrived == 0) { w.op = op; w.mask = mask; w.width = width; }
    else if (w.op != op) { fprintf(stderr, "emu: lanes of one warp reached different collectives (%d vs %d)\n", (int)w.op, (int)op); abort(); }
    w.val[lane] = value; w.param[lane] = param;
    w.arrived |= 1u << lane;
    f->state = WAIT_WARP;
    yield_to_scheduler();
    return w.res[lane];
}

static void complete_warp(Warp &w, int warp_index, int nthreads)
{
    const unsigned part = w.arrived;            // participating lanes
    for (int l = 0; l < 32; l++) {
        if (!(part >> l & 1)) continue;
        uint64_t r = 0;
        switch (w.op) {
        case OP_SHFL_IDX: case OP_SHFL_XOR: case OP_SHFL_UP: case OP_SHFL_DOWN: {
            const int wd = w.width, seg = l & ~(wd - 1);
            int src;
            if (w.op == OP_SHFL_IDX) src = seg + (w.param[l] & (wd - 1));
            else if (w.op == OP_SHFL_XOR) src = l ^ w.param[l];
            else if (w.op == OP_SHFL_UP) src = l - w.param[l];
            else src = l + w.param[l];
            bool ok = src >= seg && src < seg + wd && (part >> src & 1);
            if (w.op == OP_SHFL_XOR) ok = src >= 0 && src < 32 && (src & ~(wd - 1)) == seg && (part >> src & 1);
            r = ok ? w.val[src] : w.val[l];
            break;
        }
        case OP_BALLOT: {
            unsigned b = 0;
            for (int k = 0; k < 32; k++) if ((part >> k & 1) && w.val[k]) b |= 1u << k;
            r = b & w.mask;
            break;
        }
        case OP_MATCH: {
            unsigned b = 0;
            for (int k = 0; k < 32; k++) if ((part >> k & 1) && w.val[k] == w.val[l]) b |= 1u << k;
            r = b;
            break;
        }
        case OP_SYNC: break;
        }
        w.res[l] = r;
    }
    for (int l = 0; l < 32; l++)
        if (part >> l & 1) {
            int t = warp_index * 32 + l;
            if (t < nthreads) fibers[t].state = READY;
        }
    w.arrived = 0;
}

static void run_block(dim3 grid, dim3 block, uint3 bid)
{
    const int nthreads = (int)(block.x * block.y * block.z);
    const int nwarps = (nthreads + 31) / 32;
    fibers.resize(nthreads);
    warps.assign(nwarps, Warp());
    for (int t = 0; t < nthreads; t++) {
        Fiber &f = fibers[t];
        prepare_fiber(f, t);
        f.tc.tid.x = t % block.x; f.tc.tid.y = (t / block.x) % block.y; f.tc.tid.z = t / (block.x * block.y);
        f.tc.bid = bid; f.tc.bdim = block; f.tc.gdim = grid;
        f.tc.linear = t; f.tc.lane = t & 31; f.tc.warp = t >> 5;
        warps[t >> 5].alive |= 1u << (t & 31);
    }
    alive_threads = nthreads;
    barrier_waiting = 0;
    static std::vector<int> order;
    fill_order(order, (size_t)nthreads);
    while (alive_threads > 0) {
        bool progressed = false;
        if (order_mode() == 2) fill_order(order, (size_t)nthreads);      // a new permutation after every switch point
        for (int k = 0; k < nthreads; k++) {
            const int t = order[k];
            Fiber &f = fibers[t];
            if (f.state != READY) continue;
            running = &f; cur = &f.tc;
            emu_switch(&sched_sp, f.sp);
            progressed = true;
            if (f.state == DONE) { alive_threads--; warps[t >> 5].alive &= ~(1u << (t & 31)); }
        }
        for (int wi = 0; wi < nwarps; wi++) {
            Warp &w = warps[wi];
            // every live lane named by the mask has arrived (exited lanes count as arrived)
            if (w.arrived && (w.arrived & w.mask) == (w.alive & w.mask)) { complete_warp(w, wi, nthreads); progressed = true; }
        }
        if (barrier_waiting > 0 && barrier_waiting == alive_threads) {
            for (int t = 0; t < nthreads; t++) if (fibers[t].state == WAIT_BLOCK) fibers[t].state = READY;
            barrier_waiting = 0;
            progressed = true;
        }
        if (!progressed) {
            fprintf(stderr, "emu: deadlock in block (%u,%u,%u): %d threads alive, %d at __syncthreads\n", bid.x, bid.y, bid.z,
                    alive_threads, barrier_waiting);
            abort();
        }
    }
    running = nullptr; cur = nullptr;
}

// Guarded storage: [PROT_NONE region][bytes rounded up to pages][PROT_NONE region]; the caller places its object at the END
// of the accessible part, so the first byte past the object is already inaccessible.  The inaccessible regions are 16 MiB
// wide: shared memory is indexed with up to 16 bits times a record size, so a wild index lands megabytes -- not bytes --
// past the end (the overrun of round 1 read 43 KB past it, far beyond a single guard page).
static const size_t GUARD_BYTES = 16u << 20;
static unsigned char *guarded_pages(size_t bytes, size_t *usable)
{
    const size_t page = 4096, body = (bytes + page - 1) / page * page;
    unsigned char *m = (unsigned char *)mmap(nullptr, body + 2 * GUARD_BYTES, PROT_NONE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (m == MAP_FAILED) { perror("emu: mmap"); abort(); }
    if (mprotect(m + GUARD_BYTES, body, PROT_READ | PROT_WRITE) != 0) { perror("emu: mprotect"); abort(); }
    *usable = body;
    return m + GUARD_BYTES;
}
static const size_t DYN_ARENA = 256 * 1024;       // more than the 227 KB a CTA can have
static unsigned char *dyn_arena()
{
    static unsigned char *a = nullptr;
    if (!a) { size_t u; a = guarded_pages(DYN_ARENA, &u); }
    return a;
}
// one `__shared__` variable of static size (preprocess.py turns the declaration into a reference to this storage)
void *static_smem_alloc(size_t bytes, size_t align)
{
    size_t usable;
    unsigned char *a = guarded_pages(bytes, &usable);
    if (align < 1) align = 1;
    const size_t padded = (bytes + align - 1) / align * align;      // sizeof(T) is a multiple of alignof(T) already
    memset(a, 0xA5, usable);
    return a + usable - padded;
}

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()> &body)
{
    const size_t nthreads = (size_t)block.x * block.y * block.z;
    if (nthreads == 0 || nthreads > 1024 || (size_t)grid.x * grid.y * grid.z == 0) return;
    if (running) { fprintf(stderr, "emu: nested launch\n"); abort(); }
    if (nthreads > stack_pool_threads) {
        if (stack_pool) munmap(stack_pool, stack_pool_threads * STACK_BYTES);
        stack_pool_threads = std::max<size_t>(nthreads, 256);
        stack_pool = (unsigned char *)mmap(nullptr, stack_pool_threads * STACK_BYTES, PROT_READ | PROT_WRITE,
                                           MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (stack_pool == MAP_FAILED) { perror("emu: mmap"); abort(); }
    }
    // dynamic shared memory ends right in front of an inaccessible page (and the mapping starts behind one): a CTA that
    // reads or writes past the bytes it was launched with faults at the instruction, as it would under compute-sanitizer
    // (round 1 shipped an overrun that a heap buffer with slack silently absorbed).  The size is rounded up to 16 bytes
    // to keep the 16-byte alignment the kernels ask for.
    const size_t dyn_bytes = (smem + 15) & ~size_t(15);
    if (dyn_bytes > DYN_ARENA) { fprintf(stderr, "emu: %zu bytes of dynamic shared memory\n", smem); abort(); }
    unsigned char *arena = dyn_arena();
    dyn_smem_ptr = arena + DYN_ARENA - dyn_bytes;
    memset(dyn_smem_ptr, 0xA5, dyn_bytes);
    body_fn = &body;
    const size_t nblocks = (size_t)grid.x * grid.y * grid.z;
    std::vector<int> border;
    fill_order(border, nblocks);
    for (size_t k = 0; k < nblocks; k++) {
        const size_t b = (size_t)border[k];
        uint3 bid; bid.x = (unsigned)(b % grid.x); bid.y = (unsigned)(b / grid.x % grid.y); bid.z = (unsigned)(b / ((size_t)grid.x * grid.y));
        run_block(grid, block, bid);
    }
    body_fn = nullptr; dyn_smem_ptr = nullptr;
}

// MUFU.RCP64H / MUFU.RSQ64H model: a seed whose lower 32 mantissa bits are zero (about 2^-20 relative error); the
// Newton steps of sph_math.cuh have to recover the rest, which is what the emulated runs then check.
static double chop32(double v)
{
    uint64_t u; memcpy(&u, &v, 8); u &= 0xffffffff00000000ull; memcpy(&v, &u, 8); return v;
}
double rcp_approx_f64(double x) { return chop32(1.0 / x); }
double rsqrt_approx_f64(double x) { return chop32(1.0 / std::sqrt(x)); }

}  // namespace emu

// =====================================================================================================================
// CUDA runtime stand-in: just what libosph_b200 calls.  "Device" memory is host memory filled with a garbage pattern.
// =====================================================================================================================
static cudaError_t g_last = cudaSuccess;

// OSPH_EMU_GUARD=1: every device allocation ends right in front of an inaccessible page (and starts right behind one),
// so a kernel that reads or writes past its buffer faults at the offending instruction -- a poor man's
// compute-sanitizer memcheck.  OSPH_EMU_FILL=<byte> changes the garbage pattern fresh allocations are filled with
// (two runs with different patterns that agree bit for bit do not depend on uninitialised device memory).
//
// CUDA IPC (the peer-memory slab sequencer, slab_p2p.cu): allocations of 4 KiB and more are page-exclusive anonymous
// mappings; cudaIpcGetMemHandle turns one into a POSIX shared-memory mapping AT THE SAME ADDRESS (contents kept) and hands
// out its name, cudaIpcOpenMemHandle maps that object in the peer process.  Ranks are separate processes, so the mailbox
// kernels that spin on a peer's sequence number run against real concurrency.
struct Alloc { void *base; size_t mapped; int kind; std::string shm; };      // kind 0 heap, 1 guarded, 2 mmap, 3 shm owner, 4 shm peer
static std::map<void *, Alloc> g_allocs;
static bool guard_mode() { static int g = -1; if (g < 0) { const char *e = getenv("OSPH_EMU_GUARD"); g = e && *e == '1'; } return g == 1; }
static int fill_byte() { static int f = -1; if (f < 0) { const char *e = getenv("OSPH_EMU_FILL"); f = e ? (int)strtol(e, nullptr, 0) & 255 : 0xA5; } return f; }
static const size_t PAGE = 4096;
static void unlink_all() { for (auto &kv : g_allocs) if (kv.second.kind == 3) shm_unlink(kv.second.shm.c_str()); }

extern "C" {

// tests switch the schedule inside one process: 0 ascending, 1 reverse, 2 shuffle (seeded)
void emu_set_order(int mode, unsigned long long seed)
{
    emu::g_order_mode = mode;
    emu::g_order_state = 0x9e3779b97f4a7c15ull ^ (seed ? seed : 1);
}

cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaGetLastError(void) { cudaError_t e = g_last; g_last = cudaSuccess; return e; }
const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA error"; }

cudaError_t cudaMalloc(void **p, size_t bytes)
{
    if (guard_mode()) {
        const size_t user = (bytes + 15) & ~(size_t)15;          // 16-byte vector loads stay aligned
        const size_t body = (user + PAGE - 1) / PAGE * PAGE, total = body + 2 * PAGE;
        unsigned char *m = (unsigned char *)mmap(nullptr, total, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (m == MAP_FAILED) return cudaErrorMemoryAllocation;
        memset(m, fill_byte(), total);
        mprotect(m, PAGE, PROT_NONE);
        mprotect(m + PAGE + body, PAGE, PROT_NONE);
        unsigned char *u = m + PAGE + body - user;
        g_allocs[u] = Alloc{m, total, 1, ""};
        *p = u;
        return cudaSuccess;
    }
    if (bytes >= 4096) {
        const size_t total = (bytes + PAGE - 1) / PAGE * PAGE;
        void *m = mmap(nullptr, total, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (m == MAP_FAILED) return cudaErrorMemoryAllocation;
        memset(m, fill_byte(), total);
        g_allocs[m] = Alloc{m, total, 2, ""};
        *p = m;
        return cudaSuccess;
    }
    size_t b = (bytes + 255) & ~(size_t)255;
    if (b == 0) b = 256;
    void *q = aligned_alloc(256, b);
    if (!q) return cudaErrorMemoryAllocation;
    memset(q, fill_byte(), b);
    *p = q;
    return cudaSuccess;
}
cudaError_t cudaFree(void *p)
{
    if (!p) return cudaSuccess;
    auto it = g_allocs.find(p);
    if (it != g_allocs.end()) {
        munmap(it->second.base, it->second.mapped);
        if (it->second.kind == 3) shm_unlink(it->second.shm.c_str());
        g_allocs.erase(it);
        return cudaSuccess;
    }
    free(p);
    return cudaSuccess;
}

struct IpcName { char name[48]; unsigned long long bytes; unsigned long long magic; };
static_assert(sizeof(IpcName) == 64, "fits cudaIpcMemHandle_t");

cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *ptr)
{
    auto it = g_allocs.find(ptr);
    if (it == g_allocs.end() || (it->second.kind != 2 && it->second.kind != 3)) return cudaErrorNotSupported;
    Alloc &a = it->second;
    if (a.kind == 2) {
        static int counter = 0;
        static bool hooked = false;
        if (!hooked) { atexit(unlink_all); hooked = true; }
        char name[48];
        snprintf(name, sizeof name, "/osph_emu_%d_%d", (int)getpid(), counter++);
        int fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
        if (fd < 0 || ftruncate(fd, (off_t)a.mapped) != 0) { if (fd >= 0) { close(fd); shm_unlink(name); } return cudaErrorMemoryAllocation; }
        void *tmp = mmap(nullptr, a.mapped, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        if (tmp == MAP_FAILED) { close(fd); shm_unlink(name); return cudaErrorMemoryAllocation; }
        memcpy(tmp, a.base, a.mapped);
        munmap(tmp, a.mapped);
        void *same = mmap(a.base, a.mapped, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_FIXED, fd, 0);     // same address, now shared
        close(fd);
        if (same != a.base) { shm_unlink(name); return cudaErrorMemoryAllocation; }
        a.kind = 3; a.shm = name;
    }
    IpcName n; memset(&n, 0, sizeof n);
    snprintf(n.name, sizeof n.name, "%s", a.shm.c_str()); n.bytes = a.mapped; n.magic = 0x6f7370686d656d75ull;
    memcpy(h, &n, 64);
    return cudaSuccess;
}
cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned int)
{
    IpcName n; memcpy(&n, &h, 64);
    if (n.magic != 0x6f7370686d656d75ull) return cudaErrorInvalidValue;
    int fd = shm_open(n.name, O_RDWR, 0600);
    if (fd < 0) return cudaErrorInvalidValue;
    void *m = mmap(nullptr, (size_t)n.bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (m == MAP_FAILED) return cudaErrorMemoryAllocation;
    g_allocs[m] = Alloc{m, (size_t)n.bytes, 4, n.name};
    *p = m;
    return cudaSuccess;
}
cudaError_t cudaIpcCloseMemHandle(void *p) { return cudaFree(p); }

cudaError_t cudaStreamQuery(cudaStream_t) { return cudaSuccess; }      // launches run to completion before they return
cudaError_t cudaMallocHost(void **p, size_t bytes) { *p = malloc(bytes ? bytes : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaHostAlloc(void **p, size_t bytes, unsigned int) { return cudaMallocHost(p, bytes); }
cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { if (n) memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { if (n) memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemset(void *d, int v, size_t n) { if (n) memset(d, v, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) { if (n) memset(d, v, n); return cudaSuccess; }

cudaError_t cudaStreamCreate(cudaStream_t *s) { *s = reinterpret_cast<cudaStream_t>(malloc(8)); return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned int) { return cudaStreamCreate(s); }
cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned int) { return cudaSuccess; }
cudaError_t cudaDeviceSynchronize(void) { return cudaSuccess; }

cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = reinterpret_cast<cudaEvent_t>(malloc(8)); return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned int) { return cudaEventCreate(e); }
cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventQuery(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0.001f; return cudaSuccess; }

cudaError_t cudaFuncSetAttribute(const void *, cudaFuncAttribute, int) { return cudaSuccess; }

cudaError_t cudaDeviceCanAccessPeer(int *can, int, int) { *can = 0; return cudaSuccess; }
cudaError_t cudaDeviceEnablePeerAccess(int, unsigned int) { return cudaErrorNotSupported; }

}  // extern "C"
