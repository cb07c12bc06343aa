// hostsim.cpp -- TEST-ONLY: builds the whole C ABI (api.cpp + the kernel sources) against the SIMT
// emulator in hostsim.h.  See the header for what this is and is not.
#include "hostsim.h"

#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <mutex>

asm(R"(
.text
.globl hostsim_switch
.type hostsim_switch,@function
hostsim_switch:
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
.size hostsim_switch, .-hostsim_switch
)");

namespace hostsim {

static const size_t kStack = 96 * 1024;
static const int kMaxThreads = 1024;

struct ThreadState {
    Block* blk = nullptr;
    char* stacks = nullptr;
    Fiber fibers[kMaxThreads];
};
static thread_local ThreadState tls;

Block*& blk() { return tls.blk; }

unsigned char* dyn_smem()
{
    alignas(16) static thread_local unsigned char buf[256 * 1024];
    return buf;
}

void yield(int new_state)
{
    Block* b = tls.blk;
    Fiber* f = b->cur;
    f->state = new_state;
    hostsim_switch(&f->sp, b->sched_sp);
}

static void fiber_entry()
{
    Block* b = tls.blk;
    (*b->body)();
    b = tls.blk;
    b->cur->state = DONE;
    hostsim_switch(&b->cur->sp, b->sched_sp);
    abort();   // never resumed
}

static void run_fiber(Block* b, Fiber* f)
{
    b->cur = f;
    f->state = RUN;
    hostsim_switch(&b->sched_sp, f->sp);
}

static void run_block(Block* b)
{
    const int n = b->n;
    const int nwarps = (n + 31) / 32;
    for (;;) {
        bool any_live = false;
        for (int w = 0; w < nwarps; ++w) {
            const int lo = w * 32, hi = lo + 32 < n ? lo + 32 : n;
            for (;;) {
                for (int i = lo; i < hi; ++i)
                    if (b->fibers[i].state == RUN) run_fiber(b, &b->fibers[i]);
                bool waiting = false;
                for (int i = lo; i < hi; ++i) waiting |= b->fibers[i].state == WARP_WAIT;
                if (!waiting) break;
                // every lane of this warp is now parked: resolve the exchange among the waiting lanes
                for (int i = lo; i < hi; ++i) {
                    Fiber& f = b->fibers[i];
                    if (f.state != WARP_WAIT) continue;
                    const int src = lo + f.xsrc;
                    f.xout = (src < hi && b->fibers[src].state == WARP_WAIT) ? b->fibers[src].xin : f.xin;
                }
                for (int i = lo; i < hi; ++i)
                    if (b->fibers[i].state == WARP_WAIT) b->fibers[i].state = RUN;
            }
        }
        for (int i = 0; i < n; ++i)
            if (b->fibers[i].state == BLOCK_WAIT) {
                b->fibers[i].state = RUN;
                any_live = true;
            }
        if (!any_live) break;
    }
}

void launch(dim3 grid, dim3 block, const std::function<void()>& body)
{
    const int n = (int)(block.x * block.y * block.z);
    if (n > kMaxThreads) { fprintf(stderr, "hostsim: block too large\n"); abort(); }
    const long nblocks = (long)grid.x * grid.y * grid.z;
#pragma omp parallel
    {
        ThreadState& ts = tls;
        if (!ts.stacks) {
            ts.stacks = (char*)mmap(nullptr, kStack * kMaxThreads, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
            if (ts.stacks == MAP_FAILED) { perror("hostsim mmap"); abort(); }
        }
        Block b;
        b.grid = grid; b.block = block; b.fibers = ts.fibers; b.n = n; b.body = &body; b.cur = nullptr; b.sched_sp = nullptr;
        ts.blk = &b;
#pragma omp for schedule(dynamic, 1)
        for (long bi = 0; bi < nblocks; ++bi) {
            b.bid.x = (unsigned)(bi % grid.x);
            b.bid.y = (unsigned)((bi / grid.x) % grid.y);
            b.bid.z = (unsigned)(bi / ((long)grid.x * grid.y));
            for (int i = 0; i < n; ++i) {
                Fiber& f = ts.fibers[i];
                f.tid.x = i % block.x;
                f.tid.y = (i / block.x) % block.y;
                f.tid.z = i / (block.x * block.y);
                f.lane = i & 31;
                f.state = RUN;
                // initial frame: 6 callee-saved registers, then the entry address, then a pad slot so
                // that rsp is 8 mod 16 when fiber_entry starts (as after a call)
                uintptr_t top = ((uintptr_t)(ts.stacks + (size_t)(i + 1) * kStack)) & ~(uintptr_t)15;
                uint64_t* sp = (uint64_t*)top;
                *--sp = 0;                               // pad / fake return address
                *--sp = (uint64_t)(uintptr_t)&fiber_entry;
                for (int r = 0; r < 6; ++r) *--sp = 0;
                f.sp = sp;
            }
            run_block(&b);
        }
        ts.blk = nullptr;
    }
}

}  // namespace hostsim

// ---- device memory ---------------------------------------------------------------------------------------------------
// Default: aligned_alloc, like cudaMalloc's 256-byte alignment.  Guard modes (environment HOSTSIM_GUARD, read once):
//   end    the allocation (rounded up to 16 bytes, the widest vector access) ends at an inaccessible page
//   start  the allocation starts on a page that follows an inaccessible one
// so out-of-bounds accesses of the kernels fault (SIGSEGV) instead of reading a neighbour's bytes -- the emulator's
// stand-in for compute-sanitizer memcheck, used with the size fuzz.
bool hostsim_is_external_mapping(const void* p);
struct GuardedAllocation { void* region; size_t length; };
static std::mutex g_alloc_mutex;
static std::vector<std::pair<void*, GuardedAllocation>> g_guarded;
static int guard_mode()
{
    static int mode = -1;
    if (mode < 0) {
        const char* e = getenv("HOSTSIM_GUARD");
        mode = !e ? 0 : !strcmp(e, "end") ? 1 : !strcmp(e, "start") ? 2 : 0;
    }
    return mode;
}

static long g_live_allocations = 0;
// test hook (emulator library only): device allocations that have not been freed
extern "C" __attribute__((visibility("default"))) long hostsim_live_allocations() { return __atomic_load_n(&g_live_allocations, __ATOMIC_SEQ_CST); }

cudaError_t cudaMalloc(void** p, size_t n)
{
    const int mode = guard_mode();
    __atomic_fetch_add(&g_live_allocations, 1, __ATOMIC_SEQ_CST);
    // fresh device memory is not zero: poison it, so that nothing can depend on an allocation it never initialised
    if (!mode) {
        *p = aligned_alloc(256, (n + 255) / 256 * 256);
        if (*p) memset(*p, 0xCB, (n + 255) / 256 * 256);
        return *p ? cudaSuccess : cudaErrorInvalidValue;
    }
    const size_t page = 4096, bytes = (n + 15) / 16 * 16, body = (bytes + page - 1) / page * page, length = body + 2 * page;
    char* region = (char*)mmap(nullptr, length, PROT_NONE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (region == MAP_FAILED) return cudaErrorInvalidValue;
    if (mprotect(region + page, body, PROT_READ | PROT_WRITE) != 0) { munmap(region, length); return cudaErrorInvalidValue; }
    memset(region + page, 0xCB, body);
    *p = mode == 1 ? region + page + body - bytes : region + page;
    std::lock_guard<std::mutex> lock(g_alloc_mutex);
    g_guarded.push_back({*p, GuardedAllocation{region, length}});
    return cudaSuccess;
}

cudaError_t cudaFree(void* p)
{
    if (!p || hostsim_is_external_mapping(p)) return cudaSuccess;
    __atomic_fetch_sub(&g_live_allocations, 1, __ATOMIC_SEQ_CST);
    if (!guard_mode()) { free(p); return cudaSuccess; }
    std::lock_guard<std::mutex> lock(g_alloc_mutex);
    for (size_t i = 0; i < g_guarded.size(); ++i)
        if (g_guarded[i].first == p) {
            munmap(g_guarded[i].second.region, g_guarded[i].second.length);
            g_guarded.erase(g_guarded.begin() + i);
            return cudaSuccess;
        }
    return cudaErrorInvalidValue;
}

// ---- external memory / semaphores (the Vulkan side is tests/vkmock) -------------------------------------
// Memory: the fd is a memfd of the allocation's size, mapped shared.  Semaphore: a memfd of one page that starts
// with {magic, payload, timeline?}; both sides use atomics on the payload.
struct hostsim_external_memory_s { void* base; size_t size; };
struct hostsim_semaphore_page { uint64_t magic; uint64_t payload; uint32_t timeline; };
struct hostsim_external_semaphore_s { hostsim_semaphore_page* page; bool timeline; };
static const uint64_t kSemaphoreMagic = 0x564b4d4f434b5345ull;   // "VKMOCKSE"
static std::mutex g_ext_mutex;
static std::vector<hostsim_external_memory_s*> g_ext_mappings;

bool hostsim_is_external_mapping(const void* p)
{
    std::lock_guard<std::mutex> lock(g_ext_mutex);
    for (auto* m : g_ext_mappings)
        if ((const char*)p >= (const char*)m->base && (const char*)p < (const char*)m->base + m->size) return true;
    return false;
}

cudaError_t cudaImportExternalMemory(cudaExternalMemory_t* out, const cudaExternalMemoryHandleDesc* d)
{
    if (!out || !d || d->type != cudaExternalMemoryHandleTypeOpaqueFd) return cudaErrorInvalidValue;
    struct stat st;
    if (fstat(d->handle.fd, &st) != 0 || (unsigned long long)st.st_size < d->size || d->size == 0) return cudaErrorInvalidValue;
    void* base = mmap(nullptr, d->size, PROT_READ | PROT_WRITE, MAP_SHARED, d->handle.fd, 0);
    if (base == MAP_FAILED) return cudaErrorInvalidValue;
    close(d->handle.fd);                                   // ownership of the fd passes to the importer
    auto* m = new hostsim_external_memory_s{base, (size_t)d->size};
    { std::lock_guard<std::mutex> lock(g_ext_mutex); g_ext_mappings.push_back(m); }
    *out = m;
    return cudaSuccess;
}

cudaError_t cudaExternalMemoryGetMappedBuffer(void** ptr, cudaExternalMemory_t m, const cudaExternalMemoryBufferDesc* d)
{
    if (!ptr || !m || !d || d->offset + d->size > m->size) return cudaErrorInvalidValue;
    *ptr = (char*)m->base + d->offset;
    return cudaSuccess;
}

cudaError_t cudaDestroyExternalMemory(cudaExternalMemory_t m)
{
    if (!m) return cudaSuccess;
    {
        std::lock_guard<std::mutex> lock(g_ext_mutex);
        for (size_t i = 0; i < g_ext_mappings.size(); ++i)
            if (g_ext_mappings[i] == m) { g_ext_mappings.erase(g_ext_mappings.begin() + i); break; }
    }
    munmap(m->base, m->size);
    delete m;
    return cudaSuccess;
}

cudaError_t cudaImportExternalSemaphore(cudaExternalSemaphore_t* out, const cudaExternalSemaphoreHandleDesc* d)
{
    if (!out || !d) return cudaErrorInvalidValue;
    const bool timeline = d->type == cudaExternalSemaphoreHandleTypeTimelineSemaphoreFd;
    if (!timeline && d->type != cudaExternalSemaphoreHandleTypeOpaqueFd) return cudaErrorInvalidValue;
    void* base = mmap(nullptr, 4096, PROT_READ | PROT_WRITE, MAP_SHARED, d->handle.fd, 0);
    if (base == MAP_FAILED) return cudaErrorInvalidValue;
    auto* page = (hostsim_semaphore_page*)base;
    if (page->magic != kSemaphoreMagic || (page->timeline != 0) != timeline) { munmap(base, 4096); return cudaErrorInvalidValue; }
    close(d->handle.fd);
    *out = new hostsim_external_semaphore_s{page, timeline};
    return cudaSuccess;
}

cudaError_t cudaWaitExternalSemaphoresAsync(cudaExternalSemaphore_t* sems, const cudaExternalSemaphoreWaitParams* params, unsigned n, cudaStream_t)
{
    for (unsigned i = 0; i < n; ++i) {
        hostsim_external_semaphore_s* s = sems[i];
        const uint64_t want = s->timeline ? params[i].params.fence.value : 1;
        timespec t0; clock_gettime(CLOCK_MONOTONIC, &t0);
        while (__atomic_load_n(&s->page->payload, __ATOMIC_ACQUIRE) < want) {
            sched_yield();
            timespec t1; clock_gettime(CLOCK_MONOTONIC, &t1);
            if (t1.tv_sec - t0.tv_sec > 20) return cudaErrorNotSupported;     // a deadlock in a test must not hang the suite
        }
        if (!s->timeline) __atomic_store_n(&s->page->payload, 0, __ATOMIC_RELEASE);   // a binary semaphore is reset by its wait
    }
    return cudaSuccess;
}

cudaError_t cudaSignalExternalSemaphoresAsync(cudaExternalSemaphore_t* sems, const cudaExternalSemaphoreSignalParams* params, unsigned n, cudaStream_t)
{
    for (unsigned i = 0; i < n; ++i) {
        hostsim_external_semaphore_s* s = sems[i];
        if (!s->timeline) { __atomic_store_n(&s->page->payload, 1, __ATOMIC_RELEASE); continue; }
        const uint64_t v = params[i].params.fence.value;
        if (v <= __atomic_load_n(&s->page->payload, __ATOMIC_ACQUIRE)) return cudaErrorInvalidValue;   // timeline values must increase
        __atomic_store_n(&s->page->payload, v, __ATOMIC_RELEASE);
    }
    return cudaSuccess;
}

cudaError_t cudaDestroyExternalSemaphore(cudaExternalSemaphore_t s)
{
    if (!s) return cudaSuccess;
    munmap(s->page, 4096);
    delete s;
    return cudaSuccess;
}

// ---- the library under test -------------------------------------------------------------------------
#include "../../vulkanpbrt_b200/csrc/accumulate.cu"
#include "../../vulkanpbrt_b200/csrc/bmfr.cu"
#include "../../vulkanpbrt_b200/csrc/bfr.cu"
#include "../../vulkanpbrt_b200/csrc/taa.cu"
#include "../../vulkanpbrt_b200/csrc/halo.cu"
#include "../../vulkanpbrt_b200/csrc/debug.cu"
#include "../../vulkanpbrt_b200/csrc/convert.cu"
#include "../../vulkanpbrt_b200/csrc/api.cpp"
