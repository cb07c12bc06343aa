// runner.cpp -- executes translated reference shaders on the CPU (TEST INFRASTRUCTURE, oracle/_ref).
// One fiber per invocation of a workgroup; barrier() and the subgroup operations are scheduling points.  Workgroups
// are distributed over OpenMP threads.  See glsl.hpp for what is (and is not) defined here.
#include "glsl.hpp"

#include <stdio.h>
#include <stdlib.h>
#include <sys/mman.h>

asm(R"(
.text
.globl glslshim_switch
.type glslshim_switch,@function
glslshim_switch:
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
.size glslshim_switch, .-glslshim_switch
)");
extern "C" void glslshim_switch(void** save_sp, void* load_sp);

namespace glsl {

Binding g_bindings[32];
static ShaderEntry* g_shaders = nullptr;
void register_shader(ShaderEntry* e) { e->next = g_shaders; g_shaders = e; }

enum State { RUN = 0, SUBGROUP_WAIT = 1, BARRIER_WAIT = 2, DONE = 3 };
struct Fiber {
    void* sp;
    int state;
    Invocation inv;
    uint32_t post[3];
    int nwords;
};
static const size_t kStack = 128 * 1024;
static const int kMaxInvocations = 1024;
struct ThreadState {
    char* stacks = nullptr;
    Fiber* fibers = nullptr;
    Fiber* cur = nullptr;
    void* sched_sp = nullptr;
    void (*entry)() = nullptr;
    int n = 0;
    uint32_t snap[32][32][3];
    uint32_t snap_mask[32];
};
static thread_local ThreadState ts;

Invocation& inv() { return ts.cur->inv; }

static void yield(int st)
{
    Fiber* f = ts.cur;
    f->state = st;
    glslshim_switch(&f->sp, ts.sched_sp);
}
void barrier() { yield(BARRIER_WAIT); }
void subgroup_gather(const uint32_t* mine, int nwords, uint32_t (*all)[3], uint32_t* mask)
{
    Fiber* f = ts.cur;
    for (int i = 0; i < nwords; ++i) f->post[i] = mine[i];
    f->nwords = nwords;
    const int sg = (int)f->inv.subgroup_id;
    yield(SUBGROUP_WAIT);
    memcpy(all, ts.snap[sg], sizeof(ts.snap[sg]));
    *mask = ts.snap_mask[sg];
}

static void fiber_entry()
{
    ts.entry();
    ts.cur->state = DONE;
    glslshim_switch(&ts.cur->sp, ts.sched_sp);
    abort();
}

static void run_workgroup()
{
    // Breadth-first: every runnable invocation of the WHOLE workgroup advances to its next scheduling point
    // before any subgroup operation is resolved, i.e. subgroups progress in lockstep like on a GPU.  (The
    // reference's reductions re-use reduction[0] without a barrier between consecutive calls -- SURVEY.md
    // App. C-9; letting one subgroup run ahead would make that latent race visible, which no GPU does.)
    const int n = ts.n, nsg = (n + 31) / 32;
    for (;;) {
        bool ran = false;
        for (int i = 0; i < n; ++i)
            if (ts.fibers[i].state == RUN) {
                ts.cur = &ts.fibers[i];
                glslshim_switch(&ts.sched_sp, ts.fibers[i].sp);
                ran = true;
            }
        bool resolved = false;
        for (int w = 0; w < nsg; ++w) {
            const int lo = w * 32, hi = lo + 32 < n ? lo + 32 : n;
            uint32_t mask = 0;
            for (int i = lo; i < hi; ++i)
                if (ts.fibers[i].state == SUBGROUP_WAIT) mask |= 1u << (i - lo);
            if (!mask) continue;
            for (int i = lo; i < hi; ++i)
                if ((mask >> (i - lo)) & 1u) memcpy(ts.snap[w][i - lo], ts.fibers[i].post, sizeof(ts.fibers[i].post));
            ts.snap_mask[w] = mask;
            for (int i = lo; i < hi; ++i)
                if (ts.fibers[i].state == SUBGROUP_WAIT) ts.fibers[i].state = RUN;
            resolved = true;
        }
        if (resolved) continue;
        bool any = false;
        for (int i = 0; i < n; ++i)
            if (ts.fibers[i].state == BARRIER_WAIT) { ts.fibers[i].state = RUN; any = true; }
        if (!any) break;
        (void)ran;
    }
}

static void dispatch(ShaderEntry* e, int gx, int gy)
{
    const int lx = e->local_x, ly = e->local_y, n = lx * ly;
    if (n > kMaxInvocations) { fprintf(stderr, "glsl_shim: workgroup too large\n"); abort(); }
    const long ngroups = (long)gx * gy;
#pragma omp parallel
    {
        if (!ts.stacks) {
            ts.stacks = (char*)mmap(nullptr, kStack * kMaxInvocations, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
            ts.fibers = (Fiber*)malloc(sizeof(Fiber) * kMaxInvocations);
            if (ts.stacks == MAP_FAILED || !ts.fibers) { perror("glsl_shim"); abort(); }
        }
        ts.entry = e->main;
        ts.n = n;
#pragma omp for schedule(dynamic, 1)
        for (long g = 0; g < ngroups; ++g) {
            const unsigned wx = (unsigned)(g % gx), wy = (unsigned)(g / gx);
            for (int i = 0; i < n; ++i) {
                Fiber& f = ts.fibers[i];
                const unsigned x = i % lx, y = i / lx;
                f.inv.local_id = uvec3(x, y, 0u);
                f.inv.workgroup_id = uvec3(wx, wy, 0u);
                f.inv.num_workgroups = uvec3((unsigned)gx, (unsigned)gy, 1u);
                f.inv.global_id = uvec3(wx * lx + x, wy * ly + y, 0u);
                f.inv.local_index = (uint)i;                       // gl_LocalInvocationIndex = y * size.x + x
                f.inv.subgroup_id = (uint)i / 32u;                 // linear subgroup layout, size 32 (NVIDIA)
                f.inv.subgroup_invocation = (uint)i % 32u;
                f.inv.num_subgroups = (uint)(n + 31) / 32u;
                f.state = RUN;
                uintptr_t top = ((uintptr_t)(ts.stacks + (size_t)(i + 1) * kStack)) & ~(uintptr_t)15;
                uint64_t* sp = (uint64_t*)top;
                *--sp = 0;
                *--sp = (uint64_t)(uintptr_t)&fiber_entry;
                for (int r = 0; r < 6; ++r) *--sp = 0;
                f.sp = sp;
            }
            run_workgroup();
        }
    }
}

}  // namespace glsl

extern "C" {
struct RefBinding { void* data; int width, height, layers, format; };

__attribute__((visibility("default"))) int ref_available(const char* shader, int k0, int k1, int k2)
{
    for (glsl::ShaderEntry* e = glsl::g_shaders; e; e = e->next)
        if (!strcmp(e->name, shader) && e->key0 == k0 && e->key1 == k1 && e->key2 == k2) return 1;
    return 0;
}

// runs one dispatch of a compiled reference shader.  bindings[i] is descriptor binding i (data == NULL: unbound).
__attribute__((visibility("default"))) int ref_dispatch(const char* shader, int k0, int k1, int k2, int image_width, int image_height,
                                                        int filter_radius, const void* push, int push_size, int groups_x, int groups_y,
                                                        const RefBinding* bindings, int nbindings)
{
    for (glsl::ShaderEntry* e = glsl::g_shaders; e; e = e->next) {
        if (strcmp(e->name, shader) || e->key0 != k0 || e->key1 != k1 || e->key2 != k2) continue;
        if (e->image_width) *e->image_width = image_width;
        if (e->image_height) *e->image_height = image_height;
        if (e->filter_radius) *e->filter_radius = filter_radius;
        if (push && e->push_constants) {
            if (push_size > e->push_size) return -2;
            memcpy(e->push_constants, push, (size_t)push_size);
        }
        memset(glsl::g_bindings, 0, sizeof glsl::g_bindings);
        for (int i = 0; i < nbindings && i < 32; ++i) {
            glsl::g_bindings[i].data = bindings[i].data;
            glsl::g_bindings[i].width = bindings[i].width;
            glsl::g_bindings[i].height = bindings[i].height;
            glsl::g_bindings[i].layers = bindings[i].layers;
            glsl::g_bindings[i].format = bindings[i].format;
        }
        glsl::dispatch(e, groups_x, groups_y);
        return 0;
    }
    return -1;
}
}
