// Fiber scheduler behind tests/emu/cuda_emu.h.  TEST INFRASTRUCTURE ONLY.
#include "cuda_emu.h"

namespace emu {

uint3_emu g_threadIdx, g_blockIdx;
dim3 g_blockDim, g_gridDim;

namespace {
constexpr size_t kStack = 256 * 1024;
enum { RUN = 0, WAIT_BLOCK = 1, WAIT_WARP = 2, DONE = 3 };
struct Fiber {
    ucontext_t ctx;
    int state = RUN;
};
std::vector<Fiber> fibers;
char* stacks = nullptr;
size_t stacks_cap = 0;
ucontext_t main_ctx;
int cur = -1;
const std::function<void()>* cur_body = nullptr;
std::vector<char> smem_buf;
std::vector<uint32_t> warp_slots;

void set_tid(int i) {
    unsigned bx = g_blockDim.x, by = g_blockDim.y;
    g_threadIdx.x = i % bx;
    g_threadIdx.y = (i / bx) % by;
    g_threadIdx.z = i / (bx * by);
}

void trampoline() {
    (*cur_body)();
    fibers[cur].state = DONE;
    swapcontext(&fibers[cur].ctx, &main_ctx);
}

void yield(int st) {
    int me = cur;
    fibers[me].state = st;
    swapcontext(&fibers[me].ctx, &main_ctx);
}

void run_block(int nthreads) {
    fibers.assign(nthreads, Fiber());
    size_t need = (size_t)nthreads * kStack;
    if (need > stacks_cap) {
        if (stacks) munmap(stacks, stacks_cap);
        stacks = (char*)mmap(nullptr, need, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (stacks == MAP_FAILED) { perror("mmap"); abort(); }
        stacks_cap = need;
    }
    for (int i = 0; i < nthreads; ++i) {
        getcontext(&fibers[i].ctx);
        fibers[i].ctx.uc_stack.ss_sp = stacks + (size_t)i * kStack;
        fibers[i].ctx.uc_stack.ss_size = kStack;
        fibers[i].ctx.uc_link = &main_ctx;
        makecontext(&fibers[i].ctx, trampoline, 0);
    }
    int live = nthreads;
    int nwarps = (nthreads + 31) / 32;
    while (live > 0) {
        bool progressed = false;
        for (int i = 0; i < nthreads; ++i) {
            if (fibers[i].state != RUN) continue;
            cur = i;
            set_tid(i);
            swapcontext(&main_ctx, &fibers[i].ctx);
            progressed = true;
            if (fibers[i].state == DONE) --live;
        }
        bool released = false;
        for (int w = 0; w < nwarps; ++w) {
            int lo = w * 32, hi = std::min(nthreads, lo + 32), waiting = 0, alive = 0;
            for (int i = lo; i < hi; ++i) {
                if (fibers[i].state != DONE) ++alive;
                if (fibers[i].state == WAIT_WARP) ++waiting;
            }
            if (alive > 0 && waiting == alive) {
                for (int i = lo; i < hi; ++i) if (fibers[i].state == WAIT_WARP) fibers[i].state = RUN;
                released = true;
            }
        }
        if (!released) {
            int waiting = 0;
            for (int i = 0; i < nthreads; ++i) if (fibers[i].state == WAIT_BLOCK) ++waiting;
            if (live > 0 && waiting == live) {
                for (int i = 0; i < nthreads; ++i) if (fibers[i].state == WAIT_BLOCK) fibers[i].state = RUN;
                released = true;
            }
        }
        if (!progressed && !released && live > 0) {
            fprintf(stderr, "cuda_emu: deadlock (divergent barrier) in block (%u,%u,%u)\n", g_blockIdx.x, g_blockIdx.y, g_blockIdx.z);
            abort();
        }
    }
}
}  // namespace

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
    g_gridDim = grid;
    g_blockDim = block;
    int nthreads = block.x * block.y * block.z;
    smem_buf.assign(smem + 64, 0x7f);  // poison: uninitialised shared memory reads show up as NaN-ish garbage
    warp_slots.assign(((nthreads + 31) / 32) * 32, 0);
    cur_body = &body;
    for (unsigned z = 0; z < grid.z; ++z)
        for (unsigned y = 0; y < grid.y; ++y)
            for (unsigned x = 0; x < grid.x; ++x) {
                g_blockIdx.x = x; g_blockIdx.y = y; g_blockIdx.z = z;
                run_block(nthreads);
            }
    cur_body = nullptr;
}

void syncthreads() { yield(WAIT_BLOCK); }
void warp_barrier() { yield(WAIT_WARP); }
void* dyn_smem() {
    uintptr_t p = (uintptr_t)smem_buf.data();
    return (void*)((p + 63) & ~(uintptr_t)63);
}
uint32_t* warp_slot(int lane) { return &warp_slots[(cur / 32) * 32 + lane]; }
int lane_id() { return cur % 32; }

}  // namespace emu
