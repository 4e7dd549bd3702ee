// TEST INFRASTRUCTURE ONLY (see tests/emu/include/cuda_runtime.h): the grid runner of the kernel emulation.
// The CUDA threads of ONE block run at a time -- as fibers of the launching host thread, or with -DEMU_THREADS (the sanitizer builds) on a
// pool of host threads; blocks of a grid run one after the other.
#include <cuda_runtime.h>

#include <chrono>

namespace emu {
thread_local uint3 t_threadIdx{0, 0, 0}, t_blockIdx{0, 0, 0};
thread_local BlockCtx* t_block = nullptr;
dim3 g_blockDim(1, 1, 1), g_gridDim(1, 1, 1);
unsigned char g_dyn_smem[256 * 1024] __attribute__((aligned(128)));

#ifdef EMU_THREADS
namespace {
struct Pool {
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    std::vector<std::thread> workers;
    const std::function<void()>* body = nullptr;
    BlockCtx* block = nullptr;
    unsigned block_index = 0, n_active = 0, generation = 0, remaining = 0;
    bool stop = false;

    void worker(unsigned tid) {
        unsigned seen = 0;
        for (;;) {
            const std::function<void()>* fn;
            BlockCtx* blk;
            unsigned bidx;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_work.wait(lk, [&] { return stop || (generation != seen && tid < n_active); });
                if (stop) return;
                seen = generation;
                fn = body; blk = block; bidx = block_index;
            }
            t_threadIdx = uint3{tid, 0, 0};
            t_blockIdx = uint3{bidx, 0, 0};
            t_block = blk;
            (*fn)();
            // leaving the kernel: this lane no longer takes part in collectives or barriers
            WarpCtx& w = *blk->warps[tid >> 5];
            __atomic_fetch_and(&w.alive, ~(1u << (tid & 31u)), __ATOMIC_SEQ_CST);
            w.bar.arrive_and_drop();
            blk->bar.arrive_and_drop();
            {
                std::lock_guard<std::mutex> lk(mu);
                if (--remaining == 0) cv_done.notify_all();
            }
        }
    }
    void ensure(unsigned n) {
        while (workers.size() < n) {
            unsigned tid = (unsigned)workers.size();
            workers.emplace_back([this, tid] { worker(tid); });
        }
    }
    void run_block(unsigned bidx, unsigned n, const std::function<void()>& fn) {
        BlockCtx blk((int)n);
        {
            std::lock_guard<std::mutex> lk(mu);
            ensure(n);
            body = &fn; block = &blk; block_index = bidx; n_active = n; remaining = n;
            generation++;
        }
        cv_work.notify_all();
        std::unique_lock<std::mutex> lk(mu);
        cv_done.wait(lk, [&] { return remaining == 0; });
        n_active = 0;
    }
    ~Pool() {
        { std::lock_guard<std::mutex> lk(mu); stop = true; }
        cv_work.notify_all();
        for (auto& t : workers) t.join();
    }
};
Pool& pool() { static Pool* p = new Pool(); return *p; }  // intentionally leaked: worker threads may outlive static destruction
std::mutex g_launch_mu;
void run_block(unsigned b, unsigned block, const std::function<void()>& body) { pool().run_block(b, block, body); }
}  // namespace
#else
// ---- fibers: the CUDA threads of a block are cooperative contexts of the launching host thread --------------------------------------
// A context is a stack plus the callee-saved registers of the System V x86-64 ABI, which emu_switch pushes on the stack it leaves and
// pops from the one it enters.  A fiber runs until it has to wait at a barrier (Barrier::arrive_and_wait -> fiber_yield) or returns
// from the kernel; nothing pre-empts it, so between two barriers a thread's code runs as if alone -- which is also why a data race
// inside a block is invisible to this engine (the thread engine under ThreadSanitizer, tools/emu_sanitize.sh, is the one that sees
// them).  Compared with a host thread per CUDA thread this takes the futex wait / wake of every warp collective out of the picture:
// the CPU test suite spent four fifths of its time in the kernel for them.
extern "C" void emu_switch(void** save_sp, void* load_sp);
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
    .size emu_switch, .-emu_switch
    .section .note.GNU-stack,"",@progbits
    .text
)");

namespace {
constexpr size_t kStackBytes = 512 * 1024;  // k_texture / k_direct_step frames are ~1.5 KB on the device; -O1 host frames with inlined lobes are larger
struct Fiber {
    void* sp = nullptr;
    unsigned char* stack = nullptr;
    bool done = true;
};
std::vector<Fiber> g_fibers;  // grows to the largest block seen; stacks are reused from block to block
void* g_main_sp = nullptr;
unsigned g_cur = 0, g_n = 0, g_done = 0;
const std::function<void()>* g_body = nullptr;
BlockCtx* g_blk = nullptr;
unsigned g_block_index = 0;
std::mutex g_launch_mu;

void enter(unsigned next) {  // leave the current fiber for `next` (which is not done)
    const unsigned prev = g_cur;
    g_cur = next;
    t_threadIdx = uint3{next, 0, 0};
    emu_switch(&g_fibers[prev].sp, g_fibers[next].sp);
}
unsigned next_in_block(unsigned from) {  // the next fiber after `from`, in block order, that has not returned from the kernel
    for (unsigned k = 1; k <= g_n; ++k) {
        const unsigned c = (from + k) % g_n;
        if (!g_fibers[c].done) return c;
    }
    return from;
}
unsigned next_in_warp(unsigned from) {
    const unsigned base = from & ~31u, n = std::min(32u, g_n - base);
    for (unsigned k = 1; k <= n; ++k) {
        const unsigned c = base + ((from - base) + k) % n;
        if (!g_fibers[c].done) return c;
    }
    return from;
}
void fiber_main() {
    (*g_body)();
    // leaving the kernel: this lane no longer takes part in collectives or barriers
    const unsigned me = g_cur;
    WarpCtx& w = *g_blk->warps[me >> 5];
    w.alive &= ~(1u << (me & 31u));
    w.bar.arrive_and_drop();
    g_blk->bar.arrive_and_drop();
    g_fibers[me].done = true;
    if (++g_done == g_n) {
        void* dead;
        emu_switch(&dead, g_main_sp);
    }
    enter(next_in_block(me));
    __builtin_trap();  // a finished fiber is never entered again
}
void run_block(unsigned bidx, unsigned n, const std::function<void()>& body) {
    if (g_fibers.size() < n) g_fibers.resize(n);
    BlockCtx blk((int)n);
    g_blk = &blk; g_body = &body; g_block_index = bidx; g_n = n; g_done = 0;
    for (unsigned i = 0; i < n; ++i) {
        Fiber& f = g_fibers[i];
        if (!f.stack) {
            f.stack = static_cast<unsigned char*>(std::aligned_alloc(4096, kStackBytes));
            if (!f.stack) { std::fprintf(stderr, "emu: out of memory for fiber stacks\n"); std::abort(); }
        }
        // what emu_switch pops on first entry: six callee-saved registers, then fiber_main as the return address; one more word keeps the
        // stack pointer at 8 mod 16 on entry to fiber_main, as after a call
        uintptr_t* top = reinterpret_cast<uintptr_t*>(f.stack + kStackBytes);
        top[-1] = 0;
        top[-2] = reinterpret_cast<uintptr_t>(&fiber_main);
        for (int k = 3; k <= 8; ++k) top[-k] = 0;
        f.sp = top - 8;
        f.done = false;
    }
    t_blockIdx = uint3{bidx, 0, 0};
    t_block = &blk;
    g_cur = 0;
    t_threadIdx = uint3{0, 0, 0};
    emu_switch(&g_main_sp, g_fibers[0].sp);  // back here when the last fiber has returned from the kernel
    t_block = nullptr;
}
}  // namespace

void fiber_yield(bool block_level) {
    const unsigned next = block_level ? next_in_block(g_cur) : next_in_warp(g_cur);
    if (next != g_cur) enter(next);
}
#endif

void run_grid(unsigned grid, unsigned block, const std::function<void()>& body) {
    std::lock_guard<std::mutex> lk(g_launch_mu);  // one kernel at a time (streams are serialised)
    g_blockDim = dim3(block, 1, 1);
    g_gridDim = dim3(grid, 1, 1);
    for (unsigned b = 0; b < grid; ++b) run_block(b, block, body);
}
}  // namespace emu

double emu_now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
