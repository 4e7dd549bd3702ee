// TEST INFRASTRUCTURE ONLY (see tests/emu/include/cuda_runtime.h): the grid runner of the kernel emulation.
// A pool of host threads plays the CUDA threads of ONE block at a time; blocks of a grid run one after the other.
#include <cuda_runtime.h>

#include <chrono>

namespace emu {
thread_local uint3 t_threadIdx{0, 0, 0}, t_blockIdx{0, 0, 0};
thread_local BlockCtx* t_block = nullptr;
dim3 g_blockDim(1, 1, 1), g_gridDim(1, 1, 1);
unsigned char g_dyn_smem[256 * 1024] __attribute__((aligned(128)));

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
}  // namespace

void run_grid(unsigned grid, unsigned block, const std::function<void()>& body) {
    std::lock_guard<std::mutex> lk(g_launch_mu);  // one kernel at a time (streams are serialised)
    g_blockDim = dim3(block, 1, 1);
    g_gridDim = dim3(grid, 1, 1);
    for (unsigned b = 0; b < grid; ++b) pool().run_block(b, block, body);
}
}  // namespace emu

double emu_now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
