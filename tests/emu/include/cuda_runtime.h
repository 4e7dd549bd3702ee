// TEST INFRASTRUCTURE ONLY -- never part of the product.
//
// A stand-in for <cuda_runtime.h> that lets g++ compile rs_pbrt_b200/csrc/*.cuh and pbrt_gpu.cu (after tests/emu/build_emu.py has
// rewritten its `kernel<<<cfg>>>(args)` launches into `emu::launcher(kernel, cfg)(args)`) and run the KERNELS' SOURCE on the CPU:
// every CUDA thread of a block is a FIBER of one host thread (a cooperative context with its own stack, switched in user space:
// emu_engine.cpp; -DEMU_THREADS, which the sanitizer builds use, makes it a host thread instead), the 32 threads of a warp meet at a
// barrier for every warp collective (__ballot_sync, __shfl_*_sync, __match_any_sync, ...), __syncthreads is a block barrier, __shared__ is process-static storage (one
// block runs at a time), atomics are host atomics, the CUDA runtime calls are malloc / memcpy / no-ops.  It exists so that the logic
// of the kernels -- queues, compaction, indexing, the dimension ledger, new kernels that have not seen a GPU yet -- can be checked
// against the oracle without GPU minutes (tests/test_emu_kernels.py, tiny scenes only: it is ~10^5 x slower than a B200).
// What it cannot show: anything about the real memory model, scheduling, performance, or nvcc's code generation.
// Nothing under rs_pbrt_b200/ loads the library built from it; the product still fails loudly without a GPU.
#pragma once
#include <algorithm>
#include <atomic>
#include <barrier>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <type_traits>
#include <vector>

#define PB_HOST_EMU 1
#define __host__
#define __device__
#define __global__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))
#define __restrict__ __restrict

using std::isinf;
using std::isnan;

// ---- vector types ----------------------------------------------------------------------------------------------------------------
struct float2 { float x, y; };
struct __attribute__((aligned(16))) float4 { float x, y, z, w; };
struct uint2 { unsigned x, y; };
struct uint3 { unsigned x, y, z; };
struct __attribute__((aligned(16))) uint4 { unsigned x, y, z, w; };
struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }

// ---- execution engine ------------------------------------------------------------------------------------------------------------
#if !defined(EMU_THREADS) && !defined(__x86_64__)
#define EMU_THREADS 1  // the context switch of the fiber engine is written for x86-64
#endif
namespace emu {
#ifdef EMU_THREADS
struct Barrier {
    std::barrier<> b;
    Barrier(int n, bool) : b(n) {}
    void arrive_and_wait() { b.arrive_and_wait(); }
    void arrive_and_drop() { b.arrive_and_drop(); }
};
#else
// Cooperative barrier of the fiber engine: a fiber that has to wait hands the host thread to the next fiber that can run -- a lane of
// its own warp for a warp barrier (the others cannot complete it), any thread of the block for the block barrier -- and looks again
// when its turn comes round.  No kernel-level waiting, no system calls.
void fiber_yield(bool block_level);
struct Barrier {
    int expected, arrived = 0;
    unsigned phase = 0;
    bool block_level;
    Barrier(int n, bool blk) : expected(n), block_level(blk) {}
    void arrive_and_wait() {
        if (++arrived == expected) { arrived = 0; ++phase; return; }
        const unsigned mine = phase;
        while (phase == mine) fiber_yield(block_level);
    }
    void arrive_and_drop() {  // (std::barrier's: one thread fewer from now on, and this phase may be complete without it)
        --expected;
        if (expected > 0 && arrived == expected) { arrived = 0; ++phase; }
    }
};
#endif
struct WarpCtx {
    Barrier bar;
    uint64_t val[32];
    unsigned alive;  // bit per lane: cleared when the lane's thread leaves the kernel
    explicit WarpCtx(int n) : bar(n, false), alive(n >= 32 ? 0xffffffffu : ((1u << n) - 1u)) { std::memset(val, 0, sizeof val); }
};
struct BlockCtx {
    Barrier bar;
    std::vector<std::unique_ptr<WarpCtx>> warps;
    explicit BlockCtx(int n) : bar(n, true) {
        for (int w = 0; w * 32 < n; ++w) warps.emplace_back(new WarpCtx(std::min(32, n - w * 32)));
    }
};
extern thread_local uint3 t_threadIdx, t_blockIdx;
extern thread_local BlockCtx* t_block;
extern dim3 g_blockDim, g_gridDim;
extern unsigned char g_dyn_smem[256 * 1024] __attribute__((aligned(128)));
void run_grid(unsigned grid, unsigned block, const std::function<void()>& body);

inline WarpCtx& warp() { return *t_block->warps[t_threadIdx.x >> 5]; }
inline unsigned lane() { return t_threadIdx.x & 31u; }
// every live lane of the warp deposits a value, all meet, every lane reads what it needs, all meet again
template <typename F>
inline auto collective(uint64_t mine, F&& read) {
    WarpCtx& w = warp();
    w.val[lane()] = mine;
    w.bar.arrive_and_wait();
    auto r = read(w);
    w.bar.arrive_and_wait();
    return r;
}

template <typename K, typename... Cfg>
struct Launcher {
    K kernel;
    unsigned grid, block;
    template <typename... Args>
    void operator()(Args... args) const {
        K k = kernel;
        run_grid(grid, block, [=]() { k(args...); });
    }
};
template <typename K>
inline Launcher<K> launcher(K k, unsigned grid, unsigned block, size_t = 0, void* = nullptr) { return Launcher<K>{k, grid, block}; }
}  // namespace emu

#define threadIdx emu::t_threadIdx
#define blockIdx emu::t_blockIdx
#define blockDim emu::g_blockDim
#define gridDim emu::g_gridDim
static const int warpSize = 32;

static inline void __syncthreads() { emu::t_block->bar.arrive_and_wait(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp().bar.arrive_and_wait(); }
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }

static inline unsigned __ballot_sync(unsigned, int pred) {
    return emu::collective(pred ? 1u : 0u, [](emu::WarpCtx& w) { unsigned m = 0; for (int l = 0; l < 32; ++l) if (((w.alive >> l) & 1u) && w.val[l]) m |= 1u << l; return m; });
}
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
static inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, !pred) == 0; }
template <typename T>
static inline T __shfl_sync(unsigned, T v, int src, int = 32) {
    uint64_t bits = 0; std::memcpy(&bits, &v, sizeof(T));
    uint64_t r = emu::collective(bits, [src](emu::WarpCtx& w) { return w.val[src & 31]; });
    T out; std::memcpy(&out, &r, sizeof(T)); return out;
}
template <typename T>
static inline T __shfl_xor_sync(unsigned, T v, int mask, int = 32) {
    uint64_t bits = 0; std::memcpy(&bits, &v, sizeof(T));
    const unsigned me = emu::lane();
    uint64_t r = emu::collective(bits, [me, mask](emu::WarpCtx& w) { return w.val[(me ^ (unsigned)mask) & 31]; });
    T out; std::memcpy(&out, &r, sizeof(T)); return out;
}
template <typename T>
static inline T __shfl_down_sync(unsigned, T v, unsigned delta, int = 32) {
    uint64_t bits = 0; std::memcpy(&bits, &v, sizeof(T));
    const unsigned me = emu::lane();
    uint64_t r = emu::collective(bits, [me, delta](emu::WarpCtx& w) { return me + delta < 32 ? w.val[me + delta] : w.val[me]; });
    T out; std::memcpy(&out, &r, sizeof(T)); return out;
}
static inline unsigned __match_any_sync(unsigned, unsigned v) {
    return emu::collective(v, [v](emu::WarpCtx& w) { unsigned m = 0; for (int l = 0; l < 32; ++l) if (((w.alive >> l) & 1u) && (unsigned)w.val[l] == v) m |= 1u << l; return m; });
}

// CUDA's global min / max
template <typename A, typename B> static inline std::common_type_t<A, B> min(A a, B b) { using T = std::common_type_t<A, B>; return (T)a < (T)b ? (T)a : (T)b; }
template <typename A, typename B> static inline std::common_type_t<A, B> max(A a, B b) { using T = std::common_type_t<A, B>; return (T)a > (T)b ? (T)a : (T)b; }

// ---- scalar intrinsics -----------------------------------------------------------------------------------------------------------
static inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline int __float_as_int(float f) { int u; std::memcpy(&u, &f, 4); return u; }
static inline float __int_as_float(int u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
static inline int __double2int_rz(double d) { return (int)d; }
static inline int __float2int_rz(float x) { if (x != x) return 0; if (x >= 2147483648.0f) return 2147483647; if (x <= -2147483648.0f) return (-2147483647 - 1); return (int)x; }
static inline float __uint2float_rn(unsigned v) { return (float)v; }
static inline float __ull2float_rn(unsigned long long v) { return (float)v; }
static inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) { return (unsigned long long)(((unsigned __int128)a * b) >> 64); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline unsigned __brev(unsigned v) { unsigned r = 0; for (int i = 0; i < 32; ++i) if (v & (1u << i)) r |= 1u << (31 - i); return r; }
static inline unsigned long long __brevll(unsigned long long v) { unsigned long long r = 0; for (int i = 0; i < 64; ++i) if (v & (1ull << i)) r |= 1ull << (63 - i); return r; }
template <typename T> static inline T __ldg(const T* p) { return *p; }
// sincos(double, double*, double*) is glibc's (g++ defines _GNU_SOURCE)

template <typename T> static inline T atomicAdd(T* p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline float atomicAdd(float* p, float v) {
    unsigned* u = reinterpret_cast<unsigned*>(p);
    unsigned old = __atomic_load_n(u, __ATOMIC_SEQ_CST);
    for (;;) {
        const float nf = __uint_as_float(old) + v;
        const unsigned nu = __float_as_uint(nf);
        if (__atomic_compare_exchange_n(u, &old, nu, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) return __uint_as_float(old);
    }
}
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
template <typename T> static inline T atomicMax(T* p, T v) { T c = __atomic_load_n(p, __ATOMIC_SEQ_CST); while (c < v && !__atomic_compare_exchange_n(p, &c, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {} return c; }
template <typename T> static inline T atomicOr(T* p, T v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
template <typename T> static inline T atomicCAS(T* p, T cmp, T val) { __atomic_compare_exchange_n(p, &cmp, val, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST); return cmp; }

// ---- runtime API (host memory is "device" memory) --------------------------------------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorPeerAccessAlreadyEnabled = 704 };
typedef struct emuStream* cudaStream_t;
typedef struct emuEvent { double t; }* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum { cudaEventDisableTiming = 2, cudaStreamNonBlocking = 1 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16, cudaDevAttrL2CacheSize = 38, cudaDevAttrComputeCapabilityMajor = 75 };
struct cudaDeviceProp { int major, minor; char name[64]; };
static inline const char* cudaGetErrorString(cudaError_t) { return "emulated CUDA error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
// PB_EMU_DEVICES=<n>: pretend to have n devices (all of them this host's memory), so that the one-process multi-device entry point can
// be exercised; PB_EMU_NO_PEER=1: and no peer access between them (the staged-copy reduce)
static inline cudaError_t cudaGetDeviceCount(int* n) { const char* v = std::getenv("PB_EMU_DEVICES"); *n = v ? std::max(1, std::atoi(v)) : 1; return cudaSuccess; }
static inline cudaError_t cudaDeviceCanAccessPeer(int* can, int, int) { const char* v = std::getenv("PB_EMU_NO_PEER"); *can = (v && std::atoi(v)) ? 0 : 1; return cudaSuccess; }
static inline cudaError_t cudaDeviceEnablePeerAccess(int, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) { p->major = 10; p->minor = 0; std::strcpy(p->name, "host emulation"); return cudaSuccess; }
static inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr a, int) { *v = a == cudaDevAttrL2CacheSize ? (64 << 10) : (a == cudaDevAttrComputeCapabilityMajor ? 10 : 1); return cudaSuccess; }  // one "SM": grids stay small
// Device memory comes back uninitialised, as on the GPU; PB_EMU_POISON=<byte> fills it with that byte instead (0xff: NaNs / huge indices),
// so that a kernel that reads what no kernel wrote shows up as a parity failure or a crash rather than passing by luck.
static inline cudaError_t cudaMalloc(void** p, size_t n) {
    const size_t bytes = (n + 255) / 256 * 256;
    *p = std::aligned_alloc(256, bytes);
    if (!*p) return cudaErrorInvalidValue;
    static const char* poison = std::getenv("PB_EMU_POISON");
    if (poison) std::memset(*p, (int)std::strtol(poison, nullptr, 0), bytes);
    return cudaSuccess;
}
static inline cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
// stream-ordered allocation from the device's default pool: plain malloc / free here
typedef struct emuMemPool* cudaMemPool_t;
enum cudaMemPoolAttr { cudaMemPoolAttrReleaseThreshold = 4 };
static inline cudaError_t cudaDeviceGetDefaultMemPool(cudaMemPool_t* p, int) { *p = nullptr; return cudaSuccess; }
static inline cudaError_t cudaMemPoolSetAttribute(cudaMemPool_t, cudaMemPoolAttr, void*) { return cudaSuccess; }
static inline cudaError_t cudaMallocAsync(void** p, size_t n, cudaStream_t) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeAsync(void* p, cudaStream_t) { return cudaFree(p); }
enum cudaLimit { cudaLimitStackSize = 0 };
static inline cudaError_t cudaDeviceGetLimit(size_t* v, cudaLimit) { *v = 1024; return cudaSuccess; }
static inline cudaError_t cudaDeviceSetLimit(cudaLimit, size_t) { return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { std::memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyPeerAsync(void* d, int, const void* s, int, size_t n, cudaStream_t = nullptr) { std::memcpy(d, s, n); return cudaSuccess; }
enum { cudaHostAllocDefault = 0, cudaHostRegisterPortable = 1 };
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2 };
struct cudaPointerAttributes { cudaMemoryType type; };
// PB_EMU_PINNED=1: every host pointer counts as pinned (the direct-DMA branch of the uploader); default: none does (the staged branch)
static inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void*) { const char* v = std::getenv("PB_EMU_PINNED"); a->type = (v && std::atoi(v)) ? cudaMemoryTypeHost : cudaMemoryTypeUnregistered; return cudaSuccess; }
static inline cudaError_t cudaHostRegister(void*, size_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaHostUnregister(void*) { return cudaSuccess; }
static inline cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) { *p = std::malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorInvalidValue; }
static inline cudaError_t cudaFreeHost(void* p) { std::free(p); return cudaSuccess; }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { std::memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { std::memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = reinterpret_cast<cudaStream_t>(new int(0)); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
double emu_now_ms();
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new emuEvent{0.0}; return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) { e->t = emu_now_ms(); return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(b->t - a->t); return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
template <typename K>
static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, K, int, size_t) { *n = 2; return cudaSuccess; }
