// cuda_emul.h -- TEST INFRASTRUCTURE ONLY.  A minimal CUDA-on-CPU emulation so that the product's
// real kernel and C-ABI sources (nvorbis_b200/csrc/*.cu) can be compiled with g++ and single-stepped
// in the GPU-less build container: one OS thread per CUDA thread, std::barrier for __syncthreads /
// __syncwarp, shared memory as static storage, CTAs of a launch run one after the other.
// Nothing in nvorbis_b200/ includes or loads this; the product fails loudly without a GPU.
#pragma once
#include <atomic>
#include <barrier>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <type_traits>
#include <thread>
#include <vector>

#define NVB_CPU_SHIM 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __shared__ static
#define __launch_bounds__(...)
#define __align__(n) alignas(n)
#define __restrict__

struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
static inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
static inline float4 make_float4(float x, float y, float z, float w) { float4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(4) uchar4 { unsigned char x, y, z, w; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(8) short4 { short x, y, z, w; };
struct alignas(8) uint2 { unsigned x, y; };
static inline int4 make_int4(int x, int y, int z, int w) { int4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline int __ffsll(long long v) { return __builtin_ffsll(v); }
static inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
#define NVB_HAVE_FLOAT2 1

namespace cuemu {
struct Idx { unsigned x = 0, y = 0, z = 0; };
struct Launch {
    int nthreads = 0;
    std::unique_ptr<std::barrier<>> cta;
    std::vector<std::unique_ptr<std::barrier<>>> warp;
    std::vector<uint32_t> shfl;           // 32 words per warp
    std::atomic<int> or_acc{0};
    std::mutex named_mu;
    std::vector<std::unique_ptr<std::barrier<>>> named;
};
extern Launch* g_launch;
extern unsigned char g_dyn_smem[];
}
extern thread_local cuemu::Idx threadIdx, blockIdx;
extern cuemu::Idx blockDim, gridDim;

static inline void __syncthreads() { cuemu::g_launch->cta->arrive_and_wait(); }
// bar.sync id, nthreads: the barrier object of an id is created by whoever gets there first
static inline void cuemu_named_barrier(int id, int nthreads) {
    cuemu::Launch* L = cuemu::g_launch;
    std::barrier<>* b;
    {
        std::lock_guard<std::mutex> g(L->named_mu);
        if ((int)L->named.size() <= id) L->named.resize((size_t)id + 1);
        if (!L->named[(size_t)id]) L->named[(size_t)id] = std::make_unique<std::barrier<>>(nthreads);
        b = L->named[(size_t)id].get();
    }
    b->arrive_and_wait();
}
static inline void __syncwarp(unsigned = 0xffffffffu) { cuemu::g_launch->warp[threadIdx.x >> 5]->arrive_and_wait(); }
static inline int __syncthreads_or(int pred) {
    cuemu::Launch* L = cuemu::g_launch;
    if (pred) L->or_acc.fetch_or(1);
    L->cta->arrive_and_wait();
    int r = L->or_acc.load();
    L->cta->arrive_and_wait();
    if (threadIdx.x == 0) L->or_acc.store(0);
    L->cta->arrive_and_wait();
    return r;
}
template <class T> static inline T cuemu_shfl(T v, int src_lane_delta, bool is_xor) {
    static_assert(sizeof(T) == 4, "32-bit shuffles only");
    cuemu::Launch* L = cuemu::g_launch;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t bits; std::memcpy(&bits, &v, 4);
    L->shfl[w * 32 + lane] = bits;
    L->warp[w]->arrive_and_wait();
    int src = is_xor ? (lane ^ src_lane_delta) : (lane - src_lane_delta);
    uint32_t rb = (src >= 0 && src < 32) ? L->shfl[w * 32 + src] : bits;
    L->warp[w]->arrive_and_wait();
    T r; std::memcpy(&r, &rb, 4); return r;
}
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m) { return cuemu_shfl(v, m, true); }
template <class T> static inline T __shfl_up_sync(unsigned, T v, int d) { return cuemu_shfl(v, d, false); }
static inline unsigned __ballot_sync(unsigned, int pred) {
    cuemu::Launch* L = cuemu::g_launch;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    L->shfl[w * 32 + lane] = pred ? 1u : 0u;
    L->warp[w]->arrive_and_wait();
    unsigned m = 0;
    const int lanes = (L->nthreads - 32 * w) < 32 ? (L->nthreads - 32 * w) : 32;
    for (int i = 0; i < lanes; i++) m |= L->shfl[w * 32 + i] << i;
    L->warp[w]->arrive_and_wait();
    return m;
}
template <class T> static inline T __shfl_sync(unsigned, T v, int src) {
    static_assert(sizeof(T) == 4, "32-bit shuffles only");
    cuemu::Launch* L = cuemu::g_launch;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t bits; std::memcpy(&bits, &v, 4);
    L->shfl[w * 32 + lane] = bits;
    L->warp[w]->arrive_and_wait();
    uint32_t rb = L->shfl[w * 32 + (src & 31)];
    L->warp[w]->arrive_and_wait();
    T r; std::memcpy(&r, &rb, 4); return r;
}
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
// mbarrier emulation: bits 0-15 pending arrivals, 16-31 expected arrivals, 32-63 phase
static inline void mbar_init(uint64_t* bar, int count) { __atomic_store_n(bar, ((uint64_t)count << 16) | (uint64_t)count, __ATOMIC_SEQ_CST); }
static inline void mbar_fence_init() {}
static inline void mbar_arrive(uint64_t* bar) {
    uint64_t v = __atomic_load_n(bar, __ATOMIC_SEQ_CST), nv;
    do {
        uint64_t pending = (v & 0xffff) - 1, expected = (v >> 16) & 0xffff, phase = v >> 32;
        if (pending == 0) { pending = expected; phase++; }
        nv = (phase << 32) | (expected << 16) | pending;
    } while (!__atomic_compare_exchange_n(bar, &v, nv, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST));
}
static inline void mbar_wait(uint64_t* bar, uint32_t parity) {
    for (int spins = 0; (((__atomic_load_n(bar, __ATOMIC_SEQ_CST)) >> 32) & 1u) == parity; ++spins) {
        if (spins < 64) std::this_thread::yield(); else std::this_thread::sleep_for(std::chrono::microseconds(200));
    }
}
static inline void cnt_signal(int* c) { __atomic_fetch_add(c, 1, __ATOMIC_SEQ_CST); }
static inline void cnt_wait(const int* c, int need) {
    for (int spins = 0; __atomic_load_n(c, __ATOMIC_SEQ_CST) < need; ++spins) {
        if (spins < 64) std::this_thread::yield(); else std::this_thread::sleep_for(std::chrono::microseconds(200));
    }
}
static inline bool cnt_ready(const int* c, int need) { return __atomic_load_n(c, __ATOMIC_SEQ_CST) >= need; }
// bulk copies are emulated synchronously by the issuing thread; the arrival that announces the bytes is deferred
// until the last announced byte has been copied, as the transaction count of a real mbarrier does
namespace cuemu { inline thread_local uint64_t* tx_bar = nullptr; inline thread_local uint32_t tx_left = 0; }
static inline void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) { cuemu::tx_bar = bar; cuemu::tx_left = bytes; if (bytes == 0) mbar_arrive(bar); }
static inline void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    std::memcpy(dst, src, bytes);
    if (bar != cuemu::tx_bar || bytes > cuemu::tx_left) std::abort();
    cuemu::tx_left -= bytes;
    if (cuemu::tx_left == 0) mbar_arrive(bar);
}
static inline void fence_proxy_async() {}
static inline void cp_async_16(void* dst, const void* src) { std::memcpy(dst, src, 16); }
static inline void cp_async_wait_all() {}
static inline int __clzll(long long v) { return v == 0 ? 64 : __builtin_clzll((unsigned long long)v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __float2int_rz(float f) { return (int)f; }
static inline float __int2float_rn(int v) { return (float)v; }
static inline int atomicOr(int* p, int v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }

// ---- the few runtime calls nvb_api.cu makes -------------------------------------------------------
typedef int cudaError_t;
typedef void* cudaStream_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum { cudaStreamNonBlocking = 1, cudaHostAllocDefault = 0, cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaDevAttrMultiProcessorCount = 16 };
struct cudaDeviceProp { int major = 10, minor = 0; };
static inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) { *p = cudaDeviceProp(); return cudaSuccess; }
static inline cudaError_t cudaDeviceGetAttribute(int* v, int, int) { const char* e = std::getenv("NVB_SHIM_SMS"); *v = e ? std::atoi(e) : 4; return cudaSuccess; }
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = std::aligned_alloc(256, (n + 255) & ~size_t(255)); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
static inline cudaError_t cudaHostAlloc(void** p, size_t n, int) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeHost(void* p) { std::free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { std::memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { std::memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { std::memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t) {
    for (size_t r = 0; r < h; r++) std::memcpy((char*)d + r * dp, (const char*)s + r * sp, w);
    return cudaSuccess;
}
// every pointer of the emulation is "device memory on device 0"
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
struct cudaPointerAttributes { cudaMemoryType type; int device; };
static inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void*) { a->type = cudaMemoryTypeDevice; a->device = 0; return cudaSuccess; }
typedef void* cudaEvent_t;
enum { cudaEventDisableTiming = 2 };
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, int) { *e = nullptr; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, int) { return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, int) { *s = nullptr; return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }

namespace cuemu {
void run(int grid, int block, size_t smem, const std::function<void()>& body);
template <class K, class A> void launch(K kernel, int grid, int block, size_t smem, A arg) { run(grid, block, smem, [=]() { kernel(arg); }); }
}
#define NVB_LAUNCH(kernel, grid, block, smem, stream, arg) cuemu::launch(kernel, (int)(grid), (int)(block), (size_t)(smem), arg)
namespace cuemu { template <class K, class... A> void launchv(K kernel, int grid, int block, size_t smem, A... args) { run(grid, block, smem, [=]() { kernel(args...); }); } }
#define NVB_LAUNCHV(kernel, grid, block, smem, stream, ...) cuemu::launchv(kernel, (int)(grid), (int)(block), (size_t)(smem), __VA_ARGS__)
#define NVB_DYN_SMEM(name) unsigned char* name = cuemu::g_dyn_smem
#define nvb_grid_dep_wait() ((void)0)
#define nvb_grid_dep_launch() ((void)0)
