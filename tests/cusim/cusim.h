// cusim.h -- TEST INFRASTRUCTURE ONLY.  A tiny "CUDA source on the CPU" shim.
//
// The build container has nvcc but no GPU.  To debug kernel logic (indexing, barriers, shuffles)
// before spending scarce B200 minutes, tests compile the *same* .cu sources with g++ against this
// header: every CUDA thread of a block becomes a ucontext fiber, __syncthreads/__syncwarp/shuffles
// are cooperative barriers, and the CUDA runtime calls the host code uses map onto malloc/memcpy.
// One block runs at a time, fibers switch only at synchronisation points, so a missing barrier
// shows up as a wrong answer rather than passing by luck.
//
// This is NOT a CPU fallback: the product library (libb200gate.so) is built by nvcc only and the
// Python package never loads the simulator build.  Only tests/test_cusim_*.py do.
#pragma once
#include <ucontext.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define B200_CUSIM 1

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __restrict__ __restrict
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static

struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct double2 { double x, y; };
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
static inline float2 make_float2(float a, float b) { return float2{a, b}; }
static inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }
static inline double2 make_double2(double a, double b) { return double2{a, b}; }

namespace cusim {
struct Fiber {
    ucontext_t ctx;
    char* stack = nullptr;
    bool done = false;
    uint3 tid{0, 0, 0};
    unsigned linear = 0;
};
struct Warp {
    unsigned live = 0, arrived = 0, gen = 0;
    uint64_t slot[32];
    unsigned ballot_acc = 0, ballot_res = 0;
};
struct Block {
    std::vector<Fiber> fibers;
    std::vector<Warp> warps;
    unsigned live = 0, arrived = 0, gen = 0;
    char* smem = nullptr;
};
extern Fiber* cur;
extern Block* blk;
extern uint3 g_blockIdx;
extern dim3 g_blockDim, g_gridDim;
extern ucontext_t sched_ctx;
extern long long n_launches;

void yield();
void syncthreads();
void syncwarp();
void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()>& body);

inline unsigned lane() { return cur->linear & 31u; }
inline Warp& warp() { return blk->warps[cur->linear >> 5]; }

template <class T>
inline T shfl(T v, int src) {
    static_assert(sizeof(T) <= 8, "shfl payload");
    Warp& w = warp();
    uint64_t bits = 0;
    memcpy(&bits, &v, sizeof(T));
    w.slot[lane()] = bits;
    syncwarp();
    uint64_t r = w.slot[src & 31];
    syncwarp();
    T out;
    memcpy(&out, &r, sizeof(T));
    return out;
}
unsigned ballot(int pred);
}  // namespace cusim

#define threadIdx (cusim::cur->tid)
#define blockIdx (cusim::g_blockIdx)
#define blockDim (cusim::g_blockDim)
#define gridDim (cusim::g_gridDim)

static inline void __syncthreads() { cusim::syncthreads(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { cusim::syncwarp(); }
template <class T> static inline T __shfl_sync(unsigned, T v, int src, int = 32) { return cusim::shfl(v, src); }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) { return cusim::shfl(v, (int)(cusim::lane() ^ (unsigned)m)); }
template <class T> static inline T __shfl_down_sync(unsigned, T v, unsigned d, int = 32) {
    unsigned s = cusim::lane() + d;
    return cusim::shfl(v, (int)(s < 32 ? s : cusim::lane()));
}
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned d, int = 32) {
    int s = (int)cusim::lane() - (int)d;
    return cusim::shfl(v, s >= 0 ? s : (int)cusim::lane());
}
static inline unsigned __ballot_sync(unsigned, int pred) { return cusim::ballot(pred); }
static inline int __any_sync(unsigned, int pred) { return cusim::ballot(pred) != 0; }

static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline long long __double_as_longlong(double d) { long long v; memcpy(&v, &d, 8); return v; }
static inline double __longlong_as_double(long long v) { double d; memcpy(&d, &v, 8); return d; }
static inline int __float_as_int(float f) { int u; memcpy(&u, &f, 4); return u; }
static inline float __int_as_float(int u) { float f; memcpy(&f, &u, 4); return f; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh) {
    uint64_t v = ((uint64_t)hi << 32) | lo;
    return (unsigned)(v >> (sh & 31));
}
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline double __fma_rn(double a, double b, double c) { return fma(a, b, c); }
static inline float __frcp_rn(float a) { return 1.0f / a; }
#define __expf(a) expf(a)
static inline float __fdividef(float a, float b) { return a / b; }
static inline float __int2float_rn(int a) { return (float)a; }
static inline float __uint2float_rn(unsigned a) { return (float)a; }
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline void sincospi(double x, double* s, double* c) { *s = sin(M_PI * x); *c = cos(M_PI * x); }
static inline void sincospif(float x, float* s, float* c) { *s = sinf((float)M_PI * x); *c = cosf((float)M_PI * x); }

template <class T> static inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T> static inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <class T> static inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
template <class T> static inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }

using std::max;
using std::min;

// ---- the slice of the CUDA runtime API the host code uses ---------------------------------
typedef int cudaError_t;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocDefault = 0 };
struct cudaDeviceProp { int multiProcessorCount; size_t sharedMemPerBlockOptin; int major, minor; char name[64]; size_t totalGlobalMem; };

static inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "cusim error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline cudaError_t cudaMalloc(void** p, size_t n) {
    *p = aligned_alloc(256, (n + 255) / 256 * 256 + 256);
    if (*p) memset(*p, 0xFF, n);   // poison: uninitialised reads become NaNs / huge ints
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { *p = aligned_alloc(256, (n + 255) / 256 * 256 + 256); return cudaSuccess; }
static inline cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) { return cudaMallocHost(p, n); }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t = nullptr) {
    for (size_t r = 0; r < h; ++r) memcpy((char*)d + r * dp, (const char*)s + r * sp, w);
    return cudaSuccess;
}
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return cudaSuccess; }
static inline cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = nullptr; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = nullptr; return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = nullptr; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
    memset(p, 0, sizeof(*p));
    p->multiProcessorCount = 2; p->sharedMemPerBlockOptin = 227 * 1024; p->major = 10; p->minor = 0;
    p->totalGlobalMem = (size_t)8 << 30;
    snprintf(p->name, sizeof(p->name), "cusim");
    return cudaSuccess;
}
static inline cudaError_t cudaMemGetInfo(size_t* f, size_t* t) { *f = (size_t)8 << 30; *t = (size_t)8 << 30; return cudaSuccess; }
template <class K> static inline cudaError_t cudaFuncSetAttribute(K, int, int) { return cudaSuccess; }

#define B200_LAUNCH(kernel, grid, block, smem, stream, ...) \
    cusim::launch((grid), (block), (smem), [&]() { kernel(__VA_ARGS__); })
#define B200_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(cusim::blk->smem)
