// Host-side SIMT simulator for the ibinn_b200 CUDA-core kernels.  TEST INFRASTRUCTURE ONLY.
//
// The CPU-only container has nvcc but no GPU, and GPU minutes are budgeted.  To debug kernel
// *logic* (indexing, masks, reductions, barrier placement) without a device, tests/emu/build_emu.py
// compiles the very same .cu sources with g++ and this header force-included.  Every CUDA thread
// of a block becomes a ucontext fiber; __syncthreads / warp shuffles are cooperative barriers
// between the fibers; blocks run one after another.  Kernels that use inline PTX (tcgen05, TMA,
// mbarrier) are compiled out under IBINN_EMU and are only ever tested on a B200.
//
// Nothing under ibinn_b200/ loads the resulting library; only tests/ do (see tests/conftest.py).
#pragma once
#ifndef IBINN_EMU
#define IBINN_EMU 1
#endif
#include <ucontext.h>
#include <sys/mman.h>
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint3_emu { unsigned x, y, z; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) float2 { float x, y; };
struct alignas(16) int4 { int x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }

typedef struct CUstream_st* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }

namespace emu {
extern uint3_emu g_threadIdx, g_blockIdx;
extern dim3 g_blockDim, g_gridDim;
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);
void syncthreads();
void warp_barrier();
void* dyn_smem();
uint32_t* warp_slot(int lane);          // 32 exchange slots of the calling fiber's warp
int lane_id();
}  // namespace emu

#define threadIdx (emu::g_threadIdx)
#define blockIdx (emu::g_blockIdx)
#define blockDim (emu::g_blockDim)
#define gridDim (emu::g_gridDim)
#define warpSize 32

static inline void __syncthreads() { emu::syncthreads(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp_barrier(); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}

template <class T> static inline T emu_shfl_src(T v, int src) {
    static_assert(sizeof(T) == 4 || sizeof(T) == 8, "shfl");
    T r;
    if (sizeof(T) == 4) {
        uint32_t b; memcpy(&b, &v, 4);
        *emu::warp_slot(emu::lane_id()) = b;
        emu::warp_barrier();
        b = *emu::warp_slot(src & 31);
        emu::warp_barrier();
        memcpy(&r, &b, 4);
    } else {
        uint64_t b; memcpy(&b, &v, 8);
        uint32_t lo = (uint32_t)b, hi = (uint32_t)(b >> 32);
        lo = emu_shfl_src<uint32_t>(lo, src);
        hi = emu_shfl_src<uint32_t>(hi, src);
        b = ((uint64_t)hi << 32) | lo;
        memcpy(&r, &b, 8);
    }
    return r;
}
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) { return emu_shfl_src(v, emu::lane_id() ^ m); }
template <class T> static inline T __shfl_down_sync(unsigned, T v, int d, int = 32) {
    int s = emu::lane_id() + d; return emu_shfl_src(v, s > 31 ? emu::lane_id() : s);
}
template <class T> static inline T __shfl_sync(unsigned, T v, int src, int = 32) { return emu_shfl_src(v, src); }

template <class T> static inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
static inline unsigned atomicInc(unsigned* p, unsigned lim) { unsigned o = *p; *p = (o >= lim) ? 0 : o + 1; return o; }
template <class T> static inline T __ldg(const T* p) { return *p; }
template <class T> static inline T __ldcg(const T* p) { return *p; }
template <class T> static inline T __ldcs(const T* p) { return *p; }
template <class T> static inline void __stcs(T* p, T v) { *p = v; }
#define __expf(x) expf(x)
static inline float __fdividef(float a, float b) { return a / b; }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline double rsqrt(double x) { return 1.0 / sqrt(x); }
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline unsigned __float_as_uint(float f) { unsigned i; memcpy(&i, &f, 4); return i; }
static inline float __uint_as_float(unsigned i) { float f; memcpy(&f, &i, 4); return f; }
using std::max;
using std::min;

#define IBINN_LAUNCH_PDL IBINN_LAUNCH
static inline void ibinn_pdl_wait() {}
static inline void ibinn_pdl_trigger() {}
#define IBINN_LAUNCH(kernel, grid, block, smem, stream, ...) \
    emu::launch((grid), (block), (smem), [=]() { kernel(__VA_ARGS__); })
#define IBINN_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(emu::dyn_smem())
