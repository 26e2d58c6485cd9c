// Shared device helpers for the ibinn_b200 kernels (sm_100a only).
#pragma once
#ifndef IBINN_EMU
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
extern unsigned long long ibinn_launch_counter;
#define IBINN_LAUNCH(kernel, grid, block, smem, stream, ...) \
    (++ibinn_launch_counter, kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__))
// Programmatic dependent launch (PDL) for the latency-bound chains of small launches: a kernel launched with IBINN_LAUNCH_PDL may be
// scheduled while its predecessor in the stream is still draining (once every CTA of the predecessor has executed
// ibinn_pdl_trigger() or exited); it sets up shared memory / barriers / TMEM and then blocks in ibinn_pdl_wait() until the
// predecessor has completed and its writes are visible.  ONLY kernels that execute ibinn_pdl_wait() in every thread before their
// first global-memory access may be launched this way.  IBINN_PDL=0 turns the launch attribute off (the device calls are no-ops then).
extern int ibinn_pdl_enabled;
__device__ __forceinline__ void ibinn_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void ibinn_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <class... KArgs, class... Args>
static inline void ibinn_launch_pdl_(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = ibinn_pdl_enabled ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#define IBINN_LAUNCH_PDL(kernel, grid, block, smem, stream, ...) \
    (++ibinn_launch_counter, ibinn_launch_pdl_(kernel, (grid), (block), (smem), (stream), __VA_ARGS__))
#define IBINN_DYN_SMEM(type, name)                                  \
    extern __shared__ __align__(16) unsigned char ibinn_dyn_raw[]; \
    type* name = reinterpret_cast<type*>(ibinn_dyn_raw)
#endif
#include "../../include/ibinn_b200.h"

#define IBINN_THREADS 256

// last failure of this library in this process, for the Python wrapper's exception text
extern char ibinn_errbuf[256];
#ifdef IBINN_EMU
#define IBINN_ERRSTR(e) "emulator"
#else
#define IBINN_ERRSTR(e) cudaGetErrorString(e)
#endif

// status of the launch just enqueued: 0 or the (positive) cudaError_t
static inline int ibinn_launch_status_(const char* file, int line) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) snprintf(ibinn_errbuf, sizeof(ibinn_errbuf), "%s:%d launch failed: %s", file, line, IBINN_ERRSTR(e));
    return (int)e;
}
#define ibinn_launch_status() ibinn_launch_status_(__FILE__, __LINE__)

template <class F>
static inline int ibinn_set_smem(F kernel, size_t bytes) {
    if (bytes > 227 * 1024) {
        snprintf(ibinn_errbuf, sizeof(ibinn_errbuf), "dynamic shared memory request of %zu bytes exceeds 227 KB", bytes);
        return IBINN_EINVAL;
    }
    if (bytes > 40 * 1024) {   // static shared memory counts against the 48 KB default too
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) {
            snprintf(ibinn_errbuf, sizeof(ibinn_errbuf), "cudaFuncSetAttribute(max dynamic smem = %zu) failed: %s", bytes, IBINN_ERRSTR(e));
            return (int)e;
        }
    }
    return 0;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum over the whole block (blockDim.x multiple of 32, <= 1024); result valid in every thread.
// `scratch` needs 33 floats.  Deterministic (fixed order).
__device__ __forceinline__ float block_sum(float v, float* scratch) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[w] = v;
    __syncthreads();
    float t = 0.f;
    for (int i = 0; i < nw; ++i) t += scratch[i];
    return t;
}

// "last block done" ticket: returns true in every thread of the block that arrives last.
// counter must be zero before the launch; the caller resets it (the last block) after use.
__device__ __forceinline__ bool last_block_ticket(unsigned* counter, unsigned nblocks, int* s_flag) {
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = atomicAdd(counter, 1u);
        *s_flag = (t == nblocks - 1u) ? 1 : 0;
    }
    __syncthreads();
    bool last = (*s_flag != 0);
    if (last) __threadfence();
    return last;
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// Two fp32 lanes in one 64-bit register pair: Blackwell's FADD2 / FFMA2 issue two IEEE fp32 operations per instruction slot
// (each component rounds exactly like the scalar instruction).  The host simulator of tests/emu uses the scalar definition.
#ifdef IBINN_EMU
struct f32x2 { float lo, hi; };
__device__ __forceinline__ f32x2 f2_make(float lo, float hi) { return f32x2{lo, hi}; }
__device__ __forceinline__ f32x2 f2_sub(f32x2 a, f32x2 b) { return f32x2{a.lo - b.lo, a.hi - b.hi}; }
__device__ __forceinline__ f32x2 f2_fma(f32x2 a, f32x2 b, f32x2 c) { return f32x2{fmaf(a.lo, b.lo, c.lo), fmaf(a.hi, b.hi, c.hi)}; }
__device__ __forceinline__ float f2_lo(f32x2 a) { return a.lo; }
__device__ __forceinline__ float f2_hi(f32x2 a) { return a.hi; }
#else
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 f2_make(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ f32x2 f2_sub(f32x2 a, f32x2 b) { f32x2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 f2_fma(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float f2_lo(f32x2 a) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a)); return lo; }
__device__ __forceinline__ float f2_hi(f32x2 a) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a)); return hi; }
#endif
