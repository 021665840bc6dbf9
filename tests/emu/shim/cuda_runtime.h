// TEST INFRASTRUCTURE ONLY - logic-check harness for the CUDA sources of libfdfd_b200 (tests/emu/README.md).
//
// This header shadows <cuda_runtime.h> when tests/emu/build_emu.py compiles the kernel sources with g++ for the
// CPU: device-language keywords become no-ops, threadIdx/blockIdx come from a per-fiber context, __syncthreads,
// warp shuffles and atomics are implemented by a cooperative fiber scheduler (emu_core.cpp), and the runtime API is
// a synchronous stand-in (streams complete immediately, "device" memory is host memory).  The result
// (build/emu/libfdfd_emu.so) exists to check the INDEXING / SYNCHRONISATION LOGIC of the kernels without a GPU; it
// is loaded only by tests/test_emu_kernels_cpu.py, never by the product path (maxwellfdm.jl_b200/_lib.py loads
// libfdfd_b200.so and nothing else), says "EMULATED" in fdfd_version(), and no benchmark may time it.
#pragma once
#define FDFD_EMU 1

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>

// ---- vector types ------------------------------------------------------------------------------------------
struct alignas(16) double2 { double x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(16) int4 { int x, y, z, w; };
struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
inline float4 make_float4(float x, float y, float z, float w) { float4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
inline int4 make_int4(int x, int y, int z, int w) { int4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }

// ---- device-language keywords ----------------------------------------------------------------------------------
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __grid_constant__
#define __align__(n) alignas(n)
#define __shared__ static thread_local

// ---- emulator core (emu_core.cpp) ------------------------------------------------------------------------------
namespace emu {
struct ThreadCtx {
    uint3 tid, bid;
    dim3 bdim, gdim;
};
extern thread_local ThreadCtx *cur;
void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()> &body);
unsigned char *dyn_smem();
void yield();                 // let the other threads of the CTA run (called inside every wait loop)
void progress();              // a wait condition was satisfied / state changed (dead-lock detection)
void syncthreads();
int syncthreads_or(int v);
void syncwarp();
uint64_t shfl_xor_bits(uint64_t v, int lane_mask);
uint64_t shfl_idx_bits(uint64_t v, int src_lane);   // src_lane outside the warp: own value
int lane_id();
bool lazy_async();            // FDFD_EMU_ASYNC=lazy: bulk copies complete at the latest legal moment
// bulk-copy engine stand-in (ptx_sm100.cuh of the shim)
void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar);
void bulk_s2g(void *dst, const void *src, uint32_t bytes);
void bulk_commit();
void bulk_wait_read0();
void bulk_wait0();
// tensor-map stand-in (3-D, 8-byte elements): image = {base, dim0..2, stride1..2 (bytes), box0..2}
void tma_load_3d(void *dst, const unsigned long long *map, int c0, int c1, int c2, uint64_t *bar);
void tma_store_3d(const unsigned long long *map, int c0, int c1, int c2, const void *src);
void mbar_init(uint64_t *bar, uint32_t count);
void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes);
void mbar_wait(uint64_t *bar, uint32_t parity);
inline dim3 to_dim3(dim3 d) { return d; }
}  // namespace emu

#define threadIdx (emu::cur->tid)
#define blockIdx (emu::cur->bid)
#define blockDim (emu::cur->bdim)
#define gridDim (emu::cur->gdim)

// ---- device intrinsics -------------------------------------------------------------------------------------
inline void __syncthreads() { emu::syncthreads(); }
inline int __syncthreads_or(int v) { return emu::syncthreads_or(v); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::syncwarp(); }
template <class T>
inline T __shfl_xor_sync(unsigned, T v, int lane_mask) {
    static_assert(sizeof(T) <= 8, "shuffle of at most 64 bits");
    uint64_t b = 0;
    std::memcpy(&b, &v, sizeof(T));
    b = emu::shfl_xor_bits(b, lane_mask);
    T r;
    std::memcpy(&r, &b, sizeof(T));
    return r;
}
template <class T>
inline T emu_shfl_idx(T v, int src) {
    static_assert(sizeof(T) <= 8, "shuffle of at most 64 bits");
    uint64_t b = 0;
    std::memcpy(&b, &v, sizeof(T));
    b = emu::shfl_idx_bits(b, src);
    T r;
    std::memcpy(&r, &b, sizeof(T));
    return r;
}
template <class T> inline T __shfl_up_sync(unsigned, T v, unsigned d) { const int l = emu::lane_id(); return emu_shfl_idx(v, l - (int)d < 0 ? l : l - (int)d); }
template <class T> inline T __shfl_down_sync(unsigned, T v, unsigned d) { const int l = emu::lane_id(); return emu_shfl_idx(v, l + (int)d > 31 ? l : l + (int)d); }
template <class T> inline T __shfl_sync(unsigned, T v, int src) { return emu_shfl_idx(v, src & 31); }
inline int __all_sync(unsigned m, int pred) {
    int v = pred ? 1 : 0;
    for (int o = 16; o > 0; o >>= 1) v &= __shfl_xor_sync(m, v, o);
    return v;
}
inline int __any_sync(unsigned m, int pred) {   // warp vote as a butterfly of exchanges
    int v = pred ? 1 : 0;
    for (int o = 16; o > 0; o >>= 1) v |= __shfl_xor_sync(m, v, o);
    return v;
}
template <class T> inline T __ldg(const T *p) { return *p; }
template <class T> inline T __ldcg(const T *p) { return *p; }
inline unsigned atomicAdd(unsigned *p, unsigned v) { unsigned o = *p; *p = o + v; return o; }
inline int atomicAdd(int *p, int v) { int o = *p; *p = o + v; return o; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline double atomicAdd(double *p, double v) { double o = *p; *p = o + v; return o; }
inline void __threadfence() {}
inline void __nanosleep(unsigned) { emu::yield(); }
[[noreturn]] inline void __trap() { std::fprintf(stderr, "emu: __trap()\n"); std::abort(); }
inline size_t __cvta_generic_to_shared(const void *p) { return (size_t)p; }
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
using std::fma;

// ---- runtime API: synchronous stand-in -------------------------------------------------------------------------
typedef int cudaError_t;
enum : int {
    cudaSuccess = 0,
    cudaErrorInvalidValue = 1,
    cudaErrorMemoryAllocation = 2,
    cudaErrorInvalidConfiguration = 9,
    cudaErrorNotSupported = 801,
};
typedef struct emuStream *cudaStream_t;
typedef struct emuEvent *cudaEvent_t;
typedef struct emuGraph *cudaGraph_t;
typedef struct emuGraphExec *cudaGraphExec_t;
struct cudaIpcMemHandle_t { char reserved[64]; };
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaStreamCaptureModeThreadLocal = 1,
       cudaIpcMemLazyEnablePeerAccess = 1, cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };

const char *cudaGetErrorString(cudaError_t e);
cudaError_t cudaGetLastError();
cudaError_t cudaMalloc(void **p, size_t bytes);        // contents start as a NaN pattern (uninitialised reads show)
cudaError_t cudaFree(void *p);
cudaError_t cudaMallocHost(void **p, size_t bytes);
cudaError_t cudaFreeHost(void *p);
cudaError_t cudaMemcpy(void *dst, const void *src, size_t n, cudaMemcpyKind k);
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t n, cudaMemcpyKind k, cudaStream_t s = nullptr);
cudaError_t cudaMemset(void *p, int v, size_t n);
cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t s = nullptr);
cudaError_t cudaGetDeviceCount(int *n);
cudaError_t cudaSetDevice(int d);
cudaError_t cudaGetDevice(int *d);
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
cudaError_t cudaDeviceGetAttribute(int *v, cudaDeviceAttr a, int dev);
cudaError_t cudaDeviceSynchronize();
cudaError_t cudaDeviceGetStreamPriorityRange(int *lo, int *hi);
cudaError_t cudaStreamCreate(cudaStream_t *s);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned flags);
cudaError_t cudaStreamCreateWithPriority(cudaStream_t *s, unsigned flags, int prio);
cudaError_t cudaStreamDestroy(cudaStream_t s);
cudaError_t cudaStreamSynchronize(cudaStream_t s);
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned flags = 0);
cudaError_t cudaStreamBeginCapture(cudaStream_t s, int mode);     // not supported: callers fall back to plain launches
cudaError_t cudaStreamEndCapture(cudaStream_t s, cudaGraph_t *g);
cudaError_t cudaGraphInstantiate(cudaGraphExec_t *e, cudaGraph_t g, unsigned long long flags);
cudaError_t cudaGraphLaunch(cudaGraphExec_t e, cudaStream_t s);
cudaError_t cudaGraphDestroy(cudaGraph_t g);
cudaError_t cudaGraphExecDestroy(cudaGraphExec_t e);
cudaError_t cudaEventCreate(cudaEvent_t *e);
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned flags);
cudaError_t cudaEventDestroy(cudaEvent_t e);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s = nullptr);
cudaError_t cudaEventSynchronize(cudaEvent_t e);
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b);
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *p);
cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned flags);
cudaError_t cudaIpcCloseMemHandle(void *p);
template <class F>
inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
