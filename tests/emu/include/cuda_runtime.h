// tests/emu/include/cuda_runtime.h -- TEST INFRASTRUCTURE ONLY.
//
// A minimal CUDA-semantics shim so that the kernels and host orchestration of zkmips_b200/csrc can be
// compiled with g++ and executed on the CPU *for debugging kernel index math in a container that has no
// GPU*.  It is NOT a fallback: the product (zkmips_b200/libzkgpu.so) is built by nvcc for sm_100a only,
// the product loader never loads the emulated build, and no benchmark or parity claim uses it.
// Execution model: blocks run one after another; threads of a block run sequentially unless the launch
// is cooperative (uses __syncthreads), in which case they are OS threads with a barrier.
#pragma once
#include <algorithm>
#include <barrier>
#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#define ZK_EMU 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __constant__
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))

struct uint4 { uint32_t x, y, z, w; };
struct uint2 { uint32_t x, y; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }

namespace zkemu {
struct Dim { unsigned x = 0, y = 0, z = 0; };
extern thread_local Dim t_threadIdx, t_blockIdx, t_blockDim, t_gridDim;
extern thread_local std::barrier<>* t_barrier;
extern thread_local unsigned t_lane_base;
void launch(unsigned grid, unsigned block, const std::function<void()>& f, bool coop);
}  // namespace zkemu
#define threadIdx zkemu::t_threadIdx
#define blockIdx zkemu::t_blockIdx
#define blockDim zkemu::t_blockDim
#define gridDim zkemu::t_gridDim

static inline void __syncthreads() { if (zkemu::t_barrier) zkemu::t_barrier->arrive_and_wait(); }
static inline void __syncwarp() {}
// NOTE: warp shuffles need lock-step lanes; kernels that use them provide an emulator path (ZK_EMU)
static inline void __threadfence() {}
template <class T> static inline T __ldg(const T* p) { return *p; }
template <class T> static inline T __ldcg(const T* p) { return *p; }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline int32_t __mulhi(int32_t a, int32_t b) { return (int32_t)(((int64_t)a * b) >> 32); }
static inline uint32_t __brev(uint32_t x) {
  uint32_t r = 0;
  for (int i = 0; i < 32; i++) { r = (r << 1) | (x & 1); x >>= 1; }
  return r;
}
static inline int __clz(uint32_t x) { return x ? __builtin_clz(x) : 32; }
static inline int __popc(uint32_t x) { return __builtin_popcount(x); }
static inline uint32_t min(uint32_t a, uint32_t b) { return a < b ? a : b; }
static inline uint32_t max(uint32_t a, uint32_t b) { return a > b ? a : b; }
static inline unsigned atomicMin(unsigned* p, unsigned v) { unsigned o = *p; if (v < o) *p = v; return o; }
static inline unsigned long long atomicMin(unsigned long long* p, unsigned long long v) { auto o = *p; if (v < o) *p = v; return o; }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { unsigned o = *p; *p += v; return o; }

// ---- runtime API subset ------------------------------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2, cudaErrorPeerAccessAlreadyEnabled = 704 };
typedef void* cudaStream_t;
struct zkemu_event { std::chrono::steady_clock::time_point t; };
typedef zkemu_event* cudaEvent_t;
typedef void* cudaMemPool_t;
struct cudaDeviceProp { int major, minor, multiProcessorCount; char name[64]; };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1 };
enum { cudaMemPoolAttrReleaseThreshold = 1 };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 1 };
enum { cudaHostAllocDefault = 0 };

static inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
  p->major = 10; p->minor = 0; p->multiProcessorCount = 148; strcpy(p->name, "EMULATED"); return cudaSuccess;
}
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (void*)1; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaDeviceGetDefaultMemPool(cudaMemPool_t* p, int) { *p = nullptr; return cudaSuccess; }
static inline cudaError_t cudaMemPoolSetAttribute(cudaMemPool_t, int, void*) { return cudaSuccess; }
enum { cudaMemAllocationTypePinned = 1, cudaMemHandleTypeNone = 0, cudaMemLocationTypeDevice = 1 };
struct cudaMemLocation { int type, id; };
struct cudaMemPoolProps { int allocType, handleTypes; cudaMemLocation location; unsigned char reserved[64]; };
static inline cudaError_t cudaMemPoolCreate(cudaMemPool_t* p, const cudaMemPoolProps*) { *p = (void*)1; return cudaSuccess; }
static inline cudaError_t cudaMemPoolDestroy(cudaMemPool_t) { return cudaSuccess; }
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = calloc(1, n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
template <class T> static inline cudaError_t cudaMalloc(T** p, size_t n) { return cudaMalloc((void**)p, n); }
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMallocAsync(void** p, size_t n, cudaStream_t) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeAsync(void* p, cudaStream_t) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMallocFromPoolAsync(void** p, size_t n, cudaMemPool_t, cudaStream_t) { return cudaMalloc(p, n); }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
#define cudaMemcpyToSymbolAsync(sym, src, n, off, kind, st) (memcpy((char*)&(sym) + (off), (src), (n)), cudaSuccess)
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2 };
struct cudaPointerAttributes { cudaMemoryType type; };
static inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void*) { a->type = cudaMemoryTypeHost; return cudaSuccess; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new zkemu_event(); return cudaSuccess; }
enum { cudaEventDisableTiming = 2 };
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = new zkemu_event(); return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t wbytes, size_t rows,
                                            cudaMemcpyKind, cudaStream_t) {
  for (size_t r = 0; r < rows; r++) memcpy((char*)d + r * dp, (const char*)s + r * sp, wbytes);
  return cudaSuccess;
}
// no peer devices in the emulator: zk_ctx_set_upload_helper reports an argument error
static inline cudaError_t cudaDeviceCanAccessPeer(int* a, int, int) { *a = 0; return cudaSuccess; }
static inline cudaError_t cudaDeviceEnablePeerAccess(int, unsigned) { return cudaErrorInvalidValue; }
static inline cudaError_t cudaMemcpyPeerAsync(void* d, int, const void* s, int, size_t n, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
  *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count(); return cudaSuccess;
}
