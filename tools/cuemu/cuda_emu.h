// cuda_emu.h -- KERNEL-LOGIC EMULATOR.  DEVELOPMENT / TEST INFRASTRUCTURE ONLY.
//
// Lets the .cu sources under spwgnn_b200/csrc be compiled with g++ (-DSPW_EMU) and run on the
// host, one pthread per CUDA thread, one block at a time.  It exists because the build
// container has no GPU: indexing, barrier placement and shared-memory carve-ups are checked
// here (under AddressSanitizer) before GPU minutes are spent.  The product never uses it:
// `spwgnn_b200` loads only the nvcc-built libspwgnn.so and raises if that is missing.
#pragma once
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <functional>
#include <vector>
#include <algorithm>

struct dim3 { unsigned x, y, z; dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {} };
struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct int2 { int x, y; };
struct alignas(16) int4 { int x, y, z, w; };
struct double2 { double x, y; };
static inline float4 make_float4(float a, float b, float c, float d) { float4 v; v.x = a; v.y = b; v.z = c; v.w = d; return v; }
static inline float2 make_float2(float a, float b) { float2 v; v.x = a; v.y = b; return v; }

typedef int cudaError_t;
typedef void* cudaStream_t;
enum { cudaSuccess = 0 };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, int, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
enum { cudaMemcpyDeviceToDevice = 3, cudaMemcpyDeviceToHost = 2, cudaMemcpyHostToDevice = 1 };
template <class T> static inline cudaError_t cudaFuncSetAttribute(T, int, int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
enum { cudaDevAttrMultiProcessorCount = 16 };
static inline cudaError_t cudaDeviceGetAttribute(int* v, int, int) { *v = 4; return cudaSuccess; }   // 4 emulated "SMs"

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __shared__ static

namespace emu {
struct Block {
  pthread_barrier_t bar;
  pthread_barrier_t warp_bar[64];
  unsigned shfl_buf[64][32];
  unsigned char* dyn;
};
extern thread_local dim3 t_threadIdx, t_blockIdx;
extern dim3 g_blockDim, g_gridDim;
extern Block* g_block;
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);
}  // namespace emu
#define threadIdx (emu::t_threadIdx)
#define blockIdx (emu::t_blockIdx)
#define blockDim (emu::g_blockDim)
#define gridDim (emu::g_gridDim)

static inline void __syncthreads() { pthread_barrier_wait(&emu::g_block->bar); }
static inline void __syncwarp(unsigned = 0xffffffffu) { pthread_barrier_wait(&emu::g_block->warp_bar[threadIdx.x / 32]); }
static inline void __threadfence() { __sync_synchronize(); }

template <class T> static inline T emu_shfl(T v, int src_lane) {
  static_assert(sizeof(T) == 4, "32-bit shuffles only");
  int w = threadIdx.x / 32, l = threadIdx.x % 32;
  unsigned u; memcpy(&u, &v, 4);
  emu::g_block->shfl_buf[w][l] = u;
  pthread_barrier_wait(&emu::g_block->warp_bar[w]);
  unsigned r = emu::g_block->shfl_buf[w][src_lane & 31];
  pthread_barrier_wait(&emu::g_block->warp_bar[w]);
  T out; memcpy(&out, &r, 4); return out;
}
template <class T> static inline T __shfl_sync(unsigned, T v, int lane) { return emu_shfl(v, lane); }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m) { return emu_shfl(v, (int)(threadIdx.x % 32) ^ m); }
template <class T> static inline T __shfl_down_sync(unsigned, T v, int d) { int l = threadIdx.x % 32; return emu_shfl(v, l + d < 32 ? l + d : l); }
template <class T> static inline T __shfl_up_sync(unsigned, T v, int d) { int l = threadIdx.x % 32; return emu_shfl(v, l - d >= 0 ? l - d : l); }

static inline unsigned __ballot_sync(unsigned, int pred) {
  int w = threadIdx.x / 32, l = threadIdx.x % 32;
  emu::g_block->shfl_buf[w][l] = pred ? 1u : 0u;
  pthread_barrier_wait(&emu::g_block->warp_bar[w]);
  unsigned r = 0;
  for (int i = 0; i < 32; ++i) r |= (emu::g_block->shfl_buf[w][i] & 1u) << i;
  pthread_barrier_wait(&emu::g_block->warp_bar[w]);
  return r;
}
static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline double atomicAdd(double* p, double v) {
  static pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
  pthread_mutex_lock(&mu); double old = *p; *p = old + v; pthread_mutex_unlock(&mu); return old;
}

static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline double __dsqrt_rn(double a) { return sqrt(a); }
static inline float __ldg(const float* p) { return *p; }
static inline int __ldg(const int* p) { return *p; }
static inline float4 __ldg(const float4* p) { return *p; }
static inline float __fdividef(float a, float b) { return a / b; }

#define SPW_LAUNCH(kern, grid, block, smem, stream, ...) \
  emu::launch((grid), (block), (smem), [=]() { kern(__VA_ARGS__); })
#define SPW_DYN_SMEM(name) unsigned char* name = emu::g_block->dyn
