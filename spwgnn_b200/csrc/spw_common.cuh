// spw_common.cuh -- shared device helpers for the SPWGNN B200 kernels (sm_100a).
//
// Everything here is fp32 FFMA work on 128-row activation tiles that live in shared memory;
// weight matrices are streamed from L2 in 16-row k-tiles with cp.async double buffering.
// The same source compiles for the host-side kernel-logic emulator (tools/cuemu, -DSPW_EMU),
// which is test infrastructure only.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
#ifdef SPW_EMU
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#define SPW_LAUNCH(kern, grid, block, smem, stream, ...) kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define SPW_DYN_SMEM(name) extern __shared__ __align__(128) unsigned char name[]
#endif

// ---- programmatic dependent launch ---------------------------------------------------------------------------------------
// A training step is ~135 dependent launches, most of them persistent grids of one CTA per SM that end ragged (a CTA owns
// 2 or 3 tiles of a node-level layer).  Kernels launched with SPW_LAUNCH_PDL may become resident as soon as every CTA of the
// previous kernel has called pdl_trigger() (first statement of every such kernel) and an SM has room, i.e. when the previous
// kernel's CTA on that SM has exited; they run their prologue (barrier init, tensor-memory allocation, weight operands: data
// no triggering kernel writes) and block in pdl_wait() until the previous kernel has completed and its writes are visible.
// Nothing a predecessor reads or writes is touched before pdl_wait().  Kernels that do not trigger (packers, torch's) and
// memsets are full barriers as before.
#ifdef SPW_EMU
__device__ __forceinline__ void pdl_trigger() {}
__device__ __forceinline__ void pdl_wait() {}
#define SPW_LAUNCH_PDL SPW_LAUNCH
#else
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
template <typename... KArgs, typename... Args>
inline void spw_launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  static const bool off = getenv("SPW_NO_PDL") != nullptr;      // A/B switch: plain stream order
  cfg.attrs = at; cfg.numAttrs = off ? 0 : 1;
  cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}
#define SPW_LAUNCH_PDL(kern, grid, block, smem, stream, ...) spw_launch_pdl(kern, (grid), (block), (smem), (stream), __VA_ARGS__)
#endif

namespace spw {

constexpr int kThreads = 256;   // every tile kernel: 8 warps
constexpr int kTM = 128;        // activation-tile rows (edges or nodes)
constexpr int kKT = 16;         // weight k-tile rows per pipeline stage
constexpr int kDE = 150;        // relational width (Networks.py:46,49)
constexpr int kDEP = 152;       // ... padded to a multiple of 4 floats (16-byte rows)
constexpr int kDP = 100;        // propagation / object width (Networks.py:29,47,50)
constexpr int kLdwE = 160;      // packed weight row length for 150-wide outputs (5 cols x 32 lanes)
constexpr int kLdwP = 128;      // packed weight row length for 100-wide outputs (4 cols x 32 lanes)

__device__ __forceinline__ int imin(int a, int b) { return a < b ? a : b; }
__device__ __forceinline__ int imax(int a, int b) { return a > b ? a : b; }

// ---- cp.async (LDGSTS) 16-byte global->shared copies -----------------------------------------
__device__ __forceinline__ void cp_async16(float* sdst, const float* gsrc) {
#ifdef SPW_EMU
  memcpy(sdst, gsrc, 16);
#else
  unsigned s = (unsigned)__cvta_generic_to_shared(sdst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc));
#endif
}
__device__ __forceinline__ void cp_async_commit() {
#ifndef SPW_EMU
  asm volatile("cp.async.commit_group;\n" ::);
#endif
}
template <int N>
__device__ __forceinline__ void cp_async_wait() {
#ifndef SPW_EMU
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
#endif
}

// ---- tile GEMM: acc[r][i] += sum_k Xs[(row0+r)*ldxs + k] * W[k*LDW + lane*CN + i] ------------
// Xs: activation tile in shared memory (row-major, ldxs % 4 == 0, 16-byte aligned).
// Wg: packed weight matrix in global memory, [Kp][LDW] with LDW = 32*CN, zero padded, Kp % 4 == 0.
// Wst: shared staging, 2 * kKT * LDW floats.  All kThreads threads must call (barriers inside).
// Thread mapping: warp w owns rows [w*ROWS, w*ROWS+ROWS); lane owns columns [lane*CN, lane*CN+CN).
// X reads are warp-wide broadcasts (LDS.128 along k), W reads are conflict-free (CN odd or 4).
template <int ROWS, int CN>
__device__ __forceinline__ void gemm_tile_acc(float (&acc)[ROWS][CN], const float* Xs, int ldxs, int row0,
                                              const float* __restrict__ Wg, int Kp, float* Wst) {
  constexpr int LDW = CN * 32;
  const int tid = threadIdx.x, lane = tid & 31;
  const int nkt = (Kp + kKT - 1) / kKT;
  auto load_tile = [&](int kt, int stage) {
    const int k0 = kt * kKT;
    const int n4 = imin(kKT, Kp - k0) * (LDW / 4);
    const float* src = Wg + (size_t)k0 * LDW;
    float* dst = Wst + stage * (kKT * LDW);
    for (int i = tid; i < n4; i += kThreads) cp_async16(dst + 4 * i, src + 4 * i);
    cp_async_commit();
  };
  load_tile(0, 0);
  for (int kt = 0; kt < nkt; ++kt) {
    if (kt + 1 < nkt) {
      load_tile(kt + 1, (kt + 1) & 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const float* Ws = Wst + (kt & 1) * (kKT * LDW) + lane * CN;
    const int k0 = kt * kKT;
    const int rows = imin(kKT, Kp - k0);
    const float* Xrow = Xs + (size_t)row0 * ldxs + k0;
    for (int kk = 0; kk < rows; kk += 4) {
      float xr[ROWS][4];
#pragma unroll
      for (int r = 0; r < ROWS; ++r) {
        const float4 v = *reinterpret_cast<const float4*>(Xrow + r * ldxs + kk);
        xr[r][0] = v.x; xr[r][1] = v.y; xr[r][2] = v.z; xr[r][3] = v.w;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float wv[CN];
#pragma unroll
        for (int i = 0; i < CN; ++i) wv[i] = Ws[(kk + q) * LDW + i];
#pragma unroll
        for (int r = 0; r < ROWS; ++r)
#pragma unroll
          for (int i = 0; i < CN; ++i) acc[r][i] = fmaf(xr[r][q], wv[i], acc[r][i]);
      }
    }
    __syncthreads();
  }
}

template <int ROWS, int CN>
__device__ __forceinline__ void zero_acc(float (&acc)[ROWS][CN]) {
#pragma unroll
  for (int r = 0; r < ROWS; ++r)
#pragma unroll
    for (int i = 0; i < CN; ++i) acc[r][i] = 0.f;
}

// ---- tile weight-gradient: acc[a][b] += sum_{r<nrows} Xs[r*ldx + TA*ja + a] * Ys[r*ldy + TB*jb + b]
// ja = tid >> 4, jb = tid & 15 (16 x 16 thread grid -> a (16*TA) x (16*TB) output patch).
// Callers guarantee the tiles are readable up to column 16*TA-1 / 16*TB-1 (padding is finite
// garbage that lands in output rows/cols the reduction ignores).
template <int T>
__device__ __forceinline__ void load_vec(const float* p, float (&v)[T]) {
  if (T % 2 == 0) {
#pragma unroll
    for (int i = 0; i < T / 2; ++i) {
      const float2 t = *reinterpret_cast<const float2*>(p + 2 * i);
      v[2 * i] = t.x; v[2 * i + 1] = t.y;
    }
  } else {
#pragma unroll
    for (int i = 0; i < T; ++i) v[i] = p[i];
  }
}

template <int TA, int TB>
__device__ __forceinline__ void wgrad_tile_acc(float (&acc)[TA][TB], const float* Xs, int ldx, const float* Ys,
                                               int ldy, int nrows) {
  const int ja = threadIdx.x >> 4, jb = threadIdx.x & 15;
  const float* xp = Xs + TA * ja;
  const float* yp = Ys + TB * jb;
#pragma unroll 2
  for (int r = 0; r < nrows; ++r) {
    float xa[TA], yb[TB];
    load_vec<TA>(xp + (size_t)r * ldx, xa);
    load_vec<TB>(yp + (size_t)r * ldy, yb);
#pragma unroll
    for (int a = 0; a < TA; ++a)
#pragma unroll
      for (int b = 0; b < TB; ++b) acc[a][b] = fmaf(xa[a], yb[b], acc[a][b]);
  }
}

// acc -> per-CTA partial in global memory, THREAD-MAJOR layout: element (a, b) of thread t lives at
// part[(a*TB + b)*kThreads + t], so every load/store of the read-modify-write is a fully coalesced
// 128-byte warp access (k_reduce_parts undoes the mapping).  ACC = accumulate into the partial.
template <int TA, int TB, bool ACC>
__device__ __forceinline__ void wgrad_flush(const float (&acc)[TA][TB], float* part) {
  float* base = part + threadIdx.x;
  constexpr int H = (TA + 1) / 2;
#pragma unroll
  for (int h = 0; h < TA; h += H) {
    float old[H][TB];
    if (ACC) {
#pragma unroll
      for (int a = 0; a < H; ++a)
#pragma unroll
        for (int b = 0; b < TB; ++b)
          if (h + a < TA) old[a][b] = base[(size_t)((h + a) * TB + b) * kThreads];
    }
#pragma unroll
    for (int a = 0; a < H; ++a) {
      if (h + a >= TA) continue;
#pragma unroll
      for (int b = 0; b < TB; ++b)
        base[(size_t)((h + a) * TB + b) * kThreads] = ACC ? old[a][b] + acc[h + a][b] : acc[h + a][b];
    }
  }
}

__device__ __forceinline__ float relu_f(float v) { return v > 0.f ? v : 0.f; }

// Stateless inverted-dropout mask (Networks.py:77-78, rate 0.1, train only): element `idx` of a tensor is kept
// iff a 24-bit hash of (seed, idx) is >= rate * 2^24.  The same function is restated in numpy by the tests.
__host__ __device__ __forceinline__ uint32_t dropout_hash(uint32_t seed, uint32_t idx) {
  uint32_t h = idx * 0x9E3779B9u + seed;
  h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
  return h >> 8;
}
__device__ __forceinline__ float dropout_apply(float v, uint32_t seed, uint32_t idx, uint32_t thresh, float inv_keep) {
  return dropout_hash(seed, idx) >= thresh ? v * inv_keep : 0.f;
}

}  // namespace spw
