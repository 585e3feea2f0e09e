// probe_tmem.cu -- micro-benchmarks that size the round-2 kernel pipelines (development tool, not product code):
//   tcgen05.ld / tcgen05.st throughput with 4 and 16 warps, tcgen05.mma kind::tf32 issue-to-completion time for several N,
//   and the same MMA stream while the other warps keep loading / storing tensor memory.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I spwgnn_b200/csrc -o tools/probe/probe_tmem tools/probe/probe_tmem.cu
#include <cstdio>
#include <vector>
#include "spw_tc.cuh"
using namespace spw;
using namespace spw::tc;

__device__ __forceinline__ void mma_ts_n(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t acc) { mma_tf32_ts(d, a, bdesc, idesc, acc); }

// mode 0: LDTM  mode 1: STTM; every thread moves `cols` columns (multiple of 16 / 8) `reps` times
__global__ void __launch_bounds__(512, 1) k_tmem_bw(int mode, int cols, int reps, long long* out) {
  __shared__ uint32_t tptr;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(&tptr, 512);
  fence_before_sync(); __syncthreads(); fence_after_sync();
  const uint32_t base = tptr + ((uint32_t)(32 * (warp & 3)) << 16) + (warp >> 2) * cols;
  uint32_t sink = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    if (mode == 0) {
      for (int c = 0; c < cols; c += 16) { uint32_t v[16]; tmem_ld16(base + c, v); tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; ++i) sink ^= v[i]; }
    } else {
      for (int c = 0; c < cols; c += 8) { uint32_t v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = sink + i + r;
        tmem_st8(base + c, v); }
      tmem_wait_st();
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (tid == 0) { out[0] = t1 - t0; out[1] = sink; }
  fence_before_sync(); __syncthreads();
  if (warp == 0) tmem_dealloc(tptr, 512);
}

// LDTM without the per-instruction wait: all loads of a pass in flight
__global__ void __launch_bounds__(512, 1) k_tmem_ld_batched(int reps, long long* out) {
  __shared__ uint32_t tptr;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(&tptr, 512);
  fence_before_sync(); __syncthreads(); fence_after_sync();
  const uint32_t base = tptr + ((uint32_t)(32 * (warp & 3)) << 16) + (warp >> 2) * 48;
  uint32_t sink = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    uint32_t v0[16], v1[16], v2[16];
    tmem_ld16(base, v0); tmem_ld16(base + 16, v1); tmem_ld16(base + 32, v2);
    tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 16; ++i) sink ^= v0[i] ^ v1[i] ^ v2[i];
  }
  __syncthreads();
  const long long t1 = clock64();
  if (tid == 0) { out[0] = t1 - t0; out[1] = sink; }
  fence_before_sync(); __syncthreads();
  if (warp == 0) tmem_dealloc(tptr, 512);
}

// nmma MMAs (M128, N, K8, TS mode) issued by one thread; time from first issue to mbarrier completion.
// side: 0 nothing, 1 the other 15 warps keep loading tensor memory (D region of a second accumulator), 2 keep storing
__global__ void __launch_bounds__(512, 1) k_mma_time(int N, int nmma, int side, long long* out, const float4* gbuf, float4* gout) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* B = reinterpret_cast<float*>(smem);              // one k-step operand [2][N][4], reused by every MMA
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tptr;
  __shared__ volatile int done;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(&tptr, 512);
  if (tid == 32) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_mbar_init(); done = 0; }
  for (int i = tid; i < 2 * N * 4; i += 512) B[i] = 0.001f * (i % 7);
  fence_async_smem();
  fence_before_sync(); __syncthreads(); fence_after_sync();
  const uint32_t tb = tptr;
  const uint32_t lane_addr = tb + ((uint32_t)(32 * (warp & 3)) << 16);
  { uint32_t v[8]; for (int i = 0; i < 8; ++i) v[i] = __float_as_uint(1.0f + i);
    for (int c = 0; c < 256; c += 8) tmem_st8(lane_addr + c, v); tmem_wait_st(); }
  fence_before_sync(); __syncthreads(); fence_after_sync();
  long long t0 = 0, t1 = 0, tmid = 0;
  uint32_t sink = 0;
  if (tid == 0) {
    const uint32_t idesc = make_idesc_tf32(128, N);
    const uint64_t bd = make_b_desc(smem_u32(B), N * 16, 128);
    t0 = clock64();
    for (int i = 0; i < nmma / 2; ++i) mma_ts_n(tb + 256, tb + 8 * (i % 19), bd, idesc, i > 0);
    mma_commit(&bar[0]);
    for (int i = nmma / 2; i < nmma; ++i) mma_ts_n(tb + 256, tb + 8 * (i % 19), bd, idesc, 1u);
    mma_commit(&bar[1]);
    const long long ti = clock64();
    mbar_wait(&bar[0], 0);
    tmid = clock64();
    mbar_wait(&bar[1], 0);
    t1 = clock64();
    done = 1;
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = tmid - t0; out[2] = ti - t0; }
  } else if (side && warp >= 1) {
    int it = 0;
    while (!done && it < 100000) {
      if (side == 1) { uint32_t v[16]; tmem_ld16(lane_addr + 416 + 16 * (warp >> 2), v); tmem_wait_ld(); sink ^= v[0]; }
      else if (side == 2) { uint32_t v[8]; for (int i = 0; i < 8; ++i) v[i] = it + i; tmem_st8(lane_addr + 416 + 16 * (warp >> 2), v); tmem_wait_st(); }
      else if (side == 3) {            // streaming global loads (coalesced 128-bit), 8 in flight per thread
        float4 acc = make_float4(0, 0, 0, 0);
        const size_t base = ((size_t)blockIdx.x * 100000 + (size_t)it * 8) * 512 + tid;
#pragma unroll
        for (int u = 0; u < 8; ++u) { const float4 t = gbuf[(base + (size_t)u * 512) & ((1u << 24) - 1)]; acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w; }
        sink ^= __float_as_uint(acc.x + acc.y + acc.z + acc.w);
      } else if (side == 4) {          // FP32 pipe
        float a0 = it, a1 = it + 1, a2 = it + 2, a3 = it + 3;
#pragma unroll
        for (int u = 0; u < 64; ++u) { a0 = fmaf(a0, 1.0001f, 0.5f); a1 = fmaf(a1, 1.0001f, 0.5f); a2 = fmaf(a2, 1.0001f, 0.5f); a3 = fmaf(a3, 1.0001f, 0.5f); }
        sink ^= __float_as_uint(a0 + a1 + a2 + a3);
      } else if (side == 5) {          // streaming global stores
        const size_t base = ((size_t)blockIdx.x * 100000 + (size_t)it * 8) * 512 + tid;
#pragma unroll
        for (int u = 0; u < 8; ++u) gout[(base + (size_t)u * 512) & ((1u << 24) - 1)] = make_float4(it, u, 0, 0);
      } else if (side == 6) {          // integer ALU (what the tf32 split costs)
        uint32_t a0 = it, a1 = it + 1, a2 = it + 2, a3 = it + 3;
#pragma unroll
        for (int u = 0; u < 64; ++u) { a0 = (a0 + 0x1000u) & 0xffffe000u ^ a1; a1 = (a1 + 0x1000u) & 0xffffe000u ^ a2; a2 = (a2 + 0x1000u) & 0xffffe000u ^ a3; a3 = (a3 + 0x1000u) & 0xffffe000u ^ a0; }
        sink ^= a0 + a1 + a2 + a3;
      }
      ++it;
    }
    if (tid == 64 && blockIdx.x == 0) out[3] = it;
  }
  fence_before_sync(); __syncthreads();
  if (tid == 1) out[4] = sink;
  if (warp == 0) tmem_dealloc(tb, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 64);
  long long h[8];
  auto rd = [&]() { cudaDeviceSynchronize(); cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost); };
  for (int mode = 0; mode < 2; ++mode)
    for (int cols : {16, 48, 112}) {
      const int reps = 200;
      k_tmem_bw<<<1, 512>>>(mode, cols, reps, d); rd();
      const double bytes = 512.0 * cols * 4 * reps;
      printf("%s 16 warps x %3d cols: %lld cyc / %d reps = %.1f cyc per pass, %.1f B/cyc (err %s)\n", mode ? "STTM" : "LDTM", cols, h[0], reps,
             (double)h[0] / reps, bytes / h[0], cudaGetErrorString(cudaGetLastError()));
    }
  k_tmem_ld_batched<<<1, 512>>>(200, d); rd();
  printf("LDTM batched 16 warps x 48 cols: %.1f cyc per pass, %.1f B/cyc\n", h[0] / 200.0, 512.0 * 48 * 4 * 200 / h[0]);
  float4* gbuf; cudaMalloc(&gbuf, (size_t)(1u << 24) * 16); cudaMemset(gbuf, 0, (size_t)(1u << 24) * 16);
  float4* gout; cudaMalloc(&gout, (size_t)(1u << 24) * 16);
  for (int grid : {1, 148})
  for (int side = 0; side < 7; ++side)
    for (int N : {160}) {
      const int nmma = 1140;
      cudaFuncSetAttribute(k_mma_time, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
      k_mma_time<<<grid, 512, 2 * N * 4 * 4 + 128>>>(N, nmma, side, d, gbuf, gout); rd();
      printf("grid %3d ", grid);
      printf("MMA M128 N%3d x %d (side %d): total %lld cyc (%.1f / MMA), first half done at %lld, issue took %lld, side iters %lld (err %s)\n", N, nmma, side,
             h[0], (double)h[0] / nmma, h[1], h[2], h[3], cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
