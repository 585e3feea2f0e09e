// spw_pipe_tc.cuh -- software-pipelined tensor-core kernels of the edge step (tcgen05, 3xTF32), sm_100a only.
//
// Round-1 kernels ran stage -> MMA -> epilogue strictly in sequence per 128-row tile (tensor pipe 15-38 % busy).  Tensor
// memory cannot hold two tiles (A_hi 152 + A_lo 152 + D 160 of 512 columns), but it can be handed over PIECE BY PIECE:
//   * the 57 MMAs of a tile are issued as 38 correction products (A_lo.B_hi, A_hi.B_lo; commit -> barC) followed by the
//     19 main products (A_hi.B_hi; commit -> barM) -- the order the accumulation needs anyway (spw_tc.cuh);
//   * after barC the A_lo columns are dead: the next tile's lo words are stored while the main products still run;
//   * after barM the A_hi columns are dead and D is complete: the next tile's hi words go in, D is pulled into registers,
//     and the next tile's MMAs start; the epilogue of the finished tile then runs from registers UNDER those MMAs,
//     followed by the operand build (gather, combine, transposition) of the tile after.
// tcgen05.mma issue blocks the issuing thread for the length of the MMA stream (shallow queue), so a 17th warp does
// nothing but wait for "operands ready" (named barrier) and issue; the 16 worker warps synchronise among themselves
// with a second named barrier.  Per tile the tensor pipe idles only while hi is stored and D is read (~0.6k of ~5.3k
// cycles); everything else is hidden as long as the workers' per-tile work fits under the MMAs.
#pragma once
#ifndef SPW_EMU
#include "spw_tc.cuh"

namespace spw {
namespace tc {

constexpr int kPipeWorkers = 512;                  // 16 worker warps: thread = (row = TMEM lane, column quarter of a 64-column slab)
constexpr int kPipeThreads = kPipeWorkers + 32;    // + the MMA issuer warp

__device__ __forceinline__ void nbar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void nbar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }
constexpr int kBarOperands = 1;                    // workers arrive, issuer syncs: A of the next tile stored, D of the last one read
constexpr int kBarWorkers = 2;                     // the 512 workers among themselves

// issuer side of one tile: corrections -> barC, main products -> barM
__device__ __forceinline__ void pipe_issue_tile(uint32_t tmem_base, uint32_t bhi, uint32_t blo, uint64_t* barC, uint64_t* barM) {
  const uint32_t idesc = make_idesc_tf32(128, kN);
#pragma unroll 1
  for (int ks = 0; ks < kKS; ++ks) {
    const uint64_t dhi = make_b_desc(bhi + ks * (kBStepFloats * 4), kN * 16, 128);
    const uint64_t dlo = make_b_desc(blo + ks * (kBStepFloats * 4), kN * 16, 128);
    mma_tf32_ts(tmem_base + kColD, tmem_base + kColAlo + 8 * ks, dhi, idesc, ks > 0 ? 1u : 0u);
    mma_tf32_ts(tmem_base + kColD, tmem_base + kColAhi + 8 * ks, dlo, idesc, 1u);
  }
  mma_commit(barC);
#pragma unroll 1
  for (int ks = 0; ks < kKS; ++ks) {
    const uint64_t dhi = make_b_desc(bhi + ks * (kBStepFloats * 4), kN * 16, 128);
    mma_tf32_ts(tmem_base + kColD, tmem_base + kColAhi + 8 * ks, dhi, idesc, 1u);
  }
  mma_commit(barM);
}

// this thread's share of a tile's operand: 3 slabs x 16 columns of its row (slab 2: 16 / 8 / 0 / 0 columns for q = 0..3)
struct XRegs { float v[3][16]; };

__device__ __forceinline__ void pipe_store_lo(const XRegs& x, uint32_t lane_addr, int q) {
#pragma unroll
  for (int sl = 0; sl < 3; ++sl) {
    const int ncols = sl < 2 ? kStageCols : kDEP - 2 * kStageCols;
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      const int c = 16 * q + 8 * g;
      if (c < ncols) {                                   // warp-uniform
        uint32_t l[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { uint32_t h; split_tf32(x.v[sl][8 * g + i], h, l[i]); }
        tmem_st8(lane_addr + kColAlo + kStageCols * sl + c, l);
      }
    }
  }
}
__device__ __forceinline__ void pipe_store_hi(const XRegs& x, uint32_t lane_addr, int q) {
#pragma unroll
  for (int sl = 0; sl < 3; ++sl) {
    const int ncols = sl < 2 ? kStageCols : kDEP - 2 * kStageCols;
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      const int c = 16 * q + 8 * g;
      if (c < ncols) {
        uint32_t h[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h[i]) : "f"(x.v[sl][8 * g + i]));
        tmem_st8(lane_addr + kColAhi + kStageCols * sl + c, h);
      }
    }
  }
}
// D (160 columns) -> registers: this thread's 16 columns of each 64-column slab (slab 2 holds columns 128..159: q = 0, 1)
__device__ __forceinline__ void pipe_load_d(uint32_t (&d)[3][16], uint32_t lane_addr, int q) {
  tmem_ld16(lane_addr + kColD + 16 * q, d[0]);
  tmem_ld16(lane_addr + kColD + kStageCols + 16 * q, d[1]);
  if (q < 2) tmem_ld16(lane_addr + kColD + 2 * kStageCols + 16 * q, d[2]);
  tmem_wait_ld();
}

// =================================================================================================
// k_edge_step_p: the forward edge step (Networks.py:84-88 after the factorisation of DESIGN.md section 2), pipelined.
//   h2_e = relu(W2 . relu(A_e + S_s + R_r) + b2);  H2S_i = sum_{e -> i} h2_e (slot order);  sign bits of h1 / h2 for training.
// Same arguments, outputs and summation order as k_edge_step_tc (spw_tc.cuh): the results are bit-identical.
// =================================================================================================
constexpr size_t kEdgeStepPSmem = (size_t)(2 * kBFloats + kTM * kStagePitch + 2 * kTM + 2 * ((kTM + 8) / 2)) * sizeof(float) + 64;
static_assert(kEdgeStepPSmem <= 232448, "k_edge_step_p shared memory exceeds the 227 KB per-CTA limit");

struct StepTile { int e0, rows, n_first, nnodes; };

__global__ void __launch_bounds__(kPipeThreads, 1) k_edge_step_p(EdgeStepTcArgs a) {
  SPW_DYN_SMEM(smem_raw);
  float* Bhi_s = reinterpret_cast<float*>(smem_raw);
  float* Blo_s = Bhi_s + kBFloats;
  float* stage = Blo_s + kBFloats;
  int* srcv = reinterpret_cast<int*>(stage + kTM * kStagePitch);
  int* ssnd = srcv + kTM;
  short* snoff = reinterpret_cast<short*>(ssnd + kTM);       // [2][kTM + 8]: in_off - e0 of the tile's nodes (build and epilogue of different tiles overlap)
  uint64_t* bars = reinterpret_cast<uint64_t*>(snoff + 2 * (kTM + 8));
  uint64_t* barC = bars; uint64_t* barM = bars + 1;
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bars + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = 32 * (warp & 3) + lane, q = warp >> 2;     // workers: TMEM lane, column quarter of a slab

  if (warp == 0) tmem_alloc(tptr, kTmemCols);
  if (tid == 32) { mbar_init(barC, 1); mbar_init(barM, 1); fence_mbar_init(); }
  for (int i = tid; i < kBFloats / 4; i += kPipeThreads) {
    reinterpret_cast<float4*>(Bhi_s)[i] = reinterpret_cast<const float4*>(a.W2hi)[i];
    reinterpret_cast<float4*>(Blo_s)[i] = reinterpret_cast<const float4*>(a.W2lo)[i];
  }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tptr;
  const int ntiles = (a.E + kTM - 1) / kTM;
  const int cnt = (int)blockIdx.x < ntiles ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (warp == kPipeWorkers / 32) {
    // ---------------- MMA issuer warp ----------------
    const uint32_t bhi = smem_u32(Bhi_s), blo = smem_u32(Blo_s);
    for (int i = 0; i < cnt; ++i) {
      nbar_sync(kBarOperands, kPipeThreads);
      fence_after_sync();
      if (lane == 0) pipe_issue_tile(tmem_base, bhi, blo, barC, barM);
      __syncwarp();
    }
  } else {
    // ---------------- worker warps ----------------
    const uint32_t lane_addr = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16);
    bool failed = false;
    StepTile cur, nxt;
    XRegs x;
    SPW_PH_DECL

    auto load_tile_idx = [&](int i, StepTile& t) {             // indices of local tile i -> srcv / ssnd / snoff[i & 1]
      const int tile = blockIdx.x + i * gridDim.x;
      t.e0 = tile * kTM;
      t.rows = imin(kTM, a.E - t.e0);
      if (tid < kTM) {
        const bool valid = tid < t.rows;
        srcv[tid] = valid ? a.in_rcv[t.e0 + tid] : -1;
        ssnd[tid] = valid ? a.in_snd[t.e0 + tid] : 0;
      }
      nbar_sync(kBarWorkers, kPipeWorkers);
      t.n_first = srcv[0];
      t.nnodes = srcv[t.rows - 1] - t.n_first + 1;
      short* so = snoff + (i & 1) * (kTM + 8);
      for (int j = tid; j <= imin(t.nnodes, kTM + 7); j += kPipeWorkers)
        so[j] = (short)imax(-32000, imin(32000, a.in_off[t.n_first + j] - t.e0));
    };
    // operand of a tile: gather + combine into the slab (coalesced), row threads pick up their 16 columns; h1 sign bits
    auto build_x = [&](const StepTile& t) {
      H1Regs hreg;
      h1_slab_load(hreg, ssnd, srcv, a.A, a.S, a.R, t.e0, 0, kStageCols);
#pragma unroll
      for (int sl = 0; sl < 3; ++sl) {
        const int c0 = sl * kStageCols;
        const int ncols = imin(kStageCols, kDEP - c0);
        h1_slab_store(hreg, stage, srcv, c0, ncols);
        nbar_sync(kBarWorkers, kPipeWorkers);
        if (sl < 2) h1_slab_load(hreg, ssnd, srcv, a.A, a.S, a.R, t.e0, c0 + kStageCols, imin(kStageCols, kDEP - c0 - kStageCols));
        const int cb = 16 * q;
        uint32_t hbits = 0u;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const int c = cb + 8 * g;
          if (c < ncols) {                                    // warp-uniform
            const float2* src = reinterpret_cast<const float2*>(stage + row * kStagePitch + c);
            const float2 p0 = src[0], p1 = src[1], p2 = src[2], p3 = src[3];
            const float xx[8] = {p0.x, p0.y, p1.x, p1.y, p2.x, p2.y, p3.x, p3.y};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              x.v[sl][8 * g + i] = xx[i];
              hbits |= xx[i] > 0.f ? (1u << (8 * g + i)) : 0u;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) x.v[sl][8 * g + i] = 0.f;
          }
        }
        if (a.maskbits_h1 && row < t.rows && cb < ncols)
          reinterpret_cast<uint16_t*>(a.maskbits_h1)[(size_t)(t.e0 + row) * 16 + ((c0 + cb) >> 4)] = (uint16_t)hbits;
        nbar_sync(kBarWorkers, kPipeWorkers);
      }
    };

    if (cnt > 0) {
      load_tile_idx(0, cur);
      build_x(cur);
      pipe_store_lo(x, lane_addr, q);
      pipe_store_hi(x, lane_addr, q);
      tmem_wait_st();
      fence_before_sync();
      nbar_arrive(kBarOperands, kPipeThreads);
    }
    for (int i = 0; i < cnt; ++i) {
      const bool has_next = i + 1 < cnt;
      const uint32_t parity = (uint32_t)i & 1u;
      SPW_PH(7);
      if (has_next) {                                          // under the MMAs of tile i
        load_tile_idx(i + 1, nxt);
        SPW_PH(0);                                             // p0: indices
        build_x(nxt);
        SPW_PH(1);                                             // p1: operand build
      }
      if (!mbar_wait(barC, parity)) failed = true;             // corrections of tile i done: the A_lo columns are free
      fence_after_sync();
      SPW_PH(2);                                               // p2: wait for the correction MMAs
      if (has_next) pipe_store_lo(x, lane_addr, q);
      if (!mbar_wait(barM, parity)) failed = true;             // tile i done: A_hi free, D complete
      fence_after_sync();
      SPW_PH(3);                                               // p3: lo store + wait for the main MMAs
      if (has_next) pipe_store_hi(x, lane_addr, q);
      uint32_t d[3][16];
      pipe_load_d(d, lane_addr, q);
      if (has_next) {
        tmem_wait_st();
        fence_before_sync();
        nbar_arrive(kBarOperands, kPipeThreads);               // the issuer starts tile i + 1
      }
      SPW_PH(4);                                               // p4: hi store + D load (tensor pipe idle)
      // ---- epilogue of tile i from registers (under the MMAs of tile i + 1): relu, sign bits, receiver-segmented sum
      const int tile = blockIdx.x + i * gridDim.x;
      const short* so = snoff + (i & 1) * (kTM + 8);
      const int e0 = cur.e0, rows = cur.rows, n_first = cur.n_first, nnodes = cur.nnodes;
#pragma unroll
      for (int sl = 0; sl < 3; ++sl) {
        const int c0 = sl * kStageCols;
        const int ncols = imin(kStageCols, kN - c0);          // 64, 64, 32
        if (16 * q < ncols) {                                  // warp-uniform
          uint32_t m16 = 0u;
          float o[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            const float pre = __uint_as_float(d[sl][k]);       // bias already inside (ones column x bias row)
            const bool on = (c0 + 16 * q + k < kDE) && (pre > 0.f);
            o[k] = on ? pre : 0.f;
            m16 |= on ? (1u << k) : 0u;
          }
          float2* dst = reinterpret_cast<float2*>(stage + row * kStagePitch + 16 * q);
#pragma unroll
          for (int k = 0; k < 8; ++k) dst[k] = make_float2(o[2 * k], o[2 * k + 1]);
          if (a.maskbits && row < rows)
            reinterpret_cast<uint16_t*>(a.maskbits)[(size_t)(e0 + row) * 16 + ((c0 + 16 * q) >> 4)] = (uint16_t)m16;
        }
        nbar_sync(kBarWorkers, kPipeWorkers);
        for (int item = tid; item < nnodes * ncols; item += kPipeWorkers) {
          const int ni = item / ncols, c = item - ni * ncols;
          const int node = n_first + ni;
          int s0, s1;
          if (ni < kTM + 7) { s0 = e0 + so[ni]; s1 = e0 + so[ni + 1]; } else { s0 = a.in_off[node]; s1 = a.in_off[node + 1]; }
          const int lo = imax(s0, e0) - e0, hi = imin(s1, e0 + rows) - e0;
          if (hi <= lo) continue;
          const int col = c0 + c;
          if (col >= kDEP) continue;
          float sum = 0.f;
          if (col < kDE)
            for (int r = lo; r < hi; ++r) sum += stage[r * kStagePitch + c];
          float* dst;
          if (s0 >= e0 && s1 <= e0 + rows) dst = a.H2S + (size_t)node * kDEP;
          else if (s0 < e0) dst = a.part_first + (size_t)tile * kDEP;
          else dst = a.part_last + (size_t)tile * kDEP;
          dst[col] = sum;
        }
        nbar_sync(kBarWorkers, kPipeWorkers);
      }
      SPW_PH(5);                                               // p5: epilogue
      cur = nxt;
    }
    SPW_PH_REPORT("k_edge_step_p");
    if (failed && tid == 0) a.H2S[0] = __int_as_float(0x7fc00000);   // fail loudly: poison the output (an MMA barrier timed out)
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, kTmemCols);
}

// =================================================================================================
// k_edge_dgrad_p: data gradient of the edge step (see k_edge_dgrad_tc), pipelined the same way.
//   A = relu-bits(h2) ? dH2S[receiver] : 0;  B = W2^T;  epilogue: mask with the relu bits of h1, write DH1, write /
//   accumulate dA.  Same results as k_edge_dgrad_tc with act == null and scale == 1 (bit-identical).
// =================================================================================================
constexpr size_t kEdgeDgradPSmem = (size_t)(2 * kBFloats + kTM * kStagePitch + kTM) * sizeof(float) + 64;

__global__ void __launch_bounds__(kPipeThreads, 1) k_edge_dgrad_p(EdgeDgradTcArgs a) {
  SPW_DYN_SMEM(smem_raw);
  float* Bhi_s = reinterpret_cast<float*>(smem_raw);
  float* Blo_s = Bhi_s + kBFloats;
  float* stage = Blo_s + kBFloats;
  int* srcv = reinterpret_cast<int*>(stage + kTM * kStagePitch);
  uint64_t* bars = reinterpret_cast<uint64_t*>(srcv + kTM);
  uint64_t* barC = bars; uint64_t* barM = bars + 1;
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bars + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = 32 * (warp & 3) + lane, q = warp >> 2;

  if (warp == 0) tmem_alloc(tptr, kTmemCols);
  if (tid == 32) { mbar_init(barC, 1); mbar_init(barM, 1); fence_mbar_init(); }
  for (int i = tid; i < kBFloats / 4; i += kPipeThreads) {
    reinterpret_cast<float4*>(Bhi_s)[i] = reinterpret_cast<const float4*>(a.Whi)[i];
    reinterpret_cast<float4*>(Blo_s)[i] = reinterpret_cast<const float4*>(a.Wlo)[i];
  }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tptr;
  const int ntiles = (a.E + kTM - 1) / kTM;
  const int cnt = (int)blockIdx.x < ntiles ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (warp == kPipeWorkers / 32) {
    const uint32_t bhi = smem_u32(Bhi_s), blo = smem_u32(Blo_s);
    for (int i = 0; i < cnt; ++i) {
      nbar_sync(kBarOperands, kPipeThreads);
      fence_after_sync();
      if (lane == 0) pipe_issue_tile(tmem_base, bhi, blo, barC, barM);
      __syncwarp();
    }
  } else {
    const uint32_t lane_addr = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16);
    const int sub = lane >> 4, c4 = lane & 15;
    bool failed = false;
    XRegs x;

    auto gather_slab = [&](float4 (&v)[4], int c0) {     // dH2S[receiver] rows of one 64-column slab: half a warp per row
      const int ncols = imin(kStageCols, kDEP - c0);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int rc = srcv[warp * 8 + 2 * j + sub];
        v[j] = (rc >= 0 && 4 * c4 < ncols) ? *reinterpret_cast<const float4*>(a.dH2S + (size_t)rc * kDEP + c0 + 4 * c4)
                                           : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    // operand of local tile i: gather, slab, row threads pick up their 16 columns masked with the relu bits of h2
    auto build_x = [&](int i) {
      const int tile = blockIdx.x + i * gridDim.x;
      const int e0 = tile * kTM, rows = imin(kTM, a.E - e0);
      if (tid < kTM) srcv[tid] = tid < rows ? (a.in_rcv ? a.in_rcv[e0 + tid] : e0 + tid) : -1;
      uint32_t b2w[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) b2w[k] = 0xffffffffu;
      if (row < rows && a.maskbits) {
#pragma unroll
        for (int k = 0; k < 3; ++k) b2w[k] = a.maskbits[(size_t)(e0 + row) * 8 + 2 * k + (q >> 1)];
      }
      nbar_sync(kBarWorkers, kPipeWorkers);
      float4 v[4];
      gather_slab(v, 0);
#pragma unroll
      for (int sl = 0; sl < 3; ++sl) {
        const int c0 = sl * kStageCols;
        const int ncols = imin(kStageCols, kDEP - c0);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int r = warp * 8 + 2 * j + sub;
          if (4 * c4 < ncols) {
            float2* dst = reinterpret_cast<float2*>(stage + r * kStagePitch + 4 * c4);
            dst[0] = make_float2(v[j].x, v[j].y);
            dst[1] = make_float2(v[j].z, v[j].w);
          }
        }
        nbar_sync(kBarWorkers, kPipeWorkers);
        if (sl < 2) gather_slab(v, c0 + kStageCols);
        const int cb = 16 * q;
        const uint32_t bits = b2w[sl] >> (16 * (q & 1));
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const int c = cb + 8 * g;
          if (c < ncols) {                                 // warp-uniform
            const float2* src = reinterpret_cast<const float2*>(stage + row * kStagePitch + c);
            const float2 p0 = src[0], p1 = src[1], p2 = src[2], p3 = src[3];
            const float xx[8] = {p0.x, p0.y, p1.x, p1.y, p2.x, p2.y, p3.x, p3.y};
#pragma unroll
            for (int k = 0; k < 8; ++k) x.v[sl][8 * g + k] = ((bits >> (8 * g + k)) & 1u) ? xx[k] : 0.f;
          } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) x.v[sl][8 * g + k] = 0.f;
          }
        }
        nbar_sync(kBarWorkers, kPipeWorkers);
      }
    };

    if (cnt > 0) {
      build_x(0);
      pipe_store_lo(x, lane_addr, q);
      pipe_store_hi(x, lane_addr, q);
      tmem_wait_st();
      fence_before_sync();
      nbar_arrive(kBarOperands, kPipeThreads);
    }
    for (int i = 0; i < cnt; ++i) {
      const bool has_next = i + 1 < cnt;
      const uint32_t parity = (uint32_t)i & 1u;
      const int tile = blockIdx.x + i * gridDim.x;
      const int e0 = tile * kTM, rows = imin(kTM, a.E - e0);
      if (has_next) build_x(i + 1);                            // under the MMAs of tile i
      uint32_t b1w[3];                                         // relu bits of h1 of tile i (epilogue mask): in flight during the waits
#pragma unroll
      for (int k = 0; k < 3; ++k) b1w[k] = 0xffffffffu;
      if (row < rows && a.maskbits_h1) {
#pragma unroll
        for (int k = 0; k < 3; ++k) b1w[k] = a.maskbits_h1[(size_t)(e0 + row) * 8 + 2 * k + (q >> 1)];
      }
      if (!mbar_wait(barC, parity)) failed = true;
      fence_after_sync();
      if (has_next) pipe_store_lo(x, lane_addr, q);
      if (!mbar_wait(barM, parity)) failed = true;
      fence_after_sync();
      if (has_next) pipe_store_hi(x, lane_addr, q);
      uint32_t d[3][16];
      pipe_load_d(d, lane_addr, q);
      if (has_next) {
        tmem_wait_st();
        fence_before_sync();
        nbar_arrive(kBarOperands, kPipeThreads);
      }
      // ---- epilogue of tile i from registers: D * relu'(h1) -> slab -> DH1 (write) and dA (write or accumulate), coalesced
      const bool rmw = a.dA && !a.first;
#pragma unroll
      for (int sl = 0; sl < 3; ++sl) {
        const int c0 = sl * kStageCols;
        const int ncols = imin(kStageCols, kDEP - c0);         // columns 152..159 are never stored
        if (16 * q < imin(kStageCols, kN - c0)) {              // warp-uniform
          const int col0 = c0 + 16 * q;
          const uint32_t bits = b1w[sl] >> (16 * (q & 1));
          float o[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) o[k] = ((bits >> k) & 1u) && (col0 + k < kDE) ? __uint_as_float(d[sl][k]) : 0.f;
          float2* dst = reinterpret_cast<float2*>(stage + row * kStagePitch + 16 * q);
#pragma unroll
          for (int k = 0; k < 8; ++k) dst[k] = make_float2(o[2 * k], o[2 * k + 1]);
        }
        nbar_sync(kBarWorkers, kPipeWorkers);
        const int n4 = ncols >> 2;                             // float4 per row in this slab (16, 16, 6)
        const int total = rows * n4;
        for (int base = 0; base < total; base += 4 * kPipeWorkers) {
          float4 old[4];
          if (rmw) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int idx = base + u * kPipeWorkers + tid;
              if (idx < total) {
                const int r = idx / n4, qq = idx - r * n4;
                old[u] = *reinterpret_cast<const float4*>(a.dA + (size_t)(e0 + r) * kDEP + c0 + 4 * qq);
              }
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int idx = base + u * kPipeWorkers + tid;
            if (idx < total) {
              const int r = idx / n4, qq = idx - r * n4;
              const float2* src = reinterpret_cast<const float2*>(stage + r * kStagePitch + 4 * qq);
              const float2 p0 = src[0], p1 = src[1];
              float4 val = make_float4(p0.x, p0.y, p1.x, p1.y);
              const size_t g = (size_t)(e0 + r) * kDEP + c0 + 4 * qq;
              *reinterpret_cast<float4*>(a.DH1 + g) = val;
              if (a.dA) {
                if (rmw) { val.x += old[u].x; val.y += old[u].y; val.z += old[u].z; val.w += old[u].w; }
                *reinterpret_cast<float4*>(a.dA + g) = val;
              }
            }
          }
        }
        nbar_sync(kBarWorkers, kPipeWorkers);
      }
    }
    if (failed && tid == 0) a.poison[0] = __int_as_float(0x7fc00000);
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace tc
}  // namespace spw
#endif  // SPW_EMU
