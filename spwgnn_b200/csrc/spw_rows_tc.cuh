// spw_rows_tc.cuh -- generic fused linear layer on the tensor cores (tcgen05, 3xTF32), sm_100a only.
//
//   Y = post( act( [X0 | X1] . W + rowscale*bias + addend ) )          (same epilogue contract as k_linear)
//
// One 128-row tile per MMA group.  The row thread (TMEM lane) reads its own row from global memory, splits it
// into tf32 hi / lo and stores it straight to tensor memory (A operand, one 32-bit column per k); the packed
// weights (B operand, hi / lo, [k-step][2][NB][4]) stay resident in shared memory for the whole launch; the
// accumulator D lives in the top NB columns of tensor memory and is read back by the same row threads, which
// apply the epilogue and write 16-byte pieces of their row.
//   TMEM map: A_hi [0, 8 ks) | A_lo [8 ks, 16 ks) | D [512 - NB, 512)            (16 ks + NB <= 512)
//   Up to two row segments are concatenated along K (K0 % 4 == 0 then), e.g. [g | p] . [V1b ; V1c].
#pragma once
#ifndef SPW_EMU
#include "spw_tc.cuh"

namespace spw {
namespace tc {

// ---- batched weight packing into B operands of NB columns ------------------------------------------
// dst rows [k_off, k_lim) <- src (Keras [in][out], ld) block at (row0, col0): K valid rows, N valid columns;
// transpose: dst[k][n] = src[row0 + n][col0 + k].  Rows / columns beyond K / N are zero.
struct PackTcDesc {
  const float* src; int ld, row0, col0, K, N, transpose;
  float* hi; float* lo; int NB, k_off, k_lim;
};
constexpr int kMaxPackTc = 24;
struct PackTcArgs { PackTcDesc d[kMaxPackTc]; int n; };

__global__ void __launch_bounds__(256) k_pack_tc(PackTcArgs a) {
  const PackTcDesc d = a.d[blockIdx.y];
  const int total = (d.k_lim - d.k_off) * d.NB;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int kk = i / d.NB, n = i - kk * d.NB;          // kk: row within this descriptor
    const int k = d.k_off + kk;
    float v = 0.f;
    if (kk < d.K && n < d.N)
      v = d.transpose ? d.src[(size_t)(d.row0 + n) * d.ld + d.col0 + kk] : d.src[(size_t)(d.row0 + kk) * d.ld + d.col0 + n];
    uint32_t h, l;
    split_tf32(v, h, l);
    const size_t o = (size_t)(k >> 3) * (8 * d.NB) + (size_t)((k >> 2) & 1) * (4 * d.NB) + (size_t)n * 4 + (k & 3);
    d.hi[o] = __uint_as_float(h);
    d.lo[o] = __uint_as_float(l);
  }
}

// relation-encoder layer 0 (Networks.py:58-62,69,75; K = 2), one 16-byte piece of a row per thread:
//   X0[e][k] = relu(dx*W0[0][k] + dy*W0[1][k] + b0[k]),  [dx, dy] = pos_receiver - pos_sender;  X0[e][150] = 1, X0[e][151] = 0
__global__ void __launch_bounds__(256) k_edge_enc0(int E, const int32_t* __restrict__ in_snd, const int32_t* __restrict__ in_rcv,
                                                   const float* __restrict__ obj, const float* __restrict__ W0,
                                                   const float* __restrict__ b0, float* __restrict__ X0, uint32_t* __restrict__ bits) {
  constexpr int C4 = kDEP / 4;
  const long long total = (long long)E * C4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(i / C4), c = (int)(i - (long long)e * C4) * 4;
    const int s = in_snd[e], rc = in_rcv[e];
    const float dx = obj[3 * (size_t)rc] - obj[3 * (size_t)s], dy = obj[3 * (size_t)rc + 1] - obj[3 * (size_t)s + 1];
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = c + j;
      v[j] = k < kDE ? relu_f(fmaf(dy, __ldg(W0 + kDE + k), fmaf(dx, __ldg(W0 + k), __ldg(b0 + k)))) : (k == kDE ? 1.f : 0.f);
    }
    *reinterpret_cast<float4*>(X0 + (size_t)e * kDEP + c) = make_float4(v[0], v[1], v[2], v[3]);
  }
  if (bits) {   // sign bits of X0 (the mask of the layer's data gradient): one 32-column word per thread
    const long long nw = (long long)E * 8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nw; i += (long long)gridDim.x * blockDim.x) {
      const int e = (int)(i >> 3), w = (int)(i & 7);
      uint32_t m = 0u;
      if (w < 5) {
        const int s = in_snd[e], rc = in_rcv[e];
        const float dx = obj[3 * (size_t)rc] - obj[3 * (size_t)s], dy = obj[3 * (size_t)rc + 1] - obj[3 * (size_t)s + 1];
        for (int j = 0; j < 32; ++j) {
          const int k = 32 * w + j;
          if (k < kDE && fmaf(dy, __ldg(W0 + kDE + k), fmaf(dx, __ldg(W0 + k), __ldg(b0 + k))) > 0.f) m |= 1u << j;
        }
      }
      bits[i] = m;
    }
  }
}

struct RowsTcArgs {
  int M, nseg;
  const float* X[2]; int ldx[2]; int K[2];     // row segments: K valid columns each (K[0] % 4 == 0 when nseg == 2)
  const float* Bhi; const float* Blo;          // packed operands, ks k-steps
  int ks;                                      // ceil((K[0] + K[1]) / 8)
  int N;                                       // valid output columns (<= NB)
  const float* bias;                           // [N] or null (any alignment)
  const float* rowscale;                       // [M] multiplier of the bias or null
  const float* addend; int ld_add;             // pre-activation addend or null
  int act;                                     // 0 none, 1 relu, 2 tanh
  const float* mulsrc; int ld_mul; int mulmode;   // 1: *= [mulsrc > 0]   2: *= (1 - mulsrc^2)   3: *= bit of bits_in
  const uint32_t* bits_in;                     // [M][8] words: bit (col & 31) of word (col >> 5)   (mulmode 3)
  uint32_t* bits_out;                          // [M][8] or null: sign bits of the result, same layout
  float* Y; int ldy;                           // columns N..ldy-1 are written as 0 (ones_col: 1)
  int accumulate; float post_scale;
  uint32_t drop_thresh, drop_seed; float drop_inv_keep; int drop_stride;   // element index = row * drop_stride + col
  int ones_col;                                // >= N: Y[row][ones_col] = 1 (bias pick-up column of a later X^T.dY); -1: none
  float* poison;                               // written with NaN if an MMA barrier times out
};

template <int NB>
constexpr size_t rows_tc_smem(int ks) { return (size_t)2 * ks * 8 * NB * sizeof(float) + 32; }

// 8 consecutive columns [c, c+8) of the concatenated row of this thread -> x[8] (zeros beyond the valid columns)
__device__ __forceinline__ void rows_load8(const RowsTcArgs& a, const float* p0, const float* p1, bool valid, int c, float4& u,
                                           float4& v) {
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  u = z; v = z;
  if (!valid) return;
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    float4& dst = g ? v : u;
    int cc = c + 4 * g;
    const float* p = p0;
    int K = a.K[0];
    if (a.nseg == 2 && cc >= a.K[0]) { cc -= a.K[0]; p = p1; K = a.K[1]; }
    if (cc + 3 < K) {
      dst = *reinterpret_cast<const float4*>(p + cc);
    } else {
      if (cc < K) dst.x = p[cc];
      if (cc + 1 < K) dst.y = p[cc + 1];
      if (cc + 2 < K) dst.z = p[cc + 2];
    }
  }
}

// 16 consecutive floats row[col0 .. col0+16) -> o (zeros at columns >= lim); 16-byte loads where the pointer allows it
__device__ __forceinline__ void rows_load16(const float* __restrict__ rowp, int col0, int lim, bool vec_ok, float (&o)[16]) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const int c = col0 + 4 * g;
    if (vec_ok && c + 3 < lim) {
      const float4 q = *reinterpret_cast<const float4*>(rowp + c);
      o[4 * g] = q.x; o[4 * g + 1] = q.y; o[4 * g + 2] = q.z; o[4 * g + 3] = q.w;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) o[4 * g + i] = c + i < lim ? rowp[c + i] : 0.f;
    }
  }
}

constexpr int kRowsKpt = 13;     // k-steps per thread at most (ks <= 25, two threads per row)

template <int NB>
__global__ void __launch_bounds__(kThreads, 1) k_rows_tc(RowsTcArgs a) {
  SPW_DYN_SMEM(smem_raw);
  const int bfl = a.ks * 8 * NB;                                 // floats per hi / lo operand
  float* Bhi_s = reinterpret_cast<float*>(smem_raw);
  float* Blo_s = Bhi_s + bfl;
  uint64_t* bar = reinterpret_cast<uint64_t*>(Blo_s + bfl);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = 32 * (warp & 3) + lane, half = warp >> 2;

  for (int i = tid; i < bfl / 4; i += kThreads) {                // weights: asynchronous, overlapped with the first row loads
    cp_async16(Bhi_s + 4 * i, a.Bhi + 4 * i);
    cp_async16(Blo_s + 4 * i, a.Blo + 4 * i);
  }
  cp_async_commit();
  if (warp == 0) tmem_alloc(tptr, kTmemCols);
  if (tid == 32) { mbar_init(bar, 1); fence_mbar_init(); }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tptr;
  const uint32_t lane_addr = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16);
  const uint32_t colAhi = 0, colAlo = 8 * a.ks, colD = kTmemCols - NB;
  const uint32_t idesc = make_idesc_tf32(128, NB);
  uint32_t parity = 0;
  bool failed = false, first = true;
  const int ntiles = (a.M + kTM - 1) / kTM;
  const int ks_h0 = (a.ks + 1) >> 1;                             // k-steps [0, ks_h0) belong to half 0, the rest to half 1
  const int ks_lo = half ? ks_h0 : 0, ks_hi = half ? a.ks : ks_h0;
  const bool bias_vec = (reinterpret_cast<uintptr_t>(a.bias) & 15) == 0;

  // this thread's k-steps of its row of a tile -> registers (all loads in flight at once)
  float4 pu[kRowsKpt], pv[kRowsKpt];
  auto load_tile = [&](int tile) {
    const size_t gr = (size_t)tile * kTM + row;
    const bool ok = gr < (size_t)a.M;
    const float* p0 = a.X[0] + gr * a.ldx[0];
    const float* p1 = a.nseg == 2 ? a.X[1] + gr * a.ldx[1] : nullptr;
#pragma unroll
    for (int j = 0; j < kRowsKpt; ++j)
      if (ks_lo + j < ks_hi) rows_load8(a, p0, p1, ok, 8 * (ks_lo + j), pu[j], pv[j]);
  };
  if ((int)blockIdx.x < ntiles) load_tile(blockIdx.x);

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const size_t grow = (size_t)tile * kTM + row;
    const bool valid = grow < (size_t)a.M;
    // ---- A operand: split into tf32 hi / lo, store to tensor memory
#pragma unroll
    for (int j = 0; j < kRowsKpt; ++j) {
      if (ks_lo + j < ks_hi) {                                   // warp-uniform
        const float x[8] = {pu[j].x, pu[j].y, pu[j].z, pu[j].w, pv[j].x, pv[j].y, pv[j].z, pv[j].w};
        uint32_t h[8], l[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) split_tf32(x[i], h[i], l[i]);
        tmem_st8(lane_addr + colAhi + 8 * (ks_lo + j), h);
        tmem_st8(lane_addr + colAlo + 8 * (ks_lo + j), l);
      }
    }
    tmem_wait_st();
    if (first) { cp_async_wait<0>(); fence_async_smem(); first = false; }
    fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      fence_after_sync();
      const uint32_t bhi = smem_u32(Bhi_s), blo = smem_u32(Blo_s);
      const uint32_t d = tmem_base + colD;
      // correction products first, main products last (the tensor core truncates when it accumulates)
#pragma unroll 1
      for (int ks = 0; ks < a.ks; ++ks) {
        const uint64_t dhi = make_b_desc(bhi + ks * (8 * NB * 4), NB * 16, 128);
        const uint64_t dlo = make_b_desc(blo + ks * (8 * NB * 4), NB * 16, 128);
        mma_tf32_ts(d, tmem_base + colAlo + 8 * ks, dhi, idesc, ks > 0 ? 1u : 0u);
        mma_tf32_ts(d, tmem_base + colAhi + 8 * ks, dlo, idesc, 1u);
      }
#pragma unroll 1
      for (int ks = 0; ks < a.ks; ++ks) {
        const uint64_t dhi = make_b_desc(bhi + ks * (8 * NB * 4), NB * 16, 128);
        mma_tf32_ts(d, tmem_base + colAhi + 8 * ks, dhi, idesc, 1u);
      }
      mma_commit(bar);
    }
    // rows of this CTA's next tile: in flight while the MMAs run and the epilogue is written
    if (tile + (int)gridDim.x < ntiles) load_tile(tile + gridDim.x);
    if (!mbar_wait(bar, parity)) failed = true;
    parity ^= 1u;
    fence_after_sync();
    // ---- epilogue: 16-column blocks of this thread's row, alternating between the two halves
    const float rs = (valid && a.rowscale) ? a.rowscale[grow] : 1.f;
    for (int blk = half; blk * 16 < NB; blk += 2) {
      const int col0 = blk * 16;
      if (col0 >= a.ldy) break;                                  // warp-uniform
      uint32_t vv[16];
      tmem_ld16(lane_addr + colD + col0, vv);
      float r[16], m[16];
      if (valid && a.bias) rows_load16(a.bias, col0, a.N, bias_vec, m);
      tmem_wait_ld();
      if (!valid) continue;
#pragma unroll
      for (int i = 0; i < 16; ++i) r[i] = __uint_as_float(vv[i]);
      if (a.bias) {
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = fmaf(rs, m[i], r[i]);
      }
      if (a.addend) {
        rows_load16(a.addend + grow * a.ld_add, col0, a.N, true, m);
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] += m[i];
      }
      if (a.act == 1) {
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = relu_f(r[i]);
      } else if (a.act == 2) {
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = tanhf(r[i]);
      }
      if (a.mulmode == 3) {                                      // 16 mask bits of this block: half-word blk of the row
        const uint32_t bits = reinterpret_cast<const uint16_t*>(a.bits_in)[grow * 16 + blk];
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = ((bits >> i) & 1u) ? r[i] : 0.f;
      } else if (a.mulmode) {
        rows_load16(a.mulsrc + grow * a.ld_mul, col0, a.N, true, m);
        if (a.mulmode == 1) {
#pragma unroll
          for (int i = 0; i < 16; ++i) r[i] = m[i] > 0.f ? r[i] : 0.f;
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) r[i] *= (1.f - m[i] * m[i]);
        }
      }
      if (a.drop_thresh) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
          r[i] = dropout_apply(r[i], a.drop_seed, (uint32_t)(grow * a.drop_stride + col0 + i), a.drop_thresh, a.drop_inv_keep);
      }
      if (a.post_scale != 1.f) {
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] *= a.post_scale;
      }
      if (a.accumulate) {
        rows_load16(a.Y + grow * a.ldy, col0, a.N, true, m);
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] += m[i];
      }
      if (a.bits_out) {
        uint32_t bits = 0u;
#pragma unroll
        for (int i = 0; i < 16; ++i) bits |= (col0 + i < a.N && r[i] > 0.f) ? (1u << i) : 0u;
        reinterpret_cast<uint16_t*>(a.bits_out)[grow * 16 + blk] = (uint16_t)bits;
      }
      if (col0 + 16 > a.N) {                                     // columns beyond N: zero (or the ones column)
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (col0 + i >= a.N) r[i] = (col0 + i == a.ones_col) ? 1.f : 0.f;
      }
      float* yrow = a.Y + grow * a.ldy;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int c = col0 + 4 * g;
        if (c + 3 < a.ldy) {
          *reinterpret_cast<float4*>(yrow + c) = make_float4(r[4 * g], r[4 * g + 1], r[4 * g + 2], r[4 * g + 3]);
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (c + i < a.ldy) yrow[c + i] = r[4 * g + i];
        }
      }
    }
    fence_before_sync();
    __syncthreads();
  }
  if (failed && tid == 0) a.poison[0] = __int_as_float(0x7fc00000);
  if (first) cp_async_wait<0>();
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace tc
}  // namespace spw
#endif  // SPW_EMU
