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
  // 40 chunk-threads per row (38 write a piece of X0): the 8 threads of one 32-column word are 8 consecutive, 8-aligned lanes,
  // so the sign bits of X0 (the mask of the layer's data gradient) come from three shuffles
  constexpr int C4 = 40;
  const long long total = (long long)E * C4;
  const long long span = (long long)gridDim.x * blockDim.x;
  const long long iters = (total + span - 1) / span;
  for (long long it = 0; it < iters; ++it) {
    const long long i = it * span + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool in_range = i < total;
    const int e = in_range ? (int)(i / C4) : 0, c4 = in_range ? (int)(i - (long long)e * C4) : 0, c = 4 * c4;
    uint32_t nib = 0u;
    if (in_range && c < kDEP) {
      const int s = in_snd[e], rc = in_rcv[e];
      const float dx = obj[3 * (size_t)rc] - obj[3 * (size_t)s], dy = obj[3 * (size_t)rc + 1] - obj[3 * (size_t)s + 1];
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = c + j;
        v[j] = k < kDE ? relu_f(fmaf(dy, __ldg(W0 + kDE + k), fmaf(dx, __ldg(W0 + k), __ldg(b0 + k)))) : (k == kDE ? 1.f : 0.f);
        if (k < kDE && v[j] > 0.f) nib |= 1u << j;
      }
      *reinterpret_cast<float4*>(X0 + (size_t)e * kDEP + c) = make_float4(v[0], v[1], v[2], v[3]);
    }
    if (bits) {                                                  // block-uniform
      uint32_t w = nib << (4 * (c4 & 7));
      w |= __shfl_xor_sync(0xffffffffu, w, 1);
      w |= __shfl_xor_sync(0xffffffffu, w, 2);
      w |= __shfl_xor_sync(0xffffffffu, w, 4);
      if (in_range && (c4 & 7) == 0) bits[(size_t)e * 8 + (c4 >> 3)] = w;
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

// ---- k_rows_tc ------------------------------------------------------------------------------------
// All global traffic is coalesced (a 16-byte access per lane that touches 32 different 128-byte lines costs the L1 about
// 65 cycles; eight lanes per 128 bytes of a row cost a fifth of that): a tile's rows are fetched as 16-byte chunks in
// (row, chunk) order into registers one tile ahead, and are transposed to the row-per-thread order tensor memory needs
// through two 16 KB shared-memory slabs of 128 rows x 32 columns (chunk index XOR-swizzled with the row: conflict-free
// both ways).  The accumulator goes back the same way: row threads read D from tensor memory and drop it into a slab,
// then every thread applies the epilogue to four (row, chunk) pieces and writes them coalesced.
constexpr int kSlabCols = 32;
constexpr int kSlabFloats = kTM * kSlabCols;        // 4096 floats = 16 KB
constexpr int kRowsMaxSlabs = 7;                    // ks <= 25 -> ceil(25 / 4)
constexpr int kRowsEarlySlabs = 3;                  // correction MMAs of the first slabs are issued while the rest is staged

template <int NB>
constexpr size_t rows_tc_smem(int ks) { return (size_t)(2 * ks * 8 * NB + 2 * kSlabFloats + NB + kTM * 8) * sizeof(float) + 32; }

__device__ __forceinline__ int slab_off(int row, int ch) { return row * kSlabCols + ((ch ^ (row & 7)) << 2); }

// Sign words of a tile ([128 rows][8 words], contiguous in global memory) staged in shared memory: one coalesced 4 KB
// transfer per tile instead of a scattered 4-byte access per row and slab.
__device__ __forceinline__ void rows_bits_load(const RowsTcArgs& a, uint32_t* sbits, int r0) {
  if (a.mulmode == 3 && threadIdx.x < 256) {
    const int r = threadIdx.x >> 1;
    uint4 q = make_uint4(0u, 0u, 0u, 0u);
    if (r0 + r < a.M) q = reinterpret_cast<const uint4*>(a.bits_in + (size_t)r0 * 8)[threadIdx.x];
    reinterpret_cast<uint4*>(sbits)[threadIdx.x] = q;
  }
}
__device__ __forceinline__ void rows_bits_store(const RowsTcArgs& a, const uint32_t* sbits, int r0) {
  if (a.bits_out && threadIdx.x < 256) {
    const int r = threadIdx.x >> 1;
    if (r0 + r < a.M) reinterpret_cast<uint4*>(a.bits_out + (size_t)r0 * 8)[threadIdx.x] = reinterpret_cast<const uint4*>(sbits)[threadIdx.x];
  }
}

// Epilogue of PP (row, chunk) pieces of one 32-column accumulator slab: rows crow0 + rstride * i (i < PP) of the tile,
// chunk cch (columns 32 es + 4 cch ...).  Reads the raw accumulators from the slab, applies the LinArgs contract and
// writes 16 bytes per piece (eight lanes per 128 bytes of a row: coalesced).  Must be called by whole warps (shuffles).
template <int NB, int PP>
__device__ __forceinline__ void rows_epilogue(const RowsTcArgs& a, const float* sb, const float* sbias, uint32_t* sbits, int r0,
                                               int crow0, int rstride, int cch, int es) {
  const int col = kSlabCols * es + 4 * cch;
  const bool col_ok = col < a.ldy;
  const bool full = col + 3 < a.N;
  float r[PP][4], m[PP][4];
  bool rv[PP];
  size_t go[PP];
#pragma unroll
  for (int i = 0; i < PP; ++i) {
    const size_t grow = (size_t)r0 + crow0 + rstride * i;
    rv[i] = col_ok && grow < (size_t)a.M;
    go[i] = rv[i] ? grow : 0;
    const float4 q = *reinterpret_cast<const float4*>(sb + slab_off(crow0 + rstride * i, cch));
    r[i][0] = q.x; r[i][1] = q.y; r[i][2] = q.z; r[i][3] = q.w;
  }
  auto load_m = [&](const float* base, int ld) {             // the four (row, chunk) pieces of another array, zeros beyond N
#pragma unroll
    for (int i = 0; i < PP; ++i) {
      if (full) {
        const float4 t = *reinterpret_cast<const float4*>(base + go[i] * ld + col);
        m[i][0] = t.x; m[i][1] = t.y; m[i][2] = t.z; m[i][3] = t.w;
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) m[i][e] = (col_ok && col + e < a.N) ? base[go[i] * ld + col + e] : 0.f;
      }
    }
  };
  if (a.bias) {
    const float4 bq = *reinterpret_cast<const float4*>(sbias + (col < NB ? col : 0));
#pragma unroll
    for (int i = 0; i < PP; ++i) {
      const float rs = a.rowscale ? a.rowscale[go[i]] : 1.f;
      r[i][0] = fmaf(rs, bq.x, r[i][0]); r[i][1] = fmaf(rs, bq.y, r[i][1]);
      r[i][2] = fmaf(rs, bq.z, r[i][2]); r[i][3] = fmaf(rs, bq.w, r[i][3]);
    }
  }
  if (a.addend) {
    load_m(a.addend, a.ld_add);
#pragma unroll
    for (int i = 0; i < PP; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) r[i][e] += m[i][e];
  }
  if (a.act == 1) {
#pragma unroll
    for (int i = 0; i < PP; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) r[i][e] = relu_f(r[i][e]);
  } else if (a.act == 2) {
#pragma unroll
    for (int i = 0; i < PP; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) r[i][e] = tanhf(r[i][e]);
  }
  if (a.mulmode == 3) {                                      // mask bits: word es of the row, nibble cch
#pragma unroll
    for (int i = 0; i < PP; ++i) {
      const uint32_t bits = sbits[(crow0 + rstride * i) * 8 + es] >> (4 * cch);      // staged per tile (rows_bits_load)
#pragma unroll
      for (int e = 0; e < 4; ++e) r[i][e] = ((bits >> e) & 1u) ? r[i][e] : 0.f;
    }
  } else if (a.mulmode) {
    load_m(a.mulsrc, a.ld_mul);
#pragma unroll
    for (int i = 0; i < PP; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) r[i][e] = a.mulmode == 1 ? (m[i][e] > 0.f ? r[i][e] : 0.f) : r[i][e] * (1.f - m[i][e] * m[i][e]);
  }
  if (a.drop_thresh) {
#pragma unroll
    for (int i = 0; i < PP; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e)
        r[i][e] = dropout_apply(r[i][e], a.drop_seed, (uint32_t)(go[i] * a.drop_stride + col + e), a.drop_thresh, a.drop_inv_keep);
  }
  if (a.post_scale != 1.f) {
#pragma unroll
    for (int i = 0; i < PP; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) r[i][e] *= a.post_scale;
  }
  if (a.accumulate) {
    load_m(a.Y, a.ldy);
#pragma unroll
    for (int i = 0; i < PP; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) r[i][e] += m[i][e];
  }
  if (a.bits_out) {                                          // sign bits: the 8 chunk-threads of a row make up word es
#pragma unroll
    for (int i = 0; i < PP; ++i) {
      uint32_t nib = 0u;
#pragma unroll
      for (int e = 0; e < 4; ++e) nib |= (rv[i] && col + e < a.N && r[i][e] > 0.f) ? (1u << e) : 0u;
      uint32_t w = nib << (4 * cch);
      w |= __shfl_xor_sync(0xffffffffu, w, 1);
      w |= __shfl_xor_sync(0xffffffffu, w, 2);
      w |= __shfl_xor_sync(0xffffffffu, w, 4);
      if (cch == 0) sbits[(crow0 + rstride * i) * 8 + es] = w;                         // written out per tile (rows_bits_store)
    }
  }
  if (!full) {                                               // columns beyond N: zero (or the ones column)
#pragma unroll
    for (int i = 0; i < PP; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (col + e >= a.N) r[i][e] = (col + e == a.ones_col) ? 1.f : 0.f;
  }
#pragma unroll
  for (int i = 0; i < PP; ++i) {
    if (!rv[i]) continue;
    float* yp = a.Y + go[i] * a.ldy + col;
    if (col + 3 < a.ldy) {
      *reinterpret_cast<float4*>(yp) = make_float4(r[i][0], r[i][1], r[i][2], r[i][3]);
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (col + e < a.ldy) yp[e] = r[i][e];
    }
  }
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}

// NT threads (256 or 512): NT / 128 threads share a row in the tensor-memory view, 1024 / NT (row, chunk) pieces per thread
// and slab in the coalesced view.  More warps = less serial work per warp between the slab barriers.
template <int NB, int NT>
__global__ void __launch_bounds__(NT, 1) k_rows_tc(RowsTcArgs a) {
  constexpr int PP = 1024 / NT;            // (row, chunk) pieces per thread and slab
  constexpr int NPART = NT / 128;          // threads per row (tensor-memory view)
  constexpr int KPT = 4 / NPART;           // k-steps per thread and operand slab
  constexpr int CPT = 8 / NPART;           // 16-byte chunks per thread and accumulator slab
  constexpr int RS = NT / 8;               // row stride between the pieces of a thread
  SPW_DYN_SMEM(smem_raw);
  const int bfl = a.ks * 8 * NB;                                 // floats per hi / lo operand
  float* Bhi_s = reinterpret_cast<float*>(smem_raw);
  float* Blo_s = Bhi_s + bfl;
  float* slab = Blo_s + bfl;                                     // [2][128][32]
  float* sbias = slab + 2 * kSlabFloats;                         // [NB]
  uint32_t* sbits = reinterpret_cast<uint32_t*>(sbias + NB);     // [128][8] sign words of the tile (mask in or sign bits out)
  uint64_t* bar = reinterpret_cast<uint64_t*>(sbits + kTM * 8);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = 32 * (warp & 3) + lane, part = warp >> 2;      // tensor-memory view: thread = (row, part of the slab)
  const int crow = tid >> 3, cch = tid & 7;                      // coalesced view: rows crow + RS i (i < PP), chunk cch

  for (int i = tid; i < bfl / 4; i += NT) {                      // weights: asynchronous, overlapped with the first row loads
    cp_async16(Bhi_s + 4 * i, a.Bhi + 4 * i);
    cp_async16(Blo_s + 4 * i, a.Blo + 4 * i);
  }
  cp_async_commit();
  for (int i = tid; i < NB; i += NT) sbias[i] = (a.bias && i < a.N) ? a.bias[i] : 0.f;
  for (int i = tid; i < kTM * 8; i += NT) sbits[i] = 0u;
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
  const int ns = (a.ks + 3) >> 2;                                // operand slabs per tile (4 k-steps each)
  const int ks_early = ns > kRowsEarlySlabs ? 4 * kRowsEarlySlabs : 0;

  // ---- rows of a tile -> registers, coalesced; unconditional loads (addresses clamped), masked when stored to the slab
  float4 P[kRowsMaxSlabs][PP];
  auto load_tile = [&](int tile) {
    const float* pr0[PP]; const float* pr1[PP];
#pragma unroll
    for (int i = 0; i < PP; ++i) {
      size_t gr = (size_t)tile * kTM + crow + RS * i;
      if (gr >= (size_t)a.M) gr = (size_t)a.M - 1;
      pr0[i] = a.X[0] + gr * a.ldx[0];
      pr1[i] = a.nseg == 2 ? a.X[1] + gr * a.ldx[1] : pr0[i];
    }
#pragma unroll
    for (int sl = 0; sl < kRowsMaxSlabs; ++sl) {
      if (sl < ns) {                                             // block-uniform
        int cc = kSlabCols * sl + 4 * cch;
        bool seg1 = false;
        if (a.nseg == 2 && cc >= a.K[0]) { cc -= a.K[0]; seg1 = true; }
        if (cc >= (seg1 ? a.K[1] : a.K[0])) cc = 0;
#pragma unroll
        for (int i = 0; i < PP; ++i) P[sl][i] = *reinterpret_cast<const float4*>((seg1 ? pr1[i] : pr0[i]) + cc);
      }
    }
  };
  if ((int)blockIdx.x < ntiles) load_tile(blockIdx.x);
  SPW_PH_DECL

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    SPW_PH(7);
    const int r0 = tile * kTM;
    rows_bits_load(a, sbits, r0);                                // used after several barriers (epilogue)
    const uint64_t dhi0 = make_b_desc(smem_u32(Bhi_s), NB * 16, 128), dlo0 = make_b_desc(smem_u32(Blo_s), NB * 16, 128);
    constexpr uint64_t kStep = (8 * NB * 4) >> 4;                // descriptor start-address units (16 bytes) per k-step
    const uint32_t dcol = tmem_base + colD;
    // ---- A operand, slab by slab: registers -> slab (masked) -> row threads -> tf32 hi / lo -> tensor memory
#pragma unroll
    for (int sl = 0; sl < kRowsMaxSlabs; ++sl) {
      if (sl < ns) {
        float* sb = slab + (sl & 1) * kSlabFloats;
        {
          int cc = kSlabCols * sl + 4 * cch;
          int K = a.K[0];
          if (a.nseg == 2 && cc >= a.K[0]) { cc -= a.K[0]; K = a.K[1]; }
          const int ncol = K - cc;                               // valid columns of this chunk: >= 4 all, <= 0 none
#pragma unroll
          for (int i = 0; i < PP; ++i) {
            float4 q = P[sl][i];
            const int nval = (r0 + crow + RS * i < a.M) ? ncol : 0;
            if (nval < 4) { if (nval < 1) q.x = 0.f; if (nval < 2) q.y = 0.f; if (nval < 3) q.z = 0.f; q.w = 0.f; }
            *reinterpret_cast<float4*>(sb + slab_off(crow + RS * i, cch)) = q;
          }
        }
        const bool early = ks_early > 0 && sl == kRowsEarlySlabs;  // the first slabs are complete in tensor memory: start their MMAs
        if (early) { tmem_wait_st(); fence_before_sync(); }
        if (first && sl == 0) { cp_async_wait<0>(); fence_async_smem(); }
        __syncthreads();
        if (early && tid == 0) {                                 // correction products of k-steps [0, ks_early)
          fence_after_sync();
#pragma unroll 4
          for (int ks = 0; ks < ks_early; ++ks) {
            mma_tf32_ts(dcol, tmem_base + colAlo + 8 * ks, dhi0 + ks * kStep, idesc, ks > 0 ? 1u : 0u);
            mma_tf32_ts(dcol, tmem_base + colAhi + 8 * ks, dlo0 + ks * kStep, idesc, 1u);
          }
        }
#pragma unroll
        for (int j = 0; j < KPT; ++j) {                          // this thread's k-steps of the slab
          const int ks = 4 * sl + KPT * part + j;
          if (ks < a.ks) {                                       // warp-uniform
            const float4 u = *reinterpret_cast<const float4*>(sb + slab_off(row, 2 * (KPT * part + j)));
            const float4 v = *reinterpret_cast<const float4*>(sb + slab_off(row, 2 * (KPT * part + j) + 1));
            const float x[8] = {u.x, u.y, u.z, u.w, v.x, v.y, v.z, v.w};
            uint32_t h[8], l[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) split_tf32(x[e], h[e], l[e]);
            tmem_st8(lane_addr + colAhi + 8 * ks, h);
            tmem_st8(lane_addr + colAlo + 8 * ks, l);
          }
        }
      }
    }
    first = false;
    tmem_wait_st();
    fence_before_sync();
    SPW_PH(0);                                                   // p0: operand staging
    __syncthreads();
    SPW_PH(1);
    if (tid == 0) {
      fence_after_sync();
#pragma unroll 4
      for (int ks = ks_early; ks < a.ks; ++ks) {                 // remaining correction products
        mma_tf32_ts(dcol, tmem_base + colAlo + 8 * ks, dhi0 + ks * kStep, idesc, ks > 0 ? 1u : 0u);
        mma_tf32_ts(dcol, tmem_base + colAhi + 8 * ks, dlo0 + ks * kStep, idesc, 1u);
      }
#pragma unroll 4
      for (int ks = 0; ks < a.ks; ++ks) mma_tf32_ts(dcol, tmem_base + colAhi + 8 * ks, dhi0 + ks * kStep, idesc, 1u);   // main products last
      mma_commit(bar);
    }
    SPW_PH(2);                                                   // p2: MMA issue
    // rows of this CTA's next tile: in flight while the MMAs run and the epilogue is written
    if (tile + (int)gridDim.x < ntiles) load_tile(tile + gridDim.x);
    SPW_PH(3);                                                   // p3: issue of the next tile's loads
    if (!mbar_wait(bar, parity)) failed = true;
    parity ^= 1u;
    fence_after_sync();
    SPW_PH(4);                                                   // p4: wait for the MMAs
    // ---- epilogue, 32 accumulator columns at a time: tensor memory -> slab (row threads) -> epilogue + store (coalesced)
    constexpr int kEs = (NB + kSlabCols - 1) / kSlabCols;
#pragma unroll 1
    for (int es = 0; es < kEs; ++es) {
      if (kSlabCols * es >= a.ldy) break;                        // block-uniform
      float* sb = slab + (es & 1) * kSlabFloats;
      if (kSlabCols * es + 4 * CPT * part < NB) {                // warp-uniform
        uint32_t vv[4 * CPT];
        if (CPT == 4) tmem_ld16(lane_addr + colD + kSlabCols * es + 16 * part, reinterpret_cast<uint32_t(&)[16]>(vv[0]));
        else tmem_ld8(lane_addr + colD + kSlabCols * es + 8 * part, reinterpret_cast<uint32_t(&)[8]>(vv[0]));
        tmem_wait_ld();
#pragma unroll
        for (int g = 0; g < CPT; ++g)
          *reinterpret_cast<uint4*>(sb + slab_off(row, CPT * part + g)) = make_uint4(vv[4 * g], vv[4 * g + 1], vv[4 * g + 2], vv[4 * g + 3]);
      }
      __syncthreads();
      rows_epilogue<NB, PP>(a, sb, sbias, sbits, r0, crow, RS, cch, es);
    }
    fence_before_sync();
    SPW_PH(5);                                                   // p5: epilogue
    __syncthreads();
    rows_bits_store(a, sbits, r0);                               // sbits is next written after the next tile's staging barriers
    SPW_PH(6);
  }
  if (a.M > 100000) SPW_PH_REPORT(a.mulmode ? "k_rows_tc:enc_bwd" : "k_rows_tc:enc_fwd");
  if (failed && tid == 0) a.poison[0] = __int_as_float(0x7fc00000);
  if (first) cp_async_wait<0>();
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, kTmemCols);
}

// =================================================================================================
// CTA-pair (cta_group::2) building block, self test only this round: one cluster of two CTAs computes
//   D[256][160] = A[256][152] . B   (3xTF32)
// Each CTA keeps ITS 128 rows of A in its own tensor memory and HALF of the weight operand (80 of the 160 columns, hi and lo:
// 95 KB instead of 190 KB) in its own shared memory; the leader CTA issues tcgen05.mma.cta_group::2 (M = 256, N = 160), both
// CTAs receive the completion through a multicast commit and read their 128 rows of D from their own tensor memory.
// =================================================================================================
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols));
}
__device__ __forceinline__ void mma2_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void mma2_commit_multicast(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}

constexpr int kHalfN = kN / 2;                                   // 80 columns of the weight operand per CTA
constexpr int kB2Floats = kKS * 8 * kHalfN;                      // floats per hi / lo half operand

// Bhi / Blo: [2 halves][kKS][2][80][4] (k_pack_tc with NB = 80, one descriptor per half)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
k_tc2_selftest(const float* __restrict__ A, const float* __restrict__ Bhi, const float* __restrict__ Blo, float* __restrict__ D,
               int* __restrict__ status) {
  SPW_DYN_SMEM(smem_raw);
  float* Bhi_s = reinterpret_cast<float*>(smem_raw);
  float* Blo_s = Bhi_s + kB2Floats;
  uint64_t* bar = reinterpret_cast<uint64_t*>(Blo_s + kB2Floats);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = cluster_ctarank();
  if (warp == 0) tmem_alloc2(tptr, kTmemCols);
  if (tid == 32) { mbar_init(bar, 1); fence_mbar_init(); }
  for (int i = tid; i < kB2Floats / 4; i += 128) {               // this CTA's half of the weight columns
    reinterpret_cast<float4*>(Bhi_s)[i] = reinterpret_cast<const float4*>(Bhi + (size_t)rank * kB2Floats)[i];
    reinterpret_cast<float4*>(Blo_s)[i] = reinterpret_cast<const float4*>(Blo + (size_t)rank * kB2Floats)[i];
  }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tptr;
  const uint32_t lane_addr = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16);
  const float* arow = A + ((size_t)rank * 128 + tid) * kDEP;     // this CTA's 128 rows
#pragma unroll 1
  for (int c = 0; c < kDEP; c += 8) {
    const float4 x0 = *reinterpret_cast<const float4*>(arow + c), x1 = *reinterpret_cast<const float4*>(arow + c + 4);
    const float x[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
    uint32_t h[8], l[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) split_tf32(x[i], h[i], l[i]);
    tmem_st8(lane_addr + kColAhi + c, h);
    tmem_st8(lane_addr + kColAlo + c, l);
  }
  tmem_wait_st();
  fence_before_sync();
  cluster_sync_all();                                            // both CTAs: A in tensor memory, weights in shared memory
  if (rank == 0 && tid == 0) {
    fence_after_sync();
    const uint32_t idesc = make_idesc_tf32(256, kN);
    const uint64_t dhi0 = make_b_desc(smem_u32(Bhi_s), kHalfN * 16, 128), dlo0 = make_b_desc(smem_u32(Blo_s), kHalfN * 16, 128);
    constexpr uint64_t kStep = (8 * kHalfN * 4) >> 4;
    const uint32_t d = tmem_base + kColD;
    for (int ks = 0; ks < kKS; ++ks) {
      mma2_tf32_ts(d, tmem_base + kColAlo + 8 * ks, dhi0 + ks * kStep, idesc, ks > 0 ? 1u : 0u);
      mma2_tf32_ts(d, tmem_base + kColAhi + 8 * ks, dlo0 + ks * kStep, idesc, 1u);
    }
    for (int ks = 0; ks < kKS; ++ks) mma2_tf32_ts(d, tmem_base + kColAhi + 8 * ks, dhi0 + ks * kStep, idesc, 1u);
    mma2_commit_multicast(bar);
  }
  const bool ok = mbar_wait(bar, 0);
  fence_after_sync();
  if (!ok) {
    if (tid == 0) status[rank] = -1;
  } else {
#pragma unroll 1
    for (int c = 0; c < kN; c += 16) {
      uint32_t v[16];
      tmem_ld16(lane_addr + kColD + c, v);
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 16; ++i) D[((size_t)rank * 128 + tid) * kN + c + i] = __uint_as_float(v[i]);
    }
    if (tid == 0) status[rank] = 1;
  }
  fence_before_sync();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc2(tmem_base, kTmemCols);
}

// =================================================================================================
// k_rows_pair: the 150 -> 150 layers of the relation encoder (K = 152, N = 160, rows of 152 floats) on CTA PAIRS.
//   Each CTA of a cluster of two keeps 80 of the 160 weight columns (hi and lo, 95 KB) -- the tcgen05.mma.cta_group::2
//   stream of the leader reads both halves -- which frees the shared memory for a whole raw input tile: the next
//   tile's 128 rows (contiguous in HBM, 76 KB) arrive with ONE TMA bulk copy while the current tile is converted,
//   multiplied and written, so no warp ever blocks on issuing loads.  The epilogue is k_rows_tc's (slab + coalesced stores).
//   Per tile and CTA: wait for the bulk copy -> row threads read their row from shared memory, split, store to tensor
//   memory -> cluster barrier -> leader issues 57 MMAs (M = 256) + multicast commit -> both CTAs run their epilogue.
// =================================================================================================
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// TMA bulk copy global -> shared (16-byte aligned, size a multiple of 16), completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(void* sdst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sdst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void mbar_arrive_cta1(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// arrive on the mbarrier at the same shared-memory offset in CTA `cta` of the cluster (release at cluster scope)
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(bar)), "r"(cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
// bounded wait with acquire at cluster scope (the arrivals come from both CTAs of the pair)
__device__ __forceinline__ bool mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
#pragma unroll 1
  for (int it = 0; it < kWaitSpin; ++it) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) return true;
  }
  return false;
}

constexpr int kPairThreads = 512;
constexpr int kXbufFloats = kTM * kDEP;                          // one raw input tile
constexpr size_t kRowsPairSmem = (size_t)(2 * kB2Floats + kXbufFloats + 2 * kSlabFloats + kN + kTM * 8) * sizeof(float) + 64;

// a.Bhi / a.Blo: [2 halves][kKS][2][80][4] (k_pack_tc with NB = 80, one descriptor per half); a.ldx[0] == a.ldy == 152, nseg == 1
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPairThreads, 1) k_rows_pair(RowsTcArgs a) {
  constexpr int NB = kN, NT = kPairThreads, PP = 1024 / NT, RS = NT / 8;
  SPW_DYN_SMEM(smem_raw);
  float* Bhi_s = reinterpret_cast<float*>(smem_raw);
  float* Blo_s = Bhi_s + kB2Floats;
  float* Xbuf = Blo_s + kB2Floats;                               // [128][152] raw rows of the current tile
  float* slab = Xbuf + kXbufFloats;                              // [2][128][32] epilogue staging
  float* sbias = slab + 2 * kSlabFloats;                         // [160]
  uint32_t* sbits = reinterpret_cast<uint32_t*>(sbias + NB);     // [128][8] sign words of the tile
  uint64_t* mma_done = reinterpret_cast<uint64_t*>(sbits + kTM * 8);
  uint64_t* xfull = mma_done + 1;
  uint64_t* a_ready = xfull + 1;                                 // leader's copy is used: one arrival per CTA and tile
  uint32_t* tptr = reinterpret_cast<uint32_t*>(a_ready + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = 32 * (warp & 3) + lane, part = warp >> 2;      // tensor-memory view: thread = (row, quarter of the k-steps)
  const int crow = tid >> 3, cch = tid & 7;                      // coalesced view (epilogue)
  const uint32_t rank = cluster_ctarank();
  const int cluster = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
  const int ntiles = (a.M + kTM - 1) / kTM;
  const int npairs = (ntiles + 1) >> 1;                          // tile pair p = tiles 2p (rank 0) and 2p + 1 (rank 1)
  const int my_pairs = cluster < npairs ? (npairs - 1 - cluster) / nclusters + 1 : 0;
  auto tile_of = [&](int i) { return 2 * (cluster + i * nclusters) + (int)rank; };

  if (warp == 0) tmem_alloc2(tptr, kTmemCols);
  if (tid == 32) { mbar_init(mma_done, 1); mbar_init(xfull, 1); mbar_init(a_ready, 2); fence_mbar_init(); }
  for (int i = tid; i < kB2Floats / 4; i += NT) {                // this CTA's half of the weight columns
    cp_async16(Bhi_s + 4 * i, a.Bhi + (size_t)rank * kB2Floats + 4 * i);
    cp_async16(Blo_s + 4 * i, a.Blo + (size_t)rank * kB2Floats + 4 * i);
  }
  cp_async_commit();
  for (int i = tid; i < NB; i += NT) sbias[i] = (a.bias && i < a.N) ? a.bias[i] : 0.f;
  for (int i = tid; i < kTM * 8; i += NT) sbits[i] = 0u;
  cp_async_wait<0>();
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tptr;
  const uint32_t lane_addr = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16);
  auto issue_load = [&](int i) {                                 // one thread: bulk copy of the rows of local tile i
    const int tile = tile_of(i);
    const int rows = tile < ntiles ? imin(kTM, a.M - tile * kTM) : 0;
    if (rows > 0) {
      mbar_arrive_expect_tx(xfull, (uint32_t)rows * kDEP * 4);
      bulk_g2s(Xbuf, a.X[0] + (size_t)tile * kTM * kDEP, (uint32_t)rows * kDEP * 4, xfull);
    } else {
      mbar_arrive_cta1(xfull);
    }
  };
  if (tid == 0 && my_pairs > 0) issue_load(0);
  cluster_sync_all();                                            // barriers of both CTAs are initialised before any multicast commit
  bool failed = false;
  SPW_PH_DECL

  for (int i = 0; i < my_pairs; ++i) {
    SPW_PH(7);
    const int tile = tile_of(i);
    const int r0 = tile * kTM;
    const bool valid = r0 + row < a.M;
    rows_bits_load(a, sbits, r0);                                // used after several barriers (epilogue)
    // ---- A operand: this thread's row from the raw tile in shared memory -> tf32 hi / lo -> tensor memory
    if (!mbar_wait(xfull, (uint32_t)i & 1u)) failed = true;
    SPW_PH(0);                                                   // p0: wait for the bulk copy
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      const int ks = part + 4 * j;                               // k-steps part, part + 4, ...
      if (ks < kKS) {                                            // warp-uniform
        const float4 u = *reinterpret_cast<const float4*>(Xbuf + row * kDEP + 8 * ks);
        const float4 v = *reinterpret_cast<const float4*>(Xbuf + row * kDEP + 8 * ks + 4);
        float x[8] = {u.x, u.y, u.z, u.w, v.x, v.y, v.z, v.w};
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (!valid || 8 * ks + e >= a.K[0]) x[e] = 0.f;
        uint32_t h[8], l[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) split_tf32(x[e], h[e], l[e]);
        tmem_st8(lane_addr + kColAhi + 8 * ks, h);
        tmem_st8(lane_addr + kColAlo + 8 * ks, l);
      }
    }
    tmem_wait_st();
    fence_before_sync();
    SPW_PH(1);                                                   // p1: operand staging from shared memory
    __syncthreads();                                             // the raw tile has been consumed
    if (tid == 0) {
      if (i + 1 < my_pairs) { fence_async_smem(); issue_load(i + 1); }   // next tile: lands during the MMAs and the epilogue
      mbar_arrive_remote(a_ready, 0);                            // this CTA: A staged, D of the previous tile read
    }
    SPW_PH(2);                                                   // p2: CTA barrier
    if (rank == 0 && tid == 0) {
      if (!mbar_wait_cluster(a_ready, (uint32_t)i & 1u)) failed = true;   // ... and the other CTA of the pair too
      fence_after_sync();
      const uint32_t idesc = make_idesc_tf32(256, kN);
      const uint64_t dhi0 = make_b_desc(smem_u32(Bhi_s), kHalfN * 16, 128), dlo0 = make_b_desc(smem_u32(Blo_s), kHalfN * 16, 128);
      constexpr uint64_t kStep = (8 * kHalfN * 4) >> 4;
      const uint32_t d = tmem_base + kColD;
      // correction products first, main products last (the tensor core truncates when it accumulates)
#pragma unroll 4
      for (int ks = 0; ks < kKS; ++ks) {
        mma2_tf32_ts(d, tmem_base + kColAlo + 8 * ks, dhi0 + ks * kStep, idesc, ks > 0 ? 1u : 0u);
        mma2_tf32_ts(d, tmem_base + kColAhi + 8 * ks, dlo0 + ks * kStep, idesc, 1u);
      }
#pragma unroll 4
      for (int ks = 0; ks < kKS; ++ks) mma2_tf32_ts(d, tmem_base + kColAhi + 8 * ks, dhi0 + ks * kStep, idesc, 1u);
      mma2_commit_multicast(mma_done);
    }
    if (!mbar_wait(mma_done, (uint32_t)i & 1u)) failed = true;
    fence_after_sync();
    SPW_PH(3);                                                   // p3: MMA issue + wait
    // ---- epilogue, 32 accumulator columns at a time: tensor memory -> slab (row threads) -> epilogue + store (coalesced)
    constexpr int kEs = NB / kSlabCols;
#pragma unroll 1
    for (int es = 0; es < kEs; ++es) {
      float* sb = slab + (es & 1) * kSlabFloats;
      {
        uint32_t vv[8];
        tmem_ld8(lane_addr + kColD + kSlabCols * es + 8 * part, vv);
        tmem_wait_ld();
        *reinterpret_cast<uint4*>(sb + slab_off(row, 2 * part)) = make_uint4(vv[0], vv[1], vv[2], vv[3]);
        *reinterpret_cast<uint4*>(sb + slab_off(row, 2 * part + 1)) = make_uint4(vv[4], vv[5], vv[6], vv[7]);
      }
      __syncthreads();
      if (tile < ntiles) rows_epilogue<NB, PP>(a, sb, sbias, sbits, r0, crow, RS, cch, es);
    }
    fence_before_sync();
    SPW_PH(4);                                                   // p4: epilogue
    __syncthreads();
    if (tile < ntiles) rows_bits_store(a, sbits, r0);
    SPW_PH(5);
  }
  SPW_PH_REPORT(a.mulmode ? "k_rows_pair:enc_bwd" : "k_rows_pair:enc_fwd");
  if (failed && tid == 0) a.poison[0] = __int_as_float(0x7fc00000);
  fence_before_sync();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc2(tmem_base, kTmemCols);
}

}  // namespace tc
}  // namespace spw
#endif  // SPW_EMU
