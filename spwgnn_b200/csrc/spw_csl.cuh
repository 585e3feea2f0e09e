// spw_csl.cuh -- round-2 tensor-core data path: column-slab activations + software-pipelined tcgen05 kernels (sm_100a).
//
// LAYOUT.  Every activation array X[M][C] of the GPU path is stored COLUMN-SLAB major ("CSL"), four columns per slab:
//     element (row, c)  ->  base + (c >> 2) * slab + row * 4 + (c & 3)            (slab = rows of the allocation * 4 floats)
// i.e. [C / 4][M][4].  tcgen05 ties a tile row to a TMEM lane and a TMEM lane to a thread, so the thread that owns row r of
// a 128-row tile both stores the operand (tcgen05.st) and reads the accumulator (tcgen05.ld) of row r.  In CSL the 32 rows
// of a warp, for one group of 4 columns, are 512 CONTIGUOUS bytes: one 128-bit access per lane is a perfectly coalesced,
// full-sector warp access, so row threads load operands and store results straight from / to HBM -- no shared-memory
// transposition, no CTA-wide barriers (round 1 spent ~24k of ~29k cycles per tile there; tools/phase_probe.py).  Gathers
// by node index read 16-byte pieces of [column quad][node][4] tables: the nodes of one tower are adjacent, so a warp's
// gather touches one or two lines.
// Sign bits (relu masks) are stored byte-slab major: u8 [C / 8][M], bit (c & 7) of byte (c >> 3) of the row: a warp's
// store of one column group is 32 consecutive bytes (a full sector), and the thread that owns the columns owns the byte.
//
// PIPELINE.  Tensor memory cannot hold two tiles (A_hi + A_lo + D = 464 of 512 columns at K = 152, N = 160) but it can be
// handed over piece by piece: a tile's MMAs are issued as the 2 nks correction products (A_lo.B_hi, A_hi.B_lo; commit ->
// barC) followed by the nks main products (A_hi.B_hi; commit -> barM) -- the order the accumulation wants anyway.  After
// barC the A_lo columns are dead and the next tile's lo words go in; after barM the next tile's hi words go in, D is
// pulled into registers, and the next tile's MMAs start while the epilogue of the finished tile runs from registers.
// tcgen05.mma issue blocks the issuing thread for the length of the MMA stream (the queue holds ~5 instructions), so a
// 17th warp only waits for "operands stored" (named barrier) and issues.  The weights arrive by TMA bulk copies issued
// by that warp, overlapped with the workers' first operand build.
#pragma once
#ifndef SPW_EMU
#include "spw_tc.cuh"
#include "spw_rows_tc.cuh"

namespace spw {
namespace csl {

using namespace spw::tc;

constexpr int kWorkers = 512;                     // 16 worker warps: thread = (row = TMEM lane, quarter q of the 8-column groups)
constexpr int kThreadsC = kWorkers + 128;         // + a fifth warp group: the MMA issuer warp and three idle warps
// Registers: 20 warps = 5 per scheduler = 96 registers per thread at launch (640 x 96 = 61 440 for the CTA).  The worker warp
// groups raise their allotment with setmaxnreg once the fifth group (which only issues MMAs) has given most of its registers
// back; the CTA's pool is what it was launched with, so 512 x 112 + 128 x 24 = 60 416 fits and 512 x 120 would block forever.
__device__ __forceinline__ void regs_workers() { asm volatile("setmaxnreg.inc.sync.aligned.u32 112;" ::: "memory"); }
__device__ __forceinline__ void regs_issuer() { asm volatile("setmaxnreg.dec.sync.aligned.u32 24;" ::: "memory"); }
constexpr int kBarOps = 1;                        // workers arrive, issuer syncs
constexpr int kBarOpsCount = 512 + 32;            // ... 16 worker warps + the issuer warp
__device__ __forceinline__ void nbar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void nbar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// a view into a CSL array: element (row, c) at p + ((col0 + c) >> 2) * slab + row * 4 + ((col0 + c) & 3); col0 % 4 == 0
struct View {
  float* p; long long slab; int col0;
};
__device__ __forceinline__ float* vaddr(const View& v, long long row, int c) {      // c % 4 == 0: address of a 16-byte quad
  return v.p + (long long)((v.col0 + c) >> 2) * v.slab + row * 4;
}

// ---- L2 prefetch by the otherwise idle warps of the issuer group -------------------------------------------------------------
// The worker warps move in lock-step behind the same mbarriers, so HBM latency is not hidden by other warps: every operand or
// epilogue load that misses L2 stalls the whole SM (ncu: long_scoreboard 6-18 warps per issue).  Warps 17-19 therefore run two
// tiles ahead of the workers and pull the tiles' column-slab pieces (128 rows x 16 bytes = 2 KB per quad) into L2 with
// cp.async.bulk.prefetch.L2: one instruction per piece, no registers, no shared memory; the workers' loads then hit L2.
constexpr int kPfThreads = 96;                    // warps 17, 18, 19
__device__ __forceinline__ void l2_prefetch(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
// rows [r0, r0 + nrows) of the quads [0, nquads) of a column-slab array whose first quad starts at `base`; pl = 0 .. 95
__device__ __forceinline__ void prefetch_quads(const float* base, long long slab, int nquads, long long r0, int nrows, int pl) {
  for (int qd = pl; qd < nquads; qd += kPfThreads) l2_prefetch(base + (long long)qd * slab + r0 * 4, (uint32_t)nrows * 16u);
}
// rows [r0, r0 + 128) of the groups [0, ngroups) of a byte-slab bit array (the arrays are allocated in whole tiles)
__device__ __forceinline__ void prefetch_bits(const uint8_t* base, long long bits_rows, int ngroups, long long r0, int pl) {
  for (int g = pl; g < ngroups; g += kPfThreads) l2_prefetch(base + (long long)g * bits_rows + r0, 128u);
}

// x[0..3] += v at the L2 (one element is touched by exactly one thread per launch and launches are stream-ordered, so the
// sum is the same round-to-nearest fp32 add in the same order as a load / add / store, without the load's latency)
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// ---- weights: TMA bulk copies (issued by one thread) -------------------------------------------------------------------
__device__ __forceinline__ void bulk_load_weights(float* Bhi_s, float* Blo_s, const float* Bhi, const float* Blo, uint32_t bytes, uint64_t* bar) {
  mbar_arrive_expect_tx(bar, 2 * bytes);
  for (uint32_t off = 0; off < bytes; off += 32768) {
    const uint32_t n = bytes - off < 32768 ? bytes - off : 32768;
    bulk_g2s(reinterpret_cast<char*>(Bhi_s) + off, reinterpret_cast<const char*>(Bhi) + off, n, bar);
    bulk_g2s(reinterpret_cast<char*>(Blo_s) + off, reinterpret_cast<const char*>(Blo) + off, n, bar);
  }
}

// issuer side of one tile: nks k-steps, B operands [ks][2][NB][4] in shared memory
__device__ __forceinline__ void issue_tile(uint32_t d_tmem, uint32_t ahi, uint32_t alo, uint32_t bhi, uint32_t blo, int nks, int NB,
                                           uint64_t* barC, uint64_t* barM) {
  const uint32_t idesc = make_idesc_tf32(128, NB);
  const uint32_t step = 8 * NB * 4;
#pragma unroll 1
  for (int ks = 0; ks < nks; ++ks) {
    const uint64_t dhi = make_b_desc(bhi + ks * step, NB * 16, 128);
    const uint64_t dlo = make_b_desc(blo + ks * step, NB * 16, 128);
    mma_tf32_ts(d_tmem, alo + 8 * ks, dhi, idesc, ks > 0 ? 1u : 0u);
    mma_tf32_ts(d_tmem, ahi + 8 * ks, dlo, idesc, 1u);
  }
  mma_commit(barC);
#pragma unroll 1
  for (int ks = 0; ks < nks; ++ks) {
    const uint64_t dhi = make_b_desc(bhi + ks * step, NB * 16, 128);
    mma_tf32_ts(d_tmem, ahi + 8 * ks, dhi, idesc, 1u);
  }
  mma_commit(barM);
}

// round-to-nearest (ties away) tf32 of a finite float with integer arithmetic: what cvt.rna.tf32.f32 returns, in 2 instructions
// instead of the ~5 ptxas emits for the general (NaN / Inf aware) conversion; the split is the hot ALU work of every tile
__device__ __forceinline__ uint32_t rna_tf32(float x) { return (__float_as_uint(x) + 0x1000u) & 0xffffe000u; }
// lo = x - hi is exact in fp32; the tensor core reads only the top 19 bits of an operand word, i.e. it TRUNCATES lo to tf32:
// an error of at most 2^-10 |lo| <= 2^-21 |x| with the (random) sign of lo -- the same order as the lo.lo product the 3xTF32
// scheme drops anyway -- for one instruction instead of three.
__device__ __forceinline__ void split_fast(float x, uint32_t& hi, uint32_t& lo) {
  hi = rna_tf32(x);
  lo = __float_as_uint(x - __uint_as_float(hi));
}

// operand registers of a worker thread: its k-steps q, q + 4, ... (KJ of them), 8 floats each
template <int KJ>
struct XR { float v[KJ][8]; };

template <int KJ>
__device__ __forceinline__ void store_lo(const XR<KJ>& x, uint32_t lane_addr, uint32_t colLo, int q, int nks) {
#pragma unroll
  for (int j = 0; j < KJ; ++j) {
    const int ks = q + 4 * j;
    if (ks < nks) {                                  // warp-uniform
      uint32_t l[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) { uint32_t h; split_fast(x.v[j][i], h, l[i]); }
      tmem_st8(lane_addr + colLo + 8 * ks, l);
    }
  }
}
template <int KJ>
__device__ __forceinline__ void store_hi(const XR<KJ>& x, uint32_t lane_addr, uint32_t colHi, int q, int nks) {
#pragma unroll
  for (int j = 0; j < KJ; ++j) {
    const int ks = q + 4 * j;
    if (ks < nks) {
      uint32_t h[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) h[i] = rna_tf32(x.v[j][i]);
      tmem_st8(lane_addr + colHi + 8 * ks, h);
    }
  }
}
// accumulator groups of a worker thread: 8-column groups q, q + 4, ... (GJ of them)
template <int GJ>
__device__ __forceinline__ void load_d(uint32_t (&d)[GJ][8], uint32_t lane_addr, uint32_t colD, int q, int ngroups) {
#pragma unroll
  for (int j = 0; j < GJ; ++j) {
    const int g = q + 4 * j;
    if (g < ngroups) tmem_ld8(lane_addr + colD + 8 * g, d[j]);
  }
  tmem_wait_ld();
}

// =========================================================================================================================
// k_lin: one fused linear layer on CSL rows,  Y = post( act( X . W + rowscale * bias + addend ) ), the LinArgs contract of
// round 1 (spw_rows_tc.cuh) minus the second row segment (concatenated inputs live in ONE array: [g | p] is GP[n][200]).
//   KJ = ceil(nks / 4) <= 7 (K <= 224),  NB = 112 or 160,  16 nks + NB <= 512.
// The epilogue options are a COMPILE-TIME mask: with every option compiled in, the kernel was ~5000 SASS instructions,
// most of them address and predicate arithmetic, and the epilogue alone took 8k cycles per tile (tools/phase_lin.py).
// =========================================================================================================================
enum : uint32_t {
  EPI_BIAS = 1u, EPI_ROWSCALE = 2u, EPI_ADD = 4u, EPI_RELU = 8u, EPI_TANH = 16u, EPI_MUL_POS = 32u /* *= [mulsrc > 0] */,
  EPI_MUL_TANH = 64u /* *= 1 - mulsrc^2 */, EPI_MUL_BITS = 128u /* *= sign bit */, EPI_DROP = 256u, EPI_SCALE = 512u,
  EPI_ACC = 1024u, EPI_BITS_OUT = 2048u, EPI_ONES = 4096u
};

struct LinCArgs {
  int M, K, nks, N;                              // rows, valid input columns (a multiple of 4 columns is read), k-steps of 8, valid output columns
  View X;
  const float* Bhi; const float* Blo;            // packed operands [nks][2][NB][4] (k_pack_tc)
  const float* bias;                             // [N]                                  (EPI_BIAS)
  const float* rowscale;                         // [M] multiplier of the bias           (EPI_ROWSCALE)
  View addend;                                   //                                      (EPI_ADD)
  View mulsrc;                                   //                                      (EPI_MUL_POS / EPI_MUL_TANH)
  const uint8_t* bits_in; long long bits_in_rows;    // byte-slab sign bits [C / 8][rows] (EPI_MUL_BITS)
  uint8_t* bits_out; long long bits_out_rows;        // sign bits of the result, same layout (EPI_BITS_OUT)
  View Y;
  float post_scale;                              //                                      (EPI_SCALE)
  uint32_t drop_thresh, drop_seed; float drop_inv_keep; int drop_stride;             // (EPI_DROP)
  int ones_col;                                  // >= N: Y[row][ones_col] = 1           (EPI_ONES)
  int write_pad;                                 // 1: the columns N .. round_up(N, 8) - 1 are written (0 / ones_col); 0: a quad beyond N is never touched
  float* poison;
};

constexpr size_t lin_smem(int NB, int nks) { return (size_t)(2 * nks * 8 * NB + NB) * sizeof(float) + 64; }

// N (valid output columns: 100 or 150) and K (valid input columns: 100, 150 or 200) are compile-time as well, so that every
// column predicate folds away; a.N / a.K / a.nks must match (launch_lin checks).
template <int N, int K, uint32_t EPI>
__global__ void __launch_bounds__(kThreadsC, 1) k_lin(LinCArgs a) {
  constexpr int NB = N <= 112 ? 112 : 160;                       // MMA N
  constexpr int NKS = (K + 7) / 8, KJ = (NKS + 3) / 4;           // k-steps, k-steps per thread
  constexpr int GJ = (NB / 8 + 3) / 4;                           // accumulator groups per thread (5 or 4)
  static_assert(16 * NKS + NB <= 512, "operand + accumulator exceed tensor memory");
  SPW_DYN_SMEM(smem_raw);
  constexpr int bfl = NKS * 8 * NB;
  float* Bhi_s = reinterpret_cast<float*>(smem_raw);
  float* Blo_s = Bhi_s + bfl;
  float* sbias = Blo_s + bfl;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sbias + NB);
  uint64_t* barC = bars; uint64_t* barM = bars + 1; uint64_t* barW = bars + 2;
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bars + 3);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = 32 * (warp & 3) + lane, q = warp >> 2;

  pdl_trigger();                                                 // spw_common.cuh: the next kernel may set up behind this one
  if (warp == 0) tmem_alloc(tptr, kTmemCols);
  if (tid == 32) { mbar_init(barC, 1); mbar_init(barM, 1); mbar_init(barW, 1); fence_mbar_init(); }
  for (int i = tid; i < NB; i += kThreadsC) sbias[i] = ((EPI & EPI_BIAS) && i < N) ? a.bias[i] : 0.f;
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tptr;
  constexpr uint32_t colHi = 0, colLo = 8 * NKS, colD = kTmemCols - NB;
  const int ntiles = (a.M + kTM - 1) / kTM;
  const int cnt = (int)blockIdx.x < ntiles ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (warp >= kWorkers / 32) {
    // ---------------- MMA issuer warp (and its three idle siblings) ----------------
    regs_issuer();
    if (warp == kWorkers / 32) {
    if (lane == 0) bulk_load_weights(Bhi_s, Blo_s, a.Bhi, a.Blo, (uint32_t)bfl * 4, barW);
    bool ok = mbar_wait(barW, 0);
    for (int i = 0; i < cnt; ++i) {
      nbar_sync(kBarOps, kBarOpsCount);
      fence_after_sync();
      if (lane == 0) issue_tile(tmem_base + colD, tmem_base + colHi, tmem_base + colLo, smem_u32(Bhi_s), smem_u32(Blo_s), NKS, NB, barC, barM);
      __syncwarp();
    }
    if (!ok && lane == 0) a.poison[0] = __int_as_float(0x7fc00000);
    } else {
      // ---------------- L2 prefetch warps: tile i + 2 while the workers are in tile i ----------------
      const int pl = tid - (kWorkers + 32);
      pdl_wait();
      auto prefetch_tile = [&](int i) {
        const long long r0 = (long long)(blockIdx.x + i * gridDim.x) * kTM;
        const int nrows = a.M - r0 >= kTM ? kTM : (int)(a.M - r0);
        prefetch_quads(a.X.p + (long long)(a.X.col0 >> 2) * a.X.slab, a.X.slab, (K + 3) >> 2, r0, nrows, pl);
        if (EPI & EPI_ADD) prefetch_quads(a.addend.p + (long long)(a.addend.col0 >> 2) * a.addend.slab, a.addend.slab, (N + 3) >> 2, r0, nrows, pl);
        if (EPI & (EPI_MUL_POS | EPI_MUL_TANH)) prefetch_quads(a.mulsrc.p + (long long)(a.mulsrc.col0 >> 2) * a.mulsrc.slab, a.mulsrc.slab, (N + 3) >> 2, r0, nrows, pl);
        if (EPI & EPI_ACC) prefetch_quads(a.Y.p + (long long)(a.Y.col0 >> 2) * a.Y.slab, a.Y.slab, (N + 3) >> 2, r0, nrows, pl);
        if (EPI & EPI_MUL_BITS) prefetch_bits(a.bits_in, a.bits_in_rows, (N + 7) >> 3, r0, pl);
      };
      if (cnt > 1) prefetch_tile(1);
      for (int i = 0; i + 2 < cnt; ++i) {
        mbar_wait(barC, (uint32_t)i & 1u);
        prefetch_tile(i + 2);
      }
    }
  } else {
    // ---------------- worker warps ----------------
    regs_workers();
    const uint32_t lane_addr = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16);
    constexpr int ngroups = NB / 8;
    constexpr int nq8 = (N + 7) >> 3;                            // output column groups that exist
    bool failed = false;
    XR<KJ> x;
    SPW_PH_DECL
    // the quads of this thread: input k-step q + 4 j = quads (col0 >> 2) + 2 q + 8 j + {0, 1}; same pattern for the outputs
    const long long xs = a.X.slab, ys = a.Y.slab;
    const float* xq = a.X.p + (long long)((a.X.col0 >> 2) + 2 * q) * xs;
    float* yq = a.Y.p + (long long)((a.Y.col0 >> 2) + 2 * q) * ys;
    auto load_x = [&](int i) {                                   // operand rows of local tile i: coalesced 128-bit loads
      const long long r = (long long)(blockIdx.x + i * gridDim.x) * kTM + row;
      const bool rv = r < a.M;
      const float* xp = xq + (rv ? r : 0) * 4;
#pragma unroll
      for (int j = 0; j < KJ; ++j) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
          if (8 * (q + 4 * j) + 4 * h < K && rv) t = *reinterpret_cast<const float4*>(xp + (long long)(8 * j + h) * xs);
          x.v[j][4 * h] = t.x; x.v[j][4 * h + 1] = t.y; x.v[j][4 * h + 2] = t.z; x.v[j][4 * h + 3] = t.w;
        }
      }
    };
    pdl_wait();                                                  // the previous kernel's results are complete and visible
    if (cnt > 0) {
      load_x(0);
      store_lo<KJ>(x, lane_addr, colLo, q, NKS);
      store_hi<KJ>(x, lane_addr, colHi, q, NKS);
      tmem_wait_st();
      fence_before_sync();
      nbar_arrive(kBarOps, kBarOpsCount);
    }
    for (int i = 0; i < cnt; ++i) {
      const bool has_next = i + 1 < cnt;
      const uint32_t parity = (uint32_t)i & 1u;
      SPW_PH(7);
      if (has_next) load_x(i + 1);                               // in flight under the MMAs of tile i
      uint32_t bin[GJ];                                          // relu bits of tile i's epilogue, requested well before their use
      if (EPI & EPI_MUL_BITS) {
        const long long rb = (long long)(blockIdx.x + i * gridDim.x) * kTM + row;
#pragma unroll
        for (int j = 0; j < GJ; ++j) {
          const int g = q + 4 * j;
          bin[j] = (g < ngroups && g < nq8 && rb < a.M) ? a.bits_in[(long long)g * a.bits_in_rows + rb] : 0u;
        }
      }
      SPW_PH(0);                                                 // p0: issue of the operand loads
      if (!mbar_wait(barC, parity)) failed = true;               // corrections done: A_lo free
      fence_after_sync();
      SPW_PH(1);                                                 // p1: wait for the correction MMAs
      if (has_next) store_lo<KJ>(x, lane_addr, colLo, q, NKS);
      SPW_PH(2);                                                 // p2: lo words (first use of the loaded operand)
      if (!mbar_wait(barM, parity)) failed = true;               // tile done: A_hi free, D complete
      fence_after_sync();
      SPW_PH(3);                                                 // p3: wait for the main MMAs
      if (has_next) store_hi<KJ>(x, lane_addr, colHi, q, NKS);
      uint32_t d[GJ][8];
      load_d<GJ>(d, lane_addr, colD, q, ngroups);
      if (has_next) {
        tmem_wait_st();
        fence_before_sync();
        nbar_arrive(kBarOps, kBarOpsCount);                         // the issuer starts tile i + 1
      }
      SPW_PH(4);                                                 // p4: hi words + D load: the tensor pipe idles
      // ---- epilogue of tile i from registers, under the MMAs of tile i + 1 ----
      const long long r = (long long)(blockIdx.x + i * gridDim.x) * kTM + row;
      if (r < a.M) {
        float rs = 1.f;
        if (EPI & EPI_ROWSCALE) rs = a.rowscale[r];
        float* yp = yq + r * 4;
        const float* ap = nullptr; const float* mp = nullptr;
        if (EPI & EPI_ADD) ap = a.addend.p + (long long)((a.addend.col0 >> 2) + 2 * q) * a.addend.slab + r * 4;
        if (EPI & (EPI_MUL_POS | EPI_MUL_TANH)) mp = a.mulsrc.p + (long long)((a.mulsrc.col0 >> 2) + 2 * q) * a.mulsrc.slab + r * 4;
        // the epilogue's own loads (addend, multiplier source, old Y of an accumulation) of group j + 1 are issued BEFORE group j is
        // stored: the stores may alias them as far as the compiler knows, so left inside the loop every group paid a full L2 latency
        constexpr bool kEpiLoads = (EPI & (EPI_ADD | EPI_MUL_POS | EPI_MUL_TANH | EPI_ACC)) != 0;
        float4 ta[2][2], tm[2][2], ty[2][2];                       // [stage][quad of the group]
        auto epi_issue = [&](int j, int sg) {
          const int g = q + 4 * j;
          if (!kEpiLoads || g >= ngroups || g >= nq8) return;      // warp-uniform
          const bool h1 = 8 * g + 4 < N;
          const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
          if (EPI & EPI_ADD) {
            ta[sg][0] = *reinterpret_cast<const float4*>(ap + (long long)(8 * j) * a.addend.slab);
            ta[sg][1] = h1 ? *reinterpret_cast<const float4*>(ap + (long long)(8 * j + 1) * a.addend.slab) : z;
          }
          if (EPI & (EPI_MUL_POS | EPI_MUL_TANH)) {
            tm[sg][0] = *reinterpret_cast<const float4*>(mp + (long long)(8 * j) * a.mulsrc.slab);
            tm[sg][1] = h1 ? *reinterpret_cast<const float4*>(mp + (long long)(8 * j + 1) * a.mulsrc.slab) : z;
          }
          if (EPI & EPI_ACC) {
            ty[sg][0] = *reinterpret_cast<const float4*>(yp + (long long)(8 * j) * ys);
            ty[sg][1] = h1 ? *reinterpret_cast<const float4*>(yp + (long long)(8 * j + 1) * ys) : z;
          }
        };
        epi_issue(0, 0);
#pragma unroll
        for (int j = 0; j < GJ; ++j) {
          const int g = q + 4 * j;
          if (j + 1 < GJ) epi_issue(j + 1, (j + 1) & 1);
          if (g >= ngroups || g >= nq8) continue;                  // warp-uniform
          const int col = 8 * g, sg = j & 1;
          const bool h1 = col + 4 < N;                             // the second quad holds valid columns
          float v[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(d[j][e]);
          if (EPI & EPI_BIAS) {
            const float4 b0 = *reinterpret_cast<const float4*>(sbias + col), b1 = *reinterpret_cast<const float4*>(sbias + col + 4);
            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = (EPI & EPI_ROWSCALE) ? fmaf(rs, bb[e], v[e]) : v[e] + bb[e];
          }
          if (EPI & EPI_ADD) {
            const float4 t0 = ta[sg][0], t1 = ta[sg][1];
            v[0] += t0.x; v[1] += t0.y; v[2] += t0.z; v[3] += t0.w;
            if (h1) { v[4] += t1.x; v[5] += t1.y; v[6] += t1.z; v[7] += t1.w; }
          }
          if (EPI & EPI_RELU) {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = relu_f(v[e]);
          }
          if (EPI & EPI_TANH) {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = tanhf(v[e]);
          }
          if (EPI & EPI_MUL_BITS) {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = ((bin[j] >> e) & 1u) ? v[e] : 0.f;
          }
          if (EPI & (EPI_MUL_POS | EPI_MUL_TANH)) {
            const float4 t0 = tm[sg][0], t1 = tm[sg][1];
            const float m[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = (EPI & EPI_MUL_POS) ? (m[e] > 0.f ? v[e] : 0.f) : v[e] * (1.f - m[e] * m[e]);
          }
          if (EPI & EPI_DROP) {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = dropout_apply(v[e], a.drop_seed, (uint32_t)(r * a.drop_stride + col + e), a.drop_thresh, a.drop_inv_keep);
          }
          if (EPI & EPI_SCALE) {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] *= a.post_scale;
          }
          if (EPI & EPI_ACC) {
            const float4 t0 = ty[sg][0], t1 = ty[sg][1];
            v[0] += t0.x; v[1] += t0.y; v[2] += t0.z; v[3] += t0.w;
            if (h1) { v[4] += t1.x; v[5] += t1.y; v[6] += t1.z; v[7] += t1.w; }
          }
          if (col + 8 > N) {                                       // warp-uniform: the group straddles N
#pragma unroll
            for (int e = 0; e < 8; ++e)
              if (col + e >= N) v[e] = ((EPI & EPI_ONES) && col + e == a.ones_col) ? 1.f : 0.f;
          }
          if (EPI & EPI_BITS_OUT) {
            uint32_t b = 0u;
#pragma unroll
            for (int e = 0; e < 8; ++e) b |= (v[e] > 0.f) ? (1u << e) : 0u;
            if (col + 8 > N) b &= (1u << (N - col)) - 1u;          // the ones column is not an activation
            a.bits_out[(long long)g * a.bits_out_rows + r] = (uint8_t)b;
          }
          *reinterpret_cast<float4*>(yp + (long long)(8 * j) * ys) = make_float4(v[0], v[1], v[2], v[3]);
          if (h1 || a.write_pad)
            *reinterpret_cast<float4*>(yp + (long long)(8 * j + 1) * ys) = make_float4(v[4], v[5], v[6], v[7]);
        }
      }
      SPW_PH(5);                                                 // p5: epilogue
    }
#ifdef SPW_PHASE_TIMING
    if (a.M > 100000 && blockIdx.x == 0 && lane == 0)
      printf("k_lin warp %2d: loads %lld waitC %lld lo %lld waitM %lld hi+D %lld epi %lld loop %lld\n", warp, ph_t[0], ph_t[1], ph_t[2], ph_t[3], ph_t[4], ph_t[5], ph_t[7]);
#endif
    if (failed && tid == 0) a.poison[0] = __int_as_float(0x7fc00000);
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace csl
}  // namespace spw
#endif  // SPW_EMU
