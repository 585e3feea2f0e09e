// spw_csl_wgrad.cuh -- weight gradients on the tensor cores, column-slab inputs, software-pipelined (sm_100a only).
//
//   dW[k][n] (+)= sum_rows X[row][k] * dY[row][n],  k <= Kx: row Kx of dW is the bias gradient, picked up by a virtual ones
//   (or rowscale) feature.  The contraction runs over rows, so  A = X^T : TMEM lane = feature, TMEM column = row of a 32-row
//   chunk;  B = dY^T chunk in shared memory, K-major: [k-step of 8 rows][2][n][4 rows];  D[feature][n] in tensor memory.
//
// Round 1 ran chunk staging -> operand build -> MMA in sequence, both M-tiles (features 0..127 and the overlapping last 128) in
// one CTA with 80 running-sum registers per thread.  Here:
//   * a CTA owns ONE M-tile (b % nmt) of one row stream (b / nmt): 40 running-sum registers per thread, half the build work;
//   * operands are double-buffered (A: 2 x 64 TMEM columns, B: 2 x 40 KB shared memory) and a 17th warp issues the MMAs, so the
//     build of chunk q + 1 overlaps the MMAs of chunk q; completion of chunk q - 2's MMAs (tcgen05.commit -> barF[q & 1])
//     frees the buffers of chunk q;
//   * D is double-buffered per 128-row tile: the tensor core truncates when it accumulates, so a tile's sum is added to the
//     running sum in registers with round-to-nearest adds, one tile late, while the next tile accumulates in the other buffer;
//   * raw chunk rows arrive in a ring of stages.  Streamed column-slab arrays ([quad][row][4]) are described to the TMA as 2-D
//     tensors [quad][rows * 4 floats]: ONE cp.async.bulk.tensor per array and chunk lands a [quads][33 rows x 4] box (pitch 132
//     floats: bank-conflict-free for both operand builds; rows past the end of the array are zero-filled by the TMA),
//     completion counted in bytes on the stage's mbarrier.  Measured alternatives: a warp-wide cp.async costs the LSU ~20
//     cycles whatever it copies (1.4k cycles per chunk for the 71 quads of a plain layer, 3.8k when h1 = relu(A + S[snd] +
//     R[rcv]) was re-gathered); one plain bulk copy per 512-byte quad piece costs ~55 cycles of a serialised issue path (3.9k).
//     So the forward edge step now stores h1, and the only gather left -- d h2 = relu'(h2) * dH2S[receiver], 16 bytes per row
//     and quad -- uses cp.async.
#pragma once
#ifndef SPW_EMU
#include <cuda.h>
#include <stdio.h>
#include "spw_csl.cuh"

namespace spw {
namespace csl {

constexpr int kWgCh = 32;                         // rows per chunk
constexpr int kQPitch = 132;                      // floats per staged quad ([32 rows][4] + 4: bank-conflict-free both ways)
constexpr int kBarWork = 2;                       // named barrier of the 512 workers
constexpr uint32_t kWgColD = 0, kWgColA = 320;    // TMEM: D0 [0,160) D1 [160,320) | A buffers: hi [320 + 64 b, +32) lo [+32, +64)

struct WgradCArgs {
  int M;
  int Kx;                                                              // X features (the X array itself: tensor map tmX)
  const float* rowscale; int rsmod;                                    // value of the virtual feature Kx (null: 1)
  const int32_t* rcv;                                                  // YMODE 1: receiver of each row
  const float* dY; long long y_slab; int y_col0; int Ny;               // YMODE 0: tensor map tmY; YMODE 1: node table gathered by rcv, masked by bits
  const uint8_t* bits; long long bits_rows;                            // YMODE 1: byte-slab relu bits [19][rows]
  int nmt;                                                             // M-tiles: 1 (Kx + 1 <= 128) or 2
  float* part;                                                         // [gridDim.x / nmt][2][160][128]
  int first;                                                           // this launch initialises the partials (else it adds to them)
  int gather_tma;                                                      // k_wgrad_pair, YMODE 1: node-range tensor copies allowed (0: always cp.async)
  float* poison;
};

// stage layout (floats): X quads [nqx][132] | Y quads [nqy][132] | RS [32] | bits [20][32] bytes | RL [32] ints; each region 128-byte aligned
__host__ __device__ constexpr int wg_up32(int f) { return (f + 31) & ~31; }
__host__ __device__ constexpr int wg_stage_floats(int nqx, int nqy) { return wg_up32(nqx * kQPitch) + wg_up32(nqy * kQPitch) + 32 + 160 + 32; }

// tile-mode TMA load of a 2-D box: coordinates {c0 (innermost: floats along the rows), c1 (quad)}
__device__ __forceinline__ void tma_load_2d(void* sdst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(sdst)),
               "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
               : "memory");
}
constexpr size_t wgrad_c_smem(int nqx, int nqy, int NB, int nst) {
  return (size_t)(nst * wg_stage_floats(nqx, nqy) + 2 * 2 * (kWgCh / 8) * (2 * NB * 4) + 8) * sizeof(float) + 128 + 8 * nst;
}

#ifdef SPW_WAIT_DEBUG
#define SPW_WDBG(what, q) do { if ((threadIdx.x & 31) == 0) printf("k_wgrad_pair: block %d warp %d: wait %s timed out at chunk %d\n", (int)blockIdx.x, (int)(threadIdx.x >> 5), what, (int)(q)); } while (0)
#else
#define SPW_WDBG(what, q) do { } while (0)
#endif
constexpr size_t wgrad_pair_smem(int nqx, int NB, int nst) { return wgrad_c_smem(nqx, NB / 8, NB / 2, nst); }

// The per-chunk work of a thread is a fixed pattern; everything that does not depend on the chunk (shared-memory offsets of
// the words it reads and writes, which lanes are real / ones / padding features) is computed once, and padding is expressed as
// "read a zero word with stride 0" instead of predicates: the kernel was instruction-issue bound (ncu: 61 % issue-active).
// tmX / tmY: 2-D tensor maps [quads][rows * 4] of the X and dY arrays (base = first row and first quad of the view), boxes of
// [nqx_box][132] and [nqy][132] floats (launch code: run_wgrad_c).
// local arrive (release at CTA scope) and "arrive when this thread's earlier cp.async copies have landed" (.noinc: counted in the
// barrier's expected arrivals like an ordinary arrive)
__device__ __forceinline__ void mbar_arrive_local(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void regs_issuer32() { asm volatile("setmaxnreg.dec.sync.aligned.u32 32;" ::: "memory"); }

template <int YMODE, int NB, int NST>
__global__ void __launch_bounds__(kThreadsC, 1) k_wgrad_c(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY, WgradCArgs a, int nqx) {
  SPW_DYN_SMEM(smem_raw);
  constexpr int kUnits = ((kWgCh / 4) * NB + kWorkers - 1) / kWorkers;      // B-operand units (4 rows x 1 column) per thread
  constexpr int bfl = (kWgCh / 8) * (2 * NB * 4);                         // floats per hi or lo B operand of a chunk
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mt = (int)blockIdx.x % a.nmt, stream = (int)blockIdx.x / a.nmt, nstreams = (int)gridDim.x / a.nmt;
  const int f0 = mt == 0 ? 0 : a.Kx + 1 - 128;                   // first feature of this CTA's M-tile
  const int qlo = f0 >> 2;                                       // staged X quads: [qlo, qlo + nqx)
  const int fhi = mt == 0 ? (a.Kx < 128 ? a.Kx : 128) : a.Kx;    // features [f0, fhi) are read from X
  const int nqy = (a.Ny + 3) >> 2;                               // nqx (kernel argument): quads of the X box, the same for both M-tiles
  const int stf = wg_stage_floats(nqx, nqy);
  float* stages = reinterpret_cast<float*>(smem_raw);
  float* Bop = stages + NST * stf;                               // [2 buffers][hi | lo][bfl]
  float* zero = Bop + 4 * bfl;                                   // 8 zero words
  uint64_t* bars = reinterpret_cast<uint64_t*>(zero + 8);
  uint64_t* barF = bars; uint64_t* barT = bars + 2; uint64_t* barS = bars + 4;      // barS[NST]: stage filled (tensor copies + the producer's lanes)
  uint64_t* barE = barS + NST;                                                      // barE[NST]: stage consumed (one arrival per worker warp)
  // barR[2]: the 16 worker warps have built operand buffer `buf`.  A phase-tracked mbarrier, not the bar.arrive / bar.sync pair of the
  // pipelined kernels: without a CTA barrier per chunk a fast warp reaches chunk q + 1 (whose buffer is free once the MMAs of chunk
  // q - 1 are done) while a slow one still builds chunk q, and a counting barrier would take its arrival for the slow warp's
  uint64_t* barR = barE + NST;
  uint32_t* tptr = reinterpret_cast<uint32_t*>(barR + 2);

  pdl_trigger();
  if (warp == 0) tmem_alloc(tptr, kTmemCols);
  if (tid == 32) {
    mbar_init(barF, 1); mbar_init(barF + 1, 1); mbar_init(barT, 1); mbar_init(barT + 1, 1);
    for (int i = 0; i < NST; ++i) { mbar_init(barS + i, 1 + 2 * 32); mbar_init(barE + i, kWorkers / 32); }
    mbar_init(barR, kWorkers / 32); mbar_init(barR + 1, kWorkers / 32);
    fence_mbar_init();
  }
  if (tid < 8) zero[tid] = 0.f;
  if (!a.rowscale)                                               // the virtual feature Kx is a column of ones
    for (int i = tid; i < NST * 32; i += kThreadsC) stages[(i >> 5) * stf + wg_up32(nqx * kQPitch) + wg_up32(nqy * kQPitch) + (i & 31)] = 1.f;
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tptr;
  const int ntiles = (a.M + kTM - 1) / kTM;
  const int my_tiles = stream < ntiles ? (ntiles - 1 - stream) / nstreams + 1 : 0;
  constexpr int kCh = kTM / kWgCh;                               // chunks per tile
  const int nq = my_tiles * kCh;
  auto row0_of = [&](int q) { return (long long)(stream + (q / kCh) * nstreams) * kTM + (q % kCh) * kWgCh; };

  if (warp >= kWorkers / 32) {
    // ---------------- MMA issuer warp (and its three idle siblings) ----------------
    regs_issuer32();
    constexpr uint32_t idesc = make_idesc_tf32(128, NB);
    bool iss_ok = true;
    if (warp == kWorkers / 32)
    for (int q = 0; q < nq; ++q) {
      if (!mbar_wait(barR + (q & 1), (uint32_t)(q >> 1) & 1u)) iss_ok = false;   // the operands of chunk q are in tensor / shared memory
      fence_after_sync();
      if (lane == 0) {
        const int t = q / kCh, c = q % kCh, buf = q & 1;
        const uint32_t d = tmem_base + kWgColD + 160 * (t & 1);
        const uint32_t ahi = tmem_base + kWgColA + 64 * buf, alo = ahi + 32;
        const uint32_t bhi = smem_u32(Bop + (size_t)(2 * buf) * bfl), blo = smem_u32(Bop + (size_t)(2 * buf + 1) * bfl);
#pragma unroll 1
        for (int ks = 0; ks < kWgCh / 8; ++ks) {
          const uint64_t dhi = make_b_desc(bhi + ks * (2 * NB * 16), NB * 16, 128);
          const uint64_t dlo = make_b_desc(blo + ks * (2 * NB * 16), NB * 16, 128);
          mma_tf32_ts(d, alo + 8 * ks, dhi, idesc, (c > 0 || ks > 0) ? 1u : 0u);
          mma_tf32_ts(d, ahi + 8 * ks, dlo, idesc, 1u);
          mma_tf32_ts(d, ahi + 8 * ks, dhi, idesc, 1u);
        }
        mma_commit(barF + buf);
        if (c == kCh - 1) mma_commit(barT + (t & 1));
      }
      __syncwarp();
    }
    if (!iss_ok && lane == 0) a.poison[0] = __int_as_float(0x7fc00000);
    if (warp == kWorkers / 32 + 1) {
      // ---------------- producer warp: fills the stage ring, up to NST chunks ahead of the workers (see k_wgrad_pair) ----------------
      static_assert(YMODE == 0, "k_wgrad_c: streamed dY only (the gathered form is k_wgrad_pair's)");
      const int y_off = wg_up32(nqx * kQPitch);
      const int rs_off = y_off + wg_up32(nqy * kQPitch);
      bool ok = true;
      pdl_wait();
      int slot = 0;
      for (int qi = 0; qi < nq; ++qi) {
        if (qi >= NST && !mbar_wait(barE + slot, (uint32_t)(qi / NST - 1) & 1u)) ok = false;      // the workers are done with chunk qi - NST
        float* st = stages + slot * stf;
        uint64_t* bs = barS + slot;
        const long long r0 = row0_of(qi);
        if (lane == 0) {                                         // one tensor copy per array (rows past the end read as zero)
          mbar_arrive_expect_tx(bs, (uint32_t)((nqx + nqy) * kQPitch * 4));
          tma_load_2d(st, &tmX, (int)(r0 * 4), qlo, bs);
          tma_load_2d(st + y_off, &tmY, (int)(r0 * 4), 0, bs);
        }
        if (a.rowscale) {
          const long long r = r0 + lane;
          const bool valid = r < a.M;
          cp_async4_zfill(st + rs_off + lane, a.rowscale + (valid ? (a.rsmod ? r % a.rsmod : r) : 0), valid);
        }
        cp_async_mbar_arrive_noinc(bs);
        mbar_arrive_local(bs);
        if (++slot == NST) slot = 0;
      }
      if (!ok && lane == 0) a.poison[0] = __int_as_float(0x7fc00000);
    }
  } else {
    // ---------------- worker warps ----------------
    regs_workers();
    const int L = 32 * (warp & 3) + lane, sub = warp >> 2;      // TMEM lane = feature f0 + L; rows 8 sub .. 8 sub + 7 of a chunk
    const uint32_t lane_addr = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16);
    constexpr int ngroups = NB / 8;
    const int y_off = wg_up32(nqx * kQPitch);
    const int rs_off = y_off + wg_up32(nqy * kQPitch);           // RS (virtual feature values) inside a stage
    // A operand: word offset (within a stage) and row stride of this lane's feature; padding lanes read the zero word
    const int feat = f0 + L;
    int xa_off, xa_str;
    if (feat < fhi) { xa_off = ((feat >> 2) - qlo) * kQPitch + (feat & 3) + 32 * sub; xa_str = 4; }
    else if (feat == a.Kx) { xa_off = rs_off + 8 * sub; xa_str = 1; }
    else { xa_off = -1; xa_str = 0; }
    // B operand units of this thread: column n, row quad kc  ->  source word, bit word, destination
    int yo[kUnits], bo[kUnits], ysh[kUnits], ystr[kUnits];
#pragma unroll
    for (int u = 0; u < kUnits; ++u) {
      const int idx = tid + u * kWorkers;
      const int n = idx % NB, kc = idx / NB;
      if (idx < (kWgCh / 4) * NB && n < a.Ny) { yo[u] = y_off + (n >> 2) * kQPitch + 16 * kc + (n & 3); ystr[u] = 4; }
      else { yo[u] = -1; ystr[u] = 0; }
      bo[u] = (rs_off + 32) * 4 + (n >> 3) * 32 + 4 * kc;          // byte offset of the 4 rows' relu bytes inside a stage
      ysh[u] = n & 7;
    }
    bool failed = false;
    float acc[5][8];
#pragma unroll
    for (int j = 0; j < 5; ++j)
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[j][k] = 0.f;
    SPW_PH_DECL

    // q = 0 .. nq - 1: chunk q;  q = nq: flush of the last tile only
    pdl_wait();
    int sq = 0; uint32_t sph = 0;                                // ring slot of chunk q and the phase of its barriers
#pragma unroll 1
    for (int q = 0; q <= nq; ++q) {
      SPW_PH(7);
      if (q < nq && !mbar_wait(barS + sq, sph)) failed = true;   // the stage of chunk q is filled
      SPW_PH(0);
      SPW_PH(1);
      const int t = q / kCh, c = q % kCh, buf = q & 1;
      if ((c == 1 && t >= 1) || q == nq) {                       // D of the previous tile -> running sums (round-to-nearest adds)
        const int tf = q == nq ? my_tiles - 1 : t - 1;
        if (tf >= 0) {
          if (!mbar_wait(barT + (tf & 1), (uint32_t)(tf >> 1) & 1u)) failed = true;
          fence_after_sync();
          uint32_t d[5][8];
          load_d<5>(d, lane_addr, kWgColD + 160 * (tf & 1), sub, ngroups);
#pragma unroll
          for (int j = 0; j < 5; ++j)
#pragma unroll
            for (int k = 0; k < 8; ++k)
              if (sub + 4 * j < ngroups) acc[j][k] += __uint_as_float(d[j][k]);
          fence_before_sync();
        }
      }
      SPW_PH(2);
      if (q == nq) break;
      if (q >= 2) {                                              // operand buffers `buf` are free once chunk q - 2's MMAs are done
        if (!mbar_wait(barF + buf, (uint32_t)((q >> 1) - 1) & 1u)) failed = true;
        fence_after_sync();
      }
      SPW_PH(3);
      const float* st = stages + sq * stf;
      {   // ---- A = X^T: this lane's feature, this thread's 8 rows of the chunk as 8 TMEM columns
        const float* pa = xa_off >= 0 ? st + xa_off : zero;
        uint32_t h[8], l[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) split_fast(pa[i * xa_str], h[i], l[i]);
        const uint32_t colA = kWgColA + 64 * buf;
        tmem_st8(lane_addr + colA + 8 * sub, h);
        tmem_st8(lane_addr + colA + 32 + 8 * sub, l);
      }
      SPW_PH(4);
      {   // ---- B = dY^T: [k-step][2][n][4 rows]
        float* Bhi_s = Bop + (size_t)(2 * buf) * bfl; float* Blo_s = Bhi_s + bfl;
        const uint8_t* stb = reinterpret_cast<const uint8_t*>(st);
#pragma unroll
        for (int u = 0; u < kUnits; ++u) {
          const int idx = tid + u * kWorkers;
          if (idx < (kWgCh / 4) * NB) {
            const float* py = yo[u] >= 0 ? st + yo[u] : zero;
            uint32_t h[4], l[4];
            uint32_t bw = 0xffffffffu;
            if (YMODE == 1) bw = *reinterpret_cast<const uint32_t*>(stb + bo[u]) >> ysh[u];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float y = py[i * ystr[u]];
              if (YMODE == 1) y = ((bw >> (8 * i)) & 1u) ? y : 0.f;
              split_fast(y, h[i], l[i]);
            }
            reinterpret_cast<uint4*>(Bhi_s)[idx] = make_uint4(h[0], h[1], h[2], h[3]);
            reinterpret_cast<uint4*>(Blo_s)[idx] = make_uint4(l[0], l[1], l[2], l[3]);
          }
        }
      }
      SPW_PH(5);
      tmem_wait_st();
      fence_async_smem();
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive_local(barR + buf);             // this warp's part of chunk q is in place
      if (lane == 0) mbar_arrive_local(barE + sq);              // ... and it has read everything it needs from the stage
      if (++sq == NST) { sq = 0; sph ^= 1u; }
      SPW_PH(6);
    }
#ifdef SPW_PHASE_TIMING
    if (blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 9) && a.M > 100000)
      printf("k_wgrad_c<%d> warp %d: wait copies %lld issue copies %lld flush %lld waitMMA %lld A %lld B %lld arrive %lld loop %lld (%d chunks)\n", YMODE, warp,
             ph_t[0], ph_t[1], ph_t[2], ph_t[3], ph_t[4], ph_t[5], ph_t[6], ph_t[7], nq);
#endif
    {   // running sums -> per-CTA partial in global memory ([n][lane]: coalesced)
      float* pp = a.part + (size_t)stream * (2 * 160 * 128) + (size_t)mt * (160 * 128) + L;
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        const int g = sub + 4 * j;
        if (g < ngroups) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            float* p = pp + (size_t)(8 * g + k) * 128;
            *p = a.first ? acc[j][k] : *p + acc[j][k];
          }
        }
      }
    }
    if (failed && tid == 0) a.poison[0] = __int_as_float(0x7fc00000);
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, kTmemCols);
}

// arrive on the mbarrier at the same shared-memory offset in CTA `cta` of the cluster.  Default semantics (release at CTA scope,
// no GPU-scope memory barrier): what the arrival publishes here is tensor memory (tcgen05.st + wait::st + fence::before_thread_sync)
// and shared memory behind a fence.proxy.async, both consumed by the tensor core, not generic-proxy data.
__device__ __forceinline__ void mbar_arrive_pair(uint64_t* bar, uint32_t cta) {
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(bar)), "r"(cta));
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}

// =========================================================================================================================
// k_wgrad_pair: the same contraction on a CTA PAIR (cluster of two, tcgen05.mma.cta_group::2, M = 256), for the products with
// more than 128 features (two M-tiles).  In k_wgrad_c both CTAs of a row stream build the WHOLE dY^T operand (the largest
// phase of a chunk) and each one's MMAs read all of it from shared memory (the bandwidth wall of that kernel).  Here CTA
// `rank` owns M-tile `rank` (its 128 features in its own tensor memory) and HALF of the dY columns: it stages, gathers,
// masks, splits and stores only NB / 2 columns; the leader's MMAs read both halves.  Per chunk: workers build -> each of the
// 2 x 16 worker warps arrives on the LEADER's barR[buffer] (mapa + mbarrier.arrive: phase-tracked, unlike the bar.arrive /
// bar.sync hand-over of the single-CTA kernels, which miscounted here once the leader's issuer could block on its peer) -> the
// leader's issuer thread issues the 12 MMAs of the chunk for both CTAs and commits with a multicast to barF[buffer] (operand
// buffers free) and, at the end of a 128-row tile, to barT (accumulator complete) of BOTH CTAs.  Everything else (stage ring, TMA tensor copies, running sums in
// registers, per-CTA partials) is k_wgrad_c's.
// =========================================================================================================================
// The per-chunk work of a thread is a fixed pattern; everything that does not depend on the chunk (shared-memory offsets of
// the words it reads and writes, which lanes are real / ones / padding features) is computed once, and padding is expressed as
// "read a zero word with stride 0" instead of predicates: the kernel was instruction-issue bound (ncu: 61 % issue-active).
// tmX / tmY: 2-D tensor maps [quads][rows * 4] of the X and dY arrays (base = first row and first quad of the view), boxes of
// [nqx_box][132] and [nqy][132] floats (launch code: run_wgrad_c).
template <int YMODE, int NB, int NST>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreadsC, 1) k_wgrad_pair(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY, WgradCArgs a, int nqx) {
  SPW_DYN_SMEM(smem_raw);
  constexpr int HB = NB / 2;                                                // B columns kept by this CTA
  constexpr int kUnits = ((kWgCh / 4) * HB + kWorkers - 1) / kWorkers;      // B-operand units (4 rows x 1 column) per thread
  constexpr int bfl = (kWgCh / 8) * (2 * HB * 4);                         // floats per hi or lo half operand of a chunk
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rank = (int)cluster_ctarank();
  const int mt = rank, stream = (int)blockIdx.x >> 1, nstreams = (int)gridDim.x >> 1;
  const int n0 = rank * HB;                                      // first dY column of this CTA's half operand
  const int f0 = mt == 0 ? 0 : a.Kx + 1 - 128;                   // first feature of this CTA's M-tile
  const int qlo = f0 >> 2;                                       // staged X quads: [qlo, qlo + nqx)
  const int fhi = mt == 0 ? (a.Kx < 128 ? a.Kx : 128) : a.Kx;    // features [f0, fhi) are read from X
  constexpr int nqy = HB / 4;                                    // staged dY quads: [rank * nqy, + nqy); nqx (kernel argument): quads of the X box
  const int nqy_all = (a.Ny + 3) >> 2;
  const int stf = wg_stage_floats(nqx, nqy);
  float* stages = reinterpret_cast<float*>(smem_raw);
  float* Bop = stages + NST * stf;                               // [2 buffers][hi | lo][bfl]
  float* zero = Bop + 4 * bfl;                                   // 8 zero words
  uint64_t* bars = reinterpret_cast<uint64_t*>(zero + 8);
  uint64_t* barF = bars; uint64_t* barT = bars + 2;
  uint64_t* barR = bars + 4;      // barR[2] (the leader's copy is used): the 2 x 16 worker warps of the pair have built a buffer
  uint64_t* barS = bars + 6;      // barS[NST]: stage filled: the tensor copies' bytes + 2 arrivals of each lane of the producer warp
  uint64_t* barE = barS + NST;    // barE[NST]: stage consumed: one arrival per worker warp
  uint32_t* tptr = reinterpret_cast<uint32_t*>(barE + NST);

  pdl_trigger();
  if (warp == 0) tmem_alloc2(tptr, kTmemCols);
  if (tid == 32) {
    mbar_init(barF, 1); mbar_init(barF + 1, 1); mbar_init(barT, 1); mbar_init(barT + 1, 1); mbar_init(barR, 2 * (kWorkers / 32)); mbar_init(barR + 1, 2 * (kWorkers / 32));
    for (int i = 0; i < NST; ++i) { mbar_init(barS + i, 1 + 2 * 32); mbar_init(barE + i, kWorkers / 32); }
    fence_mbar_init();
  }
  if (tid < 8) zero[tid] = 0.f;
  if (!a.rowscale)                                               // the virtual feature Kx is a column of ones
    for (int i = tid; i < NST * 32; i += kThreadsC) stages[(i >> 5) * stf + wg_up32(nqx * kQPitch) + wg_up32(nqy * kQPitch) + (i & 31)] = 1.f;
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  cluster_sync_all();                                            // the barriers of both CTAs exist before any remote arrive / multicast commit
  const uint32_t tmem_base = *tptr;
  const int ntiles = (a.M + kTM - 1) / kTM;
  const int my_tiles = stream < ntiles ? (ntiles - 1 - stream) / nstreams + 1 : 0;
  constexpr int kCh = kTM / kWgCh;                               // chunks per tile
  const int nq = my_tiles * kCh;
  auto row0_of = [&](int q) { return (long long)(stream + (q / kCh) * nstreams) * kTM + (q % kCh) * kWgCh; };

  if (warp >= kWorkers / 32) {
    // ---------------- MMA issuer warp (and its three idle siblings) ----------------
    regs_issuer32();
    constexpr uint32_t idesc = make_idesc_tf32(256, NB);
    bool ok = true;
    if (warp == kWorkers / 32 && rank == 0 && lane == 0)         // the leader's issuer thread issues for the pair
    for (int q = 0; q < nq; ++q) {
      const int t = q / kCh, c = q % kCh, buf = q & 1;
      if (!mbar_wait(barR + buf, (uint32_t)(q >> 1) & 1u)) ok = false;   // both CTAs' operands of chunk q are in tensor / shared memory
      fence_after_sync();
      const uint32_t d = tmem_base + kWgColD + 160 * (t & 1);
      const uint32_t ahi = tmem_base + kWgColA + 64 * buf, alo = ahi + 32;
      const uint32_t bhi = smem_u32(Bop + (size_t)(2 * buf) * bfl), blo = smem_u32(Bop + (size_t)(2 * buf + 1) * bfl);
#pragma unroll 1
      for (int ks = 0; ks < kWgCh / 8; ++ks) {
        const uint64_t dhi = make_b_desc(bhi + ks * (2 * HB * 16), HB * 16, 128);
        const uint64_t dlo = make_b_desc(blo + ks * (2 * HB * 16), HB * 16, 128);
        mma2_tf32_ts(d, alo + 8 * ks, dhi, idesc, (c > 0 || ks > 0) ? 1u : 0u);
        mma2_tf32_ts(d, ahi + 8 * ks, dlo, idesc, 1u);
        mma2_tf32_ts(d, ahi + 8 * ks, dhi, idesc, 1u);
      }
      mma2_commit_multicast(barF + buf);
      if (c == kCh - 1) mma2_commit_multicast(barT + (t & 1));
    }
    if (warp == kWorkers / 32 + 1) {
      // ---------------- producer warp: fills the stage ring, up to NST chunks ahead of the workers ----------------
      // (this used to be done by the worker warps between two chunks: ~100 instructions of index arithmetic, 1.1k of the 2.6k cycles
      // of a chunk on every worker's critical path although the copies themselves are asynchronous)
      const int y_off = wg_up32(nqx * kQPitch);
      const int rs_off = y_off + wg_up32(nqy * kQPitch);
      const int rl_off = rs_off + 32 + 160;
      pdl_wait();
      int slot = 0;
      int r_nx = 0;                                              // YMODE 1: receiver of row (chunk, lane), loaded one chunk ahead of its use
      if (YMODE == 1 && nq > 0) {
        const long long r = row0_of(0) + lane;
        if (r < a.M) r_nx = a.rcv[r];
      }
      for (int qi = 0; qi < nq; ++qi) {
        if (qi >= NST && !mbar_wait(barE + slot, (uint32_t)(qi / NST - 1) & 1u)) ok = false;      // the workers are done with chunk qi - NST
        float* st = stages + slot * stf;
        uint64_t* bs = barS + slot;
        const long long r0 = row0_of(qi);
        const int nvalid = a.M - r0 >= kWgCh ? kWgCh : (a.M > r0 ? (int)(a.M - r0) : 0);
        const bool valid = lane < nvalid;
        // YMODE 1: the rows are receiver-sorted, so the receivers of a chunk are an ascending range of nodes: when it spans at most 33
        // nodes, ONE tensor copy of the node table [quads][33 nodes x 4] brings the chunk's dY values, and RL[row] = receiver - first
        // receiver says where a row's values are; otherwise (isolated nodes in between) the rows are gathered one by one and RL[row] = row.
        bool node_tma = false;
        int r_idx = 0, r_first = 0;
        if (YMODE == 1) {
          r_idx = r_nx;
          if (qi + 1 < nq) {
            const long long r = row0_of(qi + 1) + lane;
            r_nx = r < a.M ? a.rcv[r] : 0;
          }
          r_first = __shfl_sync(0xffffffffu, r_idx, 0);
          const int r_last = __shfl_sync(0xffffffffu, r_idx, nvalid > 0 ? nvalid - 1 : 0);
          node_tma = a.gather_tma && nvalid > 0 && r_last - r_first < kQPitch / 4;
        }
        if (lane == 0) {                                         // streamed arrays: one tensor copy each (rows past the end read as zero)
          mbar_arrive_expect_tx(bs, (uint32_t)((nqx + ((YMODE == 0 || node_tma) ? nqy : 0)) * kQPitch * 4));
          tma_load_2d(st, &tmX, (int)(r0 * 4), qlo, bs);
          if (YMODE == 0) tma_load_2d(st + y_off, &tmY, (int)(r0 * 4), rank * nqy, bs);         // quads past the array read as zero
          else if (node_tma) tma_load_2d(st + y_off, &tmY, r_first * 4, rank * nqy, bs);
        }
        if (YMODE == 1) {
          reinterpret_cast<int*>(st + rl_off)[lane] = node_tma ? (valid ? r_idx - r_first : 0) : lane;
          // rows past the end: X is zero-filled by the tensor copy, the virtual ones feature (bias row) is switched off here, so
          // whatever finite value such a row picks up on the dY side contributes nothing
          if (!a.rowscale) st[rs_off + lane] = valid ? 1.f : 0.f;
          if (!node_tma) {                                       // gathered rows: lane = row, 16 bytes per quad
            const float* src = a.dY + (long long)((a.y_col0 >> 2) + rank * nqy) * a.y_slab + (long long)r_idx * 4;
            float* dst = st + y_off + lane * 4;
            for (int qd = 0; qd < nqy && rank * nqy + qd < nqy_all; ++qd) {
              cp_async16_zfill(dst, src, valid);
              src += a.y_slab; dst += kQPitch;
            }
          }
          uint8_t* BT = reinterpret_cast<uint8_t*>(st + rs_off + 32);    // relu bits of the chunk: 19 groups x 32 bytes, 16-byte pieces
          for (int i = lane; i < 2 * 19; i += 32)
            cp_async16_zfill(reinterpret_cast<float*>(BT + (i >> 1) * 32 + 16 * (i & 1)),
                             reinterpret_cast<const float*>(a.bits + (long long)(i >> 1) * a.bits_rows + r0 + 16 * (i & 1)), true);
        }
        if (a.rowscale) {
          const long long r = r0 + lane;
          cp_async4_zfill(st + rs_off + lane, a.rowscale + (valid ? (a.rsmod ? r % a.rsmod : r) : 0), valid);
        }
        cp_async_mbar_arrive_noinc(bs);                          // ... when this lane's copies have landed
        mbar_arrive_local(bs);                                   // ... and its plain stores are published
        if (++slot == NST) slot = 0;
      }
    }
    if (!ok && lane == 0) a.poison[0] = __int_as_float(0x7fc00000);
  } else {
    // ---------------- worker warps ----------------
    regs_workers();
    const int L = 32 * (warp & 3) + lane, sub = warp >> 2;      // TMEM lane = feature f0 + L; rows 8 sub .. 8 sub + 7 of a chunk
    const uint32_t lane_addr = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16);
    constexpr int ngroups = NB / 8;
    const int y_off = wg_up32(nqx * kQPitch);
    const int rs_off = y_off + wg_up32(nqy * kQPitch);           // RS (virtual feature values) inside a stage
    const int rl_off = rs_off + 32 + 160;                        // RL: staged dY row of each chunk row (YMODE 1)
    // A operand: word offset (within a stage) and row stride of this lane's feature; padding lanes read the zero word
    const int feat = f0 + L;
    int xa_off, xa_str;
    if (feat < fhi) { xa_off = ((feat >> 2) - qlo) * kQPitch + (feat & 3) + 32 * sub; xa_str = 4; }
    else if (feat == a.Kx) { xa_off = rs_off + 8 * sub; xa_str = 1; }
    else { xa_off = -1; xa_str = 0; }
    // B operand units of this thread: column n, row quad kc  ->  source word, bit word, destination
    int yo[kUnits], bo[kUnits], ysh[kUnits], ystr[kUnits], kcu[kUnits];
#pragma unroll
    for (int u = 0; u < kUnits; ++u) {
      const int idx = tid + u * kWorkers;
      const int n = idx % HB, kc = idx / HB;                       // column n0 + n of dY
      // YMODE 0: rows 4 kc .. 4 kc + 3 of the staged chunk; YMODE 1: the staged rows RL[4 kc ..] (a node of the staged range, or the row itself)
      if (idx < (kWgCh / 4) * HB && n0 + n < a.Ny) { yo[u] = y_off + (n >> 2) * kQPitch + (YMODE == 1 ? 0 : 16 * kc) + (n & 3); ystr[u] = 4; }
      else { yo[u] = -1; ystr[u] = 0; }
      kcu[u] = kc < kWgCh / 4 ? kc : 0;
      bo[u] = (rs_off + 32) * 4 + ((n0 + n) >> 3) * 32 + 4 * kc;   // byte offset of the 4 rows' relu bytes inside a stage
      ysh[u] = n & 7;
    }
    bool failed = false;
    float acc[5][8];
#pragma unroll
    for (int j = 0; j < 5; ++j)
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[j][k] = 0.f;
    SPW_PH_DECL

    // q = 0 .. nq - 1: chunk q;  q = nq: flush of the last tile only
    pdl_wait();
    int sq = 0; uint32_t sph = 0;                                // ring slot of chunk q and the phase of its barriers
#pragma unroll 1
    for (int q = 0; q <= nq; ++q) {
      SPW_PH(7);
      if (q < nq && !mbar_wait(barS + sq, sph)) { failed = true; SPW_WDBG("barS", q); }   // the stage of chunk q is filled
      SPW_PH(0);
      SPW_PH(1);
      const int t = q / kCh, c = q % kCh, buf = q & 1;
      if ((c == 1 && t >= 1) || q == nq) {                       // D of the previous tile -> running sums (round-to-nearest adds)
        const int tf = q == nq ? my_tiles - 1 : t - 1;
        if (tf >= 0) {
          if (!mbar_wait(barT + (tf & 1), (uint32_t)(tf >> 1) & 1u)) { failed = true; SPW_WDBG("barT", q); }
          fence_after_sync();
          uint32_t d[5][8];
          load_d<5>(d, lane_addr, kWgColD + 160 * (tf & 1), sub, ngroups);
#pragma unroll
          for (int j = 0; j < 5; ++j)
#pragma unroll
            for (int k = 0; k < 8; ++k)
              if (sub + 4 * j < ngroups) acc[j][k] += __uint_as_float(d[j][k]);
          fence_before_sync();
        }
      }
      SPW_PH(2);
      if (q == nq) break;
      if (q >= 2) {                                              // operand buffers `buf` are free once chunk q - 2's MMAs are done
        if (!mbar_wait(barF + buf, (uint32_t)((q >> 1) - 1) & 1u)) { failed = true; SPW_WDBG("barF", q); }
        fence_after_sync();
      }
      SPW_PH(3);
      const float* st = stages + sq * stf;
      {   // ---- A = X^T: this lane's feature, this thread's 8 rows of the chunk as 8 TMEM columns
        const float* pa = xa_off >= 0 ? st + xa_off : zero;
        uint32_t h[8], l[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) split_fast(pa[i * xa_str], h[i], l[i]);
        const uint32_t colA = kWgColA + 64 * buf;
        tmem_st8(lane_addr + colA + 8 * sub, h);
        tmem_st8(lane_addr + colA + 32 + 8 * sub, l);
      }
      SPW_PH(4);
      {   // ---- B = dY^T: [k-step][2][n][4 rows]
        float* Bhi_s = Bop + (size_t)(2 * buf) * bfl; float* Blo_s = Bhi_s + bfl;
        const uint8_t* stb = reinterpret_cast<const uint8_t*>(st);
#pragma unroll
        for (int u = 0; u < kUnits; ++u) {
          const int idx = tid + u * kWorkers;
          if (idx < (kWgCh / 4) * HB) {
            const float* py = yo[u] >= 0 ? st + yo[u] : zero;
            uint32_t h[4], l[4];
            uint32_t bw = 0xffffffffu;
            int ro[4] = {0, 1, 2, 3};
            if (YMODE == 1) {
              bw = *reinterpret_cast<const uint32_t*>(stb + bo[u]) >> ysh[u];
              const int4 rr = *reinterpret_cast<const int4*>(st + rl_off + 4 * kcu[u]);
              ro[0] = rr.x; ro[1] = rr.y; ro[2] = rr.z; ro[3] = rr.w;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float y = py[ro[i] * ystr[u]];
              if (YMODE == 1) y = ((bw >> (8 * i)) & 1u) ? y : 0.f;
              split_fast(y, h[i], l[i]);
            }
            reinterpret_cast<uint4*>(Bhi_s)[idx] = make_uint4(h[0], h[1], h[2], h[3]);
            reinterpret_cast<uint4*>(Blo_s)[idx] = make_uint4(l[0], l[1], l[2], l[3]);
          }
        }
      }
      SPW_PH(5);
      tmem_wait_st();
      fence_async_smem();
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive_pair(barR + buf, 0);           // this warp's part of chunk q is in place (leader's barrier)
      if (lane == 0) mbar_arrive_local(barE + sq);              // ... and it has read everything it needs from the stage
      if (++sq == NST) { sq = 0; sph ^= 1u; }
      SPW_PH(6);
    }
#ifdef SPW_PHASE_TIMING
    if (blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 9) && a.M > 100000)
      printf("k_wgrad_pair<%d> warp %d: wait copies %lld issue copies %lld flush %lld waitMMA %lld A %lld B %lld arrive %lld loop %lld (%d chunks)\n", YMODE, warp,
             ph_t[0], ph_t[1], ph_t[2], ph_t[3], ph_t[4], ph_t[5], ph_t[6], ph_t[7], nq);
#endif
    {   // running sums -> per-CTA partial in global memory ([n][lane]: coalesced)
      float* pp = a.part + (size_t)stream * (2 * 160 * 128) + (size_t)mt * (160 * 128) + L;
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        const int g = sub + 4 * j;
        if (g < ngroups) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            float* p = pp + (size_t)(8 * g + k) * 128;
            *p = a.first ? acc[j][k] : *p + acc[j][k];
          }
        }
      }
    }
    if (failed && tid == 0) a.poison[0] = __int_as_float(0x7fc00000);
  }
  fence_before_sync();
  cluster_sync_all();                                            // the leader's MMAs read the peer's tensor and shared memory
  if (warp == 0) tmem_dealloc2(tmem_base, kTmemCols);
}

}  // namespace csl
}  // namespace spw
#endif  // SPW_EMU
