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
//   * raw chunk rows arrive by cp.async from the column-slab arrays ([quad][row][4]: 512 contiguous bytes per quad and chunk)
//     into a ring of stages; the streamed array is prefetched into L2 three chunks ahead.
#pragma once
#ifndef SPW_EMU
#include "spw_csl.cuh"

namespace spw {
namespace csl {

constexpr int kWgCh = 32;                         // rows per chunk
constexpr int kQPitch = 132;                      // floats per staged quad ([32 rows][4] + 4: bank-conflict-free both ways)
constexpr int kBarWork = 2;                       // named barrier of the 512 workers
constexpr uint32_t kWgColD = 0, kWgColA = 320;    // TMEM: D0 [0,160) D1 [160,320) | A buffers: hi [320 + 64 b, +32) lo [+32, +64)

struct WgradCArgs {
  int M;
  const float* X; long long x_slab; int x_col0; int Kx; int xmod;      // X view (XMODE 1: the A_e array); row = r % xmod if xmod
  const float* rowscale; int rsmod;                                    // value of the virtual feature Kx (null: 1)
  const float* S; const float* R; long long sr_slab;                   // XMODE 1: x = relu(X + S[snd] + R[rcv])
  const int32_t* snd; const int32_t* rcv;
  const float* dY; long long y_slab; int y_col0; int Ny;               // dY view; YMODE 1: node table gathered by rcv, masked by bits
  const uint8_t* bits; long long bits_rows;                            // YMODE 1: byte-slab relu bits [19][rows]
  int NB;                                                              // MMA N: 112 or 160 (>= Ny)
  int nmt;                                                             // M-tiles: 1 (Kx + 1 <= 128) or 2
  float* part;                                                         // [gridDim.x / nmt][2][160][128]
  int first;                                                           // this launch initialises the partials (else it adds to them)
  float* poison;
};

__device__ __forceinline__ void l2_prefetch(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// stage layout (floats): X quads [nqx][132] | (XMODE 1: S quads, R quads) | Y quads [nqy][132] | RS [32] | bits [20][32] bytes
template <int XMODE>
__host__ __device__ constexpr int wg_stage_floats(int nqx, int nqy) { return ((XMODE ? 3 : 1) * nqx + nqy) * kQPitch + 32 + 160; }
template <int XMODE>
constexpr size_t wgrad_c_smem(int nqx, int nqy, int NB, int nst) {
  return (size_t)(nst * wg_stage_floats<XMODE>(nqx, nqy) + 2 * 2 * (kWgCh / 8) * (2 * NB * 4)) * sizeof(float) + 128;
}

template <int XMODE, int YMODE, int NST>
__global__ void __launch_bounds__(kThreadsC, 1) k_wgrad_c(WgradCArgs a) {
  SPW_DYN_SMEM(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mt = (int)blockIdx.x % a.nmt, stream = (int)blockIdx.x / a.nmt, nstreams = (int)gridDim.x / a.nmt;
  const int f0 = mt == 0 ? 0 : a.Kx + 1 - 128;                   // first feature of this CTA's M-tile
  const int qlo = f0 >> 2;                                       // staged X quads: [qlo, qlo + nqx)
  const int fhi = mt == 0 ? (a.Kx < 128 ? a.Kx : 128) : a.Kx;    // features [f0, fhi) are read from X
  const int nqx = ((fhi + 3) >> 2) - qlo;
  const int nqy = (a.Ny + 3) >> 2;
  const int NB = a.NB;
  const int stf = wg_stage_floats<XMODE>(nqx, nqy);
  const int bfl = (kWgCh / 8) * (2 * NB * 4);                    // floats per hi or lo B operand of a chunk
  float* stages = reinterpret_cast<float*>(smem_raw);
  float* Bop = stages + NST * stf;                               // [2 buffers][hi | lo][bfl]
  uint64_t* bars = reinterpret_cast<uint64_t*>(Bop + 4 * bfl);
  uint64_t* barF = bars; uint64_t* barT = bars + 2;
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bars + 4);

  if (warp == 0) tmem_alloc(tptr, kTmemCols);
  if (tid == 32) { mbar_init(barF, 1); mbar_init(barF + 1, 1); mbar_init(barT, 1); mbar_init(barT + 1, 1); fence_mbar_init(); }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tptr;
  const int ntiles = (a.M + kTM - 1) / kTM;
  const int my_tiles = stream < ntiles ? (ntiles - 1 - stream) / nstreams + 1 : 0;
  constexpr int kCh = kTM / kWgCh;                               // chunks per tile
  const int nq = my_tiles * kCh;
  auto row0_of = [&](int q) { return (long long)(stream + (q / kCh) * nstreams) * kTM + (q % kCh) * kWgCh; };

  if (warp == kWorkers / 32) {
    // ---------------- MMA issuer warp ----------------
    const uint32_t idesc = make_idesc_tf32(128, NB);
    for (int q = 0; q < nq; ++q) {
      nbar_sync(kBarOps, kThreadsC);
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
  } else {
    // ---------------- worker warps ----------------
    const int L = 32 * (warp & 3) + lane, sub = warp >> 2;      // TMEM lane = feature f0 + L; rows 8 sub .. 8 sub + 7 of a chunk
    const uint32_t lane_addr = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16);
    const int feat = f0 + L;
    const int xoff = ((feat >> 2) - qlo) * kQPitch + (feat & 3);         // this feature's word inside a staged row quad
    const int ngroups = NB / 8;
    bool failed = false;
    float acc[5][8];
#pragma unroll
    for (int j = 0; j < 5; ++j)
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[j][k] = 0.f;
    int s_idx = 0, r_idx = 0;                                    // XMODE / YMODE 1: sender / receiver of row (chunk, lane), one chunk ahead
    auto load_idx = [&](int q) {
      s_idx = 0; r_idx = 0;
      if ((XMODE == 1 || YMODE == 1) && q < nq) {
        const long long r = row0_of(q) + lane;
        if (r < a.M) { if (XMODE == 1) s_idx = a.snd[r]; r_idx = a.rcv[r]; }
      }
    };
    // asynchronous copies of chunk q into its stage: thread -> (quad, row = lane); uses s_idx / r_idx loaded for chunk q
    auto issue_chunk = [&](int q) {
      if (q < nq) {
        float* st = stages + (q % NST) * stf;
        const long long r0 = row0_of(q), r = r0 + lane;
        const bool valid = r < a.M;
        const long long rr = valid ? r : 0;
        const long long xr = a.xmod ? rr % a.xmod : rr;
        const int qx0 = (a.x_col0 >> 2) + qlo;
        for (int qd = warp; qd < nqx; qd += kWorkers / 32) {
          cp_async16_zfill(st + qd * kQPitch + lane * 4, a.X + (long long)(qx0 + qd) * a.x_slab + xr * 4, valid);
          if (XMODE == 1) {
            cp_async16_zfill(st + (nqx + qd) * kQPitch + lane * 4, a.S + (long long)(qlo + qd) * a.sr_slab + (long long)s_idx * 4, valid);
            cp_async16_zfill(st + (2 * nqx + qd) * kQPitch + lane * 4, a.R + (long long)(qlo + qd) * a.sr_slab + (long long)r_idx * 4, valid);
          }
        }
        float* YD = st + (XMODE ? 3 : 1) * nqx * kQPitch;
        const int qy0 = a.y_col0 >> 2;
        for (int qd = warp; qd < nqy; qd += kWorkers / 32) {
          if (YMODE == 1) cp_async16_zfill(YD + qd * kQPitch + lane * 4, a.dY + (long long)(qy0 + qd) * a.y_slab + (long long)r_idx * 4, valid);
          else cp_async16_zfill(YD + qd * kQPitch + lane * 4, a.dY + (long long)(qy0 + qd) * a.y_slab + rr * 4, valid);
        }
        float* RS = YD + nqy * kQPitch;
        if (a.rowscale && warp == 0) cp_async4_zfill(RS + lane, a.rowscale + (a.rsmod ? rr % a.rsmod : rr), valid);
        if (YMODE == 1 && warp == 1) {                           // relu bits of the chunk: 19 groups x 32 bytes
          uint8_t* BT = reinterpret_cast<uint8_t*>(RS + 32);
          // two 16-byte pieces per group; the bit arrays are allocated in whole 128-row tiles, so the read stays in bounds
          for (int i = lane; i < 2 * 19; i += 32)
            cp_async16_zfill(reinterpret_cast<float*>(BT + (i >> 1) * 32 + 16 * (i & 1)),
                             reinterpret_cast<const float*>(a.bits + (long long)(i >> 1) * a.bits_rows + r0 + 16 * (i & 1)), true);
        }
        // stream prefetch into L2, three chunks ahead (one 512-byte piece per quad)
        if (q + 3 < nq && lane == 0) {
          const long long rp = row0_of(q + 3);
          if (rp + kWgCh <= a.M && !a.xmod)
            for (int qd = warp; qd < nqx; qd += kWorkers / 32) l2_prefetch(a.X + (long long)(qx0 + qd) * a.x_slab + rp * 4, kWgCh * 16);
          if (YMODE == 0 && rp + kWgCh <= a.M)
            for (int qd = warp; qd < nqy; qd += kWorkers / 32) l2_prefetch(a.dY + (long long)(qy0 + qd) * a.y_slab + rp * 4, kWgCh * 16);
        }
      }
      cp_async_commit();
    };
    auto flush_tile = [&](int t) {                               // D[t & 1] of local tile t -> running sums (round-to-nearest adds)
      if (!mbar_wait(barT + (t & 1), (uint32_t)(t >> 1) & 1u)) failed = true;
      fence_after_sync();
      uint32_t d[5][8];
      load_d<5>(d, lane_addr, kWgColD + 160 * (t & 1), sub, ngroups);
#pragma unroll
      for (int j = 0; j < 5; ++j)
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (sub + 4 * j < ngroups) acc[j][k] += __uint_as_float(d[j][k]);
      fence_before_sync();
    };

    // prologue: chunks 0 .. NST - 2 in flight, indices one chunk ahead
    load_idx(0);
#pragma unroll 1
    for (int p = 0; p < NST - 1; ++p) { issue_chunk(p); load_idx(p + 1); }
    for (int q = 0; q < nq; ++q) {
      const int t = q / kCh, c = q % kCh, buf = q & 1;
      cp_async_wait<NST - 2>();                                  // this thread's copies of chunk q have landed
      nbar_sync(kBarWork, kWorkers);                             // ... everybody's; and chunk q - 1 is fully consumed
      issue_chunk(q + NST - 1);                                  // refill the stage chunk q - 1 used
      load_idx(q + NST);
      if (c == 1 && t >= 1) flush_tile(t - 1);
      if (q >= 2) {                                              // operand buffers `buf` are free once chunk q - 2's MMAs are done
        if (!mbar_wait(barF + buf, (uint32_t)((q >> 1) - 1) & 1u)) failed = true;
        fence_after_sync();
      }
      const float* st = stages + (q % NST) * stf;
      const float* XA = st; const float* XS = st + nqx * kQPitch; const float* XR = st + 2 * nqx * kQPitch;
      const float* YD = st + (XMODE ? 3 : 1) * nqx * kQPitch;
      const float* RS = YD + nqy * kQPitch;
      const uint8_t* BT = reinterpret_cast<const uint8_t*>(RS + 32);
      const long long r0 = row0_of(q);
      // ---- A = X^T: this lane's feature, this thread's 8 rows of the chunk as 8 TMEM columns
      {
        uint32_t h[8], l[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int j = 8 * sub + i;
          float xv = 0.f;
          if (feat < fhi) {
            xv = XA[xoff + j * 4];
            if (XMODE == 1) xv = relu_f(xv + XS[xoff + j * 4] + XR[xoff + j * 4]);
          } else if (feat == a.Kx && r0 + j < a.M) {
            xv = a.rowscale ? RS[j] : 1.f;
          }
          split_fast(xv, h[i], l[i]);
        }
        const uint32_t colA = kWgColA + 64 * buf;
        tmem_st8(lane_addr + colA + 8 * sub, h);
        tmem_st8(lane_addr + colA + 32 + 8 * sub, l);
      }
      // ---- B = dY^T: [k-step][2][n][4 rows]
      {
        float* Bhi_s = Bop + (size_t)(2 * buf) * bfl; float* Blo_s = Bhi_s + bfl;
        for (int idx = tid; idx < (kWgCh / 4) * NB; idx += kWorkers) {
          const int n = idx % NB, kc = idx / NB;
          uint32_t h[4], l[4];
          uint32_t bw = 0xffffffffu;
          if (YMODE == 1 && n < a.Ny) bw = *reinterpret_cast<const uint32_t*>(BT + (n >> 3) * 32 + 4 * kc);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float y = 0.f;
            if (n < a.Ny) {
              y = YD[(n >> 2) * kQPitch + (4 * kc + i) * 4 + (n & 3)];
              if (YMODE == 1) y = ((bw >> (8 * i + (n & 7))) & 1u) ? y : 0.f;
            }
            split_fast(y, h[i], l[i]);
          }
          reinterpret_cast<uint4*>(Bhi_s)[idx] = make_uint4(h[0], h[1], h[2], h[3]);
          reinterpret_cast<uint4*>(Blo_s)[idx] = make_uint4(l[0], l[1], l[2], l[3]);
        }
      }
      tmem_wait_st();
      fence_async_smem();
      fence_before_sync();
      nbar_arrive(kBarOps, kThreadsC);
    }
    cp_async_wait<0>();
    if (my_tiles >= 1) flush_tile(my_tiles - 1);                 // the last tile (every earlier one was flushed one tile late)
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

}  // namespace csl
}  // namespace spw
#endif  // SPW_EMU
