// spw_csl_kernels.cuh -- the kernels of the column-slab data path besides the generic layer (spw_csl.cuh): layer-0 encoders,
// the pipelined edge step and its data gradient, segmented-sum fix-up, head, gathers of the backward pass, skinny weight
// gradients.  Layout conventions: spw_csl.cuh.  sm_100a only.
#pragma once
#ifndef SPW_EMU
#include "spw_csl.cuh"
#include <type_traits>

namespace spw {
namespace csl {

constexpr int kQE = 38;          // column quads of a 150-wide (152 allocated) edge / node array
constexpr int kQP = 25;          // column quads of a 100-wide node array

// ---- object encoder layer 0 (Networks.py:47,76; K = 2 on [y, w]): Q1 = relu(y W[0] + w W[1] + b) -------------------------
__global__ void __launch_bounds__(256) k_obj_enc0_c(const float* __restrict__ obj, int n, const float* __restrict__ W,
                                                    const float* __restrict__ b, float* __restrict__ out, long long slab) {
  pdl_trigger();
  pdl_wait();
  const long long total = (long long)n * kQP;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int qd = (int)(idx / n), i = (int)(idx - (long long)qd * n);
    const float y = obj[3 * (size_t)i + 1], w = obj[3 * (size_t)i + 2];
    float v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int c = 4 * qd + e;
      v[e] = relu_f(fmaf(w, __ldg(W + kDP + c), fmaf(y, __ldg(W + c), __ldg(b + c))));
    }
    *reinterpret_cast<float4*>(out + (long long)qd * slab + (long long)i * 4) = make_float4(v[0], v[1], v[2], v[3]);
  }
}

// ---- relation encoder layer 0 (Networks.py:58-62,69,75; K = 2): X0[e] = relu(dx W0[0] + dy W0[1] + b0), X0[e][150] = 1 ------
// thread = edge (coalesced 16-byte stores per column quad), sign bits per group of 8 columns
__global__ void __launch_bounds__(256) k_edge_enc0_c(int E, const int32_t* __restrict__ in_snd, const int32_t* __restrict__ in_rcv,
                                                     const float* __restrict__ obj, const float* __restrict__ W0,
                                                     const float* __restrict__ b0, float* __restrict__ X0, uint8_t* __restrict__ bits, long long bits_rows) {
  __shared__ float sw[3][kDEP];
  pdl_trigger();
  for (int i = threadIdx.x; i < kDEP; i += blockDim.x) {
    sw[0][i] = i < kDE ? W0[i] : 0.f; sw[1][i] = i < kDE ? W0[kDE + i] : 0.f; sw[2][i] = i < kDE ? b0[i] : 0.f;
  }
  __syncthreads();
  pdl_wait();
  const long long slab = (long long)E * 4;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += (long long)gridDim.x * blockDim.x) {
    const int s = in_snd[e], rc = in_rcv[e];
    const float dx = obj[3 * (size_t)rc] - obj[3 * (size_t)s], dy = obj[3 * (size_t)rc + 1] - obj[3 * (size_t)s + 1];
#pragma unroll 1
    for (int g = 0; g < kQE / 2; ++g) {
      float v[8];
      uint32_t b = 0u;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int c = 8 * g + k;
        v[k] = relu_f(fmaf(dy, sw[1][c], fmaf(dx, sw[0][c], sw[2][c])));
        if (c >= kDE) v[k] = c == kDE ? 1.f : 0.f;
        else if (v[k] > 0.f) b |= 1u << k;
      }
      *reinterpret_cast<float4*>(X0 + (long long)(2 * g) * slab + e * 4) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(X0 + (long long)(2 * g + 1) * slab + e * 4) = make_float4(v[4], v[5], v[6], v[7]);
      if (bits) bits[(long long)g * bits_rows + e] = (uint8_t)b;
    }
  }
}

// ---- head (Networks.py:93-96): logit_i = U5_i . V2[:,0] + c2[0]; thread = node ------------------------------------------------
__global__ void __launch_bounds__(256) k_logit_c(const float* __restrict__ U, long long slab, int n, const float* __restrict__ V2raw,
                                                 const float* __restrict__ c2raw, float* __restrict__ logits, float* __restrict__ probs) {
  __shared__ float sv[kDP];
  pdl_trigger();
  for (int i = threadIdx.x; i < kDP; i += blockDim.x) sv[i] = V2raw[(size_t)i * (kDP + 1)];
  __syncthreads();
  pdl_wait();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float s = 0.f;
#pragma unroll 5
    for (int qd = 0; qd < kQP; ++qd) {
      const float4 u = *reinterpret_cast<const float4*>(U + (long long)qd * slab + i * 4);
      s = fmaf(u.x, sv[4 * qd], s); s = fmaf(u.y, sv[4 * qd + 1], s); s = fmaf(u.z, sv[4 * qd + 2], s); s = fmaf(u.w, sv[4 * qd + 3], s);
    }
    const float z = s + c2raw[0];
    logits[i] = z;
    if (probs) probs[i] = 1.f / (1.f + expf(-z));
  }
}

// dUpre5[i][c] = dlogit_i * V2[c][0] * [U5[i][c] > 0]; thread = (quad, node)
__global__ void __launch_bounds__(256) k_logit_bwd_c(const float* __restrict__ dlogits, const float* __restrict__ U, long long u_slab, int n,
                                                     const float* __restrict__ V2raw, float* __restrict__ dU, long long du_slab) {
  pdl_trigger();
  pdl_wait();
  const long long total = (long long)n * kQP;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int qd = (int)(idx / n), i = (int)(idx - (long long)qd * n);
    const float4 u = *reinterpret_cast<const float4*>(U + (long long)qd * u_slab + (long long)i * 4);
    const float dl = dlogits[i];
    float4 o;
    o.x = u.x > 0.f ? dl * __ldg(V2raw + (size_t)(4 * qd) * (kDP + 1)) : 0.f;
    o.y = u.y > 0.f ? dl * __ldg(V2raw + (size_t)(4 * qd + 1) * (kDP + 1)) : 0.f;
    o.z = u.z > 0.f ? dl * __ldg(V2raw + (size_t)(4 * qd + 2) * (kDP + 1)) : 0.f;
    o.w = u.w > 0.f ? dl * __ldg(V2raw + (size_t)(4 * qd + 3) * (kDP + 1)) : 0.f;
    *reinterpret_cast<float4*>(dU + (long long)qd * du_slab + (long long)i * 4) = o;
  }
}

// =========================================================================================================================
// k_edge_step_c: the forward edge step (Networks.py:84-88 after the factorisation of DESIGN.md section 2), pipelined:
//   h2_e = relu(W2 . relu(A_e + S_s + R_r) + b2);  H2S_i = sum_{e -> i} h2_e;  sign bits of h1 / h2 for training.
// Operand build: thread (row, q) loads its 16-byte quads of A_e (coalesced) and of the node tables S[sender], R[receiver]
// (the nodes of a tower are adjacent: a warp touches one or two lines per quad).  Receiver sum: the rows of a warp are 32
// consecutive receiver-sorted edges, so the sum over in-edges is a SEGMENTED SCAN ACROSS LANES with warp shuffles, in a
// fixed order; a segment inside one 32-row chunk is written straight to H2S, pieces of segments that cross a chunk go to
// part_first / part_last[chunk] and are combined, in order, by k_seg_fix_c.
// =========================================================================================================================
struct EdgeStepCArgs {
  int E;
  const int32_t* in_snd; const int32_t* in_rcv; const int32_t* in_off;
  const float* A;                                       // [E][150] CSL (slab = E * 4)
  const float* S; const float* R; long long sr_slab;    // node tables [n][150] CSL
  const float* W2hi; const float* W2lo;                 // packed B operands [19][2][160][4], bias in row 150 (k_pack_umma)
  float* H2S; long long h_slab;                         // [n][150] CSL (k_seg_fix_c completes it)
  float* part_first; float* part_last;                  // [ceil(E / 32)][160] row-major
  uint8_t* bits_h2; uint8_t* bits_h1; long long bits_rows;   // byte-slab [19][bits_rows] or null
  float* H1;                                            // [E][150] CSL or null: h1 (column 150 = 1) kept for the weight gradient
  float* poison;
};
constexpr size_t kEdgeStepCSmem = (size_t)(2 * kBFloats) * sizeof(float) + 64;

__global__ void __launch_bounds__(kThreadsC, 1) k_edge_step_c(EdgeStepCArgs a) {
  constexpr int KJ = 5, GJ = 5, NKS = kKS, NB = kN;
  SPW_DYN_SMEM(smem_raw);
  float* Bhi_s = reinterpret_cast<float*>(smem_raw);
  float* Blo_s = Bhi_s + kBFloats;
  uint64_t* bars = reinterpret_cast<uint64_t*>(Blo_s + kBFloats);
  uint64_t* barC = bars; uint64_t* barM = bars + 1; uint64_t* barW = bars + 2;
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bars + 3);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = 32 * (warp & 3) + lane, q = warp >> 2;

  pdl_trigger();
  if (warp == 0) tmem_alloc(tptr, kTmemCols);
  if (tid == 32) { mbar_init(barC, 1); mbar_init(barM, 1); mbar_init(barW, 1); fence_mbar_init(); }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tptr;
  constexpr uint32_t colHi = kColAhi, colLo = kColAlo, colD = kColD;
  const int ntiles = (a.E + kTM - 1) / kTM;
  const int cnt = (int)blockIdx.x < ntiles ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (warp >= kWorkers / 32) {
    regs_issuer();
    if (warp == kWorkers / 32) {
    if (lane == 0) bulk_load_weights(Bhi_s, Blo_s, a.W2hi, a.W2lo, (uint32_t)kBFloats * 4, barW);
    bool ok = mbar_wait(barW, 0);
    for (int i = 0; i < cnt; ++i) {
      nbar_sync(kBarOps, kBarOpsCount);
      fence_after_sync();
      if (lane == 0) issue_tile(tmem_base + colD, tmem_base + colHi, tmem_base + colLo, smem_u32(Bhi_s), smem_u32(Blo_s), NKS, NB, barC, barM);
      __syncwarp();
    }
    if (!ok && lane == 0) a.poison[0] = __int_as_float(0x7fc00000);
    } else {
      // L2 prefetch warps (spw_csl.cuh): the A rows of tile i + 2 while the workers build tile i + 1
      const int pl = tid - (kWorkers + 32);
      pdl_wait();
      auto prefetch_tile = [&](int i) {
        const long long r0 = (long long)(blockIdx.x + i * gridDim.x) * kTM;
        const int nrows = a.E - r0 >= kTM ? kTM : (int)(a.E - r0);
        prefetch_quads(a.A, (long long)a.E * 4, kQE, r0, nrows, pl);
      };
      if (cnt > 1) prefetch_tile(1);
      if (cnt > 2) prefetch_tile(2);
      for (int i = 0; i + 3 < cnt; ++i) {
        mbar_wait(barC, (uint32_t)i & 1u);
        prefetch_tile(i + 3);
      }
    }
  } else {
    regs_workers();
    const uint32_t lane_addr = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16);
    const long long es = (long long)a.E * 4;
    bool failed = false;
    XR<KJ> x;
    SPW_PH_DECL
    int s_nx = 0, r_nx = -1;                             // sender / receiver of this thread's row in the tile being built
    int r_cur = -1;                                      // receiver of the row in the tile whose epilogue runs
    auto load_idx = [&](int i, int& s, int& r) {
      const long long e = (long long)(blockIdx.x + i * gridDim.x) * kTM + row;
      s = 0; r = -1;
      if (i < cnt && e < a.E) { s = a.in_snd[e]; r = a.in_rcv[e]; }
    };
    int2 io_built = make_int2(0, 0), io_cur = make_int2(0, 0);   // in_off[r], in_off[r + 1] of the row: requested a tile before the epilogue needs them
    // x = relu(A_e + S_s + R_r) for this thread's k-steps; ones column (150) picks up the bias row; h1 sign bits
    auto build_x = [&](int i, int s, int r) {
      const long long e = (long long)(blockIdx.x + i * gridDim.x) * kTM + row;
      const bool rv = r >= 0;
      const float* ap = a.A + (long long)(2 * q) * es + (rv ? e : 0) * 4;
      const float* sp = a.S + (long long)(2 * q) * a.sr_slab + (long long)s * 4;
      const float* rp = a.R + (long long)(2 * q) * a.sr_slab + (long long)(rv ? r : 0) * 4;
      // the loads of k-step j + 1 are issued BEFORE k-step j is consumed and stored (h1, sign bits): the stores may alias the loads as far
      // as the compiler knows, so it never moves a load above them by itself, and every k-step paid a full memory latency (ncu: ~40 % of
      // the worker samples on the first FADD after each load group)
      float4 va[2][2], vs[2][2], vr[2][2];               // [stage][half of the k-step]
      auto issue = [&](int j, int sg) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          va[sg][h] = make_float4(0.f, 0.f, 0.f, 0.f); vs[sg][h] = va[sg][h]; vr[sg][h] = va[sg][h];
          if (rv && q + 4 * j < NKS) {
            va[sg][h] = *reinterpret_cast<const float4*>(ap + (long long)(8 * j + h) * es);
            vs[sg][h] = *reinterpret_cast<const float4*>(sp + (long long)(8 * j + h) * a.sr_slab);
            vr[sg][h] = *reinterpret_cast<const float4*>(rp + (long long)(8 * j + h) * a.sr_slab);
          }
        }
      };
      issue(0, 0);
#pragma unroll
      for (int j = 0; j < KJ; ++j) {
        const int ks = q + 4 * j;
        if (j + 1 < KJ) issue(j + 1, (j + 1) & 1);
        if (ks < NKS) {                                  // warp-uniform
          const int sg = j & 1;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            x.v[j][4 * h] = relu_f(va[sg][h].x + vs[sg][h].x + vr[sg][h].x); x.v[j][4 * h + 1] = relu_f(va[sg][h].y + vs[sg][h].y + vr[sg][h].y);
            x.v[j][4 * h + 2] = relu_f(va[sg][h].z + vs[sg][h].z + vr[sg][h].z); x.v[j][4 * h + 3] = relu_f(va[sg][h].w + vs[sg][h].w + vr[sg][h].w);
          }
          if (ks == NKS - 1) { x.v[j][6] = rv ? 1.f : 0.f; x.v[j][7] = 0.f; }      // columns 150 (ones) and 151 (pad)
          if (a.H1 && rv) {                             // h1 for the backward pass (the weight gradient streams it)
            float* hp = a.H1 + (long long)(2 * q) * es + e * 4;
            *reinterpret_cast<float4*>(hp + (long long)(8 * j) * es) = make_float4(x.v[j][0], x.v[j][1], x.v[j][2], x.v[j][3]);
            *reinterpret_cast<float4*>(hp + (long long)(8 * j + 1) * es) = make_float4(x.v[j][4], x.v[j][5], x.v[j][6], x.v[j][7]);
          }
          if (a.bits_h1 && rv) {
            uint32_t b = 0u;
#pragma unroll
            for (int k = 0; k < 8; ++k) b |= x.v[j][k] > 0.f ? (1u << k) : 0u;
            if (ks == NKS - 1) b &= 0x3fu;
            a.bits_h1[(long long)ks * a.bits_rows + e] = (uint8_t)b;
          }
        } else {
#pragma unroll
          for (int k = 0; k < 8; ++k) x.v[j][k] = 0.f;
        }
      }
    };

    pdl_wait();
    if (cnt > 0) {
      load_idx(0, s_nx, r_nx);
      build_x(0, s_nx, r_nx);
      r_cur = r_nx;
      if (r_cur >= 0) io_cur = make_int2(a.in_off[r_cur], a.in_off[r_cur + 1]);
      load_idx(1, s_nx, r_nx);
      store_lo<KJ>(x, lane_addr, colLo, q, NKS);
      store_hi<KJ>(x, lane_addr, colHi, q, NKS);
      tmem_wait_st();
      fence_before_sync();
      nbar_arrive(kBarOps, kBarOpsCount);
    }
    for (int i = 0; i < cnt; ++i) {
      const bool has_next = i + 1 < cnt;
      const uint32_t parity = (uint32_t)i & 1u;
      int r_built = -1;
      SPW_PH(7);
      if (has_next) {                                    // under the MMAs of tile i
        if (r_nx >= 0) io_built = make_int2(a.in_off[r_nx], a.in_off[r_nx + 1]);
        build_x(i + 1, s_nx, r_nx);
        r_built = r_nx;
        load_idx(i + 2, s_nx, r_nx);                     // indices one tile ahead of the gathers that need them
      }
      SPW_PH(0);                                         // p0: operand build
      if (!mbar_wait(barC, parity)) failed = true;
      fence_after_sync();
      SPW_PH(1);
      if (has_next) store_lo<KJ>(x, lane_addr, colLo, q, NKS);
      SPW_PH(2);
      if (!mbar_wait(barM, parity)) failed = true;
      fence_after_sync();
      SPW_PH(3);
      if (has_next) store_hi<KJ>(x, lane_addr, colHi, q, NKS);
      uint32_t d[GJ][8];
      load_d<GJ>(d, lane_addr, colD, q, NB / 8);
      if (has_next) {
        tmem_wait_st();
        fence_before_sync();
        nbar_arrive(kBarOps, kBarOpsCount);
      }
      SPW_PH(4);
      // ---- epilogue of tile i: relu, sign bits, receiver-segmented scan across the lanes of the warp ----
      const long long e = (long long)(blockIdx.x + i * gridDim.x) * kTM + row;
      const int r = r_cur;
      const bool rv = r >= 0;
      const int r_prev = __shfl_up_sync(0xffffffffu, r, 1), r_next = __shfl_down_sync(0xffffffffu, r, 1);
      const bool head = lane == 0 || r != r_prev, tail = lane == 31 || r != r_next;
      const uint32_t hm = __ballot_sync(0xffffffffu, head);
      const int dist = lane - (31 - __clz(hm & (0xffffffffu >> (31 - lane))));     // rows since the head of this lane's segment
      const bool far = __ballot_sync(0xffffffffu, dist >= 16) != 0u;             // warp-uniform: some segment is longer than 16 rows
      float* dst = nullptr; long long dstep = 0;         // tail lanes: where the 4-column pieces of the segment sum go
      if (tail && rv) {
        const int i0 = io_cur.x, i1 = io_cur.y;
        const long long chunk = e >> 5;
        if (e - dist == i0 && e == i1 - 1) { dst = a.H2S + (long long)r * 4; dstep = a.h_slab; }
        else if (e - dist != i0) { dst = a.part_first + chunk * kN; dstep = 4; }
        else { dst = a.part_last + chunk * kN; dstep = 4; }
      }
      // The number of scan steps is a compile-time constant of the loop body (4: segments of at most 16 rows, else 5), chosen by ONE branch
      // per tile: with a run-time step count every shuffle sat behind its own branch (200 per tile and warp) and the 40 shuffle -> add
      // chains ran one after the other; straight-line code lets the eight values of a group travel together.
      auto epilogue = [&](auto steps_tag) {
        constexpr int STEPS = decltype(steps_tag)::value;
#pragma unroll
        for (int j = 0; j < GJ; ++j) {
          const int g = q + 4 * j;
          if (g >= NKS) continue;                        // warp-uniform: columns >= 152 do not exist
          float v[8];
          uint32_t b = 0u;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float pre = __uint_as_float(d[j][k]);  // bias already inside (ones column x bias row)
            const bool on = (8 * g + k < kDE) && pre > 0.f;
            v[k] = on ? pre : 0.f;
            b |= on ? (1u << k) : 0u;
          }
          if (a.bits_h2 && rv) a.bits_h2[(long long)g * a.bits_rows + e] = (uint8_t)b;
#pragma unroll
          for (int st = 0; st < STEPS; ++st) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const float t = __shfl_up_sync(0xffffffffu, v[k], 1 << st);
              if (dist >= (1 << st)) v[k] += t;
            }
          }
          if (dst) {
            *reinterpret_cast<float4*>(dst + (long long)(2 * g) * dstep) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(dst + (long long)(2 * g + 1) * dstep) = make_float4(v[4], v[5], v[6], v[7]);
          }
        }
      };
      if (far) epilogue(std::integral_constant<int, 5>{}); else epilogue(std::integral_constant<int, 4>{});
      r_cur = r_built; io_cur = io_built;
      SPW_PH(5);                                         // p5: epilogue
    }
#ifdef SPW_PHASE_TIMING
    if (blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 9))
      printf("k_edge_step_c warp %d: build %lld waitC %lld lo %lld waitM %lld hi+D %lld epi %lld loop %lld (%d tiles)\n", warp, ph_t[0], ph_t[1], ph_t[2],
             ph_t[3], ph_t[4], ph_t[5], ph_t[7], cnt);
#endif
    if (failed && tid == 0) a.poison[0] = __int_as_float(0x7fc00000);
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, kTmemCols);
}

// nodes whose in-edges cross a 32-row chunk: H2S_i = part_last[first chunk] + sum of whole middle chunks + part_first[last chunk]
__global__ void __launch_bounds__(256) k_seg_fix_c(int n, const int32_t* __restrict__ in_off, const float* __restrict__ part_first,
                                                   const float* __restrict__ part_last, float* __restrict__ H2S, long long h_slab) {
  pdl_trigger();
  pdl_wait();
  const long long total = (long long)n * kQE;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int qd = (int)(idx / n), i = (int)(idx - (long long)qd * n);
    const int i0 = in_off[i], i1 = in_off[i + 1];
    if (i1 <= i0) {                                      // no incoming relation: the aggregate is zero
      *reinterpret_cast<float4*>(H2S + (long long)qd * h_slab + (long long)i * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
      continue;
    }
    const int k0 = i0 >> 5, k1 = (i1 - 1) >> 5;
    if (k1 == k0) continue;
    float4 s = *reinterpret_cast<const float4*>(part_last + (long long)k0 * kN + 4 * qd);
    for (int k = k0 + 1; k <= k1; ++k) {
      const float4 t = *reinterpret_cast<const float4*>(part_first + (long long)k * kN + 4 * qd);
      s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
    }
    *reinterpret_cast<float4*>(H2S + (long long)qd * h_slab + (long long)i * 4) = s;
  }
}

// =========================================================================================================================
// k_edge_dgrad_c: data gradient of the edge step, pipelined:  d h1_pre = ((relu'(h2) * dH2S[receiver]) . W2^T) * relu'(h1);
// writes DH1 and writes / accumulates dA (both [E][150] CSL).
// =========================================================================================================================
struct EdgeDgradCArgs {
  int E;
  const int32_t* in_rcv;
  const float* dH2S; long long d_slab;                  // [n][150] CSL
  const float* Whi; const float* Wlo;                   // packed W2^T operands
  const uint8_t* bits_h2; const uint8_t* bits_h1; long long bits_rows;   // byte-slab [19][bits_rows]
  float* dA; float* DH1;                                // [E][150] CSL (slab = E * 4)
  int first;                                            // dA is written (first processed step) or accumulated
  float* poison;
};

template <bool FIRST>
__global__ void __launch_bounds__(kThreadsC, 1) k_edge_dgrad_c(EdgeDgradCArgs a) {
  constexpr int KJ = 5, GJ = 5, NKS = kKS, NB = kN;
  SPW_DYN_SMEM(smem_raw);
  float* Bhi_s = reinterpret_cast<float*>(smem_raw);
  float* Blo_s = Bhi_s + kBFloats;
  uint64_t* bars = reinterpret_cast<uint64_t*>(Blo_s + kBFloats);
  uint64_t* barC = bars; uint64_t* barM = bars + 1; uint64_t* barW = bars + 2;
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bars + 3);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = 32 * (warp & 3) + lane, q = warp >> 2;

  pdl_trigger();
  if (warp == 0) tmem_alloc(tptr, kTmemCols);
  if (tid == 32) { mbar_init(barC, 1); mbar_init(barM, 1); mbar_init(barW, 1); fence_mbar_init(); }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tptr;
  constexpr uint32_t colHi = kColAhi, colLo = kColAlo, colD = kColD;
  const int ntiles = (a.E + kTM - 1) / kTM;
  const int cnt = (int)blockIdx.x < ntiles ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (warp >= kWorkers / 32) {
    regs_issuer();
    if (warp == kWorkers / 32) {
    if (lane == 0) bulk_load_weights(Bhi_s, Blo_s, a.Whi, a.Wlo, (uint32_t)kBFloats * 4, barW);
    bool ok = mbar_wait(barW, 0);
    for (int i = 0; i < cnt; ++i) {
      nbar_sync(kBarOps, kBarOpsCount);
      fence_after_sync();
      if (lane == 0) issue_tile(tmem_base + colD, tmem_base + colHi, tmem_base + colLo, smem_u32(Bhi_s), smem_u32(Blo_s), NKS, NB, barC, barM);
      __syncwarp();
    }
    if (!ok && lane == 0) a.poison[0] = __int_as_float(0x7fc00000);
    } else {
      // L2 prefetch warps (spw_csl.cuh): the relu bytes of tile i + 2 (the gathered node table is L2-resident anyway)
      const int pl = tid - (kWorkers + 32);
      pdl_wait();
      auto prefetch_tile = [&](int i) {
        const long long r0 = (long long)(blockIdx.x + i * gridDim.x) * kTM;
        prefetch_bits(a.bits_h2, a.bits_rows, NKS, r0, pl);
        prefetch_bits(a.bits_h1, a.bits_rows, NKS, r0, pl);
      };
      if (cnt > 1) prefetch_tile(1);
      for (int i = 0; i + 2 < cnt; ++i) {
        mbar_wait(barC, (uint32_t)i & 1u);
        prefetch_tile(i + 2);
      }
    }
  } else {
    regs_workers();
    const uint32_t lane_addr = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16);
    const long long es = (long long)a.E * 4;
    bool failed = false;
    XR<KJ> x;
    int r_nx = -1;
    SPW_PH_DECL
    auto load_idx = [&](int i, int& r) {
      const long long e = (long long)(blockIdx.x + i * gridDim.x) * kTM + row;
      r = (i < cnt && e < a.E) ? a.in_rcv[e] : -1;
    };
    // Loads are issued in batches well before their first use (the kernel was long-scoreboard bound: ncu 22 stalled warps per
    // issue): the gathered operand rows and their mask bytes at the top of an iteration, consumed after the wait for the
    // correction MMAs; the old dA values and the h1 mask bytes before D is read, consumed by the epilogue.
    uint32_t b2[KJ];
    auto gather_x = [&](int i, int r) {                  // raw dH2S[receiver] quads + relu'(h2) bytes of local tile i
      const long long e = (long long)(blockIdx.x + i * gridDim.x) * kTM + row;
      const bool rv = r >= 0;
      const float* dp = a.dH2S + (long long)(2 * q) * a.d_slab + (long long)(rv ? r : 0) * 4;
      const uint8_t* bp = a.bits_h2 + (long long)q * a.bits_rows + (rv ? e : 0);
#pragma unroll
      for (int j = 0; j < KJ; ++j) {
        const int ks = q + 4 * j;
        float4 t0 = make_float4(0.f, 0.f, 0.f, 0.f), t1 = t0;
        b2[j] = 0u;
        if (ks < NKS && rv) {
          t0 = *reinterpret_cast<const float4*>(dp + (long long)(8 * j) * a.d_slab);
          t1 = *reinterpret_cast<const float4*>(dp + (long long)(8 * j + 1) * a.d_slab);
          b2[j] = bp[(long long)(4 * j) * a.bits_rows];
        }
        x.v[j][0] = t0.x; x.v[j][1] = t0.y; x.v[j][2] = t0.z; x.v[j][3] = t0.w;
        x.v[j][4] = t1.x; x.v[j][5] = t1.y; x.v[j][6] = t1.z; x.v[j][7] = t1.w;
      }
    };
    auto mask_x = [&]() {
#pragma unroll
      for (int j = 0; j < KJ; ++j)
#pragma unroll
        for (int k = 0; k < 8; ++k) x.v[j][k] = ((b2[j] >> k) & 1u) ? x.v[j][k] : 0.f;
    };
    pdl_wait();
    load_idx(0, r_nx);
#pragma unroll 1
    for (int i = -1; i < cnt; ++i) {
      const bool has_next = i + 1 < cnt;
      const uint32_t parity = (uint32_t)i & 1u;
      SPW_PH(7);
      if (has_next) {
        gather_x(i + 1, r_nx);
        load_idx(i + 2, r_nx);
      }
      SPW_PH(0);                                         // p0: issue of the gathers
      if (i >= 0) {
        if (!mbar_wait(barC, parity)) failed = true;
        fence_after_sync();
      }
      SPW_PH(1);                                         // p1: wait for the correction MMAs
      if (has_next) { mask_x(); store_lo<KJ>(x, lane_addr, colLo, q, NKS); }
      SPW_PH(2);                                         // p2: mask + lo words (first use of the gathers)
      if (i >= 0) {
        if (!mbar_wait(barM, parity)) failed = true;
        fence_after_sync();
      }
      SPW_PH(3);                                         // p3: wait for the main MMAs
      if (has_next) store_hi<KJ>(x, lane_addr, colHi, q, NKS);
      // epilogue inputs of tile i: relu'(h1) bytes and (unless this is the first processed step) the old dA values
      const long long e = (long long)(blockIdx.x + i * gridDim.x) * kTM + row;
      const bool ev = i >= 0 && e < a.E;
      float* hp = a.DH1 + (long long)(2 * q) * es + (ev ? e : 0) * 4;
      float* gp = a.dA + (long long)(2 * q) * es + (ev ? e : 0) * 4;
      uint32_t b1[GJ];
      if (ev) {
        const uint8_t* bp = a.bits_h1 + (long long)q * a.bits_rows + e;
#pragma unroll
        for (int j = 0; j < GJ; ++j) {
          b1[j] = 0u;
          if (q + 4 * j < NKS) {
            b1[j] = bp[(long long)(4 * j) * a.bits_rows];
          }
        }
      }
      uint32_t d[GJ][8];
      if (i >= 0) load_d<GJ>(d, lane_addr, colD, q, NB / 8);
      if (has_next) {
        tmem_wait_st();
        fence_before_sync();
        nbar_arrive(kBarOps, kBarOpsCount);
      }
      SPW_PH(4);                                         // p4: hi words, epilogue loads issued, D load
      // ---- epilogue of tile i: mask with relu'(h1), DH1 (write), dA (write or accumulate): coalesced 16-byte accesses
      if (ev) {
#pragma unroll
        for (int j = 0; j < GJ; ++j) {
          if (q + 4 * j >= NKS) continue;
          float v[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) v[k] = ((b1[j] >> k) & 1u) ? __uint_as_float(d[j][k]) : 0.f;
          if (a.DH1) {                                   // null in the last processed step: nothing sums d h1 by sender / receiver there
            *reinterpret_cast<float4*>(hp + (long long)(8 * j) * es) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(hp + (long long)(8 * j + 1) * es) = make_float4(v[4], v[5], v[6], v[7]);
          }
          if (FIRST) {
            *reinterpret_cast<float4*>(gp + (long long)(8 * j) * es) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(gp + (long long)(8 * j + 1) * es) = make_float4(v[4], v[5], v[6], v[7]);
          } else {                                       // dA += d h1: reduction at the L2, nothing comes back to the SM
            red_add_v4(gp + (long long)(8 * j) * es, v[0], v[1], v[2], v[3]);
            red_add_v4(gp + (long long)(8 * j + 1) * es, v[4], v[5], v[6], v[7]);
          }
        }
      }
      SPW_PH(5);                                         // p5: epilogue
    }
#ifdef SPW_PHASE_TIMING
    if (blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 9))
      printf("k_edge_dgrad_c<%d> warp %d: gathers %lld waitC %lld mask+lo %lld waitM %lld hi+D %lld epi %lld loop %lld (%d tiles)\n", (int)FIRST, warp, ph_t[0],
             ph_t[1], ph_t[2], ph_t[3], ph_t[4], ph_t[5], ph_t[7], cnt);
#endif
    if (failed && tid == 0) a.poison[0] = __int_as_float(0x7fc00000);
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, kTmemCols);
}

// dS_i = sum over out-edges, dR_i = sum over in-edges of DH1 (fixed order); thread = (quad, node)
__global__ void __launch_bounds__(256) k_gather_dsr_c(int n, int E, const int32_t* __restrict__ in_off, const int32_t* __restrict__ out_off,
                                                      const int32_t* __restrict__ out_pos, const float* __restrict__ DH1,
                                                      float* __restrict__ dS, float* __restrict__ dR, long long sr_slab) {
  pdl_trigger();
  pdl_wait();
  const long long total = (long long)n * kQE;
  const long long es = (long long)E * 4;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int qd = (int)(idx / n), i = (int)(idx - (long long)qd * n);
    const int i0 = in_off[i], i1 = in_off[i + 1], o0 = out_off[i], o1 = out_off[i + 1];
    const float* src = DH1 + (long long)qd * es;
    // fixed summation order (slot order); the loads of four relations are in flight together
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    int e = i0;
    for (; e + 4 <= i1; e += 4) {
      const float4 v0 = *reinterpret_cast<const float4*>(src + (long long)e * 4), v1 = *reinterpret_cast<const float4*>(src + (long long)(e + 1) * 4);
      const float4 v2 = *reinterpret_cast<const float4*>(src + (long long)(e + 2) * 4), v3 = *reinterpret_cast<const float4*>(src + (long long)(e + 3) * 4);
      s.x = (((s.x + v0.x) + v1.x) + v2.x) + v3.x; s.y = (((s.y + v0.y) + v1.y) + v2.y) + v3.y;
      s.z = (((s.z + v0.z) + v1.z) + v2.z) + v3.z; s.w = (((s.w + v0.w) + v1.w) + v2.w) + v3.w;
    }
    for (; e < i1; ++e) {
      const float4 v = *reinterpret_cast<const float4*>(src + (long long)e * 4);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    *reinterpret_cast<float4*>(dR + (long long)qd * sr_slab + (long long)i * 4) = s;
    s = make_float4(0.f, 0.f, 0.f, 0.f);
    int o = o0;
    for (; o + 4 <= o1; o += 4) {
      const int p0 = out_pos[o], p1 = out_pos[o + 1], p2 = out_pos[o + 2], p3 = out_pos[o + 3];
      const float4 v0 = *reinterpret_cast<const float4*>(src + (long long)p0 * 4), v1 = *reinterpret_cast<const float4*>(src + (long long)p1 * 4);
      const float4 v2 = *reinterpret_cast<const float4*>(src + (long long)p2 * 4), v3 = *reinterpret_cast<const float4*>(src + (long long)p3 * 4);
      s.x = (((s.x + v0.x) + v1.x) + v2.x) + v3.x; s.y = (((s.y + v0.y) + v1.y) + v2.y) + v3.y;
      s.z = (((s.z + v0.z) + v1.z) + v2.z) + v3.z; s.w = (((s.w + v0.w) + v1.w) + v2.w) + v3.w;
    }
    for (; o < o1; ++o) {
      const float4 v = *reinterpret_cast<const float4*>(src + (long long)out_pos[o] * 4);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    *reinterpret_cast<float4*>(dS + (long long)qd * sr_slab + (long long)i * 4) = s;
  }
}

// out[i][:] = sum over `slots` of in[slot * n + i][:]  (column-slab arrays; thread = (quad, node), fixed order)
__global__ void __launch_bounds__(256) k_sum_slots_c(int n, int slots, int nquads, const float* __restrict__ in, long long in_slab,
                                                     float* __restrict__ out, long long out_slab) {
  pdl_trigger();
  pdl_wait();
  const long long total = (long long)n * nquads;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int qd = (int)(idx / n), i = (int)(idx - (long long)qd * n);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int l = 0; l < slots; ++l) {
      const float4 v = *reinterpret_cast<const float4*>(in + (long long)qd * in_slab + ((long long)l * n + i) * 4);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    *reinterpret_cast<float4*>(out + (long long)qd * out_slab + (long long)i * 4) = s;
  }
}

// =========================================================================================================================
// Skinny weight gradients: sums over rows of  s0 * Z[row][:],  s1 * Z[row][:]  and  Z[row][:]  with per-row scalars s0, s1.
//   MODE 0  relation-encoder layer 0:  Z = G0 [E][150], (s0, s1) = pos_receiver - pos_sender       -> d rm_w0 [2][150], d rm_b0
//   MODE 1  object-encoder layer 0:    Z = dQ1 [n][100], (s0, s1) = (y, w)                           -> d om_w0 [2][100], d om_b0
//   MODE 2  head:                      Z = U5 [n][100], s0 = dlogit                                   -> d V2[:,0] [100], d c2[0]
// A warp owns 256 consecutive rows (lane = row within a 32-row chunk, 8 chunks), walks the column quads, and reduces
// across lanes with a fixed shuffle tree: partial [warp][3][ld] (MODE 2: [warp][ld], element 100 = sum of s0).
// =========================================================================================================================
constexpr int kSkinnyRows = 256;
template <int MODE>
__global__ void __launch_bounds__(256) k_skinny_c(int M, const float* __restrict__ Z, long long z_slab, int nquads, int ld,
                                                  const int32_t* __restrict__ snd, const int32_t* __restrict__ rcv,
                                                  const float* __restrict__ obj, const float* __restrict__ s0v, float* __restrict__ part) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nw = ((long long)M + kSkinnyRows - 1) / kSkinnyRows;
  if (gw >= nw) return;
  float a0[8], a1[8];
  long long rowk[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const long long r = gw * kSkinnyRows + 32 * k + lane;
    rowk[k] = r < M ? r : -1;
    a0[k] = 0.f; a1[k] = 0.f;
    if (r < M) {
      if (MODE == 0) {
        const int s = snd[r], rc = rcv[r];
        a0[k] = obj[3 * (size_t)rc] - obj[3 * (size_t)s]; a1[k] = obj[3 * (size_t)rc + 1] - obj[3 * (size_t)s + 1];
      } else if (MODE == 1) {
        a0[k] = obj[3 * (size_t)r + 1]; a1[k] = obj[3 * (size_t)r + 2];
      } else {
        a0[k] = s0v[r];
      }
    }
  }
  float* p = part + gw * (MODE == 2 ? ld : 3 * ld);
  auto wsum = [&](float v) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    return v;
  };
  // the column quads are dealt to gridDim.y blocks (each recomputes the two input features of its rows): with one block per 2048 rows the
  // largest launch (rm layer 0, all relations) ran 8 warps per SM and was latency bound
  const int qper = (nquads + (int)gridDim.y - 1) / (int)gridDim.y;
  const int q_lo = (int)blockIdx.y * qper, q_hi = q_lo + qper < nquads ? q_lo + qper : nquads;
  for (int qd = q_lo; qd < q_hi; ++qd) {
    float g0[4] = {0.f, 0.f, 0.f, 0.f}, g1[4] = {0.f, 0.f, 0.f, 0.f}, gb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (rowk[k] >= 0) {
        const float4 z = *reinterpret_cast<const float4*>(Z + (long long)qd * z_slab + rowk[k] * 4);
        const float zz[4] = {z.x, z.y, z.z, z.w};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          g0[c] = fmaf(a0[k], zz[c], g0[c]);
          if (MODE != 2) { g1[c] = fmaf(a1[k], zz[c], g1[c]); gb[c] += zz[c]; }
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float t0 = wsum(g0[c]);
      if (MODE != 2) {
        const float t1 = wsum(g1[c]), tb = wsum(gb[c]);
        if (lane == 0) { p[4 * qd + c] = t0; p[ld + 4 * qd + c] = t1; p[2 * ld + 4 * qd + c] = tb; }
      } else if (lane == 0) {
        p[4 * qd + c] = t0;
      }
    }
  }
  if (MODE == 2 && blockIdx.y == 0) {
    float sb = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) sb += a0[k];
    sb = wsum(sb);
    if (lane == 0) p[4 * nquads] = sb;
  }
}

}  // namespace csl
}  // namespace spw
#endif  // SPW_EMU
