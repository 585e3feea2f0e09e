// spw_tc.cuh -- Blackwell tensor-core (tcgen05 / TMEM) building blocks, sm_100a only.
//
// fp32-accurate GEMM on the 5th-generation tensor cores by the 3xTF32 split:
//     x = hi + lo,  hi = rna_tf32(x),  lo = rna_tf32(x - hi)       (|x - hi - lo| <= 2^-22 |x|)
//     A.B ~= A_hi.B_hi + A_lo.B_hi + A_hi.B_lo                     (fp32 accumulate in TMEM)
// Operand placement (tcgen05.mma kind::tf32, cta_group::1, M = 128, N = 160, K = 8 per instruction):
//   A (activations, 128 rows = 128 TMEM lanes): in TENSOR MEMORY, one 32-bit column per k element;
//       written straight from registers with tcgen05.st -- activations never touch shared memory;
//   B (weights): in shared memory, K-major, no swizzle ("interleave" canonical layout): per k-step
//       [chunk c = 0,1][n = 0..159][4 floats], i.e. core matrices of 8 rows x 16 bytes,
//       leading-dimension (K) byte offset 2560, stride (N) byte offset 128;
//   D (accumulator): TMEM, 160 columns.
// TMEM map (512 columns): [0,160) A_hi, [160,320) A_lo, [320,480) D.
#pragma once
#ifndef SPW_EMU
#include "spw_common.cuh"

// Per-phase clock64() accounting of the tile loops (development builds only: -DSPW_PHASE_TIMING; tools/phase_probe.py)
#ifdef SPW_PHASE_TIMING
#include <stdio.h>
#define SPW_PH_DECL long long ph_t[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long ph_last = clock64();
#define SPW_PH(i) do { const long long ph_now = clock64(); ph_t[i] += ph_now - ph_last; ph_last = ph_now; } while (0)
#define SPW_PH_REPORT(name) do { if (blockIdx.x == 0 && (threadIdx.x == 0 || threadIdx.x == 255)) \
  printf("%s tid %d: p0 %lld p1 %lld p2 %lld p3 %lld p4 %lld p5 %lld p6 %lld p7 %lld\n", name, \
         (int)threadIdx.x, ph_t[0], ph_t[1], ph_t[2], ph_t[3], ph_t[4], ph_t[5], ph_t[6], ph_t[7]); } while (0)
#else
#define SPW_PH_DECL
#define SPW_PH(i)
#define SPW_PH_REPORT(name)
#endif

namespace spw {
namespace tc {

constexpr int kN = 160;                 // MMA N (150 valid columns)
constexpr int kKS = 19;                 // k-steps of 8 (K = 152, rows 150/151 zero)
constexpr int kBStepFloats = 2 * kN * 4;            // floats per k-step of a packed B matrix (1280)
constexpr int kBFloats = kKS * kBStepFloats;        // 24320 floats = 97280 bytes per hi or lo matrix
constexpr uint32_t kColAhi = 0, kColAlo = 160, kColD = 320;
constexpr uint32_t kTmemCols = 512;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- TMEM allocation (one warp) ------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols));
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- TMEM <-> registers (thread = TMEM lane; warp w may touch lanes [32*(w%4), +32)) --------------
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- 3xTF32 split ---------------------------------------------------------------------------------
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
  const float r = x - __uint_as_float(hi);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}

// ---- descriptors ----------------------------------------------------------------------------------
// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version = 1 [46,48), layout type [61,64) = 0 (no swizzle)
__device__ __forceinline__ uint64_t make_b_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major
__device__ __forceinline__ constexpr uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[tmem] . B[smem]   (issued by ONE thread)
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
// make the mbarrier track completion of all MMAs issued so far by this thread
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- mbarrier --------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// bounded wait: returns false on timeout instead of hanging the GPU
#ifdef SPW_WAIT_DEBUG
constexpr int kWaitSpin = 1 << 16;      // development builds (tools/build_phase.sh -DSPW_WAIT_DEBUG): fail fast
#else
constexpr int kWaitSpin = 1 << 22;
#endif
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
#pragma unroll 1
  for (int it = 0; it < kWaitSpin; ++it) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
    if (ok) return true;
  }
  return false;
}

// issue the 3 x kKS MMAs of one 128-row tile: D = (A_lo.B_hi + A_hi.B_lo) + A_hi.B_hi.
// The tensor core truncates when it adds into the fp32 accumulator, so the ORDER matters: the 38
// correction products (2^-11 of the result) go first, while the accumulator is still tiny; only
// the 19 main products accumulate at full magnitude (measured: 3x smaller error than interleaving).
__device__ __forceinline__ void issue_tile_mmas(uint32_t tmem_base, const float* Bhi_s, const float* Blo_s) {
  const uint32_t idesc = make_idesc_tf32(128, kN);
  const uint32_t bhi = smem_u32(Bhi_s), blo = smem_u32(Blo_s);
#pragma unroll 1
  for (int ks = 0; ks < kKS; ++ks) {
    const uint64_t dhi = make_b_desc(bhi + ks * (kBStepFloats * 4), kN * 16, 128);
    const uint64_t dlo = make_b_desc(blo + ks * (kBStepFloats * 4), kN * 16, 128);
    mma_tf32_ts(tmem_base + kColD, tmem_base + kColAlo + 8 * ks, dhi, idesc, ks > 0 ? 1u : 0u);
    mma_tf32_ts(tmem_base + kColD, tmem_base + kColAhi + 8 * ks, dlo, idesc, 1u);
  }
#pragma unroll 1
  for (int ks = 0; ks < kKS; ++ks) {
    const uint64_t dhi = make_b_desc(bhi + ks * (kBStepFloats * 4), kN * 16, 128);
    mma_tf32_ts(tmem_base + kColD, tmem_base + kColAhi + 8 * ks, dhi, idesc, 1u);
  }
}

// ---- weight packing: Keras W[K][N] (row-major, ld) -> hi / lo B operands in the layout above -------
// bias (may be null): stored as row k == K of the operand, picked up by a ones column in A.
__global__ void __launch_bounds__(256) k_pack_umma(const float* __restrict__ W, int ld, int row0, int col0, int K, int N,
                                                   int transpose, const float* __restrict__ bias, float* __restrict__ hi,
                                                   float* __restrict__ lo) {
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < kBFloats; idx += gridDim.x * blockDim.x) {
    const int e = idx & 3, n = (idx >> 2) % kN, c = ((idx >> 2) / kN) & 1, ks = (idx >> 2) / (2 * kN);
    const int k = 8 * ks + 4 * c + e;
    float v = 0.f;
    if (k < K && n < N) v = transpose ? W[(size_t)(row0 + n) * ld + col0 + k] : W[(size_t)(row0 + k) * ld + col0 + n];
    else if (bias && k == K && n < N) v = bias[n];
    uint32_t h, l;
    split_tf32(v, h, l);
    hi[idx] = __uint_as_float(h);
    lo[idx] = __uint_as_float(l);
  }
}

// ---- self test: D[128][160] = A[128][152] . B (packed hi/lo), one CTA of 128 threads ----------------
__global__ void __launch_bounds__(128, 1) k_tc_selftest(const float* __restrict__ A, const float* __restrict__ Bhi,
                                                        const float* __restrict__ Blo, float* __restrict__ D,
                                                        int* __restrict__ status) {
  SPW_DYN_SMEM(smem_raw);
  float* Bhi_s = reinterpret_cast<float*>(smem_raw);
  float* Blo_s = Bhi_s + kBFloats;
  uint64_t* bar = reinterpret_cast<uint64_t*>(Blo_s + kBFloats);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(tptr, kTmemCols);
  if (tid == 32) { mbar_init(bar, 1); fence_mbar_init(); }
  for (int i = tid; i < kBFloats / 4; i += 128) {
    reinterpret_cast<float4*>(Bhi_s)[i] = reinterpret_cast<const float4*>(Bhi)[i];
    reinterpret_cast<float4*>(Blo_s)[i] = reinterpret_cast<const float4*>(Blo)[i];
  }
  fence_async_smem();            // generic-proxy smem writes -> visible to the tensor core (async proxy)
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tptr;
  const uint32_t lane_addr = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16);
  // A row -> hi / lo -> TMEM
  const float* arow = A + (size_t)tid * kDEP;
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
  __syncthreads();
  if (tid == 0) {
    fence_after_sync();
    issue_tile_mmas(tmem_base, Bhi_s, Blo_s);
    mma_commit(bar);
  }
  const bool ok = mbar_wait(bar, 0);
  fence_after_sync();
  if (!ok) {
    if (tid == 0) *status = -1;
  } else {
#pragma unroll 1
    for (int c = 0; c < kN; c += 16) {
      uint32_t v[16];
      tmem_ld16(lane_addr + kColD + c, v);
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 16; ++i) D[(size_t)tid * kN + c + i] = __uint_as_float(v[i]);
    }
    if (tid == 0) *status = 1;
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, kTmemCols);
}

// =================================================================================================
// K2b on the tensor cores: per step -- gather, hidden layer 2 (3xTF32 tcgen05), relu, relu bits,
// deterministic receiver-segmented sum.  Persistent, one CTA per SM, 128-edge tiles.
//   shared memory: W2 hi/lo resident for the whole kernel (190 KB) + a 128 x 64 column staging slab
//   tensor memory: A_hi | A_lo | D as in the map above
//   thread = (row = TMEM lane, column half); warps w and w+4 share a lane quarter
// =================================================================================================
constexpr int kStagePitch = 66;      // 8-byte aligned rows; row-thread 64-bit accesses are conflict-free
constexpr int kStageCols = 64;

struct EdgeStepTcArgs {
  int E;
  const int32_t* in_snd; const int32_t* in_rcv; const int32_t* in_off;
  const float* A; const float* S; const float* R;     // [E][152], [n][152], [n][152]
  const float* W2hi; const float* W2lo;               // packed B operands (k_pack_umma)
  float* H2S;                                         // [n][152]
  float* part_first; float* part_last;                // [ntiles][152] (tiles of 128 edges)
  uint32_t* maskbits;                                 // [E][8] or null: relu bits of h2, bit (col & 31) of word (col >> 5)
  uint32_t* maskbits_h1;                              // [E][8] or null: relu bits of h1 (same layout)
};

constexpr size_t kEdgeStepTcSmem = (size_t)(2 * kBFloats + kTM * kStagePitch + 2 * kTM + (kTM + 8) / 2 + kTM * 5) * sizeof(float) + 16;
static_assert(kEdgeStepTcSmem <= 232448, "k_edge_step_tc shared memory exceeds the 227 KB per-CTA limit");

// coalesced gather of one 64-column slab of h1 = relu(A_e + S_s + R_r): half a warp per row, 128-bit loads
// along the row, all eight row-pairs of the warp (24 loads per lane) in flight; split into a load half
// (registers) and a store half (slab) so that the loads of slab s+1 overlap the processing of slab s.
constexpr int kStThreads = 512;      // four threads per row (one 16-column quarter of every slab each): more warps, less serial work per warp
struct H1Regs { float4 va[4], vs[4], vr[4]; };

__device__ __forceinline__ void h1_slab_load(H1Regs& g, const int* ssnd, const int* srcv, const float* __restrict__ A,
                                             const float* __restrict__ S, const float* __restrict__ R, int e0, int c0,
                                             int ncols) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane >> 4, c4 = lane & 15;
  const bool col_ok = 4 * c4 < ncols;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int r = warp * 8 + 2 * j + sub;
    const int rc = srcv[r];
    if (rc >= 0 && col_ok) {
      g.va[j] = *reinterpret_cast<const float4*>(A + (size_t)(e0 + r) * kDEP + c0 + 4 * c4);
      g.vs[j] = *reinterpret_cast<const float4*>(S + (size_t)ssnd[r] * kDEP + c0 + 4 * c4);
      g.vr[j] = *reinterpret_cast<const float4*>(R + (size_t)rc * kDEP + c0 + 4 * c4);
    } else {
      g.va[j] = z4; g.vs[j] = z4; g.vr[j] = z4;
    }
  }
}

__device__ __forceinline__ void h1_slab_store(const H1Regs& g, float* stage, const int* srcv, int c0, int ncols) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane >> 4, c4 = lane & 15;
  if (4 * c4 >= ncols) return;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int r = warp * 8 + 2 * j + sub;
    float2* dst = reinterpret_cast<float2*>(stage + r * kStagePitch + 4 * c4);
    dst[0] = make_float2(relu_f(g.va[j].x + g.vs[j].x + g.vr[j].x), relu_f(g.va[j].y + g.vs[j].y + g.vr[j].y));
    if (c0 + 4 * c4 == kDE - 2)     // columns 150 / 151: the ones column that picks up the bias row, and the pad
      dst[1] = make_float2(srcv[r] >= 0 ? 1.f : 0.f, 0.f);
    else
      dst[1] = make_float2(relu_f(g.va[j].z + g.vs[j].z + g.vr[j].z), relu_f(g.va[j].w + g.vs[j].w + g.vr[j].w));
  }
}

__global__ void __launch_bounds__(kStThreads, 1) k_edge_step_tc(EdgeStepTcArgs a) {
  SPW_DYN_SMEM(smem_raw);
  float* Bhi_s = reinterpret_cast<float*>(smem_raw);
  float* Blo_s = Bhi_s + kBFloats;
  float* stage = Blo_s + kBFloats;
  int* srcv = reinterpret_cast<int*>(stage + kTM * kStagePitch);
  int* ssnd = srcv + kTM;
  short* snoff = reinterpret_cast<short*>(ssnd + kTM);   // in_off - e0 of the tile's nodes, clamped (kTM + 8 entries)
  uint16_t* smask = reinterpret_cast<uint16_t*>(snoff + kTM + 8);   // relu bits, 10 half-words per row
  uint64_t* bar = reinterpret_cast<uint64_t*>(smask + kTM * 10);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = 32 * (warp & 3) + lane, q = warp >> 2;  // TMEM lane, column quarter of a slab

  if (warp == 0) tmem_alloc(tptr, kTmemCols);
  if (tid == 32) { mbar_init(bar, 1); fence_mbar_init(); }
  for (int i = tid; i < kBFloats / 4; i += kStThreads) {
    reinterpret_cast<float4*>(Bhi_s)[i] = reinterpret_cast<const float4*>(a.W2hi)[i];
    reinterpret_cast<float4*>(Blo_s)[i] = reinterpret_cast<const float4*>(a.W2lo)[i];
  }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tptr;
  const uint32_t lane_addr = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16);
  uint32_t parity = 0;
  bool failed = false;
  const int ntiles = (a.E + kTM - 1) / kTM;

  SPW_PH_DECL
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    SPW_PH(7);
    const int e0 = tile * kTM;
    const int rows = imin(kTM, a.E - e0);
    if (tid < kTM) {
      const bool valid = tid < rows;
      srcv[tid] = valid ? a.in_rcv[e0 + tid] : -1;
      ssnd[tid] = valid ? a.in_snd[e0 + tid] : 0;
    }
    __syncthreads();
    const int n_first = srcv[0], n_last = srcv[rows - 1];
    const int nnodes = n_last - n_first + 1;
    for (int i = tid; i <= imin(nnodes, kTM + 7); i += kStThreads)
      snoff[i] = (short)imax(-32000, imin(32000, a.in_off[n_first + i] - e0));
    SPW_PH(0);                                            // p0: indices
    // ---- h1 = relu(A_e + S_s + R_r): coalesced gather into the slab, then row threads split it
    //      into tf32 hi/lo and store it to tensor memory (3 slabs of <= 64 columns)
    H1Regs hreg;
    h1_slab_load(hreg, ssnd, srcv, a.A, a.S, a.R, e0, 0, kStageCols);
#pragma unroll
    for (int sl = 0; sl < 3; ++sl) {
      const int c0 = sl * kStageCols;
      const int ncols = imin(kStageCols, kDEP - c0);
      h1_slab_store(hreg, stage, srcv, c0, ncols);
      __syncthreads();
      if (sl < 2) h1_slab_load(hreg, ssnd, srcv, a.A, a.S, a.R, e0, c0 + kStageCols, imin(kStageCols, kDEP - c0 - kStageCols));
      // this thread's quarter of the slab row: 16 columns (or what is left) == half a word of relu bits
      const int cb = 16 * q;
      uint32_t hbits = 0u;
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const int c = cb + 8 * g;
        if (c < ncols) {                                  // warp-uniform
          const float2* src = reinterpret_cast<const float2*>(stage + row * kStagePitch + c);
          const float2 p0 = src[0], p1 = src[1], p2 = src[2], p3 = src[3];
          const float x[8] = {p0.x, p0.y, p1.x, p1.y, p2.x, p2.y, p3.x, p3.y};
          uint32_t h[8], l[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            split_tf32(x[i], h[i], l[i]);
            hbits |= x[i] > 0.f ? (1u << (8 * g + i)) : 0u;
          }
          tmem_st8(lane_addr + kColAhi + c0 + c, h);
          tmem_st8(lane_addr + kColAlo + c0 + c, l);
        }
      }
      if (a.maskbits_h1 && row < rows && cb < ncols)
        reinterpret_cast<uint16_t*>(a.maskbits_h1)[(size_t)(e0 + row) * 16 + ((c0 + cb) >> 4)] = (uint16_t)hbits;
      __syncthreads();
    }
    tmem_wait_st();
    fence_before_sync();
    __syncthreads();
    SPW_PH(1);                                            // p1: gather + split + STTM
    if (tid == 0) {
      fence_after_sync();
      issue_tile_mmas(tmem_base, Bhi_s, Blo_s);
      mma_commit(bar);
    }
    if (!mbar_wait(bar, parity)) failed = true;
    parity ^= 1u;
    fence_after_sync();
    SPW_PH(2);                                            // p2: MMA
    // ---- epilogue: D -> +b2, relu, relu bits -> staging slab -> receiver-segmented sum
    for (int c0 = 0; c0 < kN; c0 += kStageCols) {
      const int ncols = imin(kStageCols, kN - c0);
      if (16 * q < ncols) {                               // warp-uniform: this thread's 16-column block of the slab
        uint32_t v[16];
        tmem_ld16(lane_addr + kColD + c0 + 16 * q, v);
        tmem_wait_ld();
        uint32_t m16 = 0u;
        float o[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float pre = __uint_as_float(v[i]);          // bias already inside (ones column x bias row)
          const bool on = (c0 + 16 * q + i < kDE) && (pre > 0.f);
          o[i] = on ? pre : 0.f;
          m16 |= on ? (1u << i) : 0u;
        }
        float2* dst = reinterpret_cast<float2*>(stage + row * kStagePitch + 16 * q);
#pragma unroll
        for (int i = 0; i < 8; ++i) dst[i] = make_float2(o[2 * i], o[2 * i + 1]);
        smask[row * 10 + ((c0 + 16 * q) >> 4)] = (uint16_t)m16;
      }
      __syncthreads();
      for (int item = tid; item < nnodes * ncols; item += kStThreads) {
        const int ni = item / ncols, c = item - ni * ncols;
        const int node = n_first + ni;
        int s0, s1;
        if (ni < kTM + 7) { s0 = e0 + snoff[ni]; s1 = e0 + snoff[ni + 1]; } else { s0 = a.in_off[node]; s1 = a.in_off[node + 1]; }
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
      __syncthreads();
    }
    SPW_PH(3);                                            // p3: epilogue + segmented sum
    if (a.maskbits) {
      for (int i = tid; i < rows * 5; i += kStThreads) {
        const int r = i / 5, w = i - r * 5;
        a.maskbits[(size_t)(e0 + r) * 8 + w] = (uint32_t)smask[r * 10 + 2 * w] | ((uint32_t)smask[r * 10 + 2 * w + 1] << 16);
      }
    }
    fence_before_sync();
    __syncthreads();
    SPW_PH(4);                                            // p4: mask bits + tile-end sync
  }
  SPW_PH_REPORT("k_edge_step_tc");
  if (failed && tid == 0) a.H2S[0] = __int_as_float(0x7fc00000);   // fail loudly: poison the output (MMA barrier timed out)
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, kTmemCols);
}

// =================================================================================================
// K4b (data gradient) on the tensor cores: d h1_pre = (d h2_pre . W2^T) * relu'(h1) per step.
//   A = d h2_pre = relu-bits(h2) ? dH2S[receiver] : 0   (gathered per row, split, stored to TMEM)
//   B = W2 itself, read as [N = k][K = n] (packed with transpose = 1), resident in shared memory
//   epilogue: mask with the saved relu bits of h1, stage through the slab, 128-bit coalesced
//   writes of DH1 and read-modify-write of dA.
// =================================================================================================
struct EdgeDgradTcArgs {
  int E;
  const int32_t* in_rcv;                              // row r reads dH2S[in_rcv[r]]; null: row r reads dH2S[r]
  const float* dH2S;                                  // [n][152] (or [E][152] when in_rcv is null)
  const float* Whi; const float* Wlo;                 // packed B operands ([N = out][K = in])
  const uint32_t* maskbits;                           // operand mask bits [E][8] (relu bits of h2) or null: none
  const uint32_t* maskbits_h1;                        // output mask bits [E][8] (relu bits of h1) or null: none
  const float* act;                                   // output mask by sign: keep where act[E][152] > 0 (or null)
  float scale;                                        // multiplies the output (1/keep of a dropout; else 1)
  float* dA; float* DH1;                              // DH1 [E][152] written; dA (may be null) accumulated / written
  int first;                                          // dA is written (first processed step) or accumulated
  float* poison;                                      // written with NaN if an MMA barrier times out
};

constexpr size_t kEdgeDgradTcSmem = (size_t)(2 * kBFloats + kTM * kStagePitch + kTM) * sizeof(float) + 16;

// 512 threads: four threads per row (one 16-column quarter of every 64-column slab each).  With one CTA per SM more warps
// mean less serial work per warp between the barriers.
constexpr int kDgThreads = 512;

__global__ void __launch_bounds__(kDgThreads, 1) k_edge_dgrad_tc(EdgeDgradTcArgs a) {
  SPW_DYN_SMEM(smem_raw);
  float* Bhi_s = reinterpret_cast<float*>(smem_raw);
  float* Blo_s = Bhi_s + kBFloats;
  float* stage = Blo_s + kBFloats;
  int* srcv = reinterpret_cast<int*>(stage + kTM * kStagePitch);
  uint64_t* bar = reinterpret_cast<uint64_t*>(srcv + kTM);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = 32 * (warp & 3) + lane, q = warp >> 2;        // TMEM lane, column quarter of a slab

  if (warp == 0) tmem_alloc(tptr, kTmemCols);
  if (tid == 32) { mbar_init(bar, 1); fence_mbar_init(); }
  for (int i = tid; i < kBFloats / 4; i += kDgThreads) {
    reinterpret_cast<float4*>(Bhi_s)[i] = reinterpret_cast<const float4*>(a.Whi)[i];
    reinterpret_cast<float4*>(Blo_s)[i] = reinterpret_cast<const float4*>(a.Wlo)[i];
  }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tptr;
  const uint32_t lane_addr = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16);
  uint32_t parity = 0;
  bool failed = false;
  const int ntiles = (a.E + kTM - 1) / kTM;
  const int sub = lane >> 4, c4 = lane & 15;

  auto gather_slab = [&](float4 (&v)[4], int c0) {     // dH2S[receiver] rows of one 64-column slab: half a warp per row
    const int ncols = imin(kStageCols, kDEP - c0);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int rc = srcv[warp * 8 + 2 * j + sub];
      v[j] = (rc >= 0 && 4 * c4 < ncols) ? *reinterpret_cast<const float4*>(a.dH2S + (size_t)rc * kDEP + c0 + 4 * c4)
                                         : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  SPW_PH_DECL
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    SPW_PH(7);
    const int e0 = tile * kTM;
    const int rows = imin(kTM, a.E - e0);
    if (tid < kTM) srcv[tid] = tid < rows ? (a.in_rcv ? a.in_rcv[e0 + tid] : e0 + tid) : -1;
    // relu bits of this thread's row and quarter (h2: operand mask, h1: epilogue mask): word 2 sl + (q >> 1) of slab sl
    uint32_t b2w[3], b1w[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) { b2w[i] = 0xffffffffu; b1w[i] = 0xffffffffu; }
    if (row < rows && a.maskbits) {
#pragma unroll
      for (int i = 0; i < 3; ++i) b2w[i] = a.maskbits[(size_t)(e0 + row) * 8 + 2 * i + (q >> 1)];
    }
    if (row < rows && a.maskbits_h1) {
#pragma unroll
      for (int i = 0; i < 3; ++i) b1w[i] = a.maskbits_h1[(size_t)(e0 + row) * 8 + 2 * i + (q >> 1)];
    }
    __syncthreads();
    // ---- A operand: gather dH2S[receiver] rows (coalesced, half a warp per row), mask, split, TMEM.
    //      The gather of slab s+1 is in flight while slab s is split and stored.
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
      __syncthreads();
      if (sl < 2) gather_slab(v, c0 + kStageCols);
      const int cb = 16 * q;
      if (cb < ncols) {                                    // warp-uniform
        const uint32_t bits = b2w[sl] >> (16 * (q & 1));
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const int c = cb + 8 * g;
          if (c < ncols) {
            const float2* src = reinterpret_cast<const float2*>(stage + row * kStagePitch + c);
            const float2 p0 = src[0], p1 = src[1], p2 = src[2], p3 = src[3];
            const float x[8] = {p0.x, p0.y, p1.x, p1.y, p2.x, p2.y, p3.x, p3.y};
            uint32_t h[8], l[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const bool on = (bits >> (8 * g + i)) & 1u;
              split_tf32(on ? x[i] : 0.f, h[i], l[i]);
            }
            tmem_st8(lane_addr + kColAhi + c0 + c, h);
            tmem_st8(lane_addr + kColAlo + c0 + c, l);
          }
        }
      }
      __syncthreads();
    }
    tmem_wait_st();
    fence_before_sync();
    __syncthreads();
    SPW_PH(1);                                            // p1: indices, bits, gather + split + STTM
    if (tid == 0) {
      fence_after_sync();
      issue_tile_mmas(tmem_base, Bhi_s, Blo_s);
      mma_commit(bar);
    }
    if (!mbar_wait(bar, parity)) failed = true;
    parity ^= 1u;
    fence_after_sync();
    SPW_PH(2);                                            // p2: MMA
    // ---- epilogue: D * relu'(h1) -> slab -> DH1 (write) and dA (write or accumulate), coalesced
#pragma unroll
    for (int sl = 0; sl < 3; ++sl) {
      const int c0 = sl * kStageCols;
      const int ncols = imin(kStageCols, kDEP - c0);       // columns 152..159 are never stored
      if (16 * q < imin(kStageCols, kN - c0)) {            // warp-uniform: this thread's 16-column block of the slab
        uint32_t vv[16];
        tmem_ld16(lane_addr + kColD + c0 + 16 * q, vv);
        tmem_wait_ld();
        const int col0 = c0 + 16 * q;
        const uint32_t bits = b1w[sl] >> (16 * (q & 1));
        float o[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) o[i] = ((bits >> i) & 1u) && (col0 + i < kDE) ? __uint_as_float(vv[i]) : 0.f;
        float2* dst = reinterpret_cast<float2*>(stage + row * kStagePitch + 16 * q);
#pragma unroll
        for (int i = 0; i < 8; ++i) dst[i] = make_float2(o[2 * i], o[2 * i + 1]);
      }
      __syncthreads();
      const int n4 = ncols >> 2;                           // float4 per row in this slab (16, 16, 6)
      const int total = rows * n4;
      for (int base = 0; base < total; base += 4 * kDgThreads) {
        float4 old[4], actv[4];
        const bool rmw = a.dA && !a.first;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int idx = base + u * kDgThreads + tid;
          if (idx < total) {
            const int r = idx / n4, qq = idx - r * n4;
            const size_t g = (size_t)(e0 + r) * kDEP + c0 + 4 * qq;
            if (rmw) old[u] = *reinterpret_cast<const float4*>(a.dA + g);
            if (a.act) actv[u] = *reinterpret_cast<const float4*>(a.act + g);
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int idx = base + u * kDgThreads + tid;
          if (idx < total) {
            const int r = idx / n4, qq = idx - r * n4;
            const float2* src = reinterpret_cast<const float2*>(stage + r * kStagePitch + 4 * qq);
            const float2 p0 = src[0], p1 = src[1];
            float4 val = make_float4(p0.x * a.scale, p0.y * a.scale, p1.x * a.scale, p1.y * a.scale);
            if (a.act) {
              val.x = actv[u].x > 0.f ? val.x : 0.f; val.y = actv[u].y > 0.f ? val.y : 0.f;
              val.z = actv[u].z > 0.f ? val.z : 0.f; val.w = actv[u].w > 0.f ? val.w : 0.f;
              if (c0 + 4 * qq == kDE - 2) { val.z = 0.f; val.w = 0.f; }      // columns 150 / 151 carry no gradient
            }
            const size_t g = (size_t)(e0 + r) * kDEP + c0 + 4 * qq;
            *reinterpret_cast<float4*>(a.DH1 + g) = val;
            if (a.dA) {
              if (rmw) { val.x += old[u].x; val.y += old[u].y; val.z += old[u].z; val.w += old[u].w; }
              *reinterpret_cast<float4*>(a.dA + g) = val;
            }
          }
        }
      }
      __syncthreads();
    }
    fence_before_sync();
    __syncthreads();
    SPW_PH(3);                                            // p3: epilogue
  }
  SPW_PH_REPORT("k_edge_dgrad_tc");
  if (failed && tid == 0) a.poison[0] = __int_as_float(0x7fc00000);
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, kTmemCols);
}

// =================================================================================================
// Weight gradient on the tensor cores:  dW[k][n] (+)= sum_rows X[row][k] * dY[row][n]   (k <= 150: row 150
// of dW is the bias gradient, picked up by the ones column of X).
//   The contraction runs over edge rows, so  A = X^T : TMEM lane = feature k, column = row within a 32-row
//   chunk; two overlapping M = 128 tiles cover the features: tile 0 = features 0..127, tile 1 = 23..150.
//   B = dY^T chunk in shared memory, K-major: [k-step of 8 rows][2][n = 160][4 rows].
//   D0 / D1 (TMEM) accumulate ONE 128-row tile (48 MMAs each), then are added -- with round-to-nearest
//   FADDs -- into 160 registers per thread (the tensor core truncates when it accumulates, so long accumulation
//   chains in TMEM would bias the sum); the per-CTA partial goes to global memory once, at the end
//   (layout [tile][n][lane], coalesced).
//   TMEM map: D0 [0,160) | D1 [160,320) | A0_hi [320,352) A0_lo [352,384) A1_hi [384,416) A1_lo [416,448)
// =================================================================================================
constexpr int kWgChunk = 32;
constexpr int kWgThreads = 512;              // (lane quarter, M-tile, half of the chunk rows / accumulator columns) per warp
constexpr uint32_t kWgColD0 = 0, kWgColD1 = 160, kWgColA = 320;
constexpr int kWgFeat1 = 23;                 // first feature of M-tile 1 (lane j <-> feature 23 + j)
constexpr int kWgBFloats = 4 * kBStepFloats; // 5120 floats per hi / lo chunk operand
constexpr size_t kWgPartFloats = 2 * 160 * 128;

struct WgradTcArgs {
  int M;
  int x_mode;                  // 0: X plain [M][152] (ones column included); 1: relu(A + S[snd] + R[rcv]), ones column set
  const float* X; const float* S; const float* R; const int32_t* in_snd; const int32_t* in_rcv;
  int y_mode;                  // 0: dY plain [M][152]; 1: relu-bits ? dY[rcv[row]] : 0  (dY = dH2S)
  const float* dY; const uint32_t* maskbits;
  float* part;                 // [gridDim.x][2][160][128]
  int first;                   // this launch initialises the partials (otherwise it accumulates into them)
  float* poison;
};

// shared memory: 2 stages x { XA, XS, XR, YD : [32][152] floats, bits [32][8] words, idx [64] } + B hi/lo
constexpr int kWgStageFloats = 4 * kWgChunk * kDEP + kWgChunk * 8 + 64;
constexpr size_t kWgradTcSmem = (size_t)(2 * kWgStageFloats + 2 * kWgBFloats) * sizeof(float) + 16;

__device__ __forceinline__ void cp_async16_zfill(float* sdst, const float* gsrc, bool valid) {
  const unsigned s = smem_u32(sdst);
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gsrc), "r"(sz));
}

// issue the asynchronous copies of one 32-row chunk (raw gather rows; combined later, on the fly)
template <int XMODE, int YMODE>
__device__ __forceinline__ void wg_issue_chunk(const WgradTcArgs& a, float* st, int r0) {
  constexpr int C4 = kDEP / 4;
  float* XA = st; float* XS = XA + kWgChunk * kDEP; float* XR = XS + kWgChunk * kDEP; float* YD = XR + kWgChunk * kDEP;
  uint32_t* BT = reinterpret_cast<uint32_t*>(YD + kWgChunk * kDEP);
  const int* idx = reinterpret_cast<const int*>(BT + kWgChunk * 8);      // [0,32): snd, [32,64): rcv of this chunk
  for (int i = threadIdx.x; i < kWgChunk * C4; i += kWgThreads) {
    const int r = i / C4, c = i - r * C4;
    const int row = r0 + r;
    const bool valid = row < a.M;
    const size_t rowc = valid ? (size_t)row : 0;
    cp_async16_zfill(XA + r * kDEP + 4 * c, a.X + rowc * kDEP + 4 * c, valid);
    if (XMODE == 1) {
      cp_async16_zfill(XS + r * kDEP + 4 * c, a.S + (size_t)idx[r] * kDEP + 4 * c, valid);
      cp_async16_zfill(XR + r * kDEP + 4 * c, a.R + (size_t)idx[32 + r] * kDEP + 4 * c, valid);
    }
    if (YMODE == 1) cp_async16_zfill(YD + r * kDEP + 4 * c, a.dY + (size_t)idx[32 + r] * kDEP + 4 * c, valid);
    else cp_async16_zfill(YD + r * kDEP + 4 * c, a.dY + rowc * kDEP + 4 * c, valid);
  }
  if (YMODE == 1 && threadIdx.x < kWgChunk * 2) {
    const int r = threadIdx.x >> 1, h = threadIdx.x & 1;
    const int row = r0 + r;
    const bool valid = row < a.M;
    cp_async16_zfill(reinterpret_cast<float*>(BT + r * 8 + 4 * h),
                     reinterpret_cast<const float*>(a.maskbits + (valid ? (size_t)row : 0) * 8 + 4 * h), valid);
  }
  cp_async_commit();
}

// edge indices of a chunk -> the stage's idx array (plain loads; consumed one iteration later)
__device__ __forceinline__ void wg_load_idx(const WgradTcArgs& a, float* st, int r0) {
  int* idx = reinterpret_cast<int*>(st + 4 * kWgChunk * kDEP + kWgChunk * 8);
  const int t = threadIdx.x;
  if (t < 64) {
    const int row = r0 + (t & 31);
    int v = 0;
    if (row < a.M) v = t < 32 ? a.in_snd[row] : a.in_rcv[row];
    idx[t] = v;
  }
}

template <int XMODE, int YMODE>
__global__ void __launch_bounds__(kWgThreads, 1) k_wgrad_tc(WgradTcArgs a) {
  SPW_DYN_SMEM(smem_raw);
  float* stages = reinterpret_cast<float*>(smem_raw);
  float* Bhi_s = stages + 2 * kWgStageFloats;
  float* Blo_s = Bhi_s + kWgBFloats;
  uint64_t* bar = reinterpret_cast<uint64_t*>(Blo_s + kWgBFloats);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int L = 32 * (warp & 3) + lane, mt = (warp >> 2) & 1;   // TMEM lane, M-tile handled by this thread
  const int hh = warp >> 3;                                     // half of the chunk rows (operand build) / accumulator columns (flush)
  const int feat = mt == 0 ? L : kWgFeat1 + L;                  // feature (row of dW) of this lane in its M-tile
  constexpr bool kGather = (XMODE == 1) || (YMODE == 1);

  if (warp == 0) tmem_alloc(tptr, kTmemCols);
  if (tid == 32) { mbar_init(bar, 1); fence_mbar_init(); }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tptr;
  const uint32_t lane_addr = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16);
  const uint32_t colA_hi = kWgColA + 64 * mt, colA_lo = colA_hi + 32;
  const uint32_t idesc = make_idesc_tf32(128, kN);
  uint32_t parity = 0;
  bool failed = false, pending = false;
  float* part = a.part + (size_t)blockIdx.x * kWgPartFloats;
  // D accumulates ONE tile in tensor memory; the sum over this CTA's tiles lives in registers (round-to-nearest adds):
  // this thread owns columns [80 hh, 80 hh + 80) of row L of M-tile mt
  float acc[kN / 2];
#pragma unroll
  for (int c = 0; c < kN / 2; ++c) acc[c] = 0.f;
  const int ntiles = (a.M + kTM - 1) / kTM;
  constexpr int kCh = kTM / kWgChunk;                           // chunks per tile
  // this CTA's chunks, in order: q-th chunk = (tile blockIdx.x + (q / kCh) * gridDim.x, chunk q % kCh)
  const int my_tiles = blockIdx.x < ntiles ? (ntiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;
  const int nq = my_tiles * kCh;
  auto row0_of = [&](int q) { return (blockIdx.x + (q / kCh) * gridDim.x) * kTM + (q % kCh) * kWgChunk; };
  // prologue: indices of chunks 0 and 1, copies of chunk 0
  if (nq > 0) {
    if (kGather) wg_load_idx(a, stages, row0_of(0));
    __syncthreads();
    wg_issue_chunk<XMODE, YMODE>(a, stages, row0_of(0));
    if (kGather && nq > 1) wg_load_idx(a, stages + kWgStageFloats, row0_of(1));
  }
  SPW_PH_DECL
  for (int q = 0; q < nq; ++q) {
    SPW_PH(7);
    const int ch = q % kCh;
    float* st = stages + (q & 1) * kWgStageFloats;
    float* stn = stages + ((q + 1) & 1) * kWgStageFloats;
    __syncthreads();                                   // idx of chunk q+1 visible; stage q+1 no longer read (chunk q-1 done)
    if (q + 1 < nq) {
      wg_issue_chunk<XMODE, YMODE>(a, stn, row0_of(q + 1));
      SPW_PH(0);                                       // p0: issue of the next chunk's copies
      cp_async_wait<1>();                              // chunk q has landed (this thread's copies)
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();                                   // ... and everybody else's
    SPW_PH(1);                                         // p1: wait for chunk q's copies
    if (kGather && q + 2 < nq) wg_load_idx(a, st, row0_of(q + 2));   // idx slot of stage q is free: chunk q's copies are done
    if (pending) {                                     // A / B operand regions are free once the previous MMAs are done
      if (!mbar_wait(bar, parity)) failed = true;
      parity ^= 1u;
      fence_after_sync();
      pending = false;
    }
    SPW_PH(2);                                         // p2: idx loads + wait for the previous chunk's MMAs
    const float* XA = st; const float* XS = XA + kWgChunk * kDEP; const float* XR = XS + kWgChunk * kDEP;
    const float* YD = XR + kWgChunk * kDEP;
    const uint32_t* BT = reinterpret_cast<const uint32_t*>(YD + kWgChunk * kDEP);
    const int r0 = row0_of(q);
    // A = X^T: this lane's feature, this thread's 16 rows of the chunk as 16 TMEM columns
#pragma unroll
    for (int jj = 0; jj < kWgChunk / 2; jj += 8) {
      const int j0 = 16 * hh + jj;
      uint32_t h[8], l[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int o = (j0 + i) * kDEP + feat;
        float x = XA[o];
        if (XMODE == 1) {
          x = relu_f(x + XS[o] + XR[o]);
          if (feat == kDE) x = (r0 + j0 + i < a.M) ? 1.f : 0.f;        // ones column -> bias gradient row
        }
        split_tf32(x, h[i], l[i]);
      }
      tmem_st8(lane_addr + colA_hi + j0, h);
      tmem_st8(lane_addr + colA_lo + j0, l);
    }
    SPW_PH(3);                                         // p3: A operand build
    // B = dY^T: [k-step][2][n][4 rows]
    for (int idx = tid; idx < 8 * kN; idx += kWgThreads) {
      const int n = idx % kN, kc = idx / kN;
      uint32_t h[4], l[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float y = 0.f;
        if (n < kDE) {
          y = YD[(4 * kc + i) * kDEP + n];
          if (YMODE == 1) y = ((BT[(4 * kc + i) * 8 + (n >> 5)] >> (n & 31)) & 1u) ? y : 0.f;
        }
        split_tf32(y, h[i], l[i]);
      }
      reinterpret_cast<uint4*>(Bhi_s)[idx] = make_uint4(h[0], h[1], h[2], h[3]);
      reinterpret_cast<uint4*>(Blo_s)[idx] = make_uint4(l[0], l[1], l[2], l[3]);
    }
    tmem_wait_st();
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    SPW_PH(4);                                         // p4: B operand build + sync
    if (tid == 0) {
      fence_after_sync();
      const uint32_t bhi = smem_u32(Bhi_s), blo = smem_u32(Blo_s);
#pragma unroll 1
      for (int ks = 0; ks < kWgChunk / 8; ++ks) {
        const uint64_t dhi = make_b_desc(bhi + ks * (kBStepFloats * 4), kN * 16, 128);
        const uint64_t dlo = make_b_desc(blo + ks * (kBStepFloats * 4), kN * 16, 128);
        const uint32_t acc = (ch > 0 || ks > 0) ? 1u : 0u;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const uint32_t d = tmem_base + (t ? kWgColD1 : kWgColD0);
          const uint32_t ahi = tmem_base + kWgColA + 64 * t + 8 * ks, alo = ahi + 32;
          mma_tf32_ts(d, alo, dhi, idesc, acc);
          mma_tf32_ts(d, ahi, dlo, idesc, 1u);
          mma_tf32_ts(d, ahi, dhi, idesc, 1u);
        }
      }
      mma_commit(bar);
    }
    pending = true;
    SPW_PH(5);                                         // p5: MMA issue
    if (ch == kCh - 1) {
      // tile done: wait for its MMAs, add D into the per-CTA partial with round-to-nearest adds
      if (!mbar_wait(bar, parity)) failed = true;
      parity ^= 1u;
      fence_after_sync();
      pending = false;
      const uint32_t dcol = (mt ? kWgColD1 : kWgColD0) + (kN / 2) * hh;
#pragma unroll
      for (int c = 0; c < kN / 2; c += 16) {
        uint32_t v[16];
        tmem_ld16(lane_addr + dcol + c, v);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[c + i] += __uint_as_float(v[i]);
      }
      fence_before_sync();
      SPW_PH(6);                                       // p6: tile flush
    }
  }
  {   // this CTA's sum over its tiles -> per-CTA partial in global memory ([tile][n][lane]: coalesced)
    float* pp = part + (size_t)mt * (160 * 128) + (size_t)(kN / 2) * hh * 128 + L;
    if (a.first) {
#pragma unroll
      for (int c = 0; c < kN / 2; ++c) pp[(size_t)c * 128] = acc[c];
    } else {
#pragma unroll
      for (int c0 = 0; c0 < kN / 2; c0 += 16) {
        float old[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) old[i] = pp[(size_t)(c0 + i) * 128];
#pragma unroll
        for (int i = 0; i < 16; ++i) pp[(size_t)(c0 + i) * 128] = old[i] + acc[c0 + i];
      }
    }
  }
  SPW_PH_REPORT(XMODE ? "k_wgrad_tc<1,1>" : "k_wgrad_tc<0,0>");
  if (failed && tid == 0) a.poison[0] = __int_as_float(0x7fc00000);
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, kTmemCols);
}

// =================================================================================================
// Weight gradient of a plain linear layer on the tensor cores (node-level layers and the relation encoder):
//   dW[k][n] = sum_rows X[row % xmod][k] * dY[row][n],  k < Kx;  row Kx of dW = sum_rows one[row] * dY[row][n] (bias
//   gradient; one = rowscale[row % rsmod] or 1).  Same operand roles as k_wgrad_tc: A = X^T chunk in tensor memory
//   (lane = feature), B = dY^T chunk in shared memory, D accumulates one 128-row tile and is flushed with round-to-
//   nearest adds into a per-CTA partial [2][160][128].  Kx + 1 <= 128 needs one M-tile only: then the two warp
//   groups split the chunk rows (operand build) and the accumulator columns (flush) between them.
// =================================================================================================
constexpr size_t wgrad_rows_smem(int ch, int nb) { return (size_t)(2 * kWgStageFloats + 2 * (ch / 8) * (2 * nb * 4)) * sizeof(float) + 32; }

struct WgradRowsArgs {
  int M;
  const float* X; int ldx; int Kx; int xmod;          // X row = row % xmod (xmod = 0: row); ldx % 4 == 0
  const float* rowscale; int rsmod;                   // value of the virtual column Kx (null: 1)
  const float* dY; int ldy; int Ny;                   // ldy % 4 == 0
  int NB;                                             // MMA N: 112 or 160 (>= Ny)
  float* part;                                        // [gridDim.x][2][160][128]
  float* poison;
};

__device__ __forceinline__ void cp_async4_zfill(float* sdst, const float* gsrc, bool valid) {
  const unsigned s = smem_u32(sdst);
  const int sz = valid ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(s), "l"(gsrc), "r"(sz));
}

// stage layout (floats; CH = rows per chunk, 32 or 64): XA [CH][152] | YD [CH][152] | ... | RS [CH] at 4 * 32 * 152
// (the stage has the size of k_wgrad_tc's: 64-row chunks fit because the plain mode needs two arrays, not four)
constexpr int kWgrRsOff = 4 * kWgChunk * kDEP;
template <int CH>
__device__ __forceinline__ void wgr_issue_chunk(const WgradRowsArgs& a, float* st, int r0, int cx4, int cy4) {
  float* XA = st; float* YD = st + CH * kDEP;
  float* RS = st + kWgrRsOff;
  for (int i = threadIdx.x; i < CH * cx4; i += kWgThreads) {
    const int r = i / cx4, c = i - r * cx4;
    const int row = r0 + r;
    const bool valid = row < a.M;
    const size_t xr = valid ? (size_t)(a.xmod ? row % a.xmod : row) : 0;
    cp_async16_zfill(XA + r * kDEP + 4 * c, a.X + xr * a.ldx + 4 * c, valid);
  }
  for (int i = threadIdx.x; i < CH * cy4; i += kWgThreads) {
    const int r = i / cy4, c = i - r * cy4;
    const int row = r0 + r;
    const bool valid = row < a.M;
    cp_async16_zfill(YD + r * kDEP + 4 * c, a.dY + (valid ? (size_t)row : 0) * a.ldy + 4 * c, valid);
  }
  if (a.rowscale && threadIdx.x < CH) {
    const int row = r0 + threadIdx.x;
    const bool valid = row < a.M;
    cp_async4_zfill(RS + threadIdx.x, a.rowscale + (valid ? (size_t)(a.rsmod ? row % a.rsmod : row) : 0), valid);
  }
  cp_async_commit();
}

// CH = 64 (half as many chunk iterations) needs Kx + 1 <= 128 (one M-tile: tensor-memory columns) and NB = 112 (shared memory)
template <int CH>
__global__ void __launch_bounds__(kWgThreads, 1) k_wgrad_rows_tc(WgradRowsArgs a) {
  SPW_DYN_SMEM(smem_raw);
  float* stages = reinterpret_cast<float*>(smem_raw);
  float* Bhi_s = stages + 2 * kWgStageFloats;
  float* Blo_s = Bhi_s + (CH / 8) * (2 * a.NB * 4);
  uint64_t* bar = reinterpret_cast<uint64_t*>(Blo_s + (CH / 8) * (2 * a.NB * 4));
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int L = 32 * (warp & 3) + lane, grp = warp >> 2;        // TMEM lane, warp group (0..3)
  const int nmt = a.Kx + 1 > 128 ? 2 : 1;                       // M-tiles
  const int feat1 = nmt == 2 ? a.Kx + 1 - 128 : 0;              // first feature of M-tile 1
  const int mt = nmt == 2 ? (grp & 1) : 0;                      // M-tile this thread builds / flushes
  const int sh = nmt == 2 ? (grp >> 1) : grp, nsh = nmt == 2 ? 2 : 4;   // this thread's share of the chunk rows / accumulator blocks
  const int feat = mt == 0 ? L : feat1 + L;
  const int NB = a.NB;
  const int cx4 = (a.Kx + 3) >> 2, cy4 = (a.Ny + 3) >> 2;
  const int bstep = 2 * NB * 4;                                 // floats per k-step of the B operand

  if (warp == 0) tmem_alloc(tptr, kTmemCols);
  if (tid == 32) { mbar_init(bar, 1); fence_mbar_init(); }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tptr;
  const uint32_t lane_addr = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16);
  const uint32_t colA_hi = kWgColA + 2 * CH * mt, colA_lo = colA_hi + CH;
  const uint32_t idesc = make_idesc_tf32(128, NB);
  uint32_t parity = 0;
  bool failed = false, pending = false;
  float* part = a.part + (size_t)blockIdx.x * kWgPartFloats;
  // 16-column accumulator blocks of this thread: its share (1 / nsh) of the NB / 16 blocks, at most 5
  const int nblk = NB / 16;
  const int b_lo = (sh * nblk) / nsh, b_hi = ((sh + 1) * nblk) / nsh;
  float acc[5][16];                                             // sum over this CTA's tiles (round-to-nearest adds)
#pragma unroll
  for (int b = 0; b < 5; ++b)
#pragma unroll
    for (int c = 0; c < 16; ++c) acc[b][c] = 0.f;
  const int ntiles = (a.M + kTM - 1) / kTM;
  constexpr int kCh = kTM / CH;                                 // chunks per tile
  const int my_tiles = (int)blockIdx.x < ntiles ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int nq = my_tiles * kCh;
  auto row0_of = [&](int q) { return (blockIdx.x + (q / kCh) * gridDim.x) * kTM + (q % kCh) * CH; };
  // rows of the chunk this thread turns into A columns: its share of the CH
  const int j_lo = (CH / nsh) * sh, j_hi = j_lo + CH / nsh;
  if (nq > 0) wgr_issue_chunk<CH>(a, stages, row0_of(0), cx4, cy4);
  for (int q = 0; q < nq; ++q) {
    const int ch = q % kCh;
    float* st = stages + (q & 1) * kWgStageFloats;
    float* stn = stages + ((q + 1) & 1) * kWgStageFloats;
    __syncthreads();                                   // stage q+1 no longer read (chunk q-1 done)
    if (q + 1 < nq) {
      wgr_issue_chunk<CH>(a, stn, row0_of(q + 1), cx4, cy4);
      cp_async_wait<1>();                              // chunk q has landed (this thread's copies)
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();                                   // ... and everybody else's
    if (pending) {                                     // A / B operand regions are free once the previous MMAs are done
      if (!mbar_wait(bar, parity)) failed = true;
      parity ^= 1u;
      fence_after_sync();
      pending = false;
    }
    const float* XA = st; const float* YD = st + CH * kDEP;
    const float* RS = st + kWgrRsOff;
    const int r0 = row0_of(q);
    // A = X^T: this lane's feature, rows of the chunk as TMEM columns
    for (int j0 = j_lo; j0 < j_hi; j0 += 8) {
      uint32_t h[8], l[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float x = 0.f;
        if (feat < a.Kx) x = XA[(j0 + i) * kDEP + feat];
        else if (feat == a.Kx && r0 + j0 + i < a.M) x = a.rowscale ? RS[j0 + i] : 1.f;
        split_tf32(x, h[i], l[i]);
      }
      tmem_st8(lane_addr + colA_hi + j0, h);
      tmem_st8(lane_addr + colA_lo + j0, l);
    }
    // B = dY^T: [k-step][2][n][4 rows]
    for (int idx = tid; idx < (CH / 4) * NB; idx += kWgThreads) {
      const int n = idx % NB, kc = idx / NB;
      uint32_t h[4], l[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float y = n < a.Ny ? YD[(4 * kc + i) * kDEP + n] : 0.f;
        split_tf32(y, h[i], l[i]);
      }
      reinterpret_cast<uint4*>(Bhi_s)[idx] = make_uint4(h[0], h[1], h[2], h[3]);
      reinterpret_cast<uint4*>(Blo_s)[idx] = make_uint4(l[0], l[1], l[2], l[3]);
    }
    tmem_wait_st();
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      fence_after_sync();
      const uint32_t bhi = smem_u32(Bhi_s), blo = smem_u32(Blo_s);
#pragma unroll 1
      for (int ks = 0; ks < CH / 8; ++ks) {
        const uint64_t dhi = make_b_desc(bhi + ks * (bstep * 4), NB * 16, 128);
        const uint64_t dlo = make_b_desc(blo + ks * (bstep * 4), NB * 16, 128);
        const uint32_t acc = (ch > 0 || ks > 0) ? 1u : 0u;
        for (int t = 0; t < nmt; ++t) {
          const uint32_t d = tmem_base + (t ? kWgColD1 : kWgColD0);
          const uint32_t ahi = tmem_base + kWgColA + 2 * CH * t + 8 * ks, alo = ahi + CH;
          mma_tf32_ts(d, alo, dhi, idesc, acc);
          mma_tf32_ts(d, ahi, dlo, idesc, 1u);
          mma_tf32_ts(d, ahi, dhi, idesc, 1u);
        }
      }
      mma_commit(bar);
    }
    pending = true;
    if (ch == kCh - 1) {
      // tile done: wait for its MMAs, add D into the per-CTA partial with round-to-nearest adds
      if (!mbar_wait(bar, parity)) failed = true;
      parity ^= 1u;
      fence_after_sync();
      pending = false;
      const uint32_t dcol = mt ? kWgColD1 : kWgColD0;
#pragma unroll
      for (int bi = 0; bi < 5; ++bi) {
        if (b_lo + bi < b_hi) {                        // warp-uniform
          uint32_t v[16];
          tmem_ld16(lane_addr + dcol + 16 * (b_lo + bi), v);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 16; ++i) acc[bi][i] += __uint_as_float(v[i]);
        }
      }
      fence_before_sync();
    }
  }
  {   // this CTA's sum over its tiles -> per-CTA partial in global memory ([tile][n][lane]: coalesced)
    float* pp = part + (size_t)mt * (160 * 128) + L;
#pragma unroll
    for (int bi = 0; bi < 5; ++bi) {
      if (b_lo + bi < b_hi) {
#pragma unroll
        for (int i = 0; i < 16; ++i) pp[(size_t)(16 * (b_lo + bi) + i) * 128] = acc[bi][i];
      }
    }
  }
  if (failed && tid == 0) a.poison[0] = __int_as_float(0x7fc00000);
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, kTmemCols);
}

// relation-encoder layer 0 (K = 2) gradients from G0 = d(pre-activation of X0):
//   part[cta][0][k] = sum_e dx_e G0[e][k], [1][k] with dy, [2][k] = sum_e G0[e][k]   (fixed order per CTA)
constexpr int kEnc0Threads = 640;    // 4 row quarters x 160 columns
__global__ void __launch_bounds__(kEnc0Threads, 2) k_enc0_bwd(int E, const int32_t* __restrict__ in_snd,
                                                              const int32_t* __restrict__ in_rcv, const float* __restrict__ obj,
                                                              const float* __restrict__ G0, float* __restrict__ part) {
  __shared__ float sdx[kTM], sdy[kTM];
  __shared__ float sred[3][3][160];
  const int tid = threadIdx.x, qr = tid / 160, c = tid - qr * 160;       // row quarter, column
  float g0 = 0.f, g1 = 0.f, gb = 0.f;
  const int ntiles = (E + kTM - 1) / kTM;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int e0 = tile * kTM, rows = imin(kTM, E - e0);
    __syncthreads();
    if (tid < kTM) {
      float dx = 0.f, dy = 0.f;
      if (tid < rows) {
        const int s = in_snd[e0 + tid], rc = in_rcv[e0 + tid];
        dx = obj[3 * (size_t)rc] - obj[3 * (size_t)s];
        dy = obj[3 * (size_t)rc + 1] - obj[3 * (size_t)s + 1];
      }
      sdx[tid] = dx; sdy[tid] = dy;
    }
    __syncthreads();
    if (c < kDE) {
      const int ra = 32 * qr, rb = imin(rows, ra + 32);
      const float* gp = G0 + (size_t)e0 * kDEP + c;
#pragma unroll 8
      for (int r = ra; r < rb; ++r) {
        const float d = gp[(size_t)r * kDEP];
        g0 = fmaf(sdx[r], d, g0); g1 = fmaf(sdy[r], d, g1); gb += d;
      }
    }
  }
  __syncthreads();
  if (qr > 0) { sred[qr - 1][0][c] = g0; sred[qr - 1][1][c] = g1; sred[qr - 1][2][c] = gb; }
  __syncthreads();
  if (qr == 0 && c < kDEP) {                                      // fixed order: quarter 0 + 1 + 2 + 3
    for (int k = 0; k < 3; ++k) { g0 += sred[k][0][c]; g1 += sred[k][1][c]; gb += sred[k][2][c]; }
    float* p = part + (size_t)blockIdx.x * (3 * kDEP);
    p[c] = c < kDE ? g0 : 0.f; p[kDEP + c] = c < kDE ? g1 : 0.f; p[2 * kDEP + c] = c < kDE ? gb : 0.f;
  }
}

}  // namespace tc
}  // namespace spw
#endif  // SPW_EMU
