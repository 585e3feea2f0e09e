// spw_edges.cuh -- K1: edge-index construction (replaces the numpy relation loops of
// /root/reference/src/main.py:66-81, TowerCreator.py:415-428, JengaBuilder.py:313-326).
//
// One CTA (64 threads) per tower; thread i plays block i both as sender and as receiver.
// The activity test is the reference's float64 expression bit for bit:
//     np.linalg.norm(p_m - p_j) < thr   ==   sqrt(dx*dx + dy*dy) < thr
// with separately rounded multiplies and add (no FMA contraction) -- comparing squared distances
// is NOT equivalent at the boundary.  Integer work, HBM traffic 16 B/node in, 12-20 B/edge out.
#pragma once
#include "spw_common.cuh"

namespace spw {

constexpr int kMaxNodes = 64;

__device__ __forceinline__ bool pair_active(const double* pos, int m, int j, double thr, int fully_connected) {
  if (fully_connected) return true;
  const double dx = pos[2 * m] - pos[2 * j];
  const double dy = pos[2 * m + 1] - pos[2 * j + 1];
  const double s = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
  return __dsqrt_rn(s) < thr;
}

// deg_out[i] / deg_in[i] per node; edge_off[t+1] = number of active edges of tower t (scanned later)
__global__ void __launch_bounds__(kMaxNodes) k_edges_count(const double* __restrict__ pos_xy,
                                                           const int32_t* __restrict__ node_off, double thr,
                                                           int fully_connected, int32_t* __restrict__ deg_out,
                                                           int32_t* __restrict__ deg_in,
                                                           int32_t* __restrict__ edge_off) {
  __shared__ double pos[2 * kMaxNodes];
  __shared__ int sdeg[kMaxNodes];
  const int t = blockIdx.x, i = threadIdx.x;
  const int a = node_off[t], N = node_off[t + 1] - a;
  if (i < N) { pos[2 * i] = pos_xy[2 * (size_t)(a + i)]; pos[2 * i + 1] = pos_xy[2 * (size_t)(a + i) + 1]; }
  __syncthreads();
  int dout = 0, din = 0;
  if (i < N) {
    for (int j = 0; j < N; ++j) {
      if (j == i) continue;
      dout += pair_active(pos, i, j, thr, fully_connected) ? 1 : 0;   // i sends to j
      din += pair_active(pos, j, i, thr, fully_connected) ? 1 : 0;    // j sends to i
    }
    deg_out[a + i] = dout;
    deg_in[a + i] = din;
  }
  sdeg[i] = dout;
  __syncthreads();
  if (i == 0) {
    int s = 0;
    for (int k = 0; k < N; ++k) s += sdeg[k];
    edge_off[t + 1] = s;
    if (t == 0) edge_off[0] = 0;
  }
}

// in-place inclusive scan of x[1..n] (x[0] stays 0): single CTA, 1024-wide chunks with a carry.
__global__ void __launch_bounds__(1024) k_scan_inplace(int32_t* __restrict__ x, int n) {
  __shared__ int buf[1024];
  __shared__ int carry_s;
  const int tid = threadIdx.x;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    const int idx = base + tid;
    int v = idx < n ? x[1 + idx] : 0;
    buf[tid] = v;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
      int add = tid >= off ? buf[tid - off] : 0;
      __syncthreads();
      buf[tid] += add;
      __syncthreads();
    }
    const int carry = carry_s;
    if (idx < n) x[1 + idx] = carry + buf[tid];
    __syncthreads();
    if (tid == 1023) carry_s = carry + buf[1023];
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kMaxNodes) k_edges_fill(
    const double* __restrict__ pos_xy, const int32_t* __restrict__ node_off, int n_towers, int n_nodes, double thr,
    int fully_connected, const int32_t* __restrict__ edge_off, int32_t* __restrict__ snd, int32_t* __restrict__ rcv,
    int32_t* __restrict__ slot, int32_t* __restrict__ in_off, int32_t* __restrict__ in_snd,
    int32_t* __restrict__ in_rcv, int32_t* __restrict__ out_off, int32_t* __restrict__ out_pos) {
  __shared__ double pos[2 * kMaxNodes];
  __shared__ int s_out[kMaxNodes], s_in[kMaxNodes];          // degrees, then exclusive prefixes
  __shared__ unsigned char rin[kMaxNodes * kMaxNodes];       // rank of sender m among in-edges of receiver j
  const int t = blockIdx.x, i = threadIdx.x;
  const int a = node_off[t], N = node_off[t + 1] - a;
  const int base = edge_off[t];
  if (i < N) { pos[2 * i] = pos_xy[2 * (size_t)(a + i)]; pos[2 * i + 1] = pos_xy[2 * (size_t)(a + i) + 1]; }
  __syncthreads();
  int dout = 0, din = 0;
  if (i < N) {
    for (int j = 0; j < N; ++j) {
      if (j == i) continue;
      dout += pair_active(pos, i, j, thr, fully_connected) ? 1 : 0;
      din += pair_active(pos, j, i, thr, fully_connected) ? 1 : 0;
    }
  }
  s_out[i] = dout;
  s_in[i] = din;
  __syncthreads();
  if (i == 0) {
    int so = 0, si = 0;
    for (int k = 0; k < N; ++k) {
      const int o = s_out[k], n = s_in[k];
      s_out[k] = so; s_in[k] = si;
      so += o; si += n;
    }
  }
  __syncthreads();
  if (i < N) {
    out_off[a + i] = base + s_out[i];
    in_off[a + i] = base + s_in[i];
    // as receiver: in-edges in ascending sender order (== ascending slot order)
    int c = 0;
    const int p0 = base + s_in[i];
    for (int m = 0; m < N; ++m) {
      if (m == i) continue;
      if (pair_active(pos, m, i, thr, fully_connected)) {
        rin[m * kMaxNodes + i] = (unsigned char)c;
        in_snd[p0 + c] = a + m;
        in_rcv[p0 + c] = a + i;
        ++c;
      }
    }
  }
  if (t == n_towers - 1 && i == 0) {
    out_off[n_nodes] = edge_off[n_towers];
    in_off[n_nodes] = edge_off[n_towers];
  }
  __syncthreads();
  if (i < N) {
    // as sender: slot order (main.py:69-81: m outer, j inner)
    int c = 0;
    const int e0 = base + s_out[i];
    for (int j = 0; j < N; ++j) {
      if (j == i) continue;
      if (pair_active(pos, i, j, thr, fully_connected)) {
        const int e = e0 + c;
        if (snd) snd[e] = a + i;
        if (rcv) rcv[e] = a + j;
        if (slot) slot[e] = i * (N - 1) + (j < i ? j : j - 1);
        out_pos[e] = base + s_in[j] + (int)rin[i * kMaxNodes + j];
        ++c;
      }
    }
  }
}

}  // namespace spw
