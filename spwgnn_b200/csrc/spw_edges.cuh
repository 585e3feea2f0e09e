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

// =================================================================================================
// N4: physics-free Jenga layout sampler on the device -- JengaBuilder.create_world
// (/root/reference/src/JengaBuilder.py:137-192) restated with a counter-based generator so that a 1 M-tower sweep needs
// neither host generation nor an H2D copy.  spwgnn_b200/synth.py: g_jenga_ctr is the same algorithm in numpy; the two
// agree bit for bit (integer / exact half-integer float64 arithmetic only).
// =================================================================================================
__host__ __device__ __forceinline__ uint32_t ctr_u32(uint64_t seed, uint64_t tower, uint64_t ctr) {
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * (tower + 1) + 0xD1B54A32D192ED03ull * ctr;    // splitmix64 finaliser
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (uint32_t)(z >> 32);
}
// random.randint(a, b) of the reference: uniform integer in [a, b]
__host__ __device__ __forceinline__ int ctr_randint(uint64_t seed, uint64_t tower, uint32_t& ctr, int a, int b) {
  return a + (int)(ctr_u32(seed, tower, ctr++) % (uint32_t)(b - a + 1));
}

// blocks per tower: counter 0 of each tower's stream; node_off[t + 1] = N_t (scanned afterwards), node_off[0] = 0
__global__ void __launch_bounds__(256) k_sample_sizes(uint64_t seed, int n_towers, int n_lo, int n_hi, int32_t* __restrict__ node_off) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t == 0) node_off[0] = 0;
  if (t < n_towers) {
    uint32_t ctr = 0;
    node_off[t + 1] = ctr_randint(seed, (uint64_t)t, ctr, n_lo, n_hi);
  }
}

// one thread per tower: [x, y, width] of its blocks in pixels (float64), plus the model input obj = raw / 170 (fp32,
// main.py:91) and the positions the relation test runs on (raw, or raw / 170 for the inference glue, JengaBuilder.py:309-323)
__global__ void __launch_bounds__(128) k_sample_jenga(uint64_t seed, int n_towers, const int32_t* __restrict__ node_off,
                                                      double* __restrict__ raw, float* __restrict__ obj, double* __restrict__ pos,
                                                      int inference_glue) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_towers) return;
  const int base = node_off[t], n = node_off[t + 1] - base;
  uint32_t ctr = 1;                                              // counter 0 drew the size
  const int wmin = 50, wrange = 250, wavg = 150, gap = 50, rect_h = 80, bottom = 70;
  const double left_most = 400.0, right_most = 1500.0 - 400.0;
  int left = n, layer = -1, k = 0;
  double prev_min = 0.0, prev_max = 0.0;
  auto emit = [&](double x, double y, double w) {
    const size_t o = (size_t)(base + k);
    if (raw) { raw[3 * o] = x; raw[3 * o + 1] = y; raw[3 * o + 2] = w; }
    if (obj) { obj[3 * o] = (float)(x / 170.0); obj[3 * o + 1] = (float)(y / 170.0); obj[3 * o + 2] = (float)(w / 170.0); }
    if (pos) { pos[2 * o] = inference_glue ? x / 170.0 : x; pos[2 * o + 1] = inference_glue ? y / 170.0 : y; }
    ++k;
  };
  while (left > 0) {
    ++layer;
    double r_edge, l_edge;
    if (layer == 0) { r_edge = right_most; l_edge = left_most; } else { r_edge = prev_max; l_edge = prev_min; }
    const double y = (double)(bottom + rect_h / 2 + rect_h * layer);
    double cur_min = 0.0, cur_max = 0.0;
    int cur = 0;
    auto put = [&](double x, double w) {
      emit(x, y, w);
      cur_min = cur == 0 ? x : (x < cur_min ? x : cur_min);
      cur_max = cur == 0 ? x : (x > cur_max ? x : cur_max);
      ++cur; --left;
    };
    if (r_edge == l_edge) {                                      // single block below: centre a new one on it
      const int x = ctr_randint(seed, (uint64_t)t, ctr, (int)(l_edge - wmin / 2), (int)(l_edge + wmin / 2));
      const int w = ctr_randint(seed, (uint64_t)t, ctr, wmin, wmin + wrange);
      put((double)x, (double)w);
    } else {
      if (layer > 0) l_edge -= (double)(wavg / 2);
      int w = ctr_randint(seed, (uint64_t)t, ctr, wmin, wmin + wrange);
      l_edge += (double)w;
      while (l_edge - w / 2.0 < r_edge && left > 0) {
        put(l_edge - w / 2.0, (double)w);
        l_edge += (double)ctr_randint(seed, (uint64_t)t, ctr, 0, gap);
        w = ctr_randint(seed, (uint64_t)t, ctr, wmin, wmin + wrange);
        l_edge += (double)w;
      }
      if (cur == 0) put(l_edge, (double)w);                      // degenerate draw: force one block so the loop terminates
    }
    prev_min = cur_min; prev_max = cur_max;
  }
}

// TowerCreator layouts (/root/reference/src/TowerCreator.py:106-187 create_world / create_pos_for_boxes, :265-271 drop_object):
// 150 x 80 blocks stacked in layers of shrinking size plus ONE dropped block on top, which is object 0 (TowerCreator.py:451).
// Tower t has N_t = node_off[t+1] - node_off[t] >= 2 blocks: N_t - 1 stacked + the dropped one.  synth.g_tower_ctr is the
// bit-identical numpy restatement.
__global__ void __launch_bounds__(128) k_sample_tower(uint64_t seed, int n_towers, const int32_t* __restrict__ node_off,
                                                      double* __restrict__ raw, float* __restrict__ obj, double* __restrict__ pos,
                                                      int inference_glue) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_towers) return;
  const int base = node_off[t], n = node_off[t + 1] - base - 1;   // stacked blocks
  if (n < 1) return;
  uint32_t ctr = 1;                                              // counter 0 drew the size
  const int rect_w = 150, rect_h = 80, bottom = 70;
  auto emit = [&](int k, double x, double y) {
    const size_t o = (size_t)(base + k);
    if (raw) { raw[3 * o] = x; raw[3 * o + 1] = y; raw[3 * o + 2] = (double)rect_w; }
    if (obj) { obj[3 * o] = (float)(x / 170.0); obj[3 * o + 1] = (float)(y / 170.0); obj[3 * o + 2] = (float)(rect_w / 170.0); }
    if (pos) { pos[2 * o] = inference_glue ? x / 170.0 : x; pos[2 * o + 1] = inference_glue ? y / 170.0 : y; }
  };
  int orientation = ctr_u32(seed, (uint64_t)t, ctr++) > 0x80000000u ? 1 : 0;           // rng.random() > 0.5; flipped by a failed stability check
  // layer sizes
  unsigned char layers[kMaxNodes];
  int nl = 0;
  layers[nl++] = (unsigned char)ctr_randint(seed, (uint64_t)t, ctr, 1, n / 2 > 1 ? n / 2 : 1);
  int left = n - layers[0];
  while (left > 0) {
    const int prev = layers[nl - 1];
    int r;
    if (prev == 1) {
      r = 1;
    } else {
      const int hi = prev < left ? prev : left;
      r = ctr_randint(seed, (uint64_t)t, ctr, 1, hi);
      for (int i = 0; r == 1 && left != 1 && i < 3; ++i) r = ctr_randint(seed, (uint64_t)t, ctr, 1, hi);
    }
    layers[nl++] = (unsigned char)r;
    left -= r;
  }
  // positions, layer by layer; only the previous layer's extreme x are needed
  int prev_max = 0, prev_min = 0, k = 1;                         // block 0 is the dropped one
  auto make_x = [&](int ln, int size, int idx, double mid, bool to_drop) {
    const int var = to_drop ? 75 : 45;                           // int(150 * 0.5), int(150 * 0.3)
    const int mean_range = rect_w + 2 * var;
    const double mean = mid + ((idx & 1) ? -1.0 : 1.0) * (double)(((idx + 1) / 2) * mean_range);
    if (ln > 0 && size == 1) {
      const int r = prev_max + rect_w / 2, l = prev_min - rect_w / 2;
      const int lo = l + 30, hi = r - 30;                        // int(150 * 0.2)
      return ctr_randint(seed, (uint64_t)t, ctr, lo < hi ? lo : hi, lo < hi ? hi : lo);
    }
    const int lo = (int)(mean - (double)((1 - orientation) * var)), hi = (int)(mean + (double)(orientation * var));
    return ctr_randint(seed, (uint64_t)t, ctr, lo, hi) + (size % 2 == 0 ? mean_range / 2 : 0);
  };
  auto middle = [&](int ln) { return ln == 0 ? 750.0 : (double)(int)((double)((prev_min - rect_w / 2) + (prev_max + rect_w / 2)) / 2.0); };
  // TowerCreator.is_box_stable (TowerCreator.py:250-263): with the candidate in place, the sum of int(x / count) over all
  // boxes must lie between the edges of the ground layer; otherwise ONE re-draw with the build direction flipped (:201-207)
  short xs[kMaxNodes + 1];
  int placed = 0, g_max = 0, g_min = 0;
  auto place_x = [&](int ln, int size, int idx, double mid, bool to_drop) {
    int x = make_x(ln, size, idx, mid, to_drop);
    if (ln > 0) {
      const int cnt = placed + 1;
      int com = x / cnt;
      for (int i = 0; i < placed; ++i) com += xs[i] / cnt;
      if (com < g_min - rect_w / 2 || com > g_max + rect_w / 2) {
        orientation = 1 - orientation;
        x = make_x(ln, size, idx, mid, to_drop);
      }
    }
    return x;
  };
  for (int ln = 0; ln < nl; ++ln) {
    const double mid = middle(ln);
    int cur_max = 0, cur_min = 0;
    for (int i = 0; i < layers[ln]; ++i) {
      const int x = place_x(ln, layers[ln], i, mid, false);
      cur_max = i == 0 ? x : (x > cur_max ? x : cur_max);
      cur_min = i == 0 ? x : (x < cur_min ? x : cur_min);
      xs[placed++] = (short)x;
      emit(k++, (double)x, (double)bottom + rect_h / 2.0 + (double)(rect_h * ln));
    }
    prev_max = cur_max; prev_min = cur_min;
    if (ln == 0) { g_max = cur_max; g_min = cur_min; }
  }
  const int xd = place_x(nl, 1, 0, middle(nl), true);
  emit(0, (double)xd, (double)bottom + rect_h / 2.0 + (double)(rect_h * nl));
}

// =================================================================================================
// Demolish searches: candidate towers built on the device, one packed inference batch, per-tower
// sum of block probabilities and its argmin -- instead of N (JengaBuilder.remove_to_demolish, JengaBuilder.py:236-269) or 100
// (TowerCreator.drop_to_demolish, TowerCreator.py:276-319) sequential batch-1 predict calls.
// =================================================================================================
// candidate c (0 <= c < N) = the base tower without block c, block order kept (JengaBuilder.py:245-249): N towers of N - 1 blocks
__global__ void __launch_bounds__(256) k_candidates_remove(const double* __restrict__ raw, int N, float* __restrict__ obj,
                                                           double* __restrict__ pos, int inference_glue) {
  const int total = N * (N - 1);
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int c = idx / (N - 1), k = idx - c * (N - 1);
    const int src = k < c ? k : k + 1;
    const double x = raw[3 * src], y = raw[3 * src + 1], w = raw[3 * src + 2];
    obj[3 * (size_t)idx] = (float)(x / 170.0); obj[3 * (size_t)idx + 1] = (float)(y / 170.0); obj[3 * (size_t)idx + 2] = (float)(w / 170.0);   // main.py:91
    pos[2 * (size_t)idx] = inference_glue ? x / 170.0 : x; pos[2 * (size_t)idx + 1] = inference_glue ? y / 170.0 : y;
  }
}
// candidate c (0 <= c < K) = the base tower plus a dropped block at poses[c], which is object 0 (TowerCreator.py:288-296,451):
// K towers of N + 1 blocks; the dropped block carries `width` like every block of the construction environment
__global__ void __launch_bounds__(256) k_candidates_drop(const double* __restrict__ raw, int N, const double* __restrict__ poses, int K,
                                                         double width, float* __restrict__ obj, double* __restrict__ pos, int inference_glue) {
  const long long total = (long long)K * (N + 1);
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx / (N + 1)), k = (int)(idx - (long long)c * (N + 1));
    double x, y, w;
    if (k == 0) { x = poses[2 * c]; y = poses[2 * c + 1]; w = width; }
    else { x = raw[3 * (k - 1)]; y = raw[3 * (k - 1) + 1]; w = raw[3 * (k - 1) + 2]; }
    obj[3 * idx] = (float)(x / 170.0); obj[3 * idx + 1] = (float)(y / 170.0); obj[3 * idx + 2] = (float)(w / 170.0);
    pos[2 * idx] = inference_glue ? x / 170.0 : x; pos[2 * idx + 1] = inference_glue ? y / 170.0 : y;
  }
}
// sums[t] = sum over the blocks of tower t of probs, accumulated in double in block order -- the callers' Python loop
// `stability_sum += s[0]` (JengaBuilder.py:254-256, TowerCreator.py:299-300); one thread per tower
__global__ void __launch_bounds__(256) k_tower_sums(const float* __restrict__ probs, const int32_t* __restrict__ node_off, int T,
                                                    double* __restrict__ sums) {
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < T; t += gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int i = node_off[t]; i < node_off[t + 1]; ++i) s += (double)probs[i];
    sums[t] = s;
  }
}
// np.argmin: index of the FIRST minimum (one CTA; T is a candidate count, tens to thousands)
__global__ void __launch_bounds__(256) k_argmin(const double* __restrict__ v, int T, int32_t* __restrict__ out) {
  __shared__ double sv[256];
  __shared__ int si[256];
  double best = 0.0; int bi = -1;
  for (int t = threadIdx.x; t < T; t += blockDim.x)
    if (bi < 0 || v[t] < best) { best = v[t]; bi = t; }     // ascending t per thread: keeps the first minimum
  sv[threadIdx.x] = best; si[threadIdx.x] = bi;
  __syncthreads();
  for (int off = 128; off >= 1; off >>= 1) {
    if ((int)threadIdx.x < off) {
      const int j = si[threadIdx.x + off];
      if (j >= 0 && (si[threadIdx.x] < 0 || sv[threadIdx.x + off] < sv[threadIdx.x] ||
                     (sv[threadIdx.x + off] == sv[threadIdx.x] && j < si[threadIdx.x]))) {
        sv[threadIdx.x] = sv[threadIdx.x + off]; si[threadIdx.x] = j;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = si[0];
}

}  // namespace spw
