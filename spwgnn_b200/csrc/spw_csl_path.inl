// spw_csl_path.inl -- launch sequences of the column-slab data path (included by spwgnn.cu inside its anonymous namespace;
// GPU build only).  Same math and the same reference citations as the round-1 sequences in spwgnn.cu:
// only the activation layout (spw_csl.cuh) and the kernels changed.

// ---- workspace ------------------------------------------------------------------------------------------------------------
struct LayoutC {
  size_t tcp[T_COUNT];
  size_t W2hi, W2lo, W2Thi, W2Tlo, ENCT;
  size_t degf, Q1, Q, QV, GP, U, S, R, H2S, PF, PL;
  size_t X0, X1, X2, C, A, H1;            // H1: h1 of the five steps (training)
  size_t EB, M1, M2;                      // sign bits (bytes, addressed in floats here): 4 / 5 / 5 arrays of 20 x bits_rows bytes
  size_t dU, dG, T, dS, dR, dH2S, DP, dQ, dQ1, dA, DH1, GB, partE, part0, partN;
  int slots, slotsGP;                     // per-step slots of U / S / R / H2S (5 in training, 1 otherwise) and of GP (5 / 2)
  long long bits_rows, bits_floats;       // rows (E rounded up to whole tiles) and floats of one sign-bit array
  size_t part_stride, part0_stride;       // floats between the partial buffers of two weight-gradient jobs (partN[kPartJobs], part0[3])
  size_t total;
};

constexpr int kPartJobs = 11;            // 7 node-level + 4 relation-encoder weight gradients
constexpr int kQ100 = 26, kQ150 = 38, kQ200 = 50;   // allocated column quads of 100 / 150 / 200-wide arrays

LayoutC make_layout_c(int64_t n, int64_t E, int training) {
  LayoutC L;
  size_t off = 0;
  auto take = [&](size_t floats) { size_t o = off; off = align_up(off + floats, 64); return o; };
  for (int i = 0; i < T_COUNT; ++i) L.tcp[i] = take(2 * tc_floats(i));
  L.W2hi = take(24320); L.W2lo = take(24320); L.W2Thi = take(24320); L.W2Tlo = take(24320);
  L.ENCT = take((size_t)8 * 24320);
  L.slots = training ? 5 : 1;
  L.slotsGP = training ? 5 : 2;
  L.degf = take(n);
  L.Q1 = take((size_t)kQ100 * n * 4);
  L.Q = take((size_t)kQ100 * n * 4);
  L.QV = take((size_t)kQ100 * n * 4);
  L.GP = take((size_t)kQ200 * L.slotsGP * n * 4);
  L.U = take((size_t)kQ100 * L.slots * n * 4);
  L.S = take((size_t)kQ150 * L.slots * n * 4);
  L.R = take((size_t)kQ150 * L.slots * n * 4);
  L.H2S = take((size_t)kQ150 * L.slots * n * 4);
  const size_t nchunks = (size_t)((E + 31) / 32) + 1;
  L.PF = take(nchunks * 160);
  L.PL = take(nchunks * 160);
  const size_t earr = (size_t)kQ150 * E * 4 + 64;
  L.A = take(earr);
  L.bits_rows = (long long)((E + 127) / 128) * 128 + 128;
  L.bits_floats = (20 * L.bits_rows + 3) / 4;
  if (training) {
    L.X0 = take(earr); L.X1 = take(earr); L.X2 = take(earr); L.C = take(earr);
    L.H1 = take((size_t)5 * earr);
    L.EB = take((size_t)4 * L.bits_floats); L.M1 = take((size_t)5 * L.bits_floats); L.M2 = take((size_t)5 * L.bits_floats);
    L.dU = take((size_t)kQ100 * 5 * n * 4);
    L.dG = take((size_t)kQ100 * 5 * n * 4);
    L.T = take((size_t)kQ100 * 4 * n * 4);
    L.dS = take((size_t)kQ150 * 4 * n * 4);
    L.dR = take((size_t)kQ150 * 4 * n * 4);
    L.dH2S = take((size_t)kQ150 * n * 4);
    L.DP = take((size_t)kQ100 * n * 4);
    L.dQ = take((size_t)kQ100 * n * 4);
    L.dQ1 = take((size_t)kQ100 * n * 4);
    L.dA = take(earr); L.DH1 = take(earr); L.GB = take(earr);
    L.partE = take((size_t)kMaxCtas * 2 * 160 * 128);
    L.part_stride = (size_t)kMaxCtas * 2 * 160 * 128;          // every weight gradient keeps its per-CTA partials until the ONE reduction launch
    L.partN = take((size_t)kPartJobs * L.part_stride);
    const size_t nsk = (size_t)((E > n ? E : n) + csl::kSkinnyRows - 1) / csl::kSkinnyRows + 1;
    L.part0_stride = nsk * 3 * kDEP;
    L.part0 = take(3 * L.part0_stride);
  } else {
    L.X0 = L.X2 = L.A; L.X1 = L.C = take(earr);       // inference: the encoder ping-pongs between two buffers
    L.EB = L.M1 = L.M2 = 0; L.H1 = 0;
    L.dU = L.dG = L.T = L.dS = L.dR = L.dH2S = L.DP = L.dQ = L.dQ1 = L.dA = L.DH1 = L.GB = L.partE = L.part0 = L.partN = 0;
    L.part_stride = L.part0_stride = 0;
  }
  L.total = off;
  return L;
}

// ---- helpers ----------------------------------------------------------------------------------------------------------------
// view of `rows_alloc`-row CSL array at `base`, starting at row `row0` and column `col0`
csl::View cview(float* base, long long rows_alloc, long long row0, int col0) {
  csl::View v; v.p = base + row0 * 4; v.slab = rows_alloc * 4; v.col0 = col0; return v;
}

struct LinC {
  const char* tag; int id;             // profile name, packed operand (TcId); id < 0: explicit operand pointers
  const float* Bhi = nullptr; const float* Blo = nullptr;
  int N, K; uint32_t epi;
  csl::View X, Y;
  const float* bias = nullptr; const float* rowscale = nullptr;
  csl::View addend{nullptr, 0, 0}, mulsrc{nullptr, 0, 0};
  const uint8_t* bits_in = nullptr; uint8_t* bits_out = nullptr; long long bits_rows = 0;
  float post_scale = 1.f; uint32_t drop_thresh = 0, drop_seed = 0; float drop_inv_keep = 1.f; int drop_stride = 128;
  int ones_col = -1; int write_pad = 1;
};

int run_lin_c(cudaStream_t st, float* ws, const LayoutC& L, int M, const LinC& o) {
  csl::LinCArgs a;
  memset(&a, 0, sizeof(a));
  a.M = M; a.K = o.K; a.nks = (o.K + 7) / 8; a.N = o.N;
  a.X = o.X; a.Y = o.Y;
  if (o.id >= 0) { a.Bhi = ws + L.tcp[o.id]; a.Blo = a.Bhi + tc_floats(o.id); } else { a.Bhi = o.Bhi; a.Blo = o.Blo; }
  a.bias = o.bias; a.rowscale = o.rowscale; a.addend = o.addend; a.mulsrc = o.mulsrc;
  a.bits_in = o.bits_in; a.bits_in_rows = o.bits_rows; a.bits_out = o.bits_out; a.bits_out_rows = o.bits_rows;
  a.post_scale = o.post_scale; a.drop_thresh = o.drop_thresh; a.drop_seed = o.drop_seed; a.drop_inv_keep = o.drop_inv_keep;
  a.drop_stride = o.drop_stride; a.ones_col = o.ones_col; a.write_pad = o.write_pad; a.poison = o.Y.p;
  return launch_lin(st, o.epi, a, o.tag);
}

// ---- TMA tensor maps (driver entry point resolved at run time: the library does not link libcuda) ---------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}
// 2-D map of a column-slab view: [nquads][rows * 4 floats] (quad stride = slab floats), box [box_quads][33 rows x 4 floats]
int make_csl_map(CUtensorMap* m, const csl::View& v, long long rows, int nquads, int box_quads) {
  EncodeTiledFn enc = encode_tiled();
  if (!enc) return fail(SPW_ERR_LAUNCH, "cuTensorMapEncodeTiled is not available");
  const cuuint64_t gdim[2] = {(cuuint64_t)rows * 4, (cuuint64_t)nquads};
  const cuuint64_t gstr[1] = {(cuuint64_t)v.slab * 4};
  const cuuint32_t box[2] = {(cuuint32_t)csl::kQPitch, (cuuint32_t)box_quads};
  const cuuint32_t estr[2] = {1, 1};
  float* base = v.p + (long long)(v.col0 >> 2) * v.slab;
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(SPW_ERR_LAUNCH, "cuTensorMapEncodeTiled failed (%d): rows %lld, %d quads, slab %lld", (int)r, rows, nquads, v.slab);
  return SPW_OK;
}

// dW[Kx][Ny] (+ bias row) = X^T dY over M rows, reduced in fixed order into the Keras-layout gradient.
//   gather_rcv == null: dY is a column-slab view streamed like X;  else: dY is a node table gathered by gather_rcv[row] and masked
//   with the byte-slab relu bits (the edge step's d h2).  reduce == false: the launch only updates the per-CTA partials (first:
//   initialises them) and the caller reduces them later (the five step launches of rmp layer 1).
int run_wgrad_c(cudaStream_t st, int M, const csl::View& X, int Kx, const float* rowscale, int rsmod, const csl::View& dY, int Ny, float* part,
                const WgOut& out, const char* tag, const int32_t* gather_rcv = nullptr, const uint8_t* bits = nullptr, long long bits_rows = 0,
                int first = 1, bool reduce = true, int* streams_out = nullptr, long long y_rows = 0) {
  if (M <= 0) return SPW_OK;
  csl::WgradCArgs a;
  memset(&a, 0, sizeof(a));
  a.M = M; a.Kx = Kx; a.rowscale = rowscale; a.rsmod = rsmod;
  a.dY = dY.p; a.y_slab = dY.slab; a.y_col0 = dY.col0; a.Ny = Ny; a.rcv = gather_rcv; a.bits = bits; a.bits_rows = bits_rows;
  const int NB = Ny <= 112 ? 112 : 160;
  a.nmt = Kx + 1 > 128 ? 2 : 1; a.part = part; a.first = first; a.poison = part;
  const int ntiles = (M + kTM - 1) / kTM;
  int streams = num_sms() / a.nmt;
  if (streams > kMaxCtas) streams = kMaxCtas;
  if (streams > ntiles) streams = ntiles;
  if (streams_out) *streams_out = streams;
  const int f1 = Kx + 1 - 128;
  const int nqx0 = ((Kx < 128 ? Kx : 128) + 3) >> 2, nqx1 = a.nmt == 2 ? ((Kx + 3) >> 2) - (f1 >> 2) : 0;
  const int nqx = nqx0 > nqx1 ? nqx0 : nqx1, nqy = (Ny + 3) >> 2;
  CUtensorMap tmX, tmY;
  int rc;
  if ((rc = make_csl_map(&tmX, X, M, (Kx + 3) >> 2, nqx)) != SPW_OK) return rc;
  const bool pair = a.nmt == 2;                                 // two M-tiles: one CTA pair per row stream (k_wgrad_pair)
  if (gather_rcv) {                                              // node table [nqy quads][y_rows nodes x 4]: node-range copies of the pair kernel
    static const bool no_tma = getenv("SPW_WG_NO_TMA_GATHER") != nullptr;      // A/B switch, and the way the tests reach the cp.async path
    a.gather_tma = pair && y_rows > 0 && !no_tma;
    if (!a.gather_tma) tmY = tmX;
    else if ((rc = make_csl_map(&tmY, dY, y_rows, nqy, NB / 8)) != SPW_OK) return rc;
  } else if ((rc = make_csl_map(&tmY, dY, M, nqy, pair ? NB / 8 : nqy)) != SPW_OK) return rc;
  const size_t smem = pair ? csl::wgrad_pair_smem(nqx, NB, 4) : csl::wgrad_c_smem(nqx, nqy, NB, 3);      // stage ring: 4 slots (pair) / 3
#define SPW_WG_LAUNCH(KERN, YM, NBV, NSTV)                                                                                     \
  do { auto kern = csl::KERN<YM, NBV, NSTV>; set_smem(kern, smem);                                                           \
       SPW_KLAUNCH_PDL(tag, kern, dim3(streams * a.nmt), dim3(csl::kThreadsC), smem, st, tmX, tmY, a, nqx); } while (0)
  if (gather_rcv && !pair) return fail(SPW_ERR_UNSUPPORTED, "run_wgrad_c: the gathered form exists for two M-tiles only (Kx = %d)", Kx);
  if (pair) {
    if (gather_rcv) SPW_WG_LAUNCH(k_wgrad_pair, 1, 160, 4);
    else if (NB == 160) SPW_WG_LAUNCH(k_wgrad_pair, 0, 160, 4);
    else SPW_WG_LAUNCH(k_wgrad_pair, 0, 112, 4);
  } else {
    if (NB == 160) SPW_WG_LAUNCH(k_wgrad_c, 0, 160, 3);
    else SPW_WG_LAUNCH(k_wgrad_c, 0, 112, 3);
  }
#undef SPW_WG_LAUNCH
  if (reduce) launch_reduce(st, part, streams, (int)tc::kWgPartFloats, 0, -1, f1 > 0 ? f1 : 0, Kx, Ny, out);
  return SPW_OK;
}

// sign-bit array i of a group (EB / M1 / M2)
uint8_t* bits_ptr(float* ws, size_t base, const LayoutC& L, int i) { return reinterpret_cast<uint8_t*>(ws + base + (size_t)i * L.bits_floats); }

// ---- forward ----------------------------------------------------------------------------------------------------------------
int forward_csl(const SpwParams* w, const SpwGraph* g, const float* obj, float* logits, float* probs, float* ws, size_t workspace_bytes,
                int training, float dropout_rate, uint64_t dropout_seed, cudaStream_t st) {
  const int n = g->n_nodes, E = g->n_edges;
  const bool drop = training && dropout_rate > 0.f;
  const uint32_t drop_thresh = drop ? (uint32_t)(dropout_rate * 16777216.0f) : 0u;
  const float inv_keep = drop ? 1.f / (1.f - dropout_rate) : 1.f;
  const uint32_t seed_c = (uint32_t)dropout_seed, seed_q = (uint32_t)(dropout_seed >> 32) ^ 0x5bd1e995u ^ (uint32_t)dropout_seed * 3u;
  const LayoutC L = make_layout_c(n, E, training);
  if (workspace_bytes < L.total * sizeof(float))
    return fail(SPW_ERR_WORKSPACE, "spw_forward: workspace %zu < %zu bytes", workspace_bytes, L.total * sizeof(float));
  int rc;
  // weights -> tensor-core operands
  {
    tc::PackTcArgs pa;   // same table as pack_tc (round 1), into this layout
    memset(&pa, 0, sizeof(pa));
    int k = 0;
    auto add = [&](int id, const float* src, int ld, int row0, int col0, int K, int N, int transpose) {
      tc::PackTcDesc& d = pa.d[k++];
      d.src = src; d.ld = ld; d.row0 = row0; d.col0 = col0; d.K = K; d.N = N; d.transpose = transpose;
      d.hi = ws + L.tcp[id]; d.lo = d.hi + tc_floats(id); d.NB = kTcShape[id].NB; d.k_off = 0; d.k_lim = 8 * kTcShape[id].ks;
    };
    add(T_OM1, w->om_w[1], 100, 0, 0, 100, 100, 0);
    add(T_RM1, w->rm_w[1], 150, 0, 0, 150, 150, 0);
    add(T_RM2, w->rm_w[2], 150, 0, 0, 150, 150, 0);
    add(T_RM3, w->rm_w[3], 150, 0, 0, 150, 150, 0);
    add(T_W1A, w->rmp_w[0], 150, 0, 0, 150, 150, 0);      // Networks.py:86 concat order: [rel_enc | sender | receiver]
    add(T_W1B, w->rmp_w[0], 150, 150, 0, 100, 150, 0);
    add(T_W1C, w->rmp_w[0], 150, 250, 0, 100, 150, 0);
    add(T_W3, w->rmp_w[2], 100, 0, 0, 150, 100, 0);
    add(T_V1A, w->omp_w[0], 100, 0, 0, 100, 100, 0);      // Networks.py:89 concat order: [obj_enc | effect | prop]
    add(T_V1BC, w->omp_w[0], 100, 100, 0, 200, 100, 0);   // rows [V1b ; V1c] are consecutive: one K = 200 operand for [g | p]
    add(T_V2P, w->omp_w[1], 101, 0, 1, 100, 100, 0);      // channels 1..100 (Networks.py:80)
    if (training) {
      add(T_V2PT, w->omp_w[1], 101, 0, 1, 100, 100, 1);
      add(T_V1AT, w->omp_w[0], 100, 0, 0, 100, 100, 1);
      add(T_V1BT, w->omp_w[0], 100, 100, 0, 100, 100, 1);
      add(T_V1CT, w->omp_w[0], 100, 200, 0, 100, 100, 1);
      add(T_W3T, w->rmp_w[2], 100, 0, 0, 100, 150, 1);
      add(T_W1BT, w->rmp_w[0], 150, 150, 0, 150, 100, 1);
      add(T_W1CT, w->rmp_w[0], 150, 250, 0, 150, 100, 1);
      add(T_OM1T, w->om_w[1], 100, 0, 0, 100, 100, 1);
    }
    pa.n = k;
    SPW_KLAUNCH("k_pack_tc", tc::k_pack_tc, dim3(16, k), dim3(256), 0, st, pa);
    SPW_KLAUNCH("k_pack_umma", tc::k_pack_umma, dim3(48), dim3(256), 0, st, (const float*)w->rmp_w[1], 150, 0, 0, 150, 150, 0, (const float*)w->rmp_b[1], ws + L.W2hi, ws + L.W2lo);
    if (training) {
      SPW_KLAUNCH("k_pack_umma", tc::k_pack_umma, dim3(48), dim3(256), 0, st, (const float*)w->rmp_w[1], 150, 0, 0, 150, 150, 1, (const float*)nullptr, ws + L.W2Thi, ws + L.W2Tlo);
      const float* encw[4] = {w->rmp_w[0], w->rm_w[3], w->rm_w[2], w->rm_w[1]};       // W1a, RM3, RM2, RM1
      for (int i = 0; i < 4; ++i)
        SPW_KLAUNCH("k_pack_umma", tc::k_pack_umma, dim3(48), dim3(256), 0, st, encw[i], 150, 0, 0, 150, 150, 1, (const float*)nullptr,
                    ws + L.ENCT + (size_t)(2 * i) * 24320, ws + L.ENCT + (size_t)(2 * i + 1) * 24320);
    }
  }
  SPW_KLAUNCH_PDL("k_deg_to_float", k_deg_to_float, dim3(grid_for(n, 256)), dim3(256), 0, st, g->in_off, n, ws + L.degf);

  const long long rGP = (long long)L.slotsGP * n, rS = (long long)L.slots * n;
  auto gp_slot = [&](int l) { return training ? l : (l & 1); };
  auto st_slot = [&](int l) { return training ? l : 0; };

  // object encoder (Networks.py:47,76): q1 = relu(om0([y, w])), q = relu(om1(q1)) (+ dropout, Networks.py:78); qv = q.V1a + c1
  SPW_KLAUNCH_PDL("k_obj_enc0_c", csl::k_obj_enc0_c, dim3(grid_for((int64_t)n * csl::kQP, 256)), dim3(256), 0, st, obj, n, w->om_w[0], w->om_b[0],
              ws + L.Q1, (long long)n * 4);
  {
    LinC o; o.tag = "k_lin:node"; o.id = T_OM1; o.N = 100; o.K = 100; o.epi = csl::EPI_BIAS | csl::EPI_RELU | (drop ? csl::EPI_DROP : 0u);
    o.X = cview(ws + L.Q1, n, 0, 0); o.Y = cview(ws + L.Q, n, 0, 0); o.bias = w->om_b[1];
    o.drop_thresh = drop_thresh; o.drop_seed = seed_q; o.drop_inv_keep = inv_keep;
    if ((rc = run_lin_c(st, ws, L, n, o)) != SPW_OK) return rc;
  }
  {
    LinC o; o.tag = "k_lin:node"; o.id = T_V1A; o.N = 100; o.K = 100; o.epi = csl::EPI_BIAS;
    o.X = cview(ws + L.Q, n, 0, 0); o.Y = cview(ws + L.QV, n, 0, 0); o.bias = w->omp_b[0];
    if ((rc = run_lin_c(st, ws, L, n, o)) != SPW_OK) return rc;
  }
  // relation encoder + A_e (Networks.py:46,75 and the c_e part of :86-87)
  if (E > 0) {
    float* Xs[4] = {ws + L.X0, ws + L.X1, ws + L.X2, ws + L.C};
    uint8_t* EB[4];
    for (int i = 0; i < 4; ++i) EB[i] = training ? bits_ptr(ws, L.EB, L, i) : nullptr;
    SPW_KLAUNCH_PDL("k_edge_enc0_c", csl::k_edge_enc0_c, dim3(grid_for(E, 256)), dim3(256), 0, st, E, g->in_snd, g->in_rcv, obj,
                (const float*)w->rm_w[0], (const float*)w->rm_b[0], Xs[0], EB[0], L.bits_rows);
    const int ids[3] = {T_RM1, T_RM2, T_RM3};
    const float* bs[3] = {w->rm_b[1], w->rm_b[2], w->rm_b[3]};
    for (int i = 0; i < 3; ++i) {
      LinC o; o.tag = "k_lin:enc_fwd"; o.id = ids[i]; o.N = 150; o.K = 150;
      o.epi = csl::EPI_BIAS | csl::EPI_RELU | csl::EPI_ONES | (training ? csl::EPI_BITS_OUT : 0u);
      o.X = cview(Xs[i], E, 0, 0); o.Y = cview(Xs[i + 1], E, 0, 0); o.bias = bs[i]; o.ones_col = kDE;
      o.bits_out = EB[i + 1]; o.bits_rows = L.bits_rows;
      if (i == 2 && drop) {                                       // dropout on c_e (Networks.py:77)
        o.epi |= csl::EPI_DROP; o.drop_thresh = drop_thresh; o.drop_seed = seed_c; o.drop_inv_keep = inv_keep; o.drop_stride = 160;
      }
      if ((rc = run_lin_c(st, ws, L, E, o)) != SPW_OK) return rc;
    }
    LinC o; o.tag = "k_lin:enc_fwd"; o.id = T_W1A; o.N = 150; o.K = 150; o.epi = csl::EPI_BIAS;
    o.X = cview(Xs[3], E, 0, 0); o.Y = cview(ws + L.A, E, 0, 0); o.bias = w->rmp_b[0];
    if ((rc = run_lin_c(st, ws, L, E, o)) != SPW_OK) return rc;
  }
  // p^0 = 0 (main.py:68), hence S^1 = R^1 = 0: slot 0 of the [quad][slots * n][4] arrays, one strided fill each
  cudaMemset2DAsync(ws + L.GP + (size_t)25 * rGP * 4 + (size_t)gp_slot(0) * n * 4, (size_t)rGP * 16, 0, (size_t)n * 16, 25, st);
  cudaMemset2DAsync(ws + L.S, (size_t)rS * 16, 0, (size_t)n * 16, kQ150, st);
  cudaMemset2DAsync(ws + L.R, (size_t)rS * 16, 0, (size_t)n * 16, kQ150, st);

  const int ttiles = (E + kTM - 1) / kTM;
  const int tgrid = ttiles < num_sms() ? ttiles : num_sms();
  for (int l = 0; l < SPW_N_STEPS; ++l) {                        // Networks.py:83
    const int sl = st_slot(l), gl = gp_slot(l);
    float* S = ws + L.S; float* R = ws + L.R;
    if (l > 0) {   // S = P.W1b, R = P.W1c   (sender / receiver parts of rmp layer 0, Networks.py:84-87)
      LinC o; o.tag = "k_lin:node"; o.N = 150; o.K = 100; o.epi = 0u;
      o.X = cview(ws + L.GP, rGP, (long long)gl * n, 100);
      o.id = T_W1B; o.Y = cview(S, rS, (long long)sl * n, 0);
      if ((rc = run_lin_c(st, ws, L, n, o)) != SPW_OK) return rc;
      o.id = T_W1C; o.Y = cview(R, rS, (long long)sl * n, 0);
      if ((rc = run_lin_c(st, ws, L, n, o)) != SPW_OK) return rc;
    }
    csl::View H = cview(ws + L.H2S, rS, (long long)sl * n, 0);
    if (E > 0) {
      csl::EdgeStepCArgs t;
      memset(&t, 0, sizeof(t));
      t.E = E; t.in_snd = g->in_snd; t.in_rcv = g->in_rcv; t.in_off = g->in_off; t.A = ws + L.A;
      t.S = S + (size_t)sl * n * 4; t.R = R + (size_t)sl * n * 4; t.sr_slab = rS * 4;
      t.W2hi = ws + L.W2hi; t.W2lo = ws + L.W2lo; t.H2S = H.p; t.h_slab = H.slab;
      t.part_first = ws + L.PF; t.part_last = ws + L.PL;
      t.bits_h2 = training ? bits_ptr(ws, L.M2, L, l) : nullptr; t.bits_h1 = training ? bits_ptr(ws, L.M1, L, l) : nullptr;
      t.bits_rows = L.bits_rows; t.poison = H.p;
      t.H1 = training ? ws + L.H1 + (size_t)l * ((size_t)kQ150 * E * 4 + 64) : nullptr;
      set_smem(csl::k_edge_step_c, csl::kEdgeStepCSmem);
      SPW_KLAUNCH_PDL("k_edge_step_c", csl::k_edge_step_c, dim3(tgrid), dim3(csl::kThreadsC), csl::kEdgeStepCSmem, st, t);
    }
    // segments that cross a 32-row chunk, and nodes without incoming relations (all-zero aggregate)
    SPW_KLAUNCH_PDL("k_seg_fix_c", csl::k_seg_fix_c, dim3(grid_for((int64_t)n * csl::kQE, 256)), dim3(256), 0, st, n, g->in_off,
                (const float*)(ws + L.PF), (const float*)(ws + L.PL), H.p, H.slab);
    {   // g = tanh(W3.sum h2 + deg.b3)   (Networks.py:87-88), written into columns 0..99 of [g | p]
      LinC o; o.tag = "k_lin:node"; o.id = T_W3; o.N = 100; o.K = 150; o.epi = csl::EPI_BIAS | csl::EPI_ROWSCALE | csl::EPI_TANH;
      o.X = H; o.Y = cview(ws + L.GP, rGP, (long long)gl * n, 0); o.bias = w->rmp_b[2]; o.rowscale = ws + L.degf; o.write_pad = 0;
      if ((rc = run_lin_c(st, ws, L, n, o)) != SPW_OK) return rc;
    }
    csl::View Uv = cview(ws + L.U, rS, (long long)sl * n, 0);
    {   // u = relu(V1.[q, g, p] + c1) = relu(qv + [g | p].[V1b ; V1c])    (Networks.py:89-90, hidden layer of omp)
      LinC o; o.tag = "k_lin:node"; o.id = T_V1BC; o.N = 100; o.K = 200; o.epi = csl::EPI_ADD | csl::EPI_RELU;
      o.X = cview(ws + L.GP, rGP, (long long)gl * n, 0); o.Y = Uv; o.addend = cview(ws + L.QV, n, 0, 0);
      if ((rc = run_lin_c(st, ws, L, n, o)) != SPW_OK) return rc;
    }
    if (l < SPW_N_STEPS - 1) {   // p = tanh(z[1:] + p)   (Networks.py:80,91), written into columns 100..199 of the next [g | p]
      LinC o; o.tag = "k_lin:node"; o.id = T_V2P; o.N = 100; o.K = 100; o.epi = csl::EPI_BIAS | csl::EPI_ADD | csl::EPI_TANH;
      o.X = Uv; o.Y = cview(ws + L.GP, rGP, (long long)gp_slot(l + 1) * n, 100); o.bias = w->omp_b[1] + 1;
      o.addend = cview(ws + L.GP, rGP, (long long)gl * n, 100); o.write_pad = 0;
      if ((rc = run_lin_c(st, ws, L, n, o)) != SPW_OK) return rc;
    } else {                     // head: channel 0 of the last z (Networks.py:93-96)
      SPW_KLAUNCH_PDL("k_logit_c", csl::k_logit_c, dim3(grid_for(n, 256)), dim3(256), 0, st, (const float*)Uv.p, Uv.slab, n, (const float*)w->omp_w[1],
                  (const float*)w->omp_b[1], logits, probs);
    }
  }
  return check_launch("spw_forward");
}

// ---- backward -----------------------------------------------------------------------------------------------------------------
int backward_csl(const SpwParams* w, const SpwGraph* g, const float* obj, const float* dlogits, float* ws, size_t workspace_bytes,
                 const SpwParams* grads, float dropout_rate, cudaStream_t st) {
  const int n = g->n_nodes, E = g->n_edges;
  const float inv_keep = dropout_rate > 0.f ? 1.f / (1.f - dropout_rate) : 1.f;
  const LayoutC L = make_layout_c(n, E, 1);
  if (workspace_bytes < L.total * sizeof(float))
    return fail(SPW_ERR_WORKSPACE, "spw_backward: workspace %zu < %zu bytes", workspace_bytes, L.total * sizeof(float));
  int rc;
  const long long r5 = 5LL * n, r4 = 4LL * n;
  const int etiles = (E + kTM - 1) / kTM;
  // every weight gradient keeps its own per-CTA partials; their fixed-order reductions run as ONE launch at the end of the pass
  std::vector<RedArgs> red_jobs;
  struct DeferGuard { DeferGuard(std::vector<RedArgs>* v) { t_deferred_reduces = v; } ~DeferGuard() { t_deferred_reduces = nullptr; } } defer_guard(&red_jobs);
  int part_job = 0;
  auto next_part = [&]() { return ws + L.partN + (size_t)(part_job++ % kPartJobs) * L.part_stride; };

  // head: dUpre^5 = dlogit (x) V2[:,0] * relu'
  const csl::View U5 = cview(ws + L.U, r5, 4LL * n, 0);
  SPW_KLAUNCH_PDL("k_logit_bwd_c", csl::k_logit_bwd_c, dim3(grid_for((int64_t)n * csl::kQP, 256)), dim3(256), 0, st, dlogits, (const float*)U5.p, U5.slab, n,
              (const float*)w->omp_w[1], ws + L.dU + (size_t)4 * n * 4, r5 * 4);

  int wstreams = 0;                                              // row streams of the edge-step weight gradient
  for (int l = SPW_N_STEPS - 1; l >= 0; --l) {   // step l+1 of the forward loop
    const csl::View dU = cview(ws + L.dU, r5, (long long)l * n, 0), dG = cview(ws + L.dG, r5, (long long)l * n, 0);
    const csl::View Ul = cview(ws + L.U, r5, (long long)l * n, 0);
    const csl::View Gl = cview(ws + L.GP, r5, (long long)l * n, 0), Pl = cview(ws + L.GP, r5, (long long)l * n, 100);
    const csl::View Tl = cview(ws + L.T, r4, (long long)(l < 4 ? l : 0) * n, 0);       // d(pre-tanh of p^{l+1}), l < 4
    const csl::View DP = cview(ws + L.DP, n, 0, 0), dH = cview(ws + L.dH2S, n, 0, 0);
    if (l < SPW_N_STEPS - 1) {   // dUpre = (T.V2p^T) * relu'(u)
      LinC o; o.tag = "k_lin:node"; o.id = T_V2PT; o.N = 100; o.K = 100; o.epi = csl::EPI_MUL_POS; o.X = Tl; o.Y = dU; o.mulsrc = Ul;
      if ((rc = run_lin_c(st, ws, L, n, o)) != SPW_OK) return rc;
    }
    // (dq_pre, the gradient through the q part of omp layer 0, is the same linear map of every step's dUpre: one product of their sum, below)
    {   // dg_pre = (dUpre.V1b^T) * (1 - g^2)
      LinC o; o.tag = "k_lin:node"; o.id = T_V1BT; o.N = 100; o.K = 100; o.epi = csl::EPI_MUL_TANH; o.X = dU; o.Y = dG; o.mulsrc = Gl;
      if ((rc = run_lin_c(st, ws, L, n, o)) != SPW_OK) return rc;
    }
    if (l > 0) {   // DP = dUpre.V1c^T (+ residual T)
      LinC o; o.tag = "k_lin:node"; o.id = T_V1CT; o.N = 100; o.K = 100; o.epi = l < SPW_N_STEPS - 1 ? csl::EPI_ADD : 0u; o.X = dU; o.Y = DP;
      if (l < SPW_N_STEPS - 1) o.addend = Tl;
      if ((rc = run_lin_c(st, ws, L, n, o)) != SPW_OK) return rc;
    }
    {   // d(sum h2) = dg_pre.W3^T
      LinC o; o.tag = "k_lin:node"; o.id = T_W3T; o.N = 150; o.K = 100; o.epi = 0u; o.X = dG; o.Y = dH;
      if ((rc = run_lin_c(st, ws, L, n, o)) != SPW_OK) return rc;
    }
    if (E > 0) {
      const uint8_t* m2 = bits_ptr(ws, L.M2, L, l); const uint8_t* m1 = bits_ptr(ws, L.M1, L, l);
      {   // dW2 += h1^T . d h2 with h1 kept by the forward pass and d h2 = relu'(h2) * d(sum h2)[receiver]
        const csl::View H1v = cview(ws + L.H1 + (size_t)l * ((size_t)kQ150 * E * 4 + 64), E, 0, 0);
        if ((rc = run_wgrad_c(st, E, H1v, kDE, nullptr, 0, dH, kDE, ws + L.partE, WgOut{nullptr, 0, 0, 0, nullptr, 0}, "k_wgrad_c:step", g->in_rcv, m2,
                              L.bits_rows, l == SPW_N_STEPS - 1, false, &wstreams, n)) != SPW_OK) return rc;
      }
      {
        csl::EdgeDgradCArgs t;
        memset(&t, 0, sizeof(t));
        t.E = E; t.in_rcv = g->in_rcv; t.dH2S = dH.p; t.d_slab = dH.slab; t.Whi = ws + L.W2Thi; t.Wlo = ws + L.W2Tlo;
        t.bits_h2 = m2; t.bits_h1 = m1; t.bits_rows = L.bits_rows; t.dA = ws + L.dA; t.DH1 = l > 0 ? ws + L.DH1 : nullptr;       // step 0 starts from p = 0: no d S / d R to form
        t.first = (l == SPW_N_STEPS - 1); t.poison = ws + L.dA;
        const int tgrid = etiles < num_sms() ? etiles : num_sms();
        if (t.first) {
          set_smem(csl::k_edge_dgrad_c<true>, csl::kEdgeStepCSmem);
          SPW_KLAUNCH_PDL("k_edge_dgrad_c", csl::k_edge_dgrad_c<true>, dim3(tgrid), dim3(csl::kThreadsC), csl::kEdgeStepCSmem, st, t);
        } else {
          set_smem(csl::k_edge_dgrad_c<false>, csl::kEdgeStepCSmem);
          SPW_KLAUNCH_PDL("k_edge_dgrad_c", csl::k_edge_dgrad_c<false>, dim3(tgrid), dim3(csl::kThreadsC), csl::kEdgeStepCSmem, st, t);
        }
      }
    }
    if (l > 0) {
      const csl::View dS = cview(ws + L.dS, r4, (long long)(l - 1) * n, 0), dR = cview(ws + L.dR, r4, (long long)(l - 1) * n, 0);
      if (E > 0) {
        SPW_KLAUNCH_PDL("k_gather_dsr_c", csl::k_gather_dsr_c, dim3(grid_for((int64_t)n * csl::kQE, 256)), dim3(256), 0, st, n, E, g->in_off, g->out_off,
                    g->out_pos, (const float*)(ws + L.DH1), dS.p, dR.p, r4 * 4);
      } else {
        cudaMemset2DAsync(dS.p, (size_t)r4 * 16, 0, (size_t)n * 16, kQ150, st);
        cudaMemset2DAsync(dR.p, (size_t)r4 * 16, 0, (size_t)n * 16, kQ150, st);
      }
      // T^{l} = (dS.W1b^T + dR.W1c^T + DP) * (1 - (p^l)^2)       (p^l = input state of this step)
      {
        LinC o; o.tag = "k_lin:node"; o.id = T_W1BT; o.N = 100; o.K = 150; o.epi = csl::EPI_ACC; o.X = dS; o.Y = DP;
        if ((rc = run_lin_c(st, ws, L, n, o)) != SPW_OK) return rc;
      }
      {
        LinC o; o.tag = "k_lin:node"; o.id = T_W1CT; o.N = 100; o.K = 150; o.epi = csl::EPI_ADD | csl::EPI_MUL_TANH; o.X = dR;
        o.Y = cview(ws + L.T, r4, (long long)(l - 1) * n, 0); o.addend = DP; o.mulsrc = Pl;
        if ((rc = run_lin_c(st, ws, L, n, o)) != SPW_OK) return rc;
      }
    }
  }

  // ---- node-level weight gradients, each one contraction over all steps' rows -------------------
  const csl::View dUall = cview(ws + L.dU, r5, 0, 0);
  // omp layer 0 = [V1a; V1b; V1c], bias c1
  //   rows 0..99 (V1a): q is the same in all five steps, so q^T (sum over the steps of dUpre): the sum lands in DP (free by now)
  SPW_KLAUNCH_PDL("k_sum_slots_c", csl::k_sum_slots_c, dim3(grid_for((int64_t)n * csl::kQP, 256)), dim3(256), 0, st, n, 5, csl::kQP, (const float*)(ws + L.dU), r5 * 4,
              ws + L.DP, (long long)n * 4);
  {   // dq_pre = ((sum over the steps of dUpre).V1a^T) * relu'(q), times 1/keep through the dropout on q: V1a and relu'(q) do not depend on the step
    LinC o; o.tag = "k_lin:node"; o.id = T_V1AT; o.N = 100; o.K = 100; o.epi = csl::EPI_MUL_POS | csl::EPI_SCALE;
    o.X = cview(ws + L.DP, n, 0, 0); o.Y = cview(ws + L.dQ, n, 0, 0); o.mulsrc = cview(ws + L.Q, n, 0, 0); o.post_scale = inv_keep;
    if ((rc = run_lin_c(st, ws, L, n, o)) != SPW_OK) return rc;
  }
  if ((rc = run_wgrad_c(st, n, cview(ws + L.Q, n, 0, 0), kDP, nullptr, 0, cview(ws + L.DP, n, 0, 0), kDP, next_part(), {grads->omp_w[0], 100, 0, 0, grads->omp_b[0], 0}, "k_wgrad_c:node")) != SPW_OK) return rc;
  if ((rc = run_wgrad_c(st, 5 * n, cview(ws + L.GP, r5, 0, 0), 2 * kDP, nullptr, 0, dUall, kDP, next_part(), {grads->omp_w[0], 100, 100, 0, nullptr, 0}, "k_wgrad_c:node")) != SPW_OK) return rc;
  // omp layer 1: channels 1..100 from T (steps 1..4), channel 0 from the head
  if ((rc = run_wgrad_c(st, 4 * n, cview(ws + L.U, r5, 0, 0), kDP, nullptr, 0, cview(ws + L.T, r4, 0, 0), kDP, next_part(), {grads->omp_w[1], 101, 0, 1, grads->omp_b[1], 1}, "k_wgrad_c:node")) != SPW_OK) return rc;
  {
    const int nw = (n + csl::kSkinnyRows - 1) / csl::kSkinnyRows;
    SPW_KLAUNCH_PDL("k_skinny_c", csl::k_skinny_c<2>, dim3((nw + 7) / 8, 4), dim3(256), 0, st, n, (const float*)U5.p, U5.slab, csl::kQP, 104, (const int32_t*)nullptr,
                (const int32_t*)nullptr, (const float*)nullptr, dlogits, ws + L.part0);
    launch_reduce(st, ws + L.part0, nw, 104, 1, 0, 0, kDP, 1, {grads->omp_w[1], 101, 0, 0, grads->omp_b[1], 0});
  }
  // rmp layer 2 (W3, b3 scaled by in-degree)
  if ((rc = run_wgrad_c(st, 5 * n, cview(ws + L.H2S, r5, 0, 0), kDE, ws + L.degf, n, cview(ws + L.dG, r5, 0, 0), kDP, next_part(), {grads->rmp_w[2], 100, 0, 0, grads->rmp_b[2], 0}, "k_wgrad_c:node")) != SPW_OK) return rc;
  // rmp layer 0 rows 150..349 (W1b, W1c): X = p^{l} for steps 2..5
  if ((rc = run_wgrad_c(st, 4 * n, cview(ws + L.GP, r5, n, 100), kDP, nullptr, 0, cview(ws + L.dS, r4, 0, 0), kDE, next_part(), {grads->rmp_w[0], 150, 150, 0, nullptr, 0}, "k_wgrad_c:node")) != SPW_OK) return rc;
  if ((rc = run_wgrad_c(st, 4 * n, cview(ws + L.GP, r5, n, 100), kDP, nullptr, 0, cview(ws + L.dR, r4, 0, 0), kDE, next_part(), {grads->rmp_w[0], 150, 250, 0, nullptr, 0}, "k_wgrad_c:node")) != SPW_OK) return rc;
  // object encoder
  if ((rc = run_wgrad_c(st, n, cview(ws + L.Q1, n, 0, 0), kDP, nullptr, 0, cview(ws + L.dQ, n, 0, 0), kDP, next_part(), {grads->om_w[1], 100, 0, 0, grads->om_b[1], 0}, "k_wgrad_c:node")) != SPW_OK) return rc;
  {
    LinC o; o.tag = "k_lin:node"; o.id = T_OM1T; o.N = 100; o.K = 100; o.epi = csl::EPI_MUL_POS;
    o.X = cview(ws + L.dQ, n, 0, 0); o.Y = cview(ws + L.dQ1, n, 0, 0); o.mulsrc = cview(ws + L.Q1, n, 0, 0);
    if ((rc = run_lin_c(st, ws, L, n, o)) != SPW_OK) return rc;
    const int nw = (n + csl::kSkinnyRows - 1) / csl::kSkinnyRows;
    SPW_KLAUNCH_PDL("k_skinny_c", csl::k_skinny_c<1>, dim3((nw + 7) / 8, 4), dim3(256), 0, st, n, (const float*)(ws + L.dQ1), (long long)n * 4, csl::kQP, 104,
                (const int32_t*)nullptr, (const int32_t*)nullptr, obj, (const float*)nullptr, ws + L.part0 + L.part0_stride);
    launch_reduce(st, ws + L.part0 + L.part0_stride, nw, 3 * 104, 104, 0, 0, 2, kDP, {grads->om_w[0], 100, 0, 0, grads->om_b[0], 0});
  }

  // ---- edge-level weight gradients --------------------------------------------------------------
  if (E > 0) {
    // rmp layer 1 (W2, b2) from the per-step kernel's per-CTA partials
    launch_reduce(st, ws + L.partE, wstreams, (int)tc::kWgPartFloats, 0, -1, kDE + 1 - 128, kDE, kDE, {grads->rmp_w[1], 150, 0, 0, grads->rmp_b[1], 0});
    // relation-encoder backward, layer by layer (activations X0, X1, X2, C saved by the forward pass)
    float* acts[4] = {ws + L.C, ws + L.X2, ws + L.X1, ws + L.X0};       // layer inputs: C, X2, X1, X0
    float* gw[4] = {grads->rmp_w[0], grads->rm_w[3], grads->rm_w[2], grads->rm_w[1]};
    float* gb[4] = {grads->rmp_b[0], grads->rm_b[3], grads->rm_b[2], grads->rm_b[1]};
    float* dY = ws + L.dA;
    float* gout[2] = {ws + L.DH1, ws + L.GB};
    for (int i = 0; i < 4; ++i) {
      if ((rc = run_wgrad_c(st, E, cview(acts[i], E, 0, 0), kDE, nullptr, 0, cview(dY, E, 0, 0), kDE, next_part(), {gw[i], 150, 0, 0, gb[i], 0}, "k_wgrad_c:enc")) != SPW_OK) return rc;
      // data gradient of the layer: (dY . W^T) * relu'(layer input), times 1/keep through the dropout on c_e
      LinC o; o.tag = "k_lin:enc_bwd"; o.id = -1; o.Bhi = ws + L.ENCT + (size_t)(2 * i) * 24320; o.Blo = ws + L.ENCT + (size_t)(2 * i + 1) * 24320;
      o.N = 150; o.K = 150; o.epi = csl::EPI_MUL_BITS | csl::EPI_SCALE; o.X = cview(dY, E, 0, 0); o.Y = cview(gout[i & 1], E, 0, 0);
      o.bits_in = bits_ptr(ws, L.EB, L, 3 - i); o.bits_rows = L.bits_rows; o.post_scale = i == 0 ? inv_keep : 1.f;
      if ((rc = run_lin_c(st, ws, L, E, o)) != SPW_OK) return rc;
      dY = gout[i & 1];
    }
    const int nw = (E + csl::kSkinnyRows - 1) / csl::kSkinnyRows;
    SPW_KLAUNCH_PDL("k_skinny_c", csl::k_skinny_c<0>, dim3((nw + 7) / 8, 4), dim3(256), 0, st, E, (const float*)dY, (long long)E * 4, csl::kQE, kDEP, g->in_snd, g->in_rcv,
                obj, (const float*)nullptr, ws + L.part0 + 2 * L.part0_stride);
    launch_reduce(st, ws + L.part0 + 2 * L.part0_stride, nw, 3 * kDEP, kDEP, 0, 0, 2, kDE, {grads->rm_w[0], 150, 0, 0, grads->rm_b[0], 0});
  } else {
    cudaMemsetAsync(grads->rmp_w[1], 0, 22500 * sizeof(float), st);
    cudaMemsetAsync(grads->rmp_b[1], 0, 150 * sizeof(float), st);
    cudaMemsetAsync(grads->rmp_w[0], 0, 150 * 150 * sizeof(float), st);   // rows 0..149 (rows 150.. written above)
    cudaMemsetAsync(grads->rmp_b[0], 0, 150 * sizeof(float), st);
    for (int i = 0; i < 4; ++i) {
      cudaMemsetAsync(grads->rm_w[i], 0, (i == 0 ? 300 : 22500) * sizeof(float), st);
      cudaMemsetAsync(grads->rm_b[i], 0, 150 * sizeof(float), st);
    }
  }
  flush_reduces(st, red_jobs);
  return check_launch("spw_backward");
}
