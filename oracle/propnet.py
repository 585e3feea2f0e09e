"""CPU ORACLE for the SPWGNN propagation-network hot path.  TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this module; the product (`spwgnn_b200/`) never does and fails loudly when
its CUDA library is missing.

What it restates (all citations under /root/reference/src/):
  * relation (edge) construction ............ main.py:66-81 (== TowerCreator.py:415-428,
                                              JengaBuilder.py:313-326)
  * feature normalisation .................... main.py:91
  * the Keras graph of the network ........... Networks.py:16-104 (MLPs: Blocks.py:12-91)
  * loss ..................................... Networks.py:102 (Keras binary_crossentropy)

Pinning status: the reference ships NO tests, golden vectors, weights or data, and its
arithmetic lives in un-pinned third-party Keras/TF1 which cannot be installed here
(SURVEY.md section 8c).  PARITY IS THEREFORE UNPINNED AGAINST A REAL KERAS RUN.  What pins
this file instead: `oracle/make_golden.py` imports the reference's *unmodified* Networks.py
and main.py on top of `oracle/keras_shim/` (documented Keras op semantics on torch fp64) and
stores the outputs/gradients/relation tensors the reference's own wiring produces in
`tests/golden/*.npz`; `tests/test_oracle.py` checks `forward_dense`, `forward_sparse`,
`build_relations_dense` and `edge_list` against those fixtures, plus dense == sparse in
fp64, gradcheck, and hand-computable cases.
"""
import math
import numpy as np
import torch

# ----------------------------------------------------------------------------------------
# weights: 11 Dense layers = 22 tensors, Keras layout kernel[in, out], bias[out]
#   rm   RelationalModel(2   -> 150,150,150,150)   Networks.py:46
#   om   ObjectModel    (2   -> 100,100)           Networks.py:47
#   rmp  RelationalModel(350 -> 150,150,100)       Networks.py:49
#   omp  ObjectModel    (300 -> 100,101)           Networks.py:50
# ----------------------------------------------------------------------------------------
LAYER_DIMS = {
    'rm': [(2, 150), (150, 150), (150, 150), (150, 150)],
    'om': [(2, 100), (100, 100)],
    'rmp': [(350, 150), (150, 150), (150, 100)],
    'omp': [(300, 100), (100, 101)],
}
NET_ORDER = ['rm', 'om', 'rmp', 'omp']
N_PARAMS = sum(i * o + o for net in NET_ORDER for (i, o) in LAYER_DIMS[net])   # 209501
N_STEPS = 5          # Networks.py:83
PROP_DIM = 100       # Networks.py:29,80
REL_THRESHOLD = 170  # main.py:71


def tensor_names():
    names = []
    for net in NET_ORDER:
        for li in range(len(LAYER_DIMS[net])):
            names += ['%s.w%d' % (net, li), '%s.b%d' % (net, li)]
    return names


def init_weights(seed=0, dtype=torch.float64, nonzero_bias=False):
    """Keras Dense defaults (Blocks.py:22-27): glorot_uniform kernel, zero bias.  Values are
    rounded to fp32 so the fp64 oracle and the fp32 kernels see identical parameters.
    `nonzero_bias=True` draws small random biases (tests only: a zero bias hides bias bugs)."""
    g = torch.Generator().manual_seed(seed)
    w = {}
    for net in NET_ORDER:
        for li, (i, o) in enumerate(LAYER_DIMS[net]):
            lim = math.sqrt(6.0 / (i + o))
            k = (torch.rand(i, o, generator=g, dtype=torch.float64) * 2 - 1) * lim
            if nonzero_bias:
                b = (torch.rand(o, generator=g, dtype=torch.float64) * 2 - 1) * 0.1
            else:
                b = torch.zeros(o, dtype=torch.float64)
            w['%s.w%d' % (net, li)] = k.float().to(dtype)
            w['%s.b%d' % (net, li)] = b.float().to(dtype)
    return w


_PROBE = None     # when a list: every relu records min|pre-activation| (distance to the kink)
_FORCED = None    # when a list of 0/1 tensors: relu layer k multiplies by _FORCED[k] instead (a fixed linear branch)
_FLIPS = None     # ... and records (units whose state differs from relu's own, max |pre-activation| over them)


def _relu(x):
    if _PROBE is not None and x.numel():
        _PROBE.append(float(x.detach().abs().min()))
    if _FORCED is not None:
        m = _FORCED.pop(0).to(x.dtype)
        assert m.shape == x.shape, (m.shape, x.shape)
        diff = (x.detach() > 0) != (m > 0)
        _FLIPS.append((int(diff.sum()), float(x.detach().abs()[diff].max()) if bool(diff.any()) else 0.0))
        return x * m
    return torch.relu(x)


def _mlp(w, net, x):
    """Blocks.py:20-28 / 60-68: Dense+relu for all but the last layer, last layer linear."""
    n = len(LAYER_DIMS[net])
    for li in range(n):
        x = x @ w['%s.w%d' % (net, li)] + w['%s.b%d' % (net, li)]
        if li < n - 1:
            x = _relu(x)
    return x


def min_relu_margin(w, obj, snd, rcv):
    """Smallest |pre-activation| over every relu unit of a forward pass.  Gradients of a relu network
    are discontinuous at 0: an fp32 evaluation whose rounding error exceeds this margin can switch a
    unit that the fp64 evaluation does not, and then differs by that unit's whole contribution."""
    global _PROBE
    _PROBE = []
    try:
        with torch.no_grad():
            forward_sparse(w, obj, snd, rcv)
        return min(_PROBE) if _PROBE else float('inf')
    finally:
        _PROBE = None


def kinkfree_weights(seed=0, dtype=torch.float64, scale=0.1):
    """Test weights that keep every relu unit far from its kink: glorot kernels scaled by `scale`,
    hidden biases +-1 (random sign, so both relu states occur), output biases small."""
    g = torch.Generator().manual_seed(seed + 1000)
    w = init_weights(seed, dtype=torch.float64, nonzero_bias=True)
    relu_biases = ['rm.b0', 'rm.b1', 'rm.b2', 'rm.b3', 'om.b0', 'om.b1', 'rmp.b0', 'rmp.b1', 'omp.b0']
    for k in list(w):
        if '.w' in k:
            w[k] = (w[k] * scale).float().double()
    for k in relu_biases:
        sign = (torch.rand(w[k].shape, generator=g) > 0.5).double() * 2 - 1
        w[k] = sign
    return {k: v.float().to(dtype) for k, v in w.items()}


# ----------------------------------------------------------------------------------------
# relation construction
# ----------------------------------------------------------------------------------------
def build_relations_dense(pos, thr=REL_THRESHOLD):
    """main.py:66-81 restated.  pos: (B, N, 2) float64 (frame-0 positions).
    Returns (sender_relations, receiver_relations) float64 (B, N, N(N-1))."""
    pos = np.asarray(pos, dtype=np.float64)
    B, N, _ = pos.shape
    R = N * (N - 1)
    rr = np.zeros((B, N, R), dtype=float)
    rs = np.zeros((B, N, R), dtype=float)
    cnt = 0
    for m in range(N):
        for j in range(N):
            if m != j:
                inzz = np.linalg.norm(pos[:, m, 0:2] - pos[:, j, 0:2], axis=1) < thr
                rr[inzz, j, cnt] = 1.0
                rs[inzz, m, cnt] = 1.0
                cnt += 1
    return rs, rr


def slot_of(m, j, N):
    """Closed form of the `cnt` counter in main.py:69-81 (sender m outer, receiver j inner)."""
    return m * (N - 1) + (j if j < m else j - 1)


def edge_list(pos_xy, node_off, thr=REL_THRESHOLD, fully_connected=False):
    """Sparse form of main.py:66-81 for a ragged batch.

    pos_xy: (sum N, 2) float64; node_off: (T+1,) int prefix sums of per-tower N.
    Returns int32 arrays (edge_off[T+1], snd[E], rcv[E], slot[E]); snd/rcv are GLOBAL node
    indices; edges of a tower appear in slot order.  The distance test is the exact numpy
    expression of the reference: sqrt(dx*dx + dy*dy) < thr in float64 (no fused multiply-add).
    """
    pos_xy = np.asarray(pos_xy, dtype=np.float64)
    node_off = np.asarray(node_off, dtype=np.int64)
    T = len(node_off) - 1
    edge_off = np.zeros(T + 1, dtype=np.int32)
    snd, rcv, slot = [], [], []
    for t in range(T):
        a, b = int(node_off[t]), int(node_off[t + 1])
        N = b - a
        p = pos_xy[a:b]
        if N > 1:
            d = np.linalg.norm(p[:, None, :] - p[None, :, :], axis=2)   # d[m, j]
            act = np.ones((N, N), bool) if fully_connected else (d < thr)
            np.fill_diagonal(act, False)
            m, j = np.nonzero(act)                                       # row-major == slot order
            snd.append(a + m)
            rcv.append(a + j)
            slot.append(m * (N - 1) + np.where(j < m, j, j - 1))
            edge_off[t + 1] = edge_off[t] + len(m)
        else:
            edge_off[t + 1] = edge_off[t]
    cat = lambda xs: (np.concatenate(xs) if xs else np.zeros(0)).astype(np.int32)
    return edge_off, cat(snd), cat(rcv), cat(slot)


def relations_from_edges(N, edge_off, snd, rcv, slot, node_off):
    """Expand an edge list of equal-size towers back to dense one-hots (for == tests)."""
    T = len(node_off) - 1
    R = N * (N - 1)
    rs = np.zeros((T, N, R)); rr = np.zeros((T, N, R))
    for t in range(T):
        for e in range(edge_off[t], edge_off[t + 1]):
            rs[t, snd[e] - node_off[t], slot[e]] = 1.0
            rr[t, rcv[e] - node_off[t], slot[e]] = 1.0
    return rs, rr


# ----------------------------------------------------------------------------------------
# forward: dense (line-by-line) and sparse (edge list) restatements
# ----------------------------------------------------------------------------------------
def forward_dense(w, objects, sender_relations, receiver_relations, propagation=None,
                  return_logits=False):
    """Networks.py:22-99 line by line.  objects (B,N,3); relations (B,N,R); all torch, one dtype.
    Returns per-block probabilities (B,N,1) (and logits if asked)."""
    B, N, _ = objects.shape
    if propagation is None:
        propagation = torch.zeros(B, N, PROP_DIM, dtype=objects.dtype)
    rs_t = sender_relations.permute(0, 2, 1)                 # :27
    rr_t = receiver_relations.permute(0, 2, 1)               # :28
    senders = rs_t @ objects                                 # :32
    receivers = rr_t @ objects                               # :33
    r_pos = receivers[:, :, 0:2]                             # :37,58
    s_pos = senders[:, :, 0:2]                               # :59
    diff_rs = r_pos - s_pos                                  # :62
    obj_in = torch.cat([objects[:, :, 1:2], objects[:, :, 2:3]], dim=-1)   # :65-66,71
    rel_enc = _relu(_mlp(w, 'rm', diff_rs))                  # :75
    obj_enc = _relu(_mlp(w, 'om', obj_in))                   # :76   (dropout :77-78 = identity)
    prop = propagation                                       # :79
    x = None
    for _ in range(N_STEPS):                                 # :83
        s_prop = rs_t @ prop                                 # :84
        r_prop = rr_t @ prop                                 # :85
        x = _mlp(w, 'rmp', torch.cat([rel_enc, s_prop, r_prop], dim=-1))    # :86-87
        eff = torch.tanh(receiver_relations @ x)             # :88
        x = _mlp(w, 'omp', torch.cat([obj_enc, eff, prop], dim=-1))         # :89-90
        prop = torch.tanh(x[:, :, 1:] + prop)                # :80,91
    logits = x[:, :, :1]                                     # :94
    probs = torch.sigmoid(logits)
    return (probs, logits) if return_logits else probs


def dropout_keep_mask(seed32, idx, rate):
    """numpy restatement of dropout_hash / dropout_apply (spwgnn_b200/csrc/spw_common.cuh): element idx is
    kept iff the 24-bit hash of (seed, idx) is >= float32(rate) * 2^24."""
    M = np.uint64(0xffffffff)
    h = (np.asarray(idx, dtype=np.uint64) * np.uint64(0x9E3779B9) + np.uint64(seed32)) & M
    h ^= h >> np.uint64(16); h = (h * np.uint64(0x85EBCA6B)) & M
    h ^= h >> np.uint64(13); h = (h * np.uint64(0xC2B2AE35)) & M
    h ^= h >> np.uint64(16)
    thresh = np.uint64(int(np.float32(rate) * np.float32(16777216.0)))
    return (h >> np.uint64(8)) >= thresh


def dropout_seeds(seed64):
    lo, hi = seed64 & 0xffffffff, (seed64 >> 32) & 0xffffffff
    return lo, (hi ^ 0x5bd1e995 ^ ((lo * 3) & 0xffffffff)) & 0xffffffff


def forward_sparse(w, obj, snd, rcv, return_logits=False, c_scale=None, q_scale=None):
    """Same math on an explicit edge list (F8: inactive slots contribute 0).
    obj: (sum N, 3) torch; snd/rcv: (E,) int64 torch GLOBAL node ids, slot order.
    c_scale (E,150) / q_scale (n,100): optional inverted-dropout factors (0 or 1/keep) applied to the two
    encodings like Networks.py:77-78 does in training.
    Returns per-node probabilities (sum N,) (and logits)."""
    n = obj.shape[0]
    snd = snd.long(); rcv = rcv.long()
    diff = obj[rcv, 0:2] - obj[snd, 0:2]
    c = _relu(_mlp(w, 'rm', diff))
    q = _relu(_mlp(w, 'om', obj[:, 1:3]))
    if c_scale is not None:
        c = c * c_scale
    if q_scale is not None:
        q = q * q_scale
    p = torch.zeros(n, PROP_DIM, dtype=obj.dtype)
    z = None
    for _ in range(N_STEPS):
        x = _mlp(w, 'rmp', torch.cat([c, p[snd], p[rcv]], dim=-1))
        agg = torch.zeros(n, PROP_DIM, dtype=obj.dtype).index_add(0, rcv, x)
        g = torch.tanh(agg)
        z = _mlp(w, 'omp', torch.cat([q, g, p], dim=-1))
        p = torch.tanh(z[:, 1:] + p)
    logits = z[:, 0]
    probs = torch.sigmoid(logits)
    return (probs, logits) if return_logits else probs


def bce_keras(probs, target):
    """Keras binary_crossentropy (Networks.py:102): clip to [1e-7, 1-1e-7], mean over all
    outputs (mean over the last axis, then over samples -- equal weights, so a flat mean)."""
    eps = 1e-7
    p = torch.clamp(probs, eps, 1.0 - eps)
    return -(target * torch.log(p) + (1.0 - target) * torch.log(1.0 - p)).mean()


def loss_and_grads_sparse(w, obj, snd, rcv, target, c_scale=None, q_scale=None):
    """fp64 reference gradients of the mean-BCE loss w.r.t. all 22 tensors (autograd)."""
    ws = {k: v.detach().clone().requires_grad_(True) for k, v in w.items()}
    probs, logits = forward_sparse(ws, obj, snd, rcv, return_logits=True, c_scale=c_scale, q_scale=q_scale)
    loss = bce_keras(probs, target)
    names = tensor_names()
    grads = torch.autograd.grad(loss, [ws[k] for k in names])
    return loss.detach(), probs.detach(), logits.detach(), dict(zip(names, grads))


def loss_and_grads_forced(w, obj, snd, rcv, target, masks):
    """loss_and_grads_sparse on the linear branch given by `masks` (21 tensors in the order the relu layers are evaluated:
    rm 0-3, om 0-1, then rmp 0, rmp 1, omp 0 per step): every relu is replaced by a multiplication with its mask.  Where a
    mask equals the unit's own state nothing changes; where it differs the pre-activation is within rounding of 0 if the
    masks come from an fp32 evaluation of the same network, which `flips` lets the caller verify.
    Returns (loss, logits, grads, flips) with flips = [(number of differing units, max |pre-activation| over them)] per layer."""
    global _FORCED, _FLIPS
    _FORCED, _FLIPS = [m for m in masks], []
    try:
        ws = {k: v.detach().clone().requires_grad_(True) for k, v in w.items()}
        probs, logits = forward_sparse(ws, obj, snd, rcv, return_logits=True)
        assert not _FORCED, 'unused masks'
        loss = bce_keras(probs, target)
        names = tensor_names()
        grads = torch.autograd.grad(loss, [ws[k] for k in names])
        return loss.detach(), logits.detach(), dict(zip(names, grads)), _FLIPS
    finally:
        _FORCED, _FLIPS = None, None


def normalise_objects(boxes_raw, thr=REL_THRESHOLD):
    """main.py:91: float64 divide; the fp32 cast happens at the Keras feed."""
    return np.asarray(boxes_raw, dtype=np.float64) / thr
