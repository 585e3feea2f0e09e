"""Packed, device-resident batches of tower graphs.

Replaces the O(B.N^3) float64 one-hot relation tensors of the reference
(/root/reference/src/main.py:66-81, TowerCreator.py:415-428, JengaBuilder.py:313-326) with a ragged
edge list built on the GPU by libspwgnn (spw_edges_count / spw_edges_fill).  Towers of different
sizes share one batch (the reference needs one Keras model per n_objects, Networks.py:17-18).
"""
import numpy as np
import torch

from ._capi import CApi
from ._lib import lib, require_cuda, SpwError

MAX_NODES = 64
REL_THRESHOLD = 170.0     # main.py:71


def _stream_ptr(device):
    return torch.cuda.current_stream(device).cuda_stream


def _to_dev(x, dtype, device):
    t = torch.as_tensor(x)
    if t.dtype != dtype or t.device != device:
        t = t.to(device=device, dtype=dtype, non_blocking=True)
    return t.contiguous()


class TowerBatch:
    """A batch of towers packed node-major.  All tensors live on `device`.

    node_off  int32 [T+1]   tower t owns nodes [node_off[t], node_off[t+1])
    obj       fp32  [n, 3]  [x, y, width] / 170          (main.py:91)
    snd/rcv/slot int32 [E]  active relations in the reference's slot order (optional)
    in_*, out_*             CSR views consumed by the kernels (include/spwgnn.h: SpwGraph)
    """

    def __init__(self):
        self.slot_list = None

    # -- fast path: raw poses -> edges on the GPU -------------------------------------------------
    @staticmethod
    def from_poses(obj, node_off, edge_pos=None, thr=REL_THRESHOLD, fully_connected=False, device=None,
                   want_slot_list=False, max_nodes=None):
        """obj: (n,3) normalised features.  edge_pos: (n,2) float64 positions the distance test runs on
        (training: RAW pixels, main.py:78; inference glue: normalised positions, JengaBuilder.py:309-323);
        may be None when fully_connected.  node_off: (T+1,) host array of prefix sums."""
        require_cuda()
        api = lib()
        device = torch.device(device if device is not None else 'cuda')
        b = TowerBatch()
        node_off_h = np.ascontiguousarray(np.asarray(node_off), dtype=np.int64)
        T = len(node_off_h) - 1
        n = int(node_off_h[-1]) if T >= 0 else 0
        sizes = np.diff(node_off_h) if T > 0 else np.zeros(0, np.int64)
        if max_nodes is None:
            max_nodes = int(sizes.max()) if T > 0 else 0
        if max_nodes > MAX_NODES:
            raise SpwError('tower with %d blocks; this build handles at most %d per tower' % (max_nodes, MAX_NODES))
        if n >= 2 ** 31 - 1:
            raise SpwError('batch too large for int32 node ids; split it')
        b.device, b.n_towers, b.n_nodes, b.max_nodes = device, T, n, max_nodes
        b.node_off_host = node_off_h
        b.node_off = _to_dev(node_off_h.astype(np.int32), torch.int32, device)
        b.obj = _to_dev(obj, torch.float32, device).reshape(n, 3)
        if fully_connected:
            pos = torch.zeros(max(n, 1), 2, dtype=torch.float64, device=device) if edge_pos is None \
                else _to_dev(edge_pos, torch.float64, device)
        else:
            if edge_pos is None:
                raise SpwError('edge_pos is required unless fully_connected')
            pos = _to_dev(edge_pos, torch.float64, device)
        st = _stream_ptr(device)
        deg_out = torch.empty(max(n, 1), dtype=torch.int32, device=device)
        deg_in = torch.empty(max(n, 1), dtype=torch.int32, device=device)
        edge_off = torch.empty(T + 1, dtype=torch.int32, device=device)
        api.check(api.dll.spw_edges_count(pos.data_ptr(), b.node_off.data_ptr(), T, n, max_nodes, float(thr),
                                          int(bool(fully_connected)), deg_out.data_ptr(), deg_in.data_ptr(),
                                          edge_off.data_ptr(), st))
        if fully_connected:
            E = int((sizes * (sizes - 1)).sum())
        else:
            E = int(edge_off[T].item())          # one 4-byte read-back: the only host sync of the build
        if E >= 2 ** 31 - 1:
            raise SpwError('batch has too many edges for int32 ids; split it')
        b.n_edges, b.edge_off = E, edge_off
        mk = lambda k: torch.empty(max(k, 1), dtype=torch.int32, device=device)
        b.in_off, b.out_off = mk(n + 1), mk(n + 1)
        b.in_snd, b.in_rcv, b.out_pos = mk(E), mk(E), mk(E)
        if want_slot_list:
            b.slot_list = (mk(E), mk(E), mk(E))
        sl = b.slot_list or (None, None, None)
        ptr = lambda t: 0 if t is None else t.data_ptr()
        api.check(api.dll.spw_edges_fill(pos.data_ptr(), b.node_off.data_ptr(), T, n, max_nodes, float(thr),
                                         int(bool(fully_connected)), edge_off.data_ptr(), ptr(sl[0]), ptr(sl[1]),
                                         ptr(sl[2]), b.in_off.data_ptr(), b.in_snd.data_ptr(), b.in_rcv.data_ptr(),
                                         b.out_off.data_ptr(), b.out_pos.data_ptr(), st))
        b._keep = (pos, deg_out, deg_in)
        b._finish()
        return b

    # -- device-side synthetic layouts ------------------------------
    @staticmethod
    def sample_tower(n_towers, n_lo, n_hi, seed, **kw):
        """TowerCreator layouts (TowerCreator.py:106-187, 265-271) generated on the GPU: n_lo..n_hi blocks per tower
        (>= 2: stacked blocks + the dropped one, which is object 0).  Bit-identical to synth.g_tower_ctr."""
        return TowerBatch.sample_jenga(n_towers, n_lo, n_hi, seed, _kind='tower', **kw)

    @staticmethod
    def sample_jenga(n_towers, n_lo, n_hi, seed, device=None, fully_connected=False, inference_glue=False,
                     thr=REL_THRESHOLD, want_raw=False, want_slot_list=False, _kind='jenga'):
        """Jenga-style layouts (JengaBuilder.create_world, JengaBuilder.py:137-192) generated ON the GPU and packed:
        tower t gets n_lo..n_hi blocks; poses never exist on the host.  Bit-identical to synth.g_jenga_ctr.
        Returns the TowerBatch (its .raw holds the (n, 3) float64 pixel poses when want_raw)."""
        require_cuda()
        api = lib()
        device = torch.device(device if device is not None else 'cuda')
        st = _stream_ptr(device)
        node_off = torch.empty(n_towers + 1, dtype=torch.int32, device=device)
        api.check(api.dll.spw_sample_sizes(int(seed), n_towers, n_lo, n_hi, node_off.data_ptr(), st))
        node_off_h = node_off.cpu().numpy().astype(np.int64)       # (T + 1) ints back: sizes drive the allocation
        n = int(node_off_h[-1]) if n_towers > 0 else 0
        raw = torch.empty(max(n, 1), 3, dtype=torch.float64, device=device) if want_raw else None
        obj = torch.empty(max(n, 1), 3, dtype=torch.float32, device=device)
        pos = torch.empty(max(n, 1), 2, dtype=torch.float64, device=device)
        if _kind == 'tower' and n_lo < 2:
            raise SpwError('TowerCreator layouts need at least 2 blocks per tower')
        sampler = api.dll.spw_sample_tower if _kind == 'tower' else api.dll.spw_sample_jenga
        api.check(sampler(int(seed), n_towers, node_off.data_ptr(), 0 if raw is None else raw.data_ptr(),
                          obj.data_ptr(), pos.data_ptr(), int(bool(inference_glue)), st))
        b = TowerBatch.from_poses(obj[:n], node_off_h, pos[:n], thr=thr, fully_connected=fully_connected, device=device,
                                  want_slot_list=want_slot_list, max_nodes=n_hi)
        b.raw = None if raw is None else raw[:n]
        return b

    # -- compat path: the reference's dense one-hot dict ------------------------------------------
    @staticmethod
    def from_dense_relations(objects, sender_relations, receiver_relations, device=None):
        """objects (B,N,3); sender/receiver_relations (B,N,R) one-hot columns (main.py:92 feed).
        Non-zero columns become edges; each must hold exactly one sender and one receiver."""
        require_cuda()
        device = torch.device(device if device is not None else 'cuda')
        objects = np.asarray(objects)
        B, N, D = objects.shape
        if D != 3:
            raise SpwError('object_dim=%d: only the object_dim=3 path of the reference is well defined '
                           '(Networks.py:70-73 feeds a 1-wide tensor to a 2-wide encoder otherwise)' % D)
        if N > MAX_NODES:
            raise SpwError('tower with %d blocks; limit %d' % (N, MAX_NODES))
        rs = torch.as_tensor(np.asarray(sender_relations)).to(device)
        rr = torch.as_tensor(np.asarray(receiver_relations)).to(device)
        cs, cr = (rs != 0).sum(1), (rr != 0).sum(1)           # (B, R)
        active = (cs > 0) | (cr > 0)
        if bool(((cs != 1) | (cr != 1))[active].any()):
            raise SpwError('every active relation column needs exactly one sender and one receiver')
        bi, ri = torch.nonzero(active, as_tuple=True)          # (b, slot) ascending == slot order
        snd = (rs != 0).to(torch.int8).argmax(1)[bi, ri] + bi * N
        rcv = (rr != 0).to(torch.int8).argmax(1)[bi, ri] + bi * N
        b = TowerBatch()
        b.device, b.n_towers, b.n_nodes, b.max_nodes = device, B, B * N, N
        b.node_off_host = np.arange(B + 1, dtype=np.int64) * N
        b.node_off = _to_dev(b.node_off_host.astype(np.int32), torch.int32, device)
        b.obj = _to_dev(objects.reshape(B * N, 3), torch.float32, device)
        b._set_edges(snd, rcv, slot=ri)
        return b

    def _set_edges(self, snd, rcv, slot=None):
        """Arbitrary edge list (global node ids, any order) -> CSR views, with torch ops (compat path)."""
        n, device = self.n_nodes, self.device
        E = int(snd.numel())
        snd, rcv = snd.long(), rcv.long()
        # "slot order" of the kernels: sender-major, stable
        o_s = torch.sort(snd, stable=True).indices
        snd, rcv = snd[o_s], rcv[o_s]
        if slot is not None:
            slot = slot[o_s]
        o_r = torch.sort(rcv, stable=True).indices             # receiver-major, ascending slot-order id inside
        deg_in = torch.bincount(rcv, minlength=n)
        deg_out = torch.bincount(snd, minlength=n)
        if E and int(deg_in.max()) >= 128:
            raise SpwError('a block with >= 128 incoming relations is not supported')
        z = torch.zeros(1, dtype=torch.long, device=device)
        self.n_edges = E
        i32 = lambda t: t.to(torch.int32).contiguous() if t.numel() else torch.zeros(1, dtype=torch.int32, device=device)
        self.in_off = i32(torch.cat([z, deg_in.cumsum(0)]))
        self.out_off = i32(torch.cat([z, deg_out.cumsum(0)]))
        self.in_snd, self.in_rcv = i32(snd[o_r]), i32(rcv[o_r])
        out_pos = torch.empty(E, dtype=torch.long, device=device)
        out_pos[o_r] = torch.arange(E, device=device)
        self.out_pos = i32(out_pos)
        self.edge_off = None
        self.slot_list = (i32(snd), i32(rcv), i32(slot) if slot is not None else None)
        self._finish()

    def _finish(self):
        self.c_graph = CApi.graph(self.n_towers, self.n_nodes, self.n_edges, self.node_off.data_ptr(),
                                  self.in_off.data_ptr(), self.in_snd.data_ptr(), self.in_rcv.data_ptr(),
                                  self.out_off.data_ptr(), self.out_pos.data_ptr())

    # -- convenience ----------------------------------------------------------------------------
    @staticmethod
    def from_towers(towers, thr=REL_THRESHOLD, fully_connected=False, inference_glue=False, **kw):
        """towers: list of (N_t, 3) arrays of RAW [x, y, width] in pixels.
        inference_glue=True reproduces the reference's predict-time behaviour: positions are divided
        by 170 BEFORE the distance test against 170 (JengaBuilder.py:309-323), i.e. fully connected."""
        sizes = [len(t) for t in towers]
        node_off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        raw = np.concatenate([np.asarray(t, dtype=np.float64).reshape(-1, 3) for t in towers]) if towers \
            else np.zeros((0, 3))
        obj = raw / REL_THRESHOLD
        edge_pos = obj[:, 0:2] if inference_glue else raw[:, 0:2]
        return TowerBatch.from_poses(obj, node_off, np.ascontiguousarray(edge_pos), thr, fully_connected, **kw)
