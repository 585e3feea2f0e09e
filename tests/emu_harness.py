"""TEST INFRASTRUCTURE: drive the C ABI of the kernel-logic emulator build (tools/cuemu) with numpy
arrays.  The emulator compiles the very same .cu sources for the host (one pthread per CUDA
thread) so that indexing / barrier / carve-up bugs are caught on a machine without a GPU.  It is
never used by the product: spwgnn_b200 only binds the nvcc-built libspwgnn.so."""
import ctypes
import os
import subprocess

import numpy as np

from spwgnn_b200._capi import CApi, PARAM_SPECS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_DIR = os.path.join(ROOT, 'tools', 'cuemu')
EMU_LIB = os.path.join(EMU_DIR, '_build', 'libspwgnn_emu.so')
SOURCES = [os.path.join(ROOT, 'spwgnn_b200', 'csrc', f) for f in
           ('spwgnn.cu', 'spw_common.cuh', 'spw_edges.cuh', 'spw_kernels.cuh')] + \
          [os.path.join(EMU_DIR, f) for f in ('cuda_emu.h', 'cuda_emu.cpp')] + \
          [os.path.join(ROOT, 'include', 'spwgnn.h')]


def build_emu(force=False):
    os.makedirs(os.path.dirname(EMU_LIB), exist_ok=True)
    if not force and os.path.exists(EMU_LIB) and all(os.path.getmtime(EMU_LIB) >= os.path.getmtime(s) for s in SOURCES):
        return EMU_LIB
    cmd = ['g++', '-std=c++17', '-O2', '-g', '-march=native', '-fPIC', '-shared', '-DSPW_EMU', '-I' + EMU_DIR,
           '-x', 'c++', SOURCES[0], os.path.join(EMU_DIR, 'cuda_emu.cpp'), '-o', EMU_LIB, '-lpthread']
    subprocess.run(cmd, check=True, cwd=ROOT)
    return EMU_LIB


def _p(a):
    return None if a is None else a.ctypes.data


class EmuGraph:
    pass


class Emu:
    def __init__(self):
        self.api = CApi(build_emu())

    def edges(self, pos_xy, node_off, thr=170.0, fully_connected=False):
        api = self.api
        pos_xy = np.ascontiguousarray(pos_xy, dtype=np.float64)
        node_off = np.ascontiguousarray(node_off, dtype=np.int32)
        T = len(node_off) - 1
        n = int(node_off[-1])
        max_n = int(np.diff(node_off).max()) if T else 0
        deg_out = np.zeros(max(n, 1), np.int32); deg_in = np.zeros(max(n, 1), np.int32)
        edge_off = np.full(T + 1, -7, np.int32)
        api.check(api.dll.spw_edges_count(_p(pos_xy), _p(node_off), T, n, max_n, float(thr), int(fully_connected),
                                          _p(deg_out), _p(deg_in), _p(edge_off), None))
        E = int(edge_off[T])
        g = EmuGraph()
        g.T, g.n, g.E = T, n, E
        g.node_off, g.edge_off, g.deg_out, g.deg_in = node_off, edge_off, deg_out[:n], deg_in[:n]
        mk = lambda k: np.full(max(k, 1), -9, np.int32)
        g.snd, g.rcv, g.slot = mk(E), mk(E), mk(E)
        g.in_off, g.out_off = mk(n + 1), mk(n + 1)
        g.in_snd, g.in_rcv, g.out_pos = mk(E), mk(E), mk(E)
        api.check(api.dll.spw_edges_fill(_p(pos_xy), _p(node_off), T, n, max_n, float(thr), int(fully_connected),
                                         _p(edge_off), _p(g.snd), _p(g.rcv), _p(g.slot), _p(g.in_off), _p(g.in_snd),
                                         _p(g.in_rcv), _p(g.out_off), _p(g.out_pos), None))
        for k in ('snd', 'rcv', 'slot', 'in_snd', 'in_rcv', 'out_pos'):
            setattr(g, k, getattr(g, k)[:E])
        g.c = api.graph(T, n, E, _p(node_off), _p(g.in_off), _p(g.in_snd), _p(g.in_rcv), _p(g.out_off), _p(g.out_pos))
        return g

    def sample_jenga(self, seed, n_towers, n_lo, n_hi, inference_glue=False, kind='jenga'):
        """device-side layout sampler (k_sample_sizes / k_sample_jenga) under the emulator"""
        api = self.api
        node_off = np.full(n_towers + 1, -5, np.int32)
        api.check(api.dll.spw_sample_sizes(int(seed), n_towers, n_lo, n_hi, _p(node_off), None))
        n = int(node_off[-1])
        raw = np.full((max(n, 1), 3), np.nan, np.float64)
        obj = np.full((max(n, 1), 3), np.nan, np.float32)
        pos = np.full((max(n, 1), 2), np.nan, np.float64)
        fn = api.dll.spw_sample_tower if kind == 'tower' else api.dll.spw_sample_jenga
        api.check(fn(int(seed), n_towers, _p(node_off), _p(raw), _p(obj), _p(pos), int(inference_glue), None))
        return node_off, raw[:n], obj[:n], pos[:n]

    @staticmethod
    def pack_params(w):
        """dict name -> array  =>  (list of aligned fp32 arrays in struct order, SpwParams)"""
        arrs = []
        for name, shape in PARAM_SPECS:
            a = np.zeros(int(np.prod(shape)) + 8, np.float32)
            off = (-a.ctypes.data // 4) % 4          # 16-byte align
            v = a[off:off + int(np.prod(shape))].reshape(shape)
            if w is not None:
                v[...] = np.asarray(w[name], dtype=np.float32)
            arrs.append((a, v))
        return [v for _, v in arrs], CApi.params([v.ctypes.data for _, v in arrs]), arrs

    def forward(self, w, g, obj, training=False, dropout_rate=0.0, dropout_seed=0):
        api = self.api
        wl, wp, keep = self.pack_params(w)
        obj = np.ascontiguousarray(obj, dtype=np.float32)
        nbytes = api.dll.spw_workspace_bytes(g.n, g.E, int(training))
        ws = np.zeros(nbytes // 4 + 8, np.float32)
        ws[...] = np.nan                                   # poison: reads of unwritten workspace show up
        off = (-ws.ctypes.data // 4) % 4
        wsv = ws[off:off + nbytes // 4]
        logits = np.full(max(g.n, 1), np.nan, np.float32); probs = np.full(max(g.n, 1), np.nan, np.float32)
        api.check(api.dll.spw_forward(ctypes.byref(wp), ctypes.byref(g.c), _p(obj), _p(logits), _p(probs),
                                      wsv.ctypes.data, nbytes, int(training), float(dropout_rate), int(dropout_seed), None))
        self._rate = float(dropout_rate)
        self._state = (wl, wp, keep, obj, ws, wsv, nbytes)
        return logits[:g.n], probs[:g.n]

    def bce_grad(self, logits, target, count):
        api = self.api
        n = len(logits)
        logits = np.ascontiguousarray(logits, np.float32); target = np.ascontiguousarray(target, np.float32)
        dl = np.zeros(max(n, 1), np.float32); stats = np.zeros(2, np.float64)
        api.check(api.dll.spw_bce_grad(_p(logits), _p(target), n, float(count), _p(dl), _p(stats), None))
        return dl[:n], stats

    def backward(self, g, dlogits):
        api = self.api
        wl, wp, keep, obj, ws, wsv, nbytes = self._state
        gl, gp, gkeep = self.pack_params(None)
        for v in gl:
            v[...] = np.nan
        dlogits = np.ascontiguousarray(dlogits, np.float32)
        api.check(api.dll.spw_backward(ctypes.byref(wp), ctypes.byref(g.c), _p(obj), _p(dlogits), wsv.ctypes.data,
                                       nbytes, ctypes.byref(gp), self._rate, None))
        return {name: v.copy() for (name, _), v in zip(PARAM_SPECS, gl)}
