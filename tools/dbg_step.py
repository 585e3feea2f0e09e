"""Development aid: repeated small training steps through tools/_build/libspwgnn_dbg.so (nvcc ... -DSPW_WAIT_DEBUG: short
bounded waits that print which barrier timed out); every step must reproduce the first one bit for bit."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spwgnn_b200._lib as _lib
from spwgnn_b200._capi import CApi
lib = os.environ.get('SPW_DBG_LIB', os.path.join(os.path.dirname(os.path.abspath(__file__)), '_build', 'libspwgnn_dbg.so'))
_lib._api = CApi(lib)
from spwgnn_b200.engine import Engine
from spwgnn_b200.graph import TowerBatch
from spwgnn_b200 import synth

nt = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
towers = synth.make_towers('jenga', nt, 3, n=10)
eng = Engine('cuda:0', seed=1)
batch = TowerBatch.from_towers(towers, device='cuda:0', fully_connected=True)
print('nodes', batch.n_nodes, 'edges', batch.n_edges, flush=True)
tgt = torch.zeros(batch.n_nodes, device='cuda')
ref = None
bad = 0
import time
t_all = time.time()
for it in range(steps):
    t0 = time.time()
    stats = eng.loss_and_grads(batch, tgt)
    torch.cuda.synchronize()
    if time.time() - t0 > 0.5:
        print('step', it, 'took %.1f s' % (time.time() - t0), flush=True)
    g = eng.grads.flat.clone()
    if ref is None:
        ref = g
        print('loss', float(stats[0]) / batch.n_nodes, 'grad norm', float(g.norm()), 'finite', bool(torch.isfinite(g).all()), flush=True)
    elif not torch.equal(g, ref):
        bad += 1
        print('step', it, 'differs: max abs', float((g - ref).abs().max()), 'finite', bool(torch.isfinite(g).all()), flush=True)
print('steps', steps, 'mismatching', bad, 'seconds %.1f' % (time.time() - t_all))
