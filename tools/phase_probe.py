"""Per-phase cycle counts of the tile loops (development aid): expects tools/_build/libspwgnn_phase.so
(nvcc ... -DSPW_PHASE_TIMING) and runs one training step of the C2 workload through it; the kernels print their
clock64() phase sums for CTA 0 (threads 0 and 255)."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spwgnn_b200._lib as _lib
from spwgnn_b200._capi import CApi
_lib._api = CApi(os.path.join(os.path.dirname(os.path.abspath(__file__)), '_build', 'libspwgnn_phase.so'))
from spwgnn_b200.engine import Engine
from spwgnn_b200.graph import TowerBatch
from spwgnn_b200 import synth

towers = synth.make_towers('jenga', 4096, 3, n=10)
eng = Engine('cuda:0', seed=1)
batch = TowerBatch.from_towers(towers, device='cuda:0', fully_connected=True)
print('nodes', batch.n_nodes, 'edges', batch.n_edges, flush=True)
tgt = torch.zeros(batch.n_nodes, device='cuda')
eng.loss_and_grads(batch, tgt)
torch.cuda.synchronize()
