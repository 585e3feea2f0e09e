"""Where the time of a batch-1 predict goes (C1 shape: one 7-block tower through the public API).  Usage: python tools/latency_probe.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from spwgnn_b200.Networks import PropagationNetwork
from spwgnn_b200.graph import TowerBatch
from spwgnn_b200 import synth


def wall(f, n=200):
    for _ in range(10):
        f()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6


def dev(f, n=200):
    for _ in range(10):
        f()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(n):
        f()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


def main():
    torch.cuda.set_device(0)
    model = PropagationNetwork(device='cuda:0', seed=0).getModel(7)
    eng = model.engine
    for nb in (1, 32):
        towers = synth.make_towers('tower', nb, 3, n=7)
        batch = TowerBatch.from_towers(towers, inference_glue=True, device=eng.device)
        print('batch of %d towers: %d blocks, %d relations' % (nb, batch.n_nodes, batch.n_edges))
        print('  predict_towers (public call, host-synchronous)      %8.1f us' % wall(lambda: model.predict_towers(towers)))
        print('  TowerBatch.from_towers (host + edge kernels + 1 sync) %8.1f us' % wall(lambda: TowerBatch.from_towers(towers, inference_glue=True, device=eng.device)))
        for gm, name in ((8192, 'graph replay'), (0, 'direct launches')):
            eng.graph_max_edges = gm
            print('  engine.forward, %-16s wall %8.1f us   device %8.1f us' % (name, wall(lambda: eng.forward(batch)), dev(lambda: eng.forward(batch))))
        eng.graph_max_edges = 8192
        key = [k for k in eng._graphs if k[1] == batch.n_nodes and k[2] == batch.n_edges][0]
        g = eng._graphs[key][0]
        print('  CUDAGraph.replay only                                 wall %8.1f us   device %8.1f us' % (wall(g.replay), dev(g.replay)))


if __name__ == '__main__':
    main()
