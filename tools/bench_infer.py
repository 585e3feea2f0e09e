#!/usr/bin/env python
"""Inference sweep in the shape of BASELINE config 5 (towers of 8-64 blocks, contact edges): towers are generated on
the device (spw_sample_jenga), packed (spw_edges_*) and scored (spw_forward, inference mode) chunk by chunk.  Not the
driver's bench line (bench.py measures the training step of config 2); prints one JSON line.
Usage: python tools/bench_infer.py [--towers 1000000] [--chunk 20000]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spwgnn_b200.engine import Engine          # noqa: E402
from spwgnn_b200.graph import TowerBatch       # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--towers', type=int, default=1000000)
    ap.add_argument('--chunk', type=int, default=20000)
    ap.add_argument('--lo', type=int, default=8)
    ap.add_argument('--hi', type=int, default=64)
    args = ap.parse_args()
    eng = Engine('cuda:0', seed=0)
    nchunks = (args.towers + args.chunk - 1) // args.chunk

    def run(i):
        b = TowerBatch.sample_jenga(min(args.chunk, args.towers - i * args.chunk), args.lo, args.hi, 1234 + i)
        logits, probs = eng.forward(b, training=False)
        return b.n_nodes, b.n_edges, probs

    for i in range(2):
        run(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    nodes = edges = 0
    acc = torch.zeros((), dtype=torch.float64, device='cuda')
    for i in range(nchunks):
        n, e, probs = run(i)
        nodes += n
        edges += e
        acc += probs[:n].double().sum()
    e1.record()
    torch.cuda.synchronize()
    s = e0.elapsed_time(e1) * 1e-3
    print(json.dumps({'workload': 'C5 shape: inference over %d device-generated towers of %d-%d blocks, contact edges, chunks of %d'
                      % (args.towers, args.lo, args.hi, args.chunk), 'towers_per_sec': args.towers / s, 'nodes_per_sec': nodes / s,
                      'edges_per_sec': edges / s, 'seconds': s, 'nodes': nodes, 'edges': edges,
                      'mean_probability': float(acc) / max(nodes, 1)}))


if __name__ == '__main__':
    main()
