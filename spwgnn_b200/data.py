"""Trajectory files -> training inputs.

Vectorised restatement of what /root/reference/src/main.py does with Python loops before the hot path:
  * the JSON trajectory format written by the simulators (TowerCreator.py:94-104,447-453;
    JengaBuilder.py:366-371): data[t][o][f] = [x, y] or [x, y, width];
  * frame padding: shorter trajectories repeat their last frame (main.py:49-63);
  * the stability label: block o of trajectory t is stable iff the summed frame-to-frame displacement
    over the whole trajectory is < 0.5 (main.py:8-23).
"""
import json

import numpy as np


def load_trajectories(path):
    with open(path) as f:
        data = json.load(f)
    return [d for d in data if len(d) != 0]          # main.py:44 drops empty trajectories


def pad_frames(data, n_objects, object_dim):
    """main.py:47-63 without the triple loop: boxes[t, f, o, :] float64, last frame repeated."""
    n_traj = len(data)
    n_frame = max(len(t[0]) for t in data)
    boxes = np.zeros((n_traj, n_frame, n_objects, object_dim))
    for t, traj in enumerate(data):
        for o in range(n_objects):
            fr = np.asarray(traj[o], dtype=np.float64)[:, :object_dim]      # (frames_o, D)
            k = min(len(fr), n_frame)
            boxes[t, :k, o, :] = fr[:k]
            boxes[t, k:, o, :] = fr[-1]
    return boxes


def calculate_stability(boxes, threshold=0.5):
    """main.py:8-23: y[t, o, 0] = 1 iff sum_f ||pos_f - pos_{f+1}|| < 0.5 (float64, same summation order)."""
    d = boxes[:, :-1, :, 0:2] - boxes[:, 1:, :, 0:2]                 # (T, F-1, N, 2)
    step = np.sqrt((d * d).sum(axis=3))                              # np.linalg.norm of each 2-vector
    total = np.zeros(step.shape[0:1] + step.shape[2:])
    for f in range(step.shape[1]):                                   # sequential sum over frames == the loop
        total += step[:, f]
    return (total < threshold).astype(np.float64)[:, :, None]


def training_arrays(path, n, jenga=True):
    """What main.train_gnn builds before `.fit`: (raw frame-0 boxes (T,N,D), labels (T,N,1)).
    n follows main.py:27-34: n_objects = n-1 for the jenga data, n+1 for the construction data."""
    n_objects, object_dim = (n - 1, 3) if jenga else (n + 1, 2)
    data = load_trajectories(path)
    boxes = pad_frames(data, n_objects, object_dim)
    return boxes[:, 0], calculate_stability(boxes)
