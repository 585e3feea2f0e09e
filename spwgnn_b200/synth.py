"""Seeded synthetic tower layouts (host side) -- restatements of the reference's layout samplers
without the physics engine.  Raw positions stay float64 pixels.

  g_tower(n)   TowerCreator.create_world / create_pos_for_boxes / drop_object
               (/root/reference/src/TowerCreator.py:106-187, 265-271): 150x80 blocks, one dropped
               block on top; object 0 is the dropped block (TowerCreator.py:451).
  g_jenga(n)   JengaBuilder.create_world (/root/reference/src/JengaBuilder.py:137-192): widths
               randint(50,300), gaps randint(0,50), first layer spans x in [400,1100].
  g_jenga18()  stylised 18 layers x 3 blocks with the same width/gap distributions (BASELINE config 3;
               the reference sampler collapses to one block per layer).
  g_uniform(lo,hi)  N ~ U{lo..hi}, poses from g_jenga(N).
Each returns an (N, 3) float64 array of [x, y, width] in pixels.
"""
import math
import random

import numpy as np

RECT_W, RECT_H, BOTTOM = 150, 80, 70
WINDOW_W = 1500


def g_tower(n, rng: random.Random):
    orientation = rng.random() > 0.5
    layers = [rng.randint(1, max(1, n // 2))]
    left = n - layers[0]
    while left > 0:
        prev = layers[-1]
        if prev == 1:
            r = 1
        else:
            r = rng.randint(1, min(prev, left))
            i = 0
            while r == 1 and left != 1 and i < 3:
                r = rng.randint(1, min(prev, left))
                i += 1
        layers.append(r)
        left -= r
    boxes = []           # list of layers, each a list of x positions

    def edges(layer):
        xs = boxes[layer]
        return max(xs) + RECT_W // 2, min(xs) - RECT_W // 2

    def middle(layer_num):
        if layer_num == 0:
            return WINDOW_W / 2
        r, l = edges(layer_num - 1)
        return int((l + r) / 2)

    def make_pos(layer_num, layer_size, idx, mid, to_drop=False):
        var = int(RECT_W * (0.5 if to_drop else 0.3))
        mean_range = RECT_W + 2 * var
        mean = mid + ((-1) ** idx) * math.floor((idx + 1) / 2) * mean_range
        if layer_num > 0 and layer_size == 1:
            r, l = edges(layer_num - 1)
            lo, hi = int(l) + int(RECT_W * 0.2), int(r) - int(RECT_W * 0.2)
            x = rng.randint(min(lo, hi), max(lo, hi))
        else:
            lo, hi = int(mean - (1 - orientation) * var), int(mean + orientation * var)
            x = rng.randint(lo, hi) + (int(mean_range / 2) if layer_size % 2 == 0 else 0)
        return x, BOTTOM + RECT_H / 2 + RECT_H * layer_num

    def stable(layer_num, x):
        """TowerCreator.is_box_stable (TowerCreator.py:250-263): with the candidate in place, the sum of int(x / count) over all
        boxes must lie between the edges of the ground layer."""
        if layer_num == 0:
            return True
        xs = [v for layer in boxes for v in layer] + [x]
        com = sum(int(v / len(xs)) for v in xs)
        r, l = edges(0)
        return l <= com <= r

    def place(layer_num, layer_size, idx, mid, to_drop=False):
        nonlocal orientation
        x, y = make_pos(layer_num, layer_size, idx, mid, to_drop)
        if not stable(layer_num, x):          # one re-draw with the build direction flipped (TowerCreator.py:201-207)
            orientation = not orientation
            x, y = make_pos(layer_num, layer_size, idx, mid, to_drop)
        return x, y

    out = []
    for ln, size in enumerate(layers):
        boxes.append([])
        mid = middle(ln)
        for i in range(size):
            x, y = place(ln, size, i, mid)
            boxes[ln].append(x)
            out.append([float(x), float(y), float(RECT_W)])
    ln = len(boxes)
    boxes.append([])                          # drop_object appends the new (still empty) layer before it places the block
    x, y = place(ln, 1, 0, middle(ln), to_drop=True)
    return np.array([[float(x), float(y), float(RECT_W)]] + out, dtype=np.float64)   # dropped block first


def g_jenga(n, rng: random.Random):
    wmin, wrange, wavg, gap = 50, 250, 150, 50
    left_most, right_most = 400, WINDOW_W - 400
    out, layers = [], []
    left, layer_num = n, -1
    while left > 0:
        layer_num += 1
        layers.append([])
        if layer_num == 0:
            r_edge, l_edge = right_most, left_most
        else:
            xs = layers[layer_num - 1]
            r_edge, l_edge = max(xs), min(xs)
        y = BOTTOM + RECT_H / 2 + RECT_H * layer_num
        if r_edge == l_edge:
            x = rng.randint(int(l_edge - wmin / 2), int(l_edge + wmin / 2))
            w = rng.randint(wmin, wmin + wrange)
            layers[layer_num].append(x); out.append([float(x), float(BOTTOM + int(RECT_H / 2) + RECT_H * layer_num), float(w)])
            left -= 1
            continue
        l_edge -= (layer_num > 0) * int(wavg / 2)
        w = rng.randint(wmin, wmin + wrange)
        l_edge += w
        while l_edge - w / 2 < r_edge and left > 0:
            x = l_edge - w / 2
            layers[layer_num].append(x); out.append([float(x), float(y), float(w)])
            left -= 1
            l_edge += rng.randint(0, gap)
            w = rng.randint(wmin, wmin + wrange)
            l_edge += w
        if not layers[layer_num]:           # degenerate draw: force one block so the loop terminates
            layers[layer_num].append(float(l_edge)); out.append([float(l_edge), float(y), float(w)]); left -= 1
    return np.array(out, dtype=np.float64)


def g_jenga18(rng: random.Random, layers=18, per_layer=3):
    out = []
    for ln in range(layers):
        ws = [rng.randint(50, 300) for _ in range(per_layer)]
        gaps = [rng.randint(0, 50) for _ in range(per_layer - 1)]
        total = sum(ws) + sum(gaps)
        x = 750 + rng.randint(-25, 25) - total / 2.0
        for i, w in enumerate(ws):
            out.append([x + w / 2.0, BOTTOM + RECT_H / 2 + RECT_H * ln, float(w)])
            x += w + (gaps[i] if i < per_layer - 1 else 0)
    return np.array(out, dtype=np.float64)


# ---- counter-based twin of g_jenga: the host restatement of the device sampler (csrc/spw_edges.cuh: k_sample_jenga) ----
_M64 = (1 << 64) - 1


def ctr_u32(seed, tower, ctr):
    """draw `ctr` of tower `tower`'s stream: splitmix64 finaliser, high 32 bits (ctr_u32 in spw_edges.cuh)"""
    z = (seed + 0x9E3779B97F4A7C15 * (tower + 1) + 0xD1B54A32D192ED03 * ctr) & _M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    z ^= z >> 31
    return z >> 32


class CtrRng:
    def __init__(self, seed, tower, ctr=0):
        self.seed, self.tower, self.ctr = seed, tower, ctr

    def randint(self, a, b):
        v = a + ctr_u32(self.seed, self.tower, self.ctr) % (b - a + 1)
        self.ctr += 1
        return v

    def random(self):
        v = ctr_u32(self.seed, self.tower, self.ctr) / 4294967296.0
        self.ctr += 1
        return v


def sizes_ctr(seed, n_towers, lo, hi):
    """blocks per tower of the device sampler (draw 0 of every tower's stream)"""
    return np.array([CtrRng(seed, t).randint(lo, hi) for t in range(n_towers)], dtype=np.int64)


def g_jenga_ctr(n, seed, tower):
    """g_jenga with the counter-based generator (draws 1, 2, ... of the tower's stream): what spw_sample_jenga computes"""
    return g_jenga(n, CtrRng(seed, tower, 1))


def g_tower_ctr(n_total, seed, tower):
    """g_tower with the counter-based generator: n_total - 1 stacked blocks + the dropped one (spw_sample_tower)"""
    return g_tower(n_total - 1, CtrRng(seed, tower, 1))


def g_uniform(lo, hi, rng: random.Random):
    return g_jenga(rng.randint(lo, hi), rng)


def make_towers(kind, count, seed, **kw):
    rng = random.Random(seed)
    if kind == 'tower':
        return [g_tower(kw['n'], rng) for _ in range(count)]
    if kind == 'jenga':
        return [g_jenga(kw['n'], rng) for _ in range(count)]
    if kind == 'jenga18':
        return [g_jenga18(rng) for _ in range(count)]
    if kind == 'uniform':
        return [g_uniform(kw['lo'], kw['hi'], rng) for _ in range(count)]
    raise ValueError(kind)


def pack_towers(towers):
    """list of (N_t,3) raw arrays -> (raw (n,3) float64, node_off (T+1,) int64)"""
    sizes = [len(t) for t in towers]
    node_off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    raw = np.concatenate(towers) if towers else np.zeros((0, 3))
    return np.ascontiguousarray(raw, dtype=np.float64), node_off
