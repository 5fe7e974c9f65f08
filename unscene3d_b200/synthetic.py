"""Seeded synthetic "ScanNet-shaped" scenes for parity tests and benchmarks.

A scene is an axis-aligned room (floor 5.0 m x 4.0 m, four walls 2.4 m high, 12–20 boxes), sampled
at 1 cm with 5 mm Gaussian jitter, mean-centred (so negative coordinates occur, as after
datasets/freemask_semseg.py:335) and voxelised at 2 cm (`floor(xyz / 0.02)`, datasets/utils.py:403).
The room is scaled so that the number of occupied voxels hits the requested count, then random
surplus voxels are dropped to make it exact.  Row order follows the sampling order (surface by
surface, raster within a surface) — spatially coherent like a reconstructed mesh, not sorted.

Everything is numpy on the host: this is the data generator, not the hot path.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List

import numpy as np

VOXEL = 0.02
COLOR_MEAN = np.array([0.47793125906962, 0.4303257521323044, 0.3749598901421883], dtype=np.float32)
COLOR_STD = np.array([0.2834475483823543, 0.27566157565723015, 0.27018971370874995], dtype=np.float32)


@dataclass
class Scene:
    coords: np.ndarray          # int32 [N, 3] voxel coordinates
    colors: np.ndarray          # float32 [N, 3] normalised colour features
    xyz: np.ndarray             # float32 [N, 3] raw coordinates (metres, mean-centred)
    point2segment: np.ndarray   # int64 [N] segment id of every voxel, 0..S-1
    adjacency: np.ndarray       # int64 [E, 2] undirected segment adjacency
    segment_mask: np.ndarray = field(default=None)  # bool [T, S] pseudo masks on segments
    masks: np.ndarray = field(default=None)         # bool [T, N] pseudo masks on voxels
    labels: np.ndarray = field(default=None)        # int64 [T]

    @property
    def n(self):
        return self.coords.shape[0]

    @property
    def num_segments(self):
        return int(self.point2segment.max()) + 1


def _rect(origin, u, v):
    return np.asarray(origin, float), np.asarray(u, float), np.asarray(v, float)


def _surfaces(rng, scale):
    L, W, H = 5.0 * scale, 4.0 * scale, 2.4
    s = [_rect((0, 0, 0), (L, 0, 0), (0, W, 0)),
         _rect((0, 0, 0), (L, 0, 0), (0, 0, H)), _rect((0, W, 0), (L, 0, 0), (0, 0, H)),
         _rect((0, 0, 0), (0, W, 0), (0, 0, H)), _rect((L, 0, 0), (0, W, 0), (0, 0, H))]
    for _ in range(int(rng.integers(12, 21))):
        e = rng.uniform(0.3, 1.5, size=3) * min(1.0, scale)
        against_wall = rng.random() < 0.5
        x0 = rng.uniform(0, L - e[0])
        y0 = (0.0 if rng.random() < 0.5 else W - e[1]) if against_wall else rng.uniform(0, W - e[1])
        o = np.array([x0, y0, 0.0])
        ex, ey, ez = np.array([e[0], 0, 0]), np.array([0, e[1], 0]), np.array([0, 0, e[2]])
        s += [_rect(o + ez, ex, ey), _rect(o, ex, ez), _rect(o + ey, ex, ez), _rect(o, ey, ez), _rect(o + ex, ey, ez)]
    return s


def _sample(rng, surfaces, step=0.01, jitter=0.005, cell=0.30):
    pts, seg = [], []
    seg_base, adj = 0, []
    for o, u, v in surfaces:
        lu, lv = np.linalg.norm(u), np.linalg.norm(v)
        nu, nv = max(int(lu / step), 1), max(int(lv / step), 1)
        a, b = np.meshgrid((np.arange(nu) + 0.5) * step, (np.arange(nv) + 0.5) * step, indexing="ij")
        a, b = a.reshape(-1), b.reshape(-1)
        p = o + np.outer(a / lu, u) + np.outer(b / lv, v)
        pts.append(p + rng.normal(0.0, jitter, size=p.shape))
        cu, cv = max(int(np.ceil(lu / cell)), 1), max(int(np.ceil(lv / cell)), 1)
        iu, iv = np.minimum((a / cell).astype(np.int64), cu - 1), np.minimum((b / cell).astype(np.int64), cv - 1)
        seg.append(seg_base + iu * cv + iv)
        ids = seg_base + np.arange(cu * cv).reshape(cu, cv)
        adj.append(np.stack([ids[:-1].ravel(), ids[1:].ravel()], 1))
        adj.append(np.stack([ids[:, :-1].ravel(), ids[:, 1:].ravel()], 1))
        seg_base += cu * cv
    return np.concatenate(pts), np.concatenate(seg), np.concatenate(adj)


def _voxelise(xyz, seg):
    xyz = xyz - xyz.mean(0, keepdims=True)
    c = np.floor(xyz / VOXEL).astype(np.int64)
    key = ((c[:, 0] + (1 << 17)) << 36) | ((c[:, 1] + (1 << 17)) << 18) | (c[:, 2] + (1 << 17))
    _, first = np.unique(key, return_index=True)
    first.sort()
    return c[first].astype(np.int32), xyz[first].astype(np.float32), seg[first]


def make_scene(n_voxels: int, seed: int = 0, with_masks: bool = True, num_masks: int = 20) -> Scene:
    rng = np.random.default_rng(seed)
    geo_seed = int(rng.integers(1 << 31))
    scale = (n_voxels / 200_000.0) ** 0.5
    # The room is rescaled until the voxel count lands in [n, 1.03 n]; the count is not a smooth function of the scale (box
    # sizes saturate), so the search keeps the smallest attempt that reached n and falls back to it (every seed must yield a
    # scene: ranks of a multi-GPU run use seed = rank).
    best = chosen = None
    for attempt in range(24):
        g = np.random.default_rng(geo_seed)
        xyz, seg, adj = _sample(g, _surfaces(g, scale))
        coords, xyzv, segv = _voxelise(xyz, seg)
        cnt = coords.shape[0]
        if cnt >= n_voxels and (best is None or cnt < best[0].shape[0]):
            best = (coords, xyzv, segv, adj)
        if n_voxels <= cnt <= int(n_voxels * 1.03) + 64 or (attempt == 5 and cnt >= n_voxels):
            chosen = (coords, xyzv, segv, adj)
            break
        if attempt >= 5 and best is not None:
            chosen = best
            break
        scale *= (n_voxels * (1.015 if attempt < 6 else 1.06) / cnt) ** 0.5
    assert chosen is not None, "scene generator did not reach the requested voxel count"
    coords, xyzv, segv, adj = chosen
    keep = np.sort(rng.permutation(coords.shape[0])[:n_voxels])
    coords, xyzv, segv = coords[keep], xyzv[keep], segv[keep]
    # compact segment ids, remap adjacency
    uniq, p2s = np.unique(segv, return_inverse=True)
    remap = -np.ones(int(max(adj.max(), uniq.max())) + 1, dtype=np.int64)
    remap[uniq] = np.arange(uniq.shape[0])
    adj = remap[adj]
    adj = adj[(adj >= 0).all(1)]
    colors = ((rng.uniform(0, 255, size=(n_voxels, 3)) / 255.0 - COLOR_MEAN) / COLOR_STD).astype(np.float32)
    scene = Scene(coords, colors, xyzv, p2s.astype(np.int64), adj)
    if with_masks:
        _add_pseudo_masks(scene, rng, num_masks)
    return scene


def _add_pseudo_masks(scene: Scene, rng, num_masks: int):
    """T non-overlapping masks, each a connected union of 4–40 segments (targets as built by
    datasets/utils.py:480-527: segment_mask [T, S], masks [T, N], labels all 1)."""
    S = scene.num_segments
    nbrs: List[List[int]] = [[] for _ in range(S)]
    for a, b in scene.adjacency:
        nbrs[a].append(b)
        nbrs[b].append(a)
    taken = np.zeros(S, dtype=bool)
    seg_masks = []
    for _ in range(num_masks * 4):
        if len(seg_masks) == num_masks:
            break
        free = np.nonzero(~taken)[0]
        if free.size == 0:
            break
        start = int(free[rng.integers(free.size)])
        want = int(rng.integers(4, 41))
        comp, frontier = [start], [start]
        taken[start] = True
        while frontier and len(comp) < want:
            cur = frontier.pop(0)
            for nb in nbrs[cur]:
                if not taken[nb] and len(comp) < want:
                    taken[nb] = True
                    comp.append(nb)
                    frontier.append(nb)
        m = np.zeros(S, dtype=bool)
        m[comp] = True
        seg_masks.append(m)
    scene.segment_mask = np.stack(seg_masks) if seg_masks else np.zeros((0, S), dtype=bool)
    scene.masks = scene.segment_mask[:, scene.point2segment]
    scene.labels = np.ones(scene.segment_mask.shape[0], dtype=np.int64)


def collate(scenes: List[Scene]):
    """Batched inputs exactly like ME.utils.sparse_collate: int32 [sum N, 4] (batch first) and float32
    features [sum N, 6] = (colour, raw xyz) — the trainer splits xyz off (trainer/trainer.py:110-113)."""
    coords = np.concatenate([np.concatenate([np.full((s.n, 1), b, np.int32), s.coords], 1) for b, s in enumerate(scenes)])
    feats = np.concatenate([np.concatenate([s.colors, s.xyz], 1) for s in scenes]).astype(np.float32)
    return coords, feats


def level_sizes(coords: np.ndarray, levels: int = 5):
    """Unique voxel counts at tensor strides 1, 2, 4, ... (host-side, for roofline byte accounting)."""
    out = []
    c = coords.astype(np.int64)
    for l in range(levels):
        s = 1 << l
        q = np.concatenate([c[:, :1], np.floor_divide(c[:, 1:], s)], 1)
        key = (q[:, 0] << 54) | ((q[:, 1] + (1 << 17)) << 36) | ((q[:, 2] + (1 << 17)) << 18) | (q[:, 3] + (1 << 17))
        out.append(int(np.unique(key).shape[0]))
    return out


def make_ncut_scene(n_points=300_000, n_segments=2048, seed=0):
    """per-segment random centres (sigma_within = 0.3) so that the thresholded affinity at tau = 0.6 is non-trivial;
    segments are grouped into ~40 'objects' whose centres are correlated, adjacency = ring + random chords inside an object."""
    rng = np.random.default_rng(seed)
    n_obj = 40
    obj_of = rng.integers(0, n_obj, n_segments)
    ca = rng.normal(size=(n_obj, 384)).astype(np.float32)[obj_of] + 0.35 * rng.normal(size=(n_segments, 384)).astype(np.float32)
    cb = rng.normal(size=(n_obj, 96)).astype(np.float32)[obj_of] + 0.35 * rng.normal(size=(n_segments, 96)).astype(np.float32)
    seg = rng.integers(0, n_segments, n_points)
    seg[:n_segments] = np.arange(n_segments)
    fa = ca[seg] + 0.3 * rng.normal(size=(n_points, 384)).astype(np.float32)
    fb = cb[seg] + 0.3 * rng.normal(size=(n_points, 96)).astype(np.float32)
    edges = []
    for o in range(n_obj):
        members = np.nonzero(obj_of == o)[0]
        if len(members) < 2:
            continue
        nxt = np.roll(members, -1)
        edges.append(np.stack([members, nxt], 1))
        extra = rng.integers(0, len(members), (len(members), 2))
        edges.append(members[extra])
    e = np.concatenate(edges)
    e = e[e[:, 0] != e[:, 1]]
    e = np.unique(np.concatenate([e, e[:, ::-1]]), axis=0)
    return seg.astype(np.int64), fa, fb, e.astype(np.int64)


class SyntheticFreemaskDataset:
    """Dataset stand-in with the sample layout of the reference's datasets/freemask_semseg.py (what FreeMaskVoxelizeCollate /
    freemask_voxelize consume, datasets/utils.py:370-404): (coordinates float64 [P, 3] in metres, features float32 [P, 3 colour
    + 3 raw xyz], freemasks int64 [P, 2 + T] = (label, T pseudo-mask columns, segment id), scene name, raw colours, raw normals,
    raw coordinates, index, segment connectivity).  Selected from the unmodified conf/ tree with
    `data.train_dataset._target_=unscene3d_b200.synthetic.SyntheticFreemaskDataset`; every other key of the dataset node is
    accepted and ignored."""

    def __init__(self, n_scenes: int = 4, n_voxels: int = 20000, num_masks: int = 8, voxel_size: float = 0.02, seed0: int = 0, mode="train",
                 **unused):
        self.n_scenes, self.n_voxels, self.num_masks, self.voxel_size, self.seed0 = int(n_scenes), int(n_voxels), int(num_masks), voxel_size, int(seed0)
        self.mode = mode
        self.label_info = {0: {"name": "background", "validation": True, "color": [0, 0, 0]},
                           1: {"name": "foreground", "validation": True, "color": [255, 0, 0]}}
        self._cache = {}

    def __len__(self):
        return self.n_scenes

    def __getitem__(self, idx):
        idx = int(idx) % self.n_scenes
        if idx not in self._cache:
            s = make_scene(self.n_voxels, seed=self.seed0 + idx, with_masks=True, num_masks=self.num_masks)
            xyz = (s.coords.astype(np.float64) + 0.5) * self.voxel_size          # one point per voxel, at the voxel centre
            feats = np.concatenate([s.colors, xyz.astype(np.float32)], 1).astype(np.float32)
            fm = np.concatenate([np.zeros((s.n, 1), np.int64), s.masks.T.astype(np.int64), s.point2segment[:, None].astype(np.int64)], 1)
            conn = np.concatenate([s.adjacency, s.adjacency[:, ::-1]]).astype(np.int64)
            self._cache[idx] = (xyz, feats, fm, f"synthetic{self.seed0 + idx:04d}", s.colors.copy(), np.zeros_like(s.colors), xyz.copy(), idx, conn)
        return self._cache[idx]
