"""Synthetic ScanNet-shaped scenes (SURVEY.md section 8d / BASELINE.md section 2.2).

No ScanNet data exists in the container, so the benchmark and the parity tests run on seeded
synthetic rooms: points sampled on the floor, the four walls and a few axis-aligned furniture
boxes, voxelised at 2 cm (scale=50) and shifted so every axis starts at 10 -- the coordinate
convention of the reference data loader (examples/ScanNet/datasets/scannet.py:133-135).  The
module input layout is the reference's 4-list  [coords float [P,4] (x,y,z,batch), feats [P,3],
normals (unused), batch_size]  (sparseconvnet/ioLayers.py:52-62, scannet.py:254).
"""
from __future__ import annotations

import numpy as np

PRESETS = {
    # name: (W, D, H metres, nbox)   -- nbox tuned so N0 lands within +-5 % of the nominal target
    "S100k": (3.2, 2.6, 2.2, 2),
    "S250k": (6.0, 4.4, 2.4, 3),
    "S1M": (16.0, 12.0, 3.0, 1),
    "tiny": (0.6, 0.5, 0.4, 1),
    "small": (1.2, 1.0, 0.8, 1),
}


def _plane(rng, origin, u, v, ppv, scale):
    """Uniform points on the parallelogram origin + a*u + b*v, density ppv per voxel-area."""
    area = np.linalg.norm(np.cross(u, v))
    n = max(int(area * scale * scale * ppv), 1)
    ab = rng.random((n, 2))
    return origin[None, :] + ab[:, :1] * u[None, :] + ab[:, 1:] * v[None, :]


def make_scene(preset="S250k", seed=0, scale=50, ppv=1.5, jitter=0.004):
    """Returns (coords int64 [P,3], feats float32 [P,3]); duplicates inside a voxel are kept."""
    W, D, H, nbox = PRESETS[preset] if isinstance(preset, str) else preset
    rng = np.random.default_rng(seed)
    ex, ey, ez = np.eye(3)
    o = np.zeros(3)
    parts = [
        _plane(rng, o, W * ex, D * ey, ppv, scale),                       # floor
        _plane(rng, o, W * ex, H * ez, ppv, scale),                       # wall y=0
        _plane(rng, D * ey, W * ex, H * ez, ppv, scale),                  # wall y=D
        _plane(rng, o, D * ey, H * ez, ppv, scale),                       # wall x=0
        _plane(rng, W * ex, D * ey, H * ez, ppv, scale),                  # wall x=W
    ]
    for _ in range(nbox):
        sz = np.array([rng.uniform(0.15, 0.35) * W, rng.uniform(0.15, 0.35) * D, rng.uniform(0.25, 0.6) * H])
        c = np.array([rng.uniform(0, W - sz[0]), rng.uniform(0, D - sz[1]), 0.0])
        parts += [
            _plane(rng, c + sz[2] * ez, sz[0] * ex, sz[1] * ey, ppv, scale),   # top
            _plane(rng, c, sz[0] * ex, sz[2] * ez, ppv, scale),
            _plane(rng, c + sz[1] * ey, sz[0] * ex, sz[2] * ez, ppv, scale),
            _plane(rng, c, sz[1] * ey, sz[2] * ez, ppv, scale),
            _plane(rng, c + sz[0] * ex, sz[1] * ey, sz[2] * ez, ppv, scale),
        ]
    p = np.concatenate(parts, 0)
    p = p + rng.normal(0.0, jitter, p.shape)
    q = np.floor(p * scale).astype(np.int64)
    q = q - q.min(0, keepdims=True) + 10
    order = rng.permutation(len(q))                # scan order is arbitrary in real data
    q = q[order]
    feats = rng.uniform(-1.0, 1.0, (len(q), 3)).astype(np.float32)
    return q, feats


def make_batch(preset="S250k", seeds=(0,), scale=50):
    """Concatenate scenes the way the reference collate does: coords [P,4] with the batch index in
    the last column, sorted by batch (required by CUDPPWrapper.cu:84-87)."""
    cs, fs = [], []
    for b, s in enumerate(seeds):
        c, f = make_scene(preset, seed=s, scale=scale)
        cs.append(np.concatenate([c, np.full((len(c), 1), b, np.int64)], 1))
        fs.append(f)
    return np.concatenate(cs, 0), np.concatenate(fs, 0)
