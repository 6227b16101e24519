"""Deterministic synthetic overlapping scans.

World = axis-aligned floor plan of rooms (floor, ceiling, walls with doorways,
boxes, cylinders) carrying a procedural colour texture, so that SIFT3D finds
DoG extrema well above the reference's default ``keypoint_threshold`` of 5
(intensity scale 0..255, map_merge_3d/include/map_merge_3d/map_merging.h:34).
Map *i* = points of the world inside a window, sampled independently with
5 mm noise along the surface normal, expressed in the frame of a virtual sensor
standing inside the window (random yaw, small roll/pitch).  ``truth[i]`` maps
map-i coordinates to world coordinates, so the pairwise transform the pipeline
should recover for (i -> j) is ``inv(truth[j]) @ truth[i]``.

Point layout everywhere in this repo: float32[n, 4] = x, y, z, rgba-bits
(pcl::PointXYZRGB::rgba = a<<24 | r<<16 | g<<8 | b reinterpreted as float).
"""
from __future__ import annotations

import numpy as np

__all__ = ["World", "make_maps", "pack_rgba", "unpack_rgb", "CONFIGS"]


def pack_rgba(r, g, b, a=255):
    r = np.asarray(r, np.uint32)
    g = np.asarray(g, np.uint32)
    b = np.asarray(b, np.uint32)
    return ((np.uint32(a) << np.uint32(24)) | (r << np.uint32(16)) | (g << np.uint32(8)) | b).astype(np.uint32)


def unpack_rgb(rgba):
    rgba = np.asarray(rgba, np.uint32)
    return (rgba >> 16) & 0xFF, (rgba >> 8) & 0xFF, rgba & 0xFF


class World:
    """A set of textured rectangles and vertical cylinders."""

    def __init__(self, seed: int, size_x: float, size_y: float, rooms_x: int, rooms_y: int, height: float = 3.0):
        rng = np.random.default_rng(seed * 7919 + 13)
        self.size = (size_x, size_y, height)
        rects = []  # (origin[3], u[3], v[3], lu, lv, normal[3])
        cyls = []   # (cx, cy, r, z0, z1)

        def rect(o, u, lu, v, lv):
            o = np.asarray(o, float)
            u = np.asarray(u, float)
            v = np.asarray(v, float)
            rects.append((o, u, v, float(lu), float(lv), np.cross(u, v)))

        rect((0, 0, 0), (1, 0, 0), size_x, (0, 1, 0), size_y)            # floor
        rect((0, 0, height), (1, 0, 0), size_x, (0, 1, 0), size_y)       # ceiling
        rect((0, 0, 0), (1, 0, 0), size_x, (0, 0, 1), height)            # outer walls
        rect((0, size_y, 0), (1, 0, 0), size_x, (0, 0, 1), height)
        rect((0, 0, 0), (0, 1, 0), size_y, (0, 0, 1), height)
        rect((size_x, 0, 0), (0, 1, 0), size_y, (0, 0, 1), height)
        rw, rh = size_x / rooms_x, size_y / rooms_y
        door = 1.2
        for ix in range(1, rooms_x):       # inner walls along y with one doorway per room edge
            x = ix * rw
            for iy in range(rooms_y):
                y0, y1 = iy * rh, (iy + 1) * rh
                d0 = y0 + rng.uniform(0.5, rh - door - 0.5)
                rect((x, y0, 0), (0, 1, 0), d0 - y0, (0, 0, 1), height)
                rect((x, d0 + door, 0), (0, 1, 0), y1 - d0 - door, (0, 0, 1), height)
                rect((x, d0, 2.1), (0, 1, 0), door, (0, 0, 1), height - 2.1)
        for iy in range(1, rooms_y):
            y = iy * rh
            for ix in range(rooms_x):
                x0, x1 = ix * rw, (ix + 1) * rw
                d0 = x0 + rng.uniform(0.5, rw - door - 0.5)
                rect((x0, y, 0), (1, 0, 0), d0 - x0, (0, 0, 1), height)
                rect((d0 + door, y, 0), (1, 0, 0), x1 - d0 - door, (0, 0, 1), height)
                rect((d0, y, 2.1), (1, 0, 0), door, (0, 0, 1), height - 2.1)
        for ix in range(rooms_x):
            for iy in range(rooms_y):
                nb = int(rng.integers(6, 13))
                for _ in range(nb):
                    bx, by, bz = rng.uniform(0.4, 1.6), rng.uniform(0.4, 1.6), rng.uniform(0.4, 1.8)
                    x0 = ix * rw + rng.uniform(0.3, max(0.31, rw - bx - 0.3))
                    y0 = iy * rh + rng.uniform(0.3, max(0.31, rh - by - 0.3))
                    rect((x0, y0, bz), (1, 0, 0), bx, (0, 1, 0), by)          # top
                    rect((x0, y0, 0), (1, 0, 0), bx, (0, 0, 1), bz)           # sides
                    rect((x0, y0 + by, 0), (1, 0, 0), bx, (0, 0, 1), bz)
                    rect((x0, y0, 0), (0, 1, 0), by, (0, 0, 1), bz)
                    rect((x0 + bx, y0, 0), (0, 1, 0), by, (0, 0, 1), bz)
                nc = int(rng.integers(2, 5))
                for _ in range(nc):
                    r = rng.uniform(0.15, 0.45)
                    cyls.append((ix * rw + rng.uniform(0.6, rw - 0.6), iy * rh + rng.uniform(0.6, rh - 0.6), r, 0.0,
                                 rng.uniform(1.0, height)))
        self.rects = rects
        self.cyls = cyls
        ns = len(rects) + len(cyls)
        self.phase = rng.uniform(0, 2 * np.pi, size=(ns, 6))
        self.base = rng.uniform(90, 170, size=(ns, 3))
        self.area = np.array([r[3] * r[4] for r in rects] + [2 * np.pi * c[2] * (c[4] - c[3]) for c in cyls])

    def sample(self, rng, n: int, window):
        """n points whose (x, y) lie inside window=(x0, x1, y0, y1); returns xyz, normal, rgb."""
        x0, x1, y0, y1 = window
        out_p, out_n, out_c = [], [], []
        got = 0
        frac = max(0.05, (x1 - x0) * (y1 - y0) / (self.size[0] * self.size[1]))
        prob = self.area / self.area.sum()
        nr = len(self.rects)
        O = np.array([r[0] for r in self.rects]); U = np.array([r[1] for r in self.rects]); V = np.array([r[2] for r in self.rects])
        LU = np.array([r[3] for r in self.rects]); LV = np.array([r[4] for r in self.rects]); NN = np.array([r[5] for r in self.rects])
        C = np.array(self.cyls) if self.cyls else np.zeros((0, 5))
        while got < n:
            m = int((n - got) / frac * 1.3) + 1024
            sid = rng.choice(len(prob), size=m, p=prob)
            a = rng.uniform(0, 1, m)
            b = rng.uniform(0, 1, m)
            p = np.empty((m, 3))
            nrm = np.empty((m, 3))
            uu = np.empty(m)
            vv = np.empty(m)
            isr = sid < nr
            s = sid[isr]
            uu[isr] = a[isr] * LU[s]
            vv[isr] = b[isr] * LV[s]
            p[isr] = O[s] + U[s] * uu[isr, None] + V[s] * vv[isr, None]
            nrm[isr] = NN[s]
            c = sid[~isr] - nr
            th = a[~isr] * 2 * np.pi
            p[~isr, 0] = C[c, 0] + C[c, 2] * np.cos(th)
            p[~isr, 1] = C[c, 1] + C[c, 2] * np.sin(th)
            p[~isr, 2] = C[c, 3] + b[~isr] * (C[c, 4] - C[c, 3])
            nrm[~isr, 0] = np.cos(th); nrm[~isr, 1] = np.sin(th); nrm[~isr, 2] = 0
            uu[~isr] = th * C[c, 2]
            vv[~isr] = p[~isr, 2]
            keep = (p[:, 0] >= x0) & (p[:, 0] <= x1) & (p[:, 1] >= y0) & (p[:, 1] <= y1)
            p, nrm, uu, vv, sid = p[keep], nrm[keep], uu[keep], vv[keep], sid[keep]
            ph = self.phase[sid]
            chk = (((np.floor(uu / 0.5) + np.floor(vv / 0.5)) % 2) * 2 - 1) * 40.0
            col = np.empty((len(p), 3))
            for ch in range(3):
                col[:, ch] = (self.base[sid, ch] + 60.0 * np.sin(2 * np.pi * uu / (0.7 + 0.1 * ch) + ph[:, ch]) *
                              np.sin(2 * np.pi * vv / (0.9 - 0.1 * ch) + ph[:, 3 + ch]) + chk + rng.normal(0, 4.0, len(p)))
            p = p + nrm * rng.normal(0, 0.005, len(p))[:, None]
            out_p.append(p); out_n.append(nrm); out_c.append(np.clip(col, 0, 255))
            got += len(p)
        p = np.concatenate(out_p)[:n]
        nrm = np.concatenate(out_n)[:n]
        col = np.concatenate(out_c)[:n]
        return p, nrm, col


def _rot(yaw, pitch, roll):
    cy, sy = np.cos(yaw), np.sin(yaw)
    cp, sp = np.cos(pitch), np.sin(pitch)
    cr, sr = np.cos(roll), np.sin(roll)
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    return Rz @ Ry @ Rx


def make_maps(seed: int, n_maps: int, n_points: int, size_x: float, size_y: float, rooms_x: int, rooms_y: int,
              window_frac: float = 0.6, layout: str = "chain", only=None):
    """Returns (maps, truth): maps[i] float32[n_points, 4]; truth[i] float64 4x4 map_i -> world.
    only: iterable of map indices to generate (the others come back as None; every map has its own random stream, so a
    rank that generates just its block gets the same points as a full run)."""
    world = World(seed, size_x, size_y, rooms_x, rooms_y)
    maps, truth = [], []
    wx = size_x * window_frac
    only = None if only is None else set(int(i) for i in only)
    for i in range(n_maps):
        if only is not None and i not in only:
            maps.append(None)
            truth.append(None)
            continue
        rng = np.random.default_rng(seed * 1000 + i)
        # windows slide along x so neighbours overlap >= 50 %, second neighbours >= 25 %
        if n_maps > 1:
            step = (size_x - wx) / (n_maps - 1)
            step = min(step, wx * 0.35)
        else:
            step = 0.0
        x0 = min(i * step, size_x - wx)
        if layout == "ring" and n_maps > 2:
            x0 = (size_x - wx) * 0.5 * (1 - np.cos(2 * np.pi * i / n_maps))
        window = (x0, x0 + wx, 0.0, size_y)
        p, _, col = world.sample(rng, n_points, window)
        sensor = np.array([x0 + wx * 0.5 + rng.uniform(-2, 2), size_y * 0.5 + rng.uniform(-2, 2), 1.2 + rng.uniform(-0.3, 0.3)])
        R = _rot(rng.uniform(-np.pi, np.pi), np.deg2rad(rng.uniform(-5, 5)), np.deg2rad(rng.uniform(-5, 5)))
        T = np.eye(4)
        T[:3, :3] = R
        T[:3, 3] = sensor
        local = (p - sensor) @ R  # R^T (p - t)
        m = np.empty((n_points, 4), np.float32)
        m[:, :3] = local.astype(np.float32)
        rgba = pack_rgba(col[:, 0].astype(np.uint32), col[:, 1].astype(np.uint32), col[:, 2].astype(np.uint32))
        m[:, 3] = rgba.view(np.float32)
        maps.append(m)
        truth.append(T)
    return maps, truth


# BASELINE.json configs -> generator arguments (sizes chosen so that the voxelised
# clouds land near SURVEY §8's planning numbers).
CONFIGS = {
    "c1": dict(seed=1, n_maps=2, n_points=200_000, size_x=8.0, size_y=6.0, rooms_x=1, rooms_y=1, window_frac=0.8),
    "c2": dict(seed=2, n_maps=8, n_points=500_000, size_x=28.0, size_y=20.0, rooms_x=2, rooms_y=2, window_frac=0.6),
    "c3": dict(seed=3, n_maps=32, n_points=1_000_000, size_x=40.0, size_y=30.0, rooms_x=4, rooms_y=3, window_frac=0.6),
    # configs[3]: Harris3D + SHOT variant, 16 maps, tight inlier_threshold (bench.py sets the parameters)
    "c4": dict(seed=4, n_maps=16, n_points=500_000, size_x=36.0, size_y=24.0, rooms_x=3, rooms_y=2, window_frac=0.6),
    # configs[4]: 4 maps at 10M points for output_resolution composeMaps
    "c5": dict(seed=5, n_maps=4, n_points=10_000_000, size_x=40.0, size_y=30.0, rooms_x=4, rooms_y=3, window_frac=0.6),
    "tiny": dict(seed=11, n_maps=2, n_points=40_000, size_x=6.0, size_y=5.0, rooms_x=1, rooms_y=1, window_frac=0.85),
    "small": dict(seed=12, n_maps=3, n_points=60_000, size_x=9.0, size_y=6.0, rooms_x=1, rooms_y=1, window_frac=0.75),
}
