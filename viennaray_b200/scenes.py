"""Synthetic geometries: the plane grid used by the reference's tests and the
two large BASELINE configs (C4 trench, C5 hole array).
All generators return float32 ``points (N,3)``, ``normals (N,3)`` and the
``gridDelta`` the disk radius is derived from."""
import numpy as np

_S2 = 0.70710678
_A, _B = 0.89442719, 0.44721360


def plane_grid(grid_delta, extent, direction=(0, 1, 2)):
    """Points of a plane through the origin, normal along ``direction[2]``.
    Same point ORDER as rayInternal::createPlaneGrid (rayUtil.hpp:324-351):
    outer loop over direction[0], inner over direction[1], float32
    accumulation of the coordinate."""
    d0, d1, d2 = direction
    pts = []
    a = np.float32(-extent)
    gd = np.float32(grid_delta)
    ext = np.float32(extent)
    while a <= ext:
        b = np.float32(-extent)
        while b <= ext:
            p = [np.float32(-extent)] * 3
            p[d0], p[d1], p[d2] = a, b, np.float32(0)
            pts.append(p)
            b = np.float32(b + gd)
        a = np.float32(a + gd)
    points = np.asarray(pts, np.float32)
    normals = np.zeros_like(points)
    normals[:, d2] = 1.0
    return points, normals


def _trench_profile(half_width, depth, half_extent):
    """(y, z, ny, nz) of one x-slice of a rectangular trench on the unit
    lattice, corner normals graded in three steps like examples/disk3D."""
    ys, zs, ny, nz = [], [], [], []

    def add(y, z, a, b):
        ys.append(y), zs.append(z), ny.append(a), nz.append(b)

    for side in (-1, 1):
        inward = -side  # wall normal points into the trench
        # top surface, |y| from half_width+1 to half_extent
        for k in range(half_width + 1, half_extent + 1):
            if k == half_width + 1:
                add(side * k, 0, inward * _B, _A)
            else:
                add(side * k, 0, 0.0, 1.0)
        add(side * half_width, 0, inward * _S2, _S2)  # top corner
        for k in range(1, depth):  # wall
            if k == 1 or k == depth - 1:
                add(side * half_width, -k, inward * _A, _B)
            else:
                add(side * half_width, -k, inward * 1.0, 0.0)
        add(side * half_width, -depth, inward * _S2, _S2)  # bottom corner
    for k in range(-half_width + 1, half_width):  # bottom
        if abs(k) == half_width - 1:
            add(k, -depth, (-1 if k > 0 else 1) * _B, _A)
        else:
            add(k, -depth, 0.0, 1.0)
    return (np.asarray(ys, np.float32), np.asarray(zs, np.float32), np.asarray(ny, np.float32),
            np.asarray(nz, np.float32))


def trench(num_slices=999, half_width=50, depth=200, half_extent=300):
    """C4: lattice trench along x, gridDelta 1.  Default: 999 slices x 1001
    points = 999,999 disks (the "1M-disk trench")."""
    y, z, ny, nz = _trench_profile(half_width, depth, half_extent)
    m = len(y)
    x = np.arange(num_slices, dtype=np.float32)
    points = np.empty((num_slices, m, 3), np.float32)
    normals = np.zeros((num_slices, m, 3), np.float32)
    points[:, :, 0] = x[:, None]
    points[:, :, 1] = y[None, :]
    points[:, :, 2] = z[None, :]
    normals[:, :, 1] = ny[None, :]
    normals[:, :, 2] = nz[None, :]
    return points.reshape(-1, 3), normals.reshape(-1, 3), 1.0


def hole_array(cells=10, pitch=100, radius=20, depth=240):
    """C5: cells x cells cylindrical holes of the given radius and depth on a
    square pitch, gridDelta 1; plane and bottoms on the unit lattice, walls as
    rings of arc length ~1 with radial (inward) normals.  Default ~4.02M disks."""
    size = cells * pitch
    gx, gy = np.meshgrid(np.arange(size, dtype=np.float32), np.arange(size, dtype=np.float32),
                         indexing="ij")
    cx = (np.floor(gx / pitch) + 0.5) * pitch
    cy = (np.floor(gy / pitch) + 0.5) * pitch
    rr = np.hypot(gx - cx, gy - cy)
    top = rr > radius
    bot = rr < radius - 0.5
    pts = [np.stack([gx[top], gy[top], np.zeros(top.sum(), np.float32)], 1)]
    nrm = [np.tile(np.array([[0, 0, 1]], np.float32), (int(top.sum()), 1))]
    pts.append(np.stack([gx[bot], gy[bot], np.full(bot.sum(), -depth, np.float32)], 1))
    nrm.append(np.tile(np.array([[0, 0, 1]], np.float32), (int(bot.sum()), 1)))
    nth = int(round(2 * np.pi * radius))
    th = (np.arange(nth, dtype=np.float64) + 0.5) * (2 * np.pi / nth)
    ring = np.stack([radius * np.cos(th), radius * np.sin(th)], 1)
    rn = -np.stack([np.cos(th), np.sin(th)], 1)
    zs = -(np.arange(depth, dtype=np.float64) + 0.5)
    centers = (np.arange(cells) + 0.5) * pitch
    ccx, ccy = np.meshgrid(centers, centers, indexing="ij")
    for hx, hy in zip(ccx.ravel(), ccy.ravel()):
        w = np.empty((depth, nth, 3), np.float32)
        w[:, :, 0] = hx + ring[None, :, 0]
        w[:, :, 1] = hy + ring[None, :, 1]
        w[:, :, 2] = zs[:, None]
        wn = np.zeros((depth, nth, 3), np.float32)
        wn[:, :, 0] = rn[None, :, 0]
        wn[:, :, 1] = rn[None, :, 1]
        pts.append(w.reshape(-1, 3))
        nrm.append(wn.reshape(-1, 3))
    return (np.ascontiguousarray(np.concatenate(pts), np.float32),
            np.ascontiguousarray(np.concatenate(nrm), np.float32), 1.0)
