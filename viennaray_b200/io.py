"""Readers for the reference's text data formats.

* grid file  -- ``N gridDelta`` then N points, then N normals
  (format read by rayInternal::readGridFromFile, rayUtil.hpp:353-372);
* mesh file  -- ``grid_delta``/``n_nodes``/``n_elements`` header, ``n x y z``
  and ``e i j k`` records (rayInternal::readMeshFromFile, rayUtil.hpp:374-414).
"""
import numpy as np


def read_grid(path):
    with open(path) as f:
        tok = f.read().split()
    n = int(tok[0])
    grid_delta = float(tok[1])
    vals = np.asarray(tok[2:2 + 6 * n], dtype=np.float64).astype(np.float32)
    points = vals[:3 * n].reshape(n, 3).copy()
    normals = vals[3 * n:].reshape(n, 3).copy()
    return grid_delta, points, normals


def read_mesh(path, dim=3):
    grid_delta = None
    nodes, elems = [], []
    with open(path) as f:
        for line in f:
            t = line.split()
            if not t:
                continue
            if t[0] == "grid_delta":
                grid_delta = float(t[1])
            elif t[0] == "n":
                nodes.append([float(x) for x in t[1:4]])
            elif t[0] == "e":
                elems.append([int(x) for x in t[1:1 + dim]])
    return grid_delta, np.asarray(nodes, np.float32), np.asarray(elems, np.uint32)
