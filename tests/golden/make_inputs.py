"""Packs the reference's example / test input geometries (configs C1-C3 and the
sphere grids) into tests/golden/inputs.npz so GPU-box runs do not need
/root/reference.  Run here (where /root/reference exists):
    python tests/golden/make_inputs.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from viennaray_b200 import io  # noqa: E402

REF = os.environ.get("REF", "/root/reference")
out = {}
for key, path in [("disk3D", "examples/disk3D/trenchGrid3D.dat"),
                  ("disk2D", "examples/disk2D/trenchGrid2D.dat"),
                  ("sphere3D", "tests/Resources/sphereGrid3D_R1.dat"),
                  ("sphere2D", "tests/Resources/sphereGrid2D_R1.dat")]:
    gd, p, n = io.read_grid(os.path.join(REF, path))
    out[key + "_gridDelta"] = np.float32(gd)
    out[key + "_points"] = p
    out[key + "_normals"] = n
gd, nodes, tris = io.read_mesh(os.path.join(REF, "examples/triangle3D/trenchMesh.dat"))
out["triangle3D_gridDelta"] = np.float32(gd)
out["triangle3D_nodes"] = nodes
out["triangle3D_triangles"] = tris
np.savez_compressed(os.path.join(ROOT, "tests/golden/inputs.npz"), **out)
print({k: getattr(v, "shape", v) for k, v in out.items()})
