"""Generates tests/golden/builder_small.json from the PYTHON builder / simplifier oracles (oracle/pybuilder.py,
oracle/pysimplify.py).

meshoptimizer, which the reference delegates these steps to, is neither under /root/reference nor installed, and the
reference's tests hold no mesh-build vectors: the fixture pins the oracles (and, through tests/test_simplifier_cpu.py, the
product) AGAINST REGRESSION only — "parity unpinned" (DESIGN.md §2).

    python tests/golden/make_golden_builder.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.dirname(HERE))
import pybuilder  # noqa: E402
import pysimplify  # noqa: E402
from test_builder_cpu import torus  # noqa: E402

MESH = dict(nu=28, nv=14)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def generate():
    pos, nrm, uv, i0, _ = torus(MESH["nu"], MESH["nv"])
    built = pybuilder.build(pos, [(i0, 0.0)], normals=nrm, texcoords=uv, auto_lods=True)
    lods = [dict(index_count=int(len(l["indices"])), error=float(np.float32(l["error"])), meshlets=int(len(l["meshlets"])),
                 indices_sha=sha(l["indices"]), meshlets_sha=sha(l["meshlets"]), micro_sha=sha(l["micro"]),
                 vertex_indices_sha=sha(l["vertex_indices"]),
                 bounds_sha=sha(np.array([list(c) + list(axy) + list(e) + [az, cut] for (c, axy, e, az, cut) in l["bounds"]], dtype=np.int64)))
            for l in built["lods"]]
    half, half_err = pysimplify.simplify(i0, pos, None, len(i0) // 2 // 3 * 3)
    return dict(mesh=MESH, vertex_count=int(built["vertex_count"]), positions_sha=sha(built["positions_q"]), normals_sha=sha(built["normals_q"]),
                lods=lods, positions_only_half=dict(index_count=len(half), error=float(half_err), indices_sha=sha(np.array(half, dtype=np.uint32))))


if __name__ == "__main__":
    out = os.path.join(HERE, "builder_small.json")
    json.dump(generate(), open(out, "w"), indent=1)
    print("wrote", out)
