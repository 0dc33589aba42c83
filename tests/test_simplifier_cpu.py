"""LOD simplifier of the mesh builder (oxb_simplify / OxbMeshInput::auto_lods, csrc/host/mesh_simplifier.cpp — the role of
meshopt_simplifyWithAttributes in AssetManager_GLTF.cpp:596-641) vs its independent pure-Python oracle
(oracle/pysimplify.py): identical index buffers and errors, plus the properties the LOD chain relies on.  Host only: runs
without a GPU."""
from collections import Counter

import numpy as np
import pytest

from oxylus_b200 import capi
from test_builder_cpu import parse, torus

FLT_MAX = 3.4028234663852886e38


# ------------------------------------------------------------------------------------------------ meshes
def seam_torus(nu, nv):
    """the torus with ring i = 0 duplicated: the strip that closes the loop uses copies with tilted normals, so ring 0 is an
    attribute seam (two wedges per position, like a UV seam or a crease)"""
    pos, nrm, uv, i0, _ = torus(nu, nv)
    pad = 5
    ring = pad + np.arange(nv)                      # vertices (i = 0, j)
    dup = len(pos) + np.arange(nv)
    pos = np.concatenate([pos, pos[ring]])
    tilted = nrm[ring] + np.array([0.0, 0.0, 0.8], dtype=np.float32)
    nrm = np.concatenate([nrm, (tilted / np.linalg.norm(tilted, axis=1, keepdims=True)).astype(np.float32)])
    tris = i0.reshape(-1, 3).copy()
    strip = np.arange((nu - 1) * nv * 2, nu * nv * 2)   # triangles of i = nu - 1: they reach ring 0 through (i + 1) % nu
    sub = tris[strip]
    for j in range(nv):
        sub[sub == ring[j]] = dup[j]
    tris[strip] = sub
    return pos, nrm, tris.reshape(-1)


def hard_cube(n):
    """six n x n grids with face normals: cube edges are seams (2 wedges), corners have 3 wedges (locked)"""
    pos, nrm, tris = [], [], []
    axes = [((1, 0, 0), (0, 1, 0), (0, 0, 1)), ((0, 1, 0), (0, 0, 1), (1, 0, 0)), ((0, 0, 1), (1, 0, 0), (0, 1, 0))]
    for (u, v, w) in axes:
        for sign in (1.0, -1.0):
            base = len(pos)
            uu, vv, ww = np.array(u, float), np.array(v, float), np.array(w, float) * sign
            if sign < 0:
                uu, vv = vv, uu                     # keep the winding outward
            for a in range(n + 1):
                for b in range(n + 1):
                    pos.append((2 * a / n - 1) * uu + (2 * b / n - 1) * vv + ww)
                    nrm.append(ww)
            for a in range(n):
                for b in range(n):
                    p, q, r, s = base + a * (n + 1) + b, base + (a + 1) * (n + 1) + b, base + (a + 1) * (n + 1) + b + 1, base + a * (n + 1) + b + 1
                    tris += [(p, q, r), (p, r, s)]
    return np.array(pos, dtype=np.float32), np.array(nrm, dtype=np.float32), np.array(tris, dtype=np.uint32).reshape(-1)


def open_grid(n, bump=0.0, seed=2):
    """(n+1)^2 height-field patch with a border"""
    g = np.arange(n + 1) / n
    x, y = np.meshgrid(g, g, indexing="ij")
    z = bump * np.sin(3 * x) * np.cos(2 * y) + bump * 0.05 * np.random.default_rng(seed).standard_normal(x.shape)
    pos = np.stack([x, y, z], axis=2).reshape(-1, 3).astype(np.float32)
    nrm = np.tile(np.array([0, 0, 1], dtype=np.float32), (len(pos), 1))
    tris = []
    for a in range(n):
        for b in range(n):
            p, q, r, s = a * (n + 1) + b, (a + 1) * (n + 1) + b, (a + 1) * (n + 1) + b + 1, a * (n + 1) + b + 1
            tris += [(p, q, r), (p, r, s)]
    return pos, nrm, np.array(tris, dtype=np.uint32).reshape(-1)


# ------------------------------------------------------------------------------------------------ topology helpers
def position_ids(pos):
    ids, out = {}, np.zeros(len(pos), dtype=np.int64)
    for v, p in enumerate(np.asarray(pos, dtype=np.float32)):
        out[v] = ids.setdefault(p.tobytes(), len(ids))
    return out


def directed_edges(idx, ids):
    c = Counter()
    t = ids[np.asarray(idx, dtype=np.int64)].reshape(-1, 3)
    for a, b, cc in t:
        c[(a, b)] += 1
        c[(b, cc)] += 1
        c[(cc, a)] += 1
    return c


def is_closed_manifold(idx, pos):
    c = directed_edges(idx, position_ids(pos))
    return all(n == 1 and c.get((b, a), 0) == 1 for (a, b), n in c.items())


def border_edges(idx, pos):
    c = directed_edges(idx, position_ids(pos))
    return {(a, b) for (a, b) in c if (b, a) not in c}


def point_triangle_distance(p, a, b, c):
    """distances of the points p (N,3) to the triangles (a,b,c) (T,3 each): (N,T), by projecting on the plane and clamping to
    the edges (brute force; test sizes only)"""
    def seg(p, a, b):
        ab = b - a
        t = np.clip(np.einsum("ntk,tk->nt", p[:, None, :] - a[None], ab) / np.maximum(np.einsum("tk,tk->t", ab, ab), 1e-300), 0, 1)
        return np.linalg.norm(p[:, None, :] - (a[None] + t[..., None] * ab[None]), axis=2)
    n = np.cross(b - a, c - a)
    ln = np.linalg.norm(n, axis=1)
    ok = ln > 0
    n = n / np.maximum(ln, 1e-300)[:, None]
    d = np.einsum("ntk,tk->nt", p[:, None, :] - a[None], n)
    q = p[:, None, :] - d[..., None] * n[None]
    def side(u, v):
        return np.einsum("ntk,tk->nt", np.cross((v - u)[None], q - u[None]), n)
    inside = (side(a, b) >= 0) & (side(b, c) >= 0) & (side(c, a) >= 0) & ok[None]
    edge = np.minimum(np.minimum(seg(p, a, b), seg(p, b, c)), seg(p, c, a))
    return np.where(inside, np.abs(d), edge)


@pytest.fixture(scope="module")
def pys():
    import pysimplify

    return pysimplify


CASES = {
    "torus": lambda: torus(24, 16)[:2] + (torus(24, 16)[3],),
    "torus_positions_only": lambda: (torus(20, 12)[0], None, torus(20, 12)[3]),
    "seam_torus": lambda: seam_torus(20, 12),
    "hard_cube": lambda: hard_cube(6),
    "open_grid": lambda: open_grid(14, bump=0.2),
}


# ------------------------------------------------------------------------------------------------ parity with the oracle
@pytest.mark.parametrize("name", list(CASES))
def test_simplify_matches_oracle(pys, name):
    pos, nrm, idx = CASES[name]()
    for div in (2, 5):
        target = len(idx) // 3 // div * 3
        want, want_err = pys.simplify(idx, pos, nrm, target)
        got, got_err = capi.simplify(idx, pos, nrm, target)
        np.testing.assert_array_equal(got, np.array(want, dtype=np.uint32))
        assert got_err == want_err and got_err.dtype == np.float32
        assert len(got) < len(idx)


def test_simplify_error_limit_and_trivial_targets(pys):
    pos, nrm, _, idx, _ = torus(24, 16)
    same, err = capi.simplify(idx, pos, nrm, len(idx))
    np.testing.assert_array_equal(same, idx)
    assert err == 0.0
    # a finite target error stops the passes early, identically in both implementations
    for limit in (1e-3, 4e-3):
        want, want_err = pys.simplify(idx, pos, nrm, 30, np.float32(limit))
        got, got_err = capi.simplify(idx, pos, nrm, 30, limit)
        np.testing.assert_array_equal(got, np.array(want, dtype=np.uint32))
        assert got_err == want_err and got_err <= np.float32(limit) and len(got) > 30
    lib = capi.load()
    dst = np.zeros(8, dtype=np.uint32)
    assert lib.oxb_simplify(capi._ptr(dst), capi._ptr(idx), 7, capi._ptr(pos), None, len(pos), 3, 1.0, None) < 0   # not a triangle list
    bad = np.array([0, 1, len(pos)], dtype=np.uint32)
    assert lib.oxb_simplify(capi._ptr(dst), capi._ptr(bad), 3, capi._ptr(pos), None, len(pos), 0, 1.0, None) < 0   # index out of range


def test_simplify_random_soups_match_oracle(pys):
    """robustness: triangle soups with coincident positions, degenerate, duplicated and non-manifold triangles never crash and
    still agree with the oracle bit for bit (everything irregular is classified "locked")"""
    rng = np.random.default_rng(0)
    for it in range(60):
        V, T = int(rng.integers(4, 40)), int(rng.integers(1, 80))
        pos = rng.standard_normal((V, 3)).astype(np.float32) if it % 3 == 0 else rng.integers(0, 4, size=(V, 3)).astype(np.float32)
        nrm = rng.standard_normal((V, 3)).astype(np.float32) if it % 2 else None
        idx = rng.integers(0, V, size=T * 3).astype(np.uint32)
        target = int(rng.integers(0, T + 1)) * 3
        want, want_err = pys.simplify(idx, pos, nrm, target)
        got, got_err = capi.simplify(idx, pos, nrm, target)
        np.testing.assert_array_equal(got, np.array(want, dtype=np.uint32))
        assert got_err == want_err


# ------------------------------------------------------------------------------------------------ properties
def test_closed_surfaces_stay_closed_and_unflipped():
    pos, nrm, _, idx, _ = torus(48, 24)
    errs = []
    for div in (2, 4, 8, 16):
        out, err = capi.simplify(idx, pos, nrm, len(idx) // 3 // div * 3)
        errs.append(float(err))
        assert set(out.tolist()) <= set(idx.tolist())                          # existing vertices only
        assert len(out) <= 1.5 * (len(idx) // 3 // div * 3)                     # gets (close to) the target
        assert is_closed_manifold(out, pos)                                     # no holes, no fins
        t = out.reshape(-1, 3)
        n = np.cross(pos[t[:, 1]] - pos[t[:, 0]], pos[t[:, 2]] - pos[t[:, 0]]).astype(np.float64)
        m = nrm[t].sum(axis=1).astype(np.float64)
        facing = np.einsum("ij,ij->i", n, m) / (np.linalg.norm(n, axis=1) * np.linalg.norm(m, axis=1))
        # no folds: a face never turns more than acos(0.25) away from its original normal (the surface is noisy: the shading
        # normals of its corners are a looser reference than that)
        assert facing.min() > -0.25 and (facing > 0.2).mean() > 0.98
        # the reported error is the relative deviation: every ORIGINAL vertex is within a small multiple of it from the new surface
        used = np.unique(idx)
        d = point_triangle_distance(pos[used].astype(np.float64), *(pos[t[:, k]].astype(np.float64) for k in range(3))).min(axis=1)
        extent = float((pos[5:].max(axis=0) - pos[5:].min(axis=0)).max())       # (the 5 pad vertices at 99 are part of the buffer ...
        full_extent = float((pos.max(axis=0) - pos.min(axis=0)).max())          #  ... and therefore of the simplifier's scale)
        assert d.max() <= 4.0 * float(err) * full_extent + 1e-6 and d.max() < 0.2 * extent
    assert errs == sorted(errs) and 0 < errs[0] and errs[-1] < 0.5             # coarser target, larger error


def test_borders_are_locked():
    pos, nrm, idx = open_grid(20, bump=0.15)
    before = border_edges(idx, pos)
    out, err = capi.simplify(idx, pos, nrm, len(idx) // 4 // 3 * 3)
    assert border_edges(out, pos) == before and len(before) == 80              # meshopt_SimplifyLockBorder: every border edge survives
    assert len(out) <= 1.5 * (len(idx) // 4 // 3 * 3)
    # a flat patch simplifies for free
    pos, nrm, idx = open_grid(20, bump=0.0)
    out, err = capi.simplify(idx, pos, nrm, len(idx) // 6 // 3 * 3)
    assert err < 1e-6 and border_edges(out, pos) == border_edges(idx, pos)
    t = out.reshape(-1, 3)
    area = 0.5 * np.linalg.norm(np.cross(pos[t[:, 1]] - pos[t[:, 0]], pos[t[:, 2]] - pos[t[:, 0]]), axis=1).sum()
    assert abs(area - 1.0) < 1e-5                                               # still covers the unit square exactly once


def test_attribute_seams_stay_closed():
    for pos, nrm, idx in (seam_torus(32, 16), hard_cube(8)):
        assert is_closed_manifold(idx, pos)
        out, err = capi.simplify(idx, pos, nrm, len(idx) // 4 // 3 * 3)
        assert len(out) < 0.5 * len(idx)
        assert is_closed_manifold(out, pos)                                     # both wedges of a seam vertex collapse together: no crack
        t = out.reshape(-1, 3)
        if len(pos) == 6 * 81:
            # cube: wedges are never mixed — the three corner normals of every triangle are still its face normal
            n = np.cross(pos[t[:, 1]] - pos[t[:, 0]], pos[t[:, 2]] - pos[t[:, 0]]).astype(np.float64)
            n /= np.linalg.norm(n, axis=1, keepdims=True)
            assert np.einsum("ij,ikj->ik", n, nrm[t].astype(np.float64)).min() > 0.999
        else:
            # torus: a triangle uses the seam ring's original vertices or their tilted copies, never both, and some of each survive
            ring, dup = set(range(5, 5 + 16)), set(range(len(pos) - 16, len(pos)))
            uses_ring = np.array([bool(ring & set(tri)) for tri in t.tolist()])
            uses_dup = np.array([bool(dup & set(tri)) for tri in t.tolist()])
            assert not np.any(uses_ring & uses_dup) and uses_ring.any() and uses_dup.any()
            # both wedges of a surviving seam position survive
            assert {v - 5 for v in ring & set(out.tolist())} == {v - (len(pos) - 16) for v in dup & set(out.tolist())}
    # the cube's faces are flat and its creases straight: everything but the 8 corners can go at zero error
    pos, nrm, idx = hard_cube(8)
    out, err = capi.simplify(idx, pos, nrm, 36)
    assert err < 1e-6 and len(out) <= 1.5 * 36 * 2
    ids = position_ids(pos)
    corners = {ids[v] for v in range(len(pos)) if np.all(np.abs(pos[v]) == 1.0)}
    assert corners <= set(ids[out].tolist()) and len(corners) == 8


# ------------------------------------------------------------------------------------------------ the LOD chain in the builder
def test_auto_lods_match_oracle_chain(pys):
    import pybuilder

    pos, nrm, uv, i0, _ = torus(32, 16)
    built = capi.BuiltMesh(pos, [(i0, 0.0)], normals=nrm, texcoords=uv, auto_lods=True)
    got = parse(built)
    want = pybuilder.build(pos, [(i0, 0.0)], normals=nrm, texcoords=uv, auto_lods=True)
    assert len(got["lods"]) == len(want["lods"]) >= 4
    for g, w in zip(got["lods"], want["lods"]):
        np.testing.assert_array_equal(g["indices"], w["indices"])
        np.testing.assert_array_equal(g["meshlets"], w["meshlets"])
        np.testing.assert_array_equal(g["micro"], w["micro"])
        np.testing.assert_array_equal(g["vertex_indices"], w["vertex_indices"])
        assert np.float32(g["error"]) == np.float32(w["error"])
    counts = [len(g["indices"]) for g in got["lods"]]
    errors = [g["error"] for g in got["lods"]]
    # AssetManager_GLTF.cpp:601,628-634: each LOD has (about) half the indices of the previous one, the error accumulates
    for a, b in zip(counts, counts[1:]):
        target = (a + 5) // 6 * 3
        assert 6 <= b <= target + target // 2
    assert errors[0] == 0.0 and all(0 < b - a <= 0.5 for a, b in zip(errors, errors[1:]))   # a step above 0.5 ends the chain
    assert int(got["mesh"]["lod_count"]) == len(counts) <= 8
    built.close()
    # the chain uses the vertex buffer of LOD 0: no LOD references a vertex LOD 0 does not
    assert all(g["indices"].max() < got["vertex_count"] for g in got["lods"])
    with pytest.raises(capi.OxcError):
        capi.BuiltMesh(pos, [(i0, 0.0), (i0, 0.1)], normals=nrm, auto_lods=True)   # auto_lods takes LOD 0 only


def test_auto_lods_feed_the_cull(orc):
    """the generated chain through the oracle's cull_meshes: a far instance selects a generated coarse LOD"""
    pos, nrm, uv, i0, _ = torus(64, 32)
    built = capi.BuiltMesh(pos, [(i0, 0.0)], normals=nrm, texcoords=uv, auto_lods=True, spatial=True)
    xf = np.tile(np.eye(4, dtype=np.float32).T.reshape(16), (3, 1))
    xf[0, 12:15] = (0.0, 0.0, -8.0)
    xf[1, 12:15] = (1.0, 1.0, -40.0)
    xf[2, 12:15] = (-2.0, -1.0, -200.0)
    sc = capi.assemble_scene([built], [0, 0, 0], xf, 320, 180)
    hs = orc.HostScene(sc)
    cam = sc.camera()
    mask = np.zeros((sc.max_meshlet_instance_count + 31) // 32, dtype=np.uint32)
    r = orc.frame(hs, cam, sc.width, sc.height, mask, None)
    lods = list(hs.mesh_instances["lod_index"])
    assert lods[0] == 0 and lods[0] <= lods[1] <= lods[2] and lods[2] >= 2
    assert r["late"] > 10


def test_golden_builder_fixture():
    """tests/golden/builder_small.json (written by make_golden_builder.py from the Python oracles) still describes what the
    oracles AND the product build today: regression pin for the whole builder (remap, quantisation, simplification chain,
    meshlets, bounds)."""
    import hashlib
    import importlib.util
    import json
    import os

    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_golden_builder", os.path.join(here, "golden", "make_golden_builder.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    want = json.load(open(os.path.join(here, "golden", "builder_small.json")))
    assert mod.generate() == want                                               # the oracles
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()  # noqa: E731
    pos, nrm, uv, i0, _ = torus(want["mesh"]["nu"], want["mesh"]["nv"])
    built = capi.BuiltMesh(pos, [(i0, 0.0)], normals=nrm, texcoords=uv, auto_lods=True)
    got = parse(built)                                                          # the product
    assert got["vertex_count"] == want["vertex_count"] and sha(got["positions_q"]) == want["positions_sha"] and sha(got["normals_q"]) == want["normals_sha"]
    assert len(got["lods"]) == len(want["lods"])
    for g, w in zip(got["lods"], want["lods"]):
        assert len(g["indices"]) == w["index_count"] and np.float32(g["error"]) == np.float32(w["error"]) and len(g["meshlets"]) == w["meshlets"]
        assert sha(g["indices"]) == w["indices_sha"] and sha(g["meshlets"]) == w["meshlets_sha"] and sha(g["micro"]) == w["micro_sha"]
        assert sha(g["vertex_indices"]) == w["vertex_indices_sha"]
        b = g["bounds"]
        rows = np.array([list(map(int, r["aabb_center"])) + list(map(int, r["cone_axis_xy"])) + list(map(int, r["aabb_extent"])) +
                         [int(r["cone_axis_z"]), int(r["cone_cutoff"])] for r in b], dtype=np.int64)
        assert sha(rows) == w["bounds_sha"]
    half, err = capi.simplify(i0, pos, None, len(i0) // 2 // 3 * 3)
    assert len(half) == want["positions_only_half"]["index_count"] and float(err) == want["positions_only_half"]["error"]
    assert sha(half) == want["positions_only_half"]["indices_sha"]
    built.close()
