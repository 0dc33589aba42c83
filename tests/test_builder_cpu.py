"""Mesh builder (host C++ in liboxcull.so, include/oxcull.h oxb_*) vs its independent Python oracle
(oracle/pybuilder.py): identical blobs field by field, plus the properties a cull relies on.  Runs without a GPU."""
import numpy as np
import pytest

from oxylus_b200 import abi, capi


def torus(nu, nv, R=2.0, r=0.7, seed=3):
    """(positions, normals, uvs, LOD0 indices, LOD1 indices): nu x nv grid torus with a few unused and duplicated
    vertices; LOD1 re-triangulates every second grid line (a vertex subset, as the simplifier would produce)."""
    u, v = np.meshgrid(np.arange(nu) / nu * 2 * np.pi, np.arange(nv) / nv * 2 * np.pi, indexing="ij")
    rng = np.random.default_rng(seed)
    rr = r * (1 + 0.05 * rng.standard_normal(u.shape))
    pos = np.stack([(R + rr * np.cos(v)) * np.cos(u), (R + rr * np.cos(v)) * np.sin(u), rr * np.sin(v)], axis=2).reshape(-1, 3)
    nrm = np.stack([np.cos(v) * np.cos(u), np.cos(v) * np.sin(u), np.sin(v)], axis=2).reshape(-1, 3)
    uv = np.stack([u / (2 * np.pi), v / (2 * np.pi)], axis=2).reshape(-1, 2)

    def grid_indices(step):
        tris = []
        for i in range(0, nu, step):
            for j in range(0, nv, step):
                a, b = i * nv + j, ((i + step) % nu) * nv + j
                c, d = i * nv + (j + step) % nv, ((i + step) % nu) * nv + (j + step) % nv
                tris += [(a, b, d), (a, d, c)]
        return np.array(tris, dtype=np.uint32).reshape(-1)

    # prepend 5 unused vertices so the fetch remap has something to drop / renumber
    pad = 5
    pos = np.concatenate([np.full((pad, 3), 99.0), pos]).astype(np.float32)
    nrm = np.concatenate([np.zeros((pad, 3)), nrm]).astype(np.float32)
    uv = np.concatenate([np.zeros((pad, 2)), uv]).astype(np.float32)
    return pos, nrm, uv, grid_indices(1) + pad, grid_indices(2) + pad


def parse(built):
    """blob -> the same structure oracle/pybuilder.build returns"""
    blob = np.zeros(built.blob_size, dtype=np.uint8)
    mesh = built.emit(0, blob)[0]
    vc = int(mesh["vertex_count"])
    out = dict(vertex_count=vc, mesh=mesh, blob=blob)
    out["positions_q"] = np.frombuffer(blob, dtype=np.uint16, count=vc * 4, offset=int(mesh["vertex_positions"])).reshape(vc, 4)
    out["normals_q"] = np.frombuffer(blob, dtype=np.uint32, count=vc, offset=int(mesh["vertex_normals"])) if mesh["vertex_normals"] else None
    out["texcoords_q"] = (np.frombuffer(blob, dtype=np.uint16, count=vc * 2, offset=int(mesh["texture_coords"])).reshape(vc, 2)
                          if mesh["texture_coords"] else None)
    lods = np.frombuffer(blob, dtype=abi.MESH_LOD_DT, count=int(mesh["lod_count"]), offset=int(mesh["lods"]))
    out["lods"] = []
    for d in lods:
        mc = int(d["meshlet_count"])
        out["lods"].append(dict(
            indices=np.frombuffer(blob, dtype=np.uint32, count=int(d["indices_count"]), offset=int(d["indices"])),
            meshlets=np.frombuffer(blob, dtype=np.uint32, count=mc * 4, offset=int(d["meshlets"])).reshape(mc, 4),
            bounds=np.frombuffer(blob, dtype=abi.MESHLET_BOUNDS_DT, count=mc, offset=int(d["meshlet_bounds"])),
            micro=np.frombuffer(blob, dtype=np.uint8, count=int(d["local_triangle_indices_count"]), offset=int(d["local_triangle_indices"])),
            vertex_indices=np.frombuffer(blob, dtype=np.uint32, count=int(d["indirect_vertex_indices_count"]),
                                         offset=int(d["indirect_vertex_indices"])),
            error=float(d["error"]), rec=d))
    return out


@pytest.fixture(scope="module")
def pyb():
    import pybuilder

    return pybuilder


def test_quantizers_known_answers(pyb):
    q = pyb.quantize_half(np.array([0.0, 1.0, -2.0, 65504.0, 1e6, 6e-5, 6.103515625e-05, 1.0009765625, 1.00048828125, np.nan], dtype=np.float32))
    # 6e-5 < 2^-14 = 6.1035e-5 flushes to zero; 2^-14 is the smallest normal; ties round up (add-then-truncate)
    assert list(q[:6]) == [0x0000, 0x3C00, 0xC000, 0x7BFF, 0x7C00, 0x0000]
    assert q[6] == 0x0400 and q[7] == 0x3C01 and q[8] == 0x3C01 and q[9] == 0x7E00
    # agrees with IEEE round-to-nearest-even everywhere except exact ties and the denormal range
    x = np.random.default_rng(1).standard_normal(20000).astype(np.float32) * 8
    assert np.array_equal(pyb.quantize_half(x), x.astype(np.float16).view(np.uint16))
    assert [pyb.quantize_snorm(v, 8) for v in (0.0, 1.0, -1.0, 0.5, -0.5, 2.0, 0.00394)] == [0, 127, -127, 64, -64, 127, 1]
    assert pyb.quantize_snorm(1.0, 10) + 511 == 1022 and pyb.quantize_snorm(-1.0, 10) + 511 == 0


def test_builder_matches_oracle(pyb):
    pos, nrm, uv, i0, i1 = torus(48, 24)
    built = capi.BuiltMesh(pos, [(i0, 0.0), (i1, 0.02)], normals=nrm, texcoords=uv)
    got = parse(built)
    want = pyb.build(pos, [(i0, 0.0), (i1, 0.02)], normals=nrm, texcoords=uv)
    assert got["vertex_count"] == want["vertex_count"] == 48 * 24
    np.testing.assert_array_equal(got["positions_q"], want["positions_q"])
    np.testing.assert_array_equal(got["normals_q"], want["normals_q"])
    np.testing.assert_array_equal(got["texcoords_q"], want["texcoords_q"])
    np.testing.assert_array_equal(got["mesh"]["bounds"]["aabb_center"], want["bounds_center"])
    np.testing.assert_array_equal(got["mesh"]["bounds"]["aabb_extent"], want["bounds_extent"])
    assert len(got["lods"]) == len(want["lods"]) == 2
    for g, w in zip(got["lods"], want["lods"]):
        np.testing.assert_array_equal(g["indices"], w["indices"])
        np.testing.assert_array_equal(g["meshlets"], w["meshlets"])
        np.testing.assert_array_equal(g["micro"], w["micro"])
        np.testing.assert_array_equal(g["vertex_indices"], w["vertex_indices"])
        assert g["error"] == w["error"]
        for b, (c, axy, e, az, cut) in zip(g["bounds"], w["bounds"]):
            assert (tuple(b["aabb_center"]), tuple(b["cone_axis_xy"]), tuple(b["aabb_extent"]), int(b["cone_axis_z"]), int(b["cone_cutoff"])) == \
                   (c, axy, e, az, cut)
    assert built.lod0_meshlet_count == len(want["lods"][0]["meshlets"])
    # layout: everything the kernels load with 128-bit accesses is 16-byte aligned (oxc_set_scene checks the same)
    for g in got["lods"]:
        assert g["rec"]["meshlets"] % 16 == 0 and g["rec"]["meshlet_bounds"] % 16 == 0
    assert built.blob_size % 16 == 0


def test_builder_properties():
    pos, nrm, uv, i0, i1 = torus(64, 32)
    built = capi.BuiltMesh(pos, [(i0, 0.0), (i1, 0.05)], normals=nrm)
    got = parse(built)
    assert got["texcoords_q"] is None and got["mesh"]["texture_coords"] == 0
    posq = got["positions_q"][:, :3].copy().view(np.float16).astype(np.float64)
    for lod, src in zip(got["lods"], (i0, i1)):
        ml, micro, vi = lod["meshlets"], lod["micro"], lod["vertex_indices"]
        assert ml[:, 2].max() <= 64 and ml[:, 3].max() <= 64 and ml[:, 3].min() >= 1
        # every input triangle exactly once, in order
        tris = np.concatenate([vi[vo + micro[to: to + 3 * tc].astype(np.int64)] for vo, to, vc, tc in ml])
        np.testing.assert_array_equal(tris, lod["indices"])
        assert len(tris) == len(src)
        # micro-index runs are padded to 4 bytes and never overlap; vertex runs are unique within a meshlet
        assert np.all(ml[:, 1] % 4 == 0)
        for vo, to, vc, tc in ml:
            assert len(np.unique(vi[vo: vo + vc])) == vc and micro[to: to + 3 * tc].max() < vc
        # decoded AABB (centre +- extent/2, half precision) contains every vertex of the meshlet up to half rounding
        b = lod["bounds"]
        c = b["aabb_center"].copy().view(np.float16).astype(np.float64)
        e = b["aabb_extent"].copy().view(np.float16).astype(np.float64)
        for k, (vo, to, vc, tc) in enumerate(ml):
            p = posq[vi[vo: vo + vc]]
            tol = 2e-3 * (np.abs(c[k]) + e[k]) + 1e-6
            assert np.all(p >= c[k] - e[k] / 2 - tol) and np.all(p <= c[k] + e[k] / 2 + tol)
        # the s8 cone is conservative: no camera position that sees a front face of the meshlet is culled by
        # dot(center - cam, axis) >= cutoff * |center - cam| + radius  (cull.slang:173-175)
        rng = np.random.default_rng(5)
        cams = rng.standard_normal((64, 3)) * 6
        for k, (vo, to, vc, tc) in enumerate(ml[:: max(1, len(ml) // 40)]):
            k = k * max(1, len(ml) // 40)
            cut = int(b["cone_cutoff"][k])
            if cut >= 127:
                continue
            axis = np.array([b["cone_axis_xy"][k][0], b["cone_axis_xy"][k][1], b["cone_axis_z"][k]], dtype=np.float64) / 127.0
            radius = np.linalg.norm(e[k] / 2)
            loc = micro[to: to + 3 * tc].astype(np.int64).reshape(-1, 3)
            tp = posq[vi[vo + loc]]
            n = np.cross(tp[:, 1] - tp[:, 0], tp[:, 2] - tp[:, 0])
            for cam in cams:
                d = c[k] - cam
                culled = np.dot(d, axis) >= cut / 127.0 * np.linalg.norm(d) + radius
                if culled:
                    # culled => every triangle faces away from the camera (n . (p0 - cam) >= 0)
                    assert np.all(np.einsum("ij,ij->i", n, tp[:, 0] - cam) >= -1e-9)


def test_builder_errors_and_degenerates():
    pos = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [5, 5, 5]], dtype=np.float32)
    with pytest.raises(capi.OxcError):
        capi.BuiltMesh(pos, [(np.array([0, 1, 9], dtype=np.uint32), 0.0)])       # index out of range
    with pytest.raises(capi.OxcError):
        capi.BuiltMesh(pos, [(np.array([0, 1], dtype=np.uint32), 0.0)])          # not a triangle list
    with pytest.raises(capi.OxcError):
        capi.BuiltMesh(pos, [(np.array([0, 1, 2], dtype=np.uint32), 0.0), (np.array([0, 1, 3], dtype=np.uint32), 0.1)])  # LOD1 uses a new vertex
    # a degenerate (zero-area) triangle next to a real one: cone from the real one only; a fully degenerate meshlet
    # gets the cone disabled
    b = capi.BuiltMesh(pos, [(np.array([0, 1, 2, 0, 0, 1], dtype=np.uint32), 0.0)])
    g = parse(b)
    assert g["vertex_count"] == 3 and len(g["lods"][0]["meshlets"]) == 1 and g["lods"][0]["meshlets"][0][3] == 2
    bd = g["lods"][0]["bounds"][0]
    assert (int(bd["cone_axis_xy"][0]), int(bd["cone_axis_xy"][1]), int(bd["cone_axis_z"])) == (0, 0, 127) and int(bd["cone_cutoff"]) == 1
    b = capi.BuiltMesh(pos, [(np.array([0, 0, 1], dtype=np.uint32), 0.0)])
    assert int(parse(b)["lods"][0]["bounds"][0]["cone_cutoff"]) == 127


def test_built_scene_runs_through_the_oracle_pipeline(orc):
    """builder output -> scene tables -> the oracle's two-pass frame: the torus is visible and decodes"""
    pos, nrm, uv, i0, i1 = torus(96, 48)
    built = capi.BuiltMesh(pos, [(i0, 0.0), (i1, 0.05)], normals=nrm, texcoords=uv)
    xf = np.tile(np.eye(4, dtype=np.float32).T.reshape(16), (3, 1))
    xf[0, 12:15] = (0.0, 0.0, -8.0)
    xf[1, 12:15] = (3.0, 1.0, -14.0)
    xf[2, 12:15] = (-2.0, -1.0, -60.0)   # far: picks the coarse LOD
    sc = capi.assemble_scene([built], [0, 0, 0], xf, 320, 180)
    hs = orc.HostScene(sc)
    cam = sc.camera()
    mask = np.zeros((sc.max_meshlet_instance_count + 31) // 32, dtype=np.uint32)
    r = orc.frame(hs, cam, sc.width, sc.height, mask, None)
    assert r["late"] > 10 and r["ntri_late"] > 500
    assert list(hs.mesh_instances["lod_index"]) == [0, 0, 1]
    v32, depth = orc.resolve(r["vis64"])
    covered = v32 != 0xFFFFFFFF
    assert covered.sum() > 2000
    d = orc.decode_visbuffer(hs, r["meshlet_instances"], int(r["visibility"]["total"][0]), cam, v32)
    assert np.array_equal(d["lambda_"][:, :, 3] == 1.0, covered)
    uvs = d["uv_normal"][covered][:, :2]
    assert uvs.min() > -0.05 and uvs.max() < 1.05


def _lod0_meshlet_triangles(got):
    """per-meshlet triangle arrays (blob vertex numbering) of LOD 0 from parse()'s structure"""
    lod = got["lods"][0]
    ml, micro, vi = lod["meshlets"], lod["micro"], lod["vertex_indices"]
    return [vi[vo + micro[to: to + 3 * tc].astype(np.int64)].reshape(-1, 3) for vo, to, vc, tc in ml]


def test_spatial_clusteriser_properties():
    """cluster_mode 1 (the role of meshopt_buildMeshlets): on a torus whose triangles arrive in a SHUFFLED order the spatial
    clusteriser covers every triangle exactly once within the 64 / 64 limits, fills its meshlets and makes them compact — the
    caller-order scan on the same input produces many more, scattered meshlets.  Deterministic."""
    pos, nrm, uv, i0, _ = torus(64, 32)
    rng = np.random.default_rng(3)
    shuffled = i0.reshape(-1, 3)[rng.permutation(len(i0) // 3)].reshape(-1)
    spatial = capi.BuiltMesh(pos, [(shuffled, 0.0)], normals=nrm, texcoords=uv, spatial=True)
    linear = capi.BuiltMesh(pos, [(shuffled, 0.0)], normals=nrm, texcoords=uv, spatial=False)
    again = capi.BuiltMesh(pos, [(shuffled, 0.0)], normals=nrm, texcoords=uv, spatial=True)
    gs, gl, ga = parse(spatial), parse(linear), parse(again)
    assert np.array_equal(gs["blob"], ga["blob"])  # deterministic
    ts, tl = _lod0_meshlet_triangles(gs), _lod0_meshlet_triangles(gl)
    n_tri = len(shuffled) // 3
    ps = gs["positions_q"][:, :3].copy().view(np.float16).astype(np.float64)
    pl = gl["positions_q"][:, :3].copy().view(np.float16).astype(np.float64)

    def canon(tris, p):  # triangles as position triples: the vertex numbering is the same (fetch remap precedes clustering)
        allt = np.concatenate(tris)
        return {tuple(sorted(map(tuple, p[t]))) for t in allt}

    assert sum(len(t) for t in ts) == n_tri and canon(ts, ps) == canon(tl, pl) and len(canon(ts, ps)) == n_tri
    for t in ts:
        assert 1 <= len(t) <= 64 and len(np.unique(t)) <= 64

    def radius(tris, p):
        return float(np.mean([np.linalg.norm(p[np.unique(t)] - p[np.unique(t)].mean(axis=0), axis=1).max() for t in tris]))

    assert len(ts) <= 1.3 * np.ceil(n_tri / 64)              # nearly full meshlets
    assert len(tl) >= 2 * len(ts)                            # the shuffled scan runs out of vertices long before 64 triangles
    assert radius(ts, ps) < 0.35 * radius(tl, pl)            # ... and its meshlets span the whole mesh
    for b in (spatial, linear, again):
        b.close()


def test_spatial_clusteriser_equals_scan_of_its_own_order(pyb):
    """"order + scan" is the whole definition of cluster_mode 1: feeding the triangle order it produced back through the
    caller-order path gives the same meshlets, and the independent Python builder reproduces that blob bit for bit."""
    pos, nrm, uv, i0, _ = torus(24, 20)
    rng = np.random.default_rng(5)
    shuffled = i0.reshape(-1, 3)[rng.permutation(len(i0) // 3)].reshape(-1)
    spatial = capi.BuiltMesh(pos, [(shuffled, 0.0)], normals=nrm, texcoords=uv, spatial=True)
    gs = parse(spatial)
    ts = _lod0_meshlet_triangles(gs)
    # blob vertex numbering == order of first use in the shuffled buffer: map the clustered order back to input numbering
    first_use = {}
    for v in shuffled:
        first_use.setdefault(int(v), len(first_use))
    inverse = np.zeros(len(first_use), dtype=np.uint32)
    for v, r in first_use.items():
        inverse[r] = v
    ordered = inverse[np.concatenate(ts).reshape(-1)]
    again = capi.BuiltMesh(pos, [(ordered, 0.0)], normals=nrm, texcoords=uv, spatial=False)
    ga = parse(again)
    ta = _lod0_meshlet_triangles(ga)
    assert len(ta) == len(ts)
    ps = gs["positions_q"][:, :3]
    pa = ga["positions_q"][:, :3]
    for a, b in zip(ta, ts):
        np.testing.assert_array_equal(pa[a], ps[b])          # same triangles, same corner order, meshlet by meshlet
    want = pyb.build(pos, [(ordered, 0.0)], normals=nrm, texcoords=uv)
    np.testing.assert_array_equal(ga["lods"][0]["meshlets"], want["lods"][0]["meshlets"])
    np.testing.assert_array_equal(ga["lods"][0]["micro"], want["lods"][0]["micro"])
    np.testing.assert_array_equal(ga["lods"][0]["vertex_indices"], want["lods"][0]["vertex_indices"])
    spatial.close(); again.close()
