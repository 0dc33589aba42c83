"""ORACLE (test infrastructure only) for the mesh builder — an independent numpy / pure-Python restatement of
build_gltf_mesh (Oxylus/src/Asset/AssetManager_GLTF.cpp:481-771) with the same substitutions the product makes for
the meshoptimizer v1.2 calls (xmake/packages.lua:9; library absent from /root/reference and from this image):
first-use fetch remap, meshopt_quantizeHalf / quantizeSnorm as published, Ritter-sphere normal cone as in
meshopt_computeMeshletBounds, linear-scan meshlets instead of meshopt_buildMeshlets, the LOD chain of :596-641 over the
edge-collapse simplifier restated in oracle/pysimplify.py.

PARITY UNPINNED: nothing in the reference's tests pins meshlet build output, and meshoptimizer's own clustering is
not reproduced.  Every float op below is an explicit np.float32 operation in the product's order (small inputs only:
pure-Python loops)."""
import numpy as np

F = np.float32
NONE = 0xFFFFFFFF


def quantize_half(v):
    """meshopt_quantizeHalf: vectorised over an f32 array"""
    ui = np.asarray(v, dtype=np.float32).view(np.uint32).astype(np.int64)
    s = (ui >> 16) & 0x8000
    em = ui & 0x7FFFFFFF
    h = (em - (112 << 23) + (1 << 12)) >> 13
    h = np.where(em < (113 << 23), 0, h)
    h = np.where(em >= (143 << 23), 0x7C00, h)
    h = np.where(em > (255 << 23), 0x7E00, h)
    return (s | h).astype(np.uint16)


def quantize_snorm(v, bits):
    v = F(v)
    scale = F((1 << (bits - 1)) - 1)
    rnd = F(0.5) if v >= 0 else F(-0.5)
    v = v if v >= F(-1) else F(-1)
    v = v if v <= F(1) else F(1)
    return int(F(F(v * scale) + rnd))  # C float -> int conversion truncates toward zero, as int() does


def bounding_sphere(pts):
    """Ritter's sphere over an (n, 3) f32 array; returns (centre[3], radius) in f32, op for op as the product."""
    n = len(pts)
    pmin, pmax = [0, 0, 0], [0, 0, 0]
    for i in range(n):
        for a in range(3):
            if pts[i][a] < pts[pmin[a]][a]:
                pmin[a] = i
            if pts[i][a] > pts[pmax[a]][a]:
                pmax[a] = i
    best, axis = F(0), 0
    for a in range(3):
        p1, p2 = pts[pmin[a]], pts[pmax[a]]
        d = [F(p2[k] - p1[k]) for k in range(3)]
        d2 = F(F(F(d[0] * d[0]) + F(d[1] * d[1])) + F(d[2] * d[2]))
        if d2 > best:
            best, axis = d2, a
    p1, p2 = pts[pmin[axis]], pts[pmax[axis]]
    c = [F(F(p1[k] + p2[k]) / F(2)) for k in range(3)]
    r = F(np.sqrt(best) / F(2))
    for i in range(n):
        p = pts[i]
        d = [F(p[k] - c[k]) for k in range(3)]
        d2 = F(F(F(d[0] * d[0]) + F(d[1] * d[1])) + F(d[2] * d[2]))
        if d2 > F(r * r):
            dd = F(np.sqrt(d2))
            k = F(F(0.5) + F(F(r / dd) / F(2)))
            omk = F(F(1) - k)
            c = [F(F(c[j] * k) + F(p[j] * omk)) for j in range(3)]
            r = F(F(r + dd) / F(2))
    return c, r


def build(positions, lods, normals=None, texcoords=None, max_vertices=64, max_triangles=64, auto_lods=False):
    """Returns dict(vertex_count, positions_q (V,4) u16, normals_q (V,) u32 | None, texcoords_q (V,2) u16 | None,
    remap, bounds_center/extent (f32 mesh bounds), lods=[dict(indices, meshlets (M,4) u32, bounds (M,) records as tuples,
    micro u8, vertex_indices u32, error)])."""
    pos_in = np.asarray(positions, dtype=np.float32).reshape(-1, 3)
    remap = np.full(len(pos_in), NONE, dtype=np.uint32)
    vc = 0
    for i in np.asarray(lods[0][0], dtype=np.uint32).reshape(-1):
        if remap[i] == NONE:
            remap[i] = vc
            vc += 1
    used = remap != NONE
    pos = np.zeros((vc, 3), dtype=np.float32)
    pos[remap[used]] = pos_in[used]
    out = dict(vertex_count=vc, remap=remap)
    pq = np.zeros((vc, 4), dtype=np.uint16)
    pq[:, :3] = quantize_half(pos)
    out["positions_q"] = pq
    out["normals_q"] = None
    if normals is not None:
        nin = np.asarray(normals, dtype=np.float32).reshape(-1, 3)
        nq = np.zeros(vc, dtype=np.uint32)
        for v in np.nonzero(used)[0]:
            x, y, z = (quantize_snorm(nin[v][k], 10) + 511 for k in range(3))
            nq[remap[v]] = (x << 20) | (y << 10) | z
        out["normals_q"] = nq
    out["texcoords_q"] = None
    if texcoords is not None:
        tin = np.asarray(texcoords, dtype=np.float32).reshape(-1, 2)
        tq = np.zeros((vc, 2), dtype=np.uint16)
        tq[remap[used]] = quantize_half(tin[used])
        out["texcoords_q"] = tq
    fmax, flow = np.finfo(np.float32).max, np.finfo(np.float32).min
    mesh_min, mesh_max = np.full(3, fmax, dtype=np.float32), np.full(3, flow, dtype=np.float32)
    out["lods"] = []
    lods = [(remap[np.asarray(idx_in, dtype=np.uint32).reshape(-1)], err) for (idx_in, err) in lods]
    if auto_lods:  # AssetManager_GLTF.cpp:596-641 in blob vertex numbering
        import pysimplify

        assert len(lods) == 1
        nrm_blob = None
        if normals is not None:
            nrm_blob = np.zeros((vc, 3), dtype=np.float32)
            nrm_blob[remap[used]] = np.asarray(normals, dtype=np.float32).reshape(-1, 3)[used]
        lods = [(np.array(i, dtype=np.uint32), e) for (i, e) in pysimplify.lod_chain(lods[0][0], pos, nrm_blob, error0=lods[0][1])]
    for l, (indices, err) in enumerate(lods):
        meshlets, vertex_indices, micro = [], [], []
        slot = {}
        cur = [0, 0, 0, 0]  # vertex_offset, triangle_offset, vertex_count, triangle_count

        def close():
            nonlocal cur, slot
            if cur[3] == 0:
                return
            want = cur[1] + ((cur[3] * 3 + 3) & ~3)
            micro.extend([0] * (want - len(micro)))
            meshlets.append(tuple(cur))
            slot = {}
            cur = [len(vertex_indices), len(micro), 0, 0]

        for t in range(0, len(indices) - 2, 3):
            tri = [int(indices[t]), int(indices[t + 1]), int(indices[t + 2])]
            extra = len({v for v in tri if v not in slot})
            if cur[2] + extra > max_vertices or cur[3] >= max_triangles:
                close()
            for v in tri:
                if v not in slot:
                    slot[v] = cur[2]
                    cur[2] += 1
                    vertex_indices.append(v)
                micro.append(slot[v])
            cur[3] += 1
        close()
        if not meshlets:
            break
        bounds = []
        for (vo, to, _, tc) in meshlets:
            corners = np.array([vertex_indices[vo + micro[to + k]] for k in range(tc * 3)], dtype=np.int64)
            p = pos[corners]  # (3*tc, 3) f32
            bmin, bmax = p.min(axis=0), p.max(axis=0)
            nrm = []
            for t in range(tc):
                p0, p1, p2 = p[3 * t], p[3 * t + 1], p[3 * t + 2]
                e1 = [F(p1[k] - p0[k]) for k in range(3)]
                e2 = [F(p2[k] - p0[k]) for k in range(3)]
                nx = F(F(e1[1] * e2[2]) - F(e1[2] * e2[1]))
                ny = F(F(e1[2] * e2[0]) - F(e1[0] * e2[2]))
                nz = F(F(e1[0] * e2[1]) - F(e1[1] * e2[0]))
                area = F(np.sqrt(F(F(F(nx * nx) + F(ny * ny)) + F(nz * nz))))
                if area == 0:
                    continue
                nrm.append([F(nx / area), F(ny / area), F(nz / area)])
            axis_s8, cutoff_s8 = [0, 0, 0], 127
            if nrm:
                c, _ = bounding_sphere(nrm)
                ln = F(np.sqrt(F(F(F(c[0] * c[0]) + F(c[1] * c[1])) + F(c[2] * c[2]))))
                inv = F(0) if ln == 0 else F(F(1) / ln)
                axis = [F(c[k] * inv) for k in range(3)]
                mindp = F(1)
                for n_ in nrm:
                    dp = F(F(F(n_[0] * axis[0]) + F(n_[1] * axis[1])) + F(n_[2] * axis[2]))
                    if dp < mindp:
                        mindp = dp
                if mindp > F(0.1):
                    cutoff = F(np.sqrt(F(F(1) - F(mindp * mindp))))
                    e = F(0)
                    for k in range(3):
                        axis_s8[k] = quantize_snorm(axis[k], 8)
                        e = F(e + np.abs(F(F(F(axis_s8[k]) / F(127)) - axis[k])))
                    cutoff_s8 = min(127, int(F(F(F(127) * F(cutoff + e)) + F(1))))
            center = quantize_half((bmax + bmin) * F(0.5))
            extent = quantize_half(bmax - bmin)
            bounds.append((tuple(int(x) for x in center), tuple(axis_s8[:2]), tuple(int(x) for x in extent), axis_s8[2], cutoff_s8))
            if l == 0:
                mesh_min, mesh_max = np.minimum(mesh_min, bmin), np.maximum(mesh_max, bmax)
        out["lods"].append(dict(indices=indices.astype(np.uint32), meshlets=np.array(meshlets, dtype=np.uint32),
                                bounds=bounds, micro=np.array(micro, dtype=np.uint8),
                                vertex_indices=np.array(vertex_indices, dtype=np.uint32), error=float(np.float32(err))))
    out["bounds_center"] = (mesh_max + mesh_min) * F(0.5)
    out["bounds_extent"] = mesh_max - mesh_min
    return out
