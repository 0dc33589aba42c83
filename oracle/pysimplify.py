"""ORACLE (test infrastructure only) for the mesh builder's LOD simplifier — an independent pure-Python restatement of
oxb_simplify (oxylus_b200/csrc/host/mesh_simplifier.cpp), the step build_gltf_mesh delegates to
meshopt_simplifyWithAttributes (Oxylus/src/Asset/AssetManager_GLTF.cpp:604-637: normals as attributes with weight 1,
meshopt_SimplifyLockBorder, target error FLT_MAX, target index count = half, relative result error).

meshoptimizer v1.2 (xmake/packages.lua:9) is neither under /root/reference nor installed: the scheme below is the library's
published one — edge collapses onto EXISTING vertices, area-weighted plane quadrics per position, attribute quadrics with
per-triangle gradients per wedge, seam-edge quadrics, locked borders, collapses ranked once per pass (position + attribute
error) and applied in order with both endpoints locked for the rest of the pass, the reported error being the largest
POSITION error of a performed collapse (relative to the mesh extent: what MeshLOD::error feeds into the LOD selection of
cull_meshes.slang:35-57) — with two additions that make the result checkable: the link
condition (a 2-manifold stays a 2-manifold) and exact, stable ordering.  It is NOT bit-compatible with meshoptimizer.
PARITY UNPINNED.

Arithmetic: IEEE binary64, one rounding per operation, the operation order written here (Python floats; the product is
compiled with -ffp-contract=off), so product and oracle agree bit for bit.  Small inputs only: pure-Python loops."""
import math

import numpy as np

MANIFOLD, SEAM, LOCKED = 0, 1, 2
NONE = -1
SEAM_EDGE_WEIGHT = 1.0


def _sub(a, b):
    return (a[0] - b[0], a[1] - b[1], a[2] - b[2])


def _dot(a, b):
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]


def _cross(a, b):
    return (a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])


# quadric layout: a00 a11 a22 a10 a20 a21 b0 b1 b2 c w
def _q_plane(a, b, c, d, w):
    aw, bw, cw, dw = a * w, b * w, c * w, d * w
    return [a * aw, b * bw, c * cw, a * bw, a * cw, b * cw, a * dw, b * dw, c * dw, d * dw, w]


def _q_add(q, r):
    for i in range(11):
        q[i] = q[i] + r[i]


def _q_eval(q, p):
    x, y, z = p
    rx = q[0] * x + q[3] * y + q[4] * z
    ry = q[3] * x + q[1] * y + q[5] * z
    rz = q[4] * x + q[5] * y + q[2] * z
    r = rx * x + ry * y + rz * z
    r = r + 2.0 * (q[6] * x + q[7] * y + q[8] * z)
    return r + q[9]


def _q_triangle(p0, p1, p2):
    n = _cross(_sub(p1, p0), _sub(p2, p0))
    ln = math.sqrt(_dot(n, n))
    if ln == 0.0:
        return None
    n = (n[0] / ln, n[1] / ln, n[2] / ln)
    return _q_plane(n[0], n[1], n[2], -_dot(n, p0), math.sqrt(ln))


def _q_seam_edge(p0, p1, p2):
    """plane through the edge p0-p1, perpendicular to the triangle: keeps the seam line in place"""
    e = _sub(p1, p0)
    ee = _dot(e, e)
    if ee == 0.0:
        return None
    f = _sub(p2, p0)
    t = _dot(f, e) / ee
    n = (f[0] - e[0] * t, f[1] - e[1] * t, f[2] - e[2] * t)
    ln = math.sqrt(_dot(n, n))
    if ln == 0.0:
        return None
    n = (n[0] / ln, n[1] / ln, n[2] / ln)
    return _q_plane(n[0], n[1], n[2], -_dot(n, p0), math.sqrt(ee) * SEAM_EDGE_WEIGHT)


def _q_attributes(p0, p1, p2, a0, a1, a2):
    """(quadric, gradients[k] = w * (gx, gy, gz, d)) of the attributes' linear interpolants over the triangle"""
    e1, e2 = _sub(p1, p0), _sub(p2, p0)
    n = _cross(e1, e2)
    ln = math.sqrt(_dot(n, n))
    w = math.sqrt(ln)
    d00, d01, d11 = _dot(e1, e1), _dot(e1, e2), _dot(e2, e2)
    den = d00 * d11 - d01 * d01
    inv = 0.0 if den == 0.0 else 1.0 / den
    q = [0.0] * 11
    grads = []
    for k in range(len(a0)):
        da1, da2 = a1[k] - a0[k], a2[k] - a0[k]
        c1 = (da1 * d11 - da2 * d01) * inv
        c2 = (da2 * d00 - da1 * d01) * inv
        g = (c1 * e1[0] + c2 * e2[0], c1 * e1[1] + c2 * e2[1], c1 * e1[2] + c2 * e2[2])
        d = a0[k] - _dot(g, p0)
        _q_add(q, _q_plane(g[0], g[1], g[2], d, w))
        grads.append([g[0] * w, g[1] * w, g[2] * w, d * w])
    q[10] = w  # one weight per triangle, not one per attribute
    return q, grads


def simplify(indices, positions, normals, target_index_count, target_error=3.4028234663852886e38):
    """-> (new index list, result_error as np.float32).  positions (V,3) f32, normals (V,3) f32 or None."""
    idx = [int(i) for i in np.asarray(indices).reshape(-1)]
    pos32 = np.asarray(positions, dtype=np.float32).reshape(-1, 3)
    V = len(pos32)
    nrm = None if normals is None else [tuple(float(x) for x in r) for r in np.asarray(normals, dtype=np.float32).reshape(-1, 3)]
    if len(idx) <= target_index_count:
        return idx, np.float32(0.0)

    # ---- positions rescaled into the unit cube (extent of ALL vertices, like the LOD chain's common scale) ----
    lo = [float(pos32[:, a].min()) for a in range(3)]
    hi = [float(pos32[:, a].max()) for a in range(3)]
    extent = max(hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2])
    inv = 0.0 if extent == 0.0 else 1.0 / extent
    P = [((float(p[0]) - lo[0]) * inv, (float(p[1]) - lo[1]) * inv, (float(p[2]) - lo[2]) * inv) for p in pos32]

    # ---- wedges: used vertices with bit-identical positions ----
    used = [False] * V
    for i in idx:
        used[i] = True
    remap, wedge = list(range(V)), list(range(V))
    first, last = {}, {}
    for v in range(V):
        if not used[v]:
            continue
        key = pos32[v].tobytes()
        if key in first:
            remap[v] = first[key]
            wedge[last[key]] = v
            wedge[v] = first[key]
        else:
            first[key] = v
        last[key] = v

    # ---- classification ----
    wedge_edges, pos_edges = {}, {}
    T = len(idx) // 3
    for t in range(T):
        for e in range(3):
            a, b = idx[3 * t + e], idx[3 * t + (e + 1) % 3]
            wedge_edges[(a, b)] = wedge_edges.get((a, b), 0) + 1
            pos_edges[(remap[a], remap[b])] = pos_edges.get((remap[a], remap[b]), 0) + 1
    loop, loopback = [NONE] * V, [NONE] * V
    outc, inc = [0] * V, [0] * V
    complex_, pborder = [False] * V, [False] * V
    seam_edges = []  # (i0, i1, i2): open at wedge level, closed at position level
    for t in range(T):
        for e in range(3):
            a, b, c = idx[3 * t + e], idx[3 * t + (e + 1) % 3], idx[3 * t + (e + 2) % 3]
            ra, rb = remap[a], remap[b]
            if ra == rb or pos_edges[(ra, rb)] > 1:
                complex_[ra] = complex_[rb] = True
            if (rb, ra) not in pos_edges:
                pborder[ra] = pborder[rb] = True
            elif (b, a) not in wedge_edges:
                seam_edges.append((a, b, c))
            if (b, a) not in wedge_edges:
                if outc[a] == 0:
                    loop[a] = b
                if inc[b] == 0:
                    loopback[b] = a
                outc[a] += 1
                inc[b] += 1
    kind = [LOCKED] * V
    for r in range(V):
        if not used[r] or remap[r] != r or complex_[r] or pborder[r]:
            continue
        w = wedge[r]
        if w == r:
            if outc[r] == 0 and inc[r] == 0:
                kind[r] = MANIFOLD
        elif wedge[w] == r:
            if outc[r] == 1 and inc[r] == 1 and outc[w] == 1 and inc[w] == 1 and \
               remap[loop[r]] == remap[loopback[w]] and remap[loopback[r]] == remap[loop[w]]:
                kind[r] = SEAM

    # ---- quadrics ----
    Q = [[0.0] * 11 for _ in range(V)]
    for t in range(T):
        i0, i1, i2 = idx[3 * t], idx[3 * t + 1], idx[3 * t + 2]
        q = _q_triangle(P[i0], P[i1], P[i2])
        if q is not None:
            _q_add(Q[remap[i0]], q)
            _q_add(Q[remap[i1]], q)
            _q_add(Q[remap[i2]], q)
    for (i0, i1, i2) in seam_edges:
        q = _q_seam_edge(P[i0], P[i1], P[i2])
        if q is not None:
            _q_add(Q[remap[i0]], q)
            _q_add(Q[remap[i1]], q)
    QA = G = None
    if nrm is not None:
        QA = [[0.0] * 11 for _ in range(V)]
        G = [[[0.0] * 4 for _ in range(3)] for _ in range(V)]
        for t in range(T):
            i0, i1, i2 = idx[3 * t], idx[3 * t + 1], idx[3 * t + 2]
            q, grads = _q_attributes(P[i0], P[i1], P[i2], nrm[i0], nrm[i1], nrm[i2])
            for v in (i0, i1, i2):
                _q_add(QA[v], q)
                for k in range(3):
                    for j in range(4):
                        G[v][k][j] = G[v][k][j] + grads[k][j]

    def pos_error(v0, v1):
        q = Q[remap[v0]]
        return abs(_q_eval(q, P[v1])) * (0.0 if q[10] == 0.0 else 1.0 / q[10])

    def attr_error(v0, v1):
        q, p, a = QA[v0], P[v1], nrm[v1]
        r = _q_eval(q, p)
        for k in range(3):
            g = G[v0][k][0] * p[0] + G[v0][k][1] * p[1] + G[v0][k][2] * p[2] + G[v0][k][3]
            r = r + a[k] * (a[k] * q[10] - 2.0 * g)
        return abs(r) * (0.0 if q[10] == 0.0 else 1.0 / q[10])

    def seam_partner(v0, v1):
        s0 = wedge[v0]
        s1 = loopback[s0] if loop[v0] == v1 else loop[s0]
        return (s0, s1) if s1 != NONE and remap[s1] == remap[v1] else None

    def collapse_error(v0, v1):
        e = pos_error(v0, v1)
        if nrm is not None:
            e = e + attr_error(v0, v1)
            if kind[remap[v0]] == SEAM:
                s = seam_partner(v0, v1)
                e = e + attr_error(s[0], s[1])
        return math.inf if e != e else e

    def allowed(v0, v1):
        k0 = kind[remap[v0]]
        if k0 == MANIFOLD:
            return True
        if k0 == SEAM and kind[remap[v1]] == SEAM and (loop[v0] == v1 or loopback[v0] == v1):
            return seam_partner(v0, v1) is not None
        return False

    result = idx
    # each triangle remembers its ORIGINAL normal: a corner may move many times, the face may never turn away from where it started
    orig_n = [_cross(_sub(P[idx[3 * t + 1]], P[idx[3 * t]]), _sub(P[idx[3 * t + 2]], P[idx[3 * t]])) for t in range(T)]
    result_error = 0.0
    error_limit = float(target_error) * float(target_error)
    while len(result) > target_index_count:
        T = len(result) // 3
        tris_of = [[] for _ in range(V)]
        for t in range(T):
            for e in range(3):
                tris_of[remap[result[3 * t + e]]].append(t)
        cands = []
        for t in range(T):
            for e in range(3):
                i0, i1 = result[3 * t + e], result[3 * t + (e + 1) % 3]
                r0, r1 = remap[i0], remap[i1]
                if r0 == r1 or r1 > r0:  # the opposite half-edge generates this pair
                    continue
                a01, a10 = allowed(i0, i1), allowed(i1, i0)
                if not a01 and not a10:
                    continue
                if a01 and a10:
                    e01, e10 = collapse_error(i0, i1), collapse_error(i1, i0)
                    cands.append((e10, i1, i0, pos_error(i1, i0)) if e10 < e01 else (e01, i0, i1, pos_error(i0, i1)))
                elif a01:
                    cands.append((collapse_error(i0, i1), i0, i1, pos_error(i0, i1)))
                else:
                    cands.append((collapse_error(i1, i0), i1, i0, pos_error(i1, i0)))
        if not cands:
            break
        order = sorted(range(len(cands)), key=lambda c: cands[c][0])  # stable
        goal = (len(result) - target_index_count) // 3
        edge_goal = goal // 2
        error_goal = 1.5 * cands[order[edge_goal]][0] if edge_goal < len(cands) else math.inf
        collapse_remap = list(range(V))
        pos_collapse = list(range(V))
        locked = [False] * V
        collapsed_tris = collapses = 0

        def ring(r):
            """current triangles of position r as position triples, r first, degenerate ones dropped"""
            out = []
            for t in tris_of[r]:
                a, b, c = (pos_collapse[remap[result[3 * t + k]]] for k in range(3))
                if a == b or b == c or c == a:
                    continue
                if b == r:
                    a, b, c = b, c, a
                elif c == r:
                    a, b, c = c, a, b
                out.append((a, b, c, t))
            return out

        for ci in order:
            err, v0, v1 = cands[ci][:3]
            if err > error_limit or collapsed_tris >= goal:
                break
            if err > error_goal and collapsed_tris > goal // 6:
                break
            r0, r1 = remap[v0], remap[v1]
            if locked[r0] or locked[r1]:
                continue
            ring0 = ring(r0)
            shared = [c if b == r1 else b for (_, b, c, _) in ring0 if b == r1 or c == r1]
            if len(shared) != 2 or shared[0] == shared[1]:
                continue
            n0 = {x for (_, b, c, _) in ring0 for x in (b, c)}
            n1 = {x for (_, b, c, _) in ring(r1) for x in (b, c)}
            if (n0 & n1) != set(shared):
                continue  # link condition: the collapse would pinch the surface
            flip = False
            for (_, b, c, t) in ring0:
                if b == r1 or c == r1:
                    continue
                n_old = _cross(_sub(P[b], P[r0]), _sub(P[c], P[r0]))
                n_new = _cross(_sub(P[b], P[r1]), _sub(P[c], P[r1]))
                nn = _dot(n_new, n_new)
                if _dot(n_old, n_new) <= 0.25 * math.sqrt(_dot(n_old, n_old) * nn) or \
                   _dot(orig_n[t], n_new) <= 0.25 * math.sqrt(_dot(orig_n[t], orig_n[t]) * nn):
                    flip = True
                    break
            if flip:
                continue
            pairs = [(v0, v1)]
            if kind[r0] == SEAM:
                pairs.append(seam_partner(v0, v1))
            for (a, b) in pairs:
                collapse_remap[a] = b
                if QA is not None:
                    _q_add(QA[b], QA[a])
                    for k in range(3):
                        for j in range(4):
                            G[b][k][j] = G[b][k][j] + G[a][k][j]
            _q_add(Q[r1], Q[r0])
            pos_collapse[r0] = r1
            locked[r0] = locked[r1] = True
            collapsed_tris += 2
            collapses += 1
            perr = cands[ci][3]
            if perr > result_error:
                result_error = perr
        if collapses == 0:
            break
        new, new_n = [], []
        for t in range(T):
            a, b, c = (collapse_remap[result[3 * t + k]] for k in range(3))
            if remap[a] != remap[b] and remap[b] != remap[c] and remap[c] != remap[a]:
                new += [a, b, c]
                new_n.append(orig_n[t])
        result, orig_n = new, new_n
        for tbl in (loop, loopback):
            upd = list(tbl)
            for i in range(V):
                l = tbl[i]
                if l == NONE or collapse_remap[i] != i:
                    continue
                r = collapse_remap[l]
                if r == i:
                    l2 = tbl[l]
                    r = NONE if l2 == NONE else collapse_remap[l2]
                    r = NONE if r == i else r
                upd[i] = r
            tbl[:] = upd
    return result, np.float32(math.sqrt(result_error))


def lod_chain(indices0, positions, normals, error0=0.0, max_lods=8):
    """AssetManager_GLTF.cpp:596-641: [(indices, error)] with LOD l simplified from LOD l-1 to half its index count."""
    lods = [([int(i) for i in np.asarray(indices0).reshape(-1)], np.float32(error0))]
    for _ in range(1, max_lods):
        last, last_err = lods[-1]
        target = ((len(last) + 5) // 6) * 3
        simp, err = simplify(last, positions, normals, target)
        cur_err = np.float32(last_err + err)
        if len(simp) > target + target // 2 or err > np.float32(0.5) or len(simp) < 6:
            break
        lods.append((simp, cur_err))
    return lods
