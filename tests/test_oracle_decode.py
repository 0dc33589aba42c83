"""Oracle known answers / properties for the two §8f rows next to the path: the vis-buffer decode
(visbuffer_decode.slang:42-183, geometry part) and the hierarchical page bitmap (rmvsm_downsample_hpb.slang)."""
import numpy as np

from oxylus_b200 import abi
from tests.helpers_scene import quad_scene


def oct_decode(e):
    """inverse of common/encoding.slang:17-21 (test-side only)"""
    x, y = float(e[0]), float(e[1])
    z = 1.0 - abs(x) - abs(y)
    if z < 0:
        x, y = (1 - abs(y)) * np.sign(x if x != 0 else 1), (1 - abs(x)) * np.sign(y if y != 0 else 1)
    v = np.array([x, y, z])
    return v / np.linalg.norm(v)


def test_decode_known_answer_quad(orc):
    W = H = 8
    sc, cam = quad_scene(W, H, attributes=True)
    hs = orc.HostScene(sc)
    mi, vis, _ = orc.cull_meshes(hs, cam, abi.CULL_TEST_ALL)
    assert int(vis["total"][0]) == 1
    v32 = np.full((H, W), 0xFFFFFFFF, dtype=np.uint32)
    v32[3, 5] = (0 << 8) | 0           # meshlet instance 0, triangle 0 = vertices (0, 2, 1)
    v32[4, 2] = (0 << 8) | 1           # triangle 1 = vertices (0, 3, 2)
    v32[0, 0] = (0xFFFFFE << 8) | 3    # terrain sentinel (visbuffer.slang:16-20) -> discarded
    v32[7, 7] = (5 << 8) | 0           # meshlet instance out of range -> discarded (guard)
    d = orc.decode_visbuffer(hs, mi, 1, cam, v32)
    lam = d["lambda_"]
    assert lam[0, 0, 3] == 0.0 and lam[7, 7, 3] == 0.0 and lam[1, 1, 3] == 0.0 and not lam[0, 0, :3].any()
    # pixel (5,3): ndc = (0.375, -0.125); p - v0 = a (v2 - v0) + b (v1 - v0): a = y + 0.5, b = x - y
    np.testing.assert_allclose(lam[3, 5], [0.125, 0.375, 0.5, 1.0], atol=1e-6)
    np.testing.assert_allclose(d["ddx"][3, 5, :3], [-2.0 / W, 0.0, 2.0 / W], atol=1e-6)
    # the reference scales ddy by -2/res.y (visbuffer_decode.slang:76): derivative towards -y in texture space
    np.testing.assert_allclose(d["ddy"][3, 5, :3], [0.0, -2.0 / H, 2.0 / H], atol=1e-6)
    # uv = xy + 0.5 per vertex -> interpolated uv == ndc + 0.5; gradients follow
    np.testing.assert_allclose(d["uv_normal"][3, 5, :2], [0.875, 0.375], atol=1e-3)
    np.testing.assert_allclose(d["uv_grad"][3, 5], [2.0 / W, 0.0, 0.0, -2.0 / H], atol=1e-3)
    # normal (0,0,1) (1022/511 - 1 == 1 exactly), identity world -> oct (0,0)
    np.testing.assert_allclose(d["uv_normal"][3, 5, 2:], [0.0, 0.0], atol=1e-6)
    # pixel (2,4): ndc = (-0.375, 0.125), triangle (0,3,2): p - v0 = a (v3 - v0) + b (v2 - v0): b = x + 0.5, a = y - x
    np.testing.assert_allclose(lam[4, 2], [0.375, 0.5, 0.125, 1.0], atol=1e-6)


def test_decode_without_attributes_and_bad_index(orc):
    W = H = 8
    sc, cam = quad_scene(W, H)
    hs = orc.HostScene(sc)
    mi, _, _ = orc.cull_meshes(hs, cam, abi.CULL_TEST_ALL)
    v32 = np.full((H, W), 0xFFFFFFFF, dtype=np.uint32)
    v32[3, 5] = 0
    d = orc.decode_visbuffer(hs, mi, 1, cam, v32)
    np.testing.assert_allclose(d["lambda_"][3, 5], [0.125, 0.375, 0.5, 1.0], atol=1e-6)
    assert not d["uv_grad"][3, 5].any() and not d["uv_normal"][3, 5, :2].any()  # Mesh::texture_coords == nullptr -> {}
    assert np.isnan(d["uv_normal"][3, 5, 2:]).all()                               # normalize(0) -> NaN, as the shader would
    sc.meshes["vertex_count"] = 2   # visbuffer_decode.slang:115-117: any index > vertex_count - 1 -> zero output
    d = orc.decode_visbuffer(orc.HostScene(sc), mi, 1, cam, v32)
    assert d["lambda_"][3, 5, 3] == 2.0 and not d["lambda_"][3, 5, :3].any() and not d["ddx"][3, 5].any()


def test_decode_frame_properties(orc, small_scene):
    sc = small_scene
    hs = orc.HostScene(sc)
    mask = np.zeros((sc.max_meshlet_instance_count + 31) // 32, dtype=np.uint32)
    cam = sc.camera()
    r = orc.frame(hs, cam, sc.width, sc.height, mask, sc.occluder_depth)
    total = int(r["visibility"]["total"][0])
    v32, depth = orc.resolve(r["vis64"])
    d = orc.decode_visbuffer(hs, r["meshlet_instances"], total, cam, v32)
    lam, status = d["lambda_"][:, :, :3].astype(np.float64), d["lambda_"][:, :, 3]
    covered = v32 != 0xFFFFFFFF
    assert covered.sum() > 500
    assert np.array_equal(status == 1.0, covered) and np.array_equal(status == 0.0, ~covered)
    ys, xs = np.nonzero(covered)
    L = lam[ys, xs]
    assert np.all(np.isfinite(L))
    # barycentrics of a pixel the rasteriser assigned to this triangle: partition of unity, inside up to the
    # sub-pixel snapping of the raster (24.8 fixed point) and f32 rounding
    np.testing.assert_allclose(L.sum(axis=1), 1.0, atol=2e-4)
    assert L.min() > -0.2 and L.max() < 1.2  # tiny triangles: vertex snapping is a visible fraction of their size
    # re-projection: sum lambda_i * world_i projects back onto the pixel centre
    pv = cam["projection_view"][0].reshape(4, 4).T.astype(np.float64)
    mi = r["meshlet_instances"]
    worst = 0.0
    sel = np.linspace(0, len(ys) - 1, 300).astype(int)
    for k in sel:
        y, x = ys[k], xs[k]
        t = int(v32[y, x])
        wp = world_positions(sc, mi, t >> 8, t & 0xFF)
        P = (L[k][:, None] * wp).sum(axis=0)
        c = pv @ np.append(P, 1.0)
        ndc = c[:2] / c[3]
        want = np.array([(x + 0.5) / sc.width * 2 - 1, (y + 0.5) / sc.height * 2 - 1])
        worst = max(worst, float(np.abs(ndc - want).max()))
        # reverse-Z depth of the reconstructed point == the rasterised depth (different interpolation paths)
        assert abs(c[2] / c[3] - depth[y, x]) < 2e-4 + 1e-3 * depth[y, x]
    assert worst < 1e-3
    # ddx / ddy == barycentrics of the SAME triangle evaluated one pixel towards +x / towards -y (the decode does
    # not care whether the raster covered that pixel): decode the image shifted by one pixel
    dxs = orc.decode_visbuffer(hs, r["meshlet_instances"], total, cam, np.roll(v32, 1, axis=1))["lambda_"][:, :, :3].astype(np.float64)
    dys = orc.decode_visbuffer(hs, r["meshlet_instances"], total, cam, np.roll(v32, -1, axis=0))["lambda_"][:, :, :3].astype(np.float64)
    inner = covered.copy()
    inner[:, -1] = False
    inner[0, :] = False
    yy, xx = np.nonzero(inner)
    assert len(yy) > 500
    scale = np.abs(lam[yy, xx]).max(axis=1, keepdims=True) + np.abs(dxs[yy, xx + 1]).max(axis=1, keepdims=True)
    assert np.all(np.abs(d["ddx"][yy, xx, :3] - (dxs[yy, xx + 1] - lam[yy, xx])) < 1e-3 * scale + 1e-4)
    scale = np.abs(lam[yy, xx]).max(axis=1, keepdims=True) + np.abs(dys[yy - 1, xx]).max(axis=1, keepdims=True)
    assert np.all(np.abs(d["ddy"][yy, xx, :3] - (dys[yy - 1, xx] - lam[yy, xx])) < 1e-3 * scale + 1e-4)
    # interpolated uv stays in the patch's [0,1] parameter range; the normal is unit length and faces the camera side
    uv = d["uv_normal"][ys, xs, :2]
    assert uv.min() > -0.1 and uv.max() < 1.1
    for k in sel[:100]:
        y, x = ys[k], xs[k]
        n = oct_decode(d["uv_normal"][y, x, 2:])
        t = int(v32[y, x])
        wp = world_positions(sc, mi, t >> 8, t & 0xFF)
        g = np.cross(wp[1] - wp[0], wp[2] - wp[0])
        g /= np.linalg.norm(g)
        assert abs(np.dot(n, g)) > 0.7


def world_positions(sc, mi, mii, tri):
    """test-side re-fetch of one triangle's world positions (f64)"""
    m = mi[mii]
    inst = sc.mesh_instances[m["mesh_instance_index"]]
    mesh = sc.meshes[inst["mesh_index"]]
    lod = np.frombuffer(sc.blob, dtype=abi.MESH_LOD_DT, count=int(mesh["lod_count"]), offset=int(mesh["lods"]))[inst["lod_index"]]
    ml = np.frombuffer(sc.blob, dtype=abi.MESHLET_DT, count=int(lod["meshlet_count"]), offset=int(lod["meshlets"]))[m["meshlet_index"]]
    micro = sc.blob[int(lod["local_triangle_indices"]) + int(ml["local_triangle_index_offset"]) + 3 * tri:][:3]
    vidx = np.frombuffer(sc.blob, dtype=np.uint32, count=int(lod["indirect_vertex_indices_count"]), offset=int(lod["indirect_vertex_indices"]))
    idx = vidx[int(ml["indirect_vertex_index_offset"]) + micro.astype(np.int64)]
    pos = np.frombuffer(sc.blob, dtype=np.float16, count=int(mesh["vertex_count"]) * 4, offset=int(mesh["vertex_positions"])).reshape(-1, 4)
    p = pos[idx.astype(np.int64), :3].astype(np.float64)
    w = sc.transforms["world"][inst["transform_index"]].reshape(4, 4).T.astype(np.float64)
    return (w[:3, :3] @ p.T).T + w[:3, 3]


def test_build_hpb_known_answer(orc):
    V, D, B = abi.VSM_PAGE_VISIBLE, abi.VSM_PAGE_DIRTY, abi.VSM_PAGE_BACKED
    pt = np.zeros((2, 4, 4), dtype=np.uint32)
    pt[0, 1, 2] = V | D | B | (77 << 16)   # cached (physical address bits ignored)
    pt[0, 0, 0] = V | B                    # not dirty
    pt[0, 3, 3] = D | B                    # not visible
    pt[1, 2, 1] = V | D                    # not backed
    pt[1, 3, 0] = V | D | B | 8            # + Invalidated flag: still cached
    hpb = orc.build_hpb(pt, 3)
    offs, total = orc.hpb_layout(4, 2, 3)
    assert total == 2 * 16 + 2 * 4 + 2 and offs == [0, 32, 40]
    l0 = hpb[:32].reshape(2, 4, 4)
    want0 = np.zeros((2, 4, 4), dtype=np.uint8)
    want0[0, 1, 2] = 1
    want0[1, 3, 0] = 1
    assert np.array_equal(l0, want0)
    l1 = hpb[32:40].reshape(2, 2, 2)
    assert np.array_equal(l1, np.array([[[0, 1], [0, 0]], [[0, 0], [1, 0]]], dtype=np.uint8))
    assert list(hpb[40:42]) == [1, 1]
    # more levels than log2(size)+1: the extra 1x1 levels repeat the top (out-of-range taps read 0)
    hpb5 = orc.build_hpb(pt, 5)
    assert list(hpb5[40:]) == [1, 1, 1, 1, 1, 1]


def test_build_hpb_matches_numpy(orc):
    rng = np.random.default_rng(7)
    pt = rng.integers(0, 32, size=(3, 32, 32)).astype(np.uint32) | (rng.integers(0, 65536, size=(3, 32, 32)).astype(np.uint32) << 16)
    hpb = orc.build_hpb(pt, 6)
    offs, _ = orc.hpb_layout(32, 3, 6)
    cur = ((pt & 7) == 7).astype(np.uint8)
    for l in range(6):
        s = max(1, 32 >> l)
        assert np.array_equal(hpb[offs[l]: offs[l] + 3 * s * s].reshape(3, s, s), cur)
        if s > 1:
            cur = cur[:, 0::2, 0::2] | cur[:, 1::2, 0::2] | cur[:, 0::2, 1::2] | cur[:, 1::2, 1::2]
