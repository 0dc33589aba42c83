"""Specification tests of the CLIPPED raster (oracle only — the CUDA raster still drops triangles with a vertex at w <= 0;
DESIGN.md §8 item 5): near / side-plane clipping of exactly the triangles the plain spec drops."""
import numpy as np

from oxylus_b200 import abi, capi


def ground_scene(cells, width=320, height=180, split=False):
    """a 100 x 110 ground quad one unit below the camera, reaching 10 units BEHIND it (crosses the near plane and both side
    planes), as cells x cells quads.  split=True: one mesh per triangle of the 1x1 version (for the watertightness check)."""
    xs = np.linspace(-50.0, 50.0, cells + 1)
    zs = np.linspace(10.0, -100.0, cells + 1)
    gx, gz = np.meshgrid(xs, zs, indexing="ij")
    pos = np.stack([gx, np.full_like(gx, -1.0), gz], axis=2).reshape(-1, 3).astype(np.float32)
    i, j = np.meshgrid(np.arange(cells), np.arange(cells), indexing="ij")
    a, b = i * (cells + 1) + j, (i + 1) * (cells + 1) + j
    c, d = i * (cells + 1) + j + 1, (i + 1) * (cells + 1) + j + 1
    tris = np.stack([a, d, c, a, b, d], axis=2).reshape(-1, 3).astype(np.uint32)
    if split:
        built = [capi.BuiltMesh(pos, [(t.reshape(-1), 0.0)]) for t in tris]
    else:
        built = [capi.BuiltMesh(pos, [(tris.reshape(-1), 0.0)])]
    xf = np.tile(np.eye(4, dtype=np.float32).reshape(16), (len(built), 1))
    return capi.assemble_scene(built, np.arange(len(built)), xf, width, height)


def render(orc, sc, clip):
    hs = orc.HostScene(sc)
    cam = sc.camera()
    mi, vis, _ = orc.cull_meshes(hs, cam, abi.CULL_TEST_ALL)
    total = int(vis["total"][0])
    assert total > 0
    img = orc.clear_visbuffer(sc.width, sc.height)
    ids = np.arange(total, dtype=np.uint32)
    if clip:
        ntri, nclip = orc.raster_clip(hs, mi, ids, 0, total, cam, img)
    else:
        ntri, nclip = orc.raster(hs, mi, ids, 0, total, cam, img), 0
    v32, depth = orc.resolve(img)
    return v32, depth, ntri, nclip


def test_front_facing_ground_is_dropped_without_clipping_and_drawn_with_it(orc):
    coarse = ground_scene(1)
    v_plain, _, ntri, _ = render(orc, coarse, clip=False)
    assert ntri == 2, "both triangles must pass cull_triangles (front facing from above)"
    assert (v_plain != 0xFFFFFFFF).sum() == 0          # the plain spec drops them: a vertex is behind the camera
    v_clip, d_clip, ntri, nclip = render(orc, coarse, clip=True)
    assert (ntri, nclip) == (2, 2)
    cov = v_clip != 0xFFFFFFFF
    # ground truth: the same plane as 160 x 160 small quads — the visible ones never cross a clip plane and are drawn by
    # the plain rules (the camera is 1 above the ground: the nearest visible ground is 1.7 units away, near plane 0.1)
    fine = ground_scene(160)
    v_fine, d_fine, _, nclip_fine = render(orc, fine, clip=True)
    v_fine_plain, _, _, _ = render(orc, fine, clip=False)
    ref = v_fine != 0xFFFFFFFF
    assert ref.sum() > 0.4 * ref.size                  # the ground fills the lower half of the image
    # clipping only adds the few small quads around the camera that are (partly) behind it; none of them is on screen
    assert np.array_equal(ref, v_fine_plain != 0xFFFFFFFF)
    diff = cov ^ ref
    assert diff.sum() <= 2 * (coarse.width + coarse.height), f"{diff.sum()} pixels differ"   # silhouette pixels only
    ys, xs = np.nonzero(diff)
    if len(ys):  # every differing pixel touches the boundary of the reference coverage (horizon line / image border)
        pad = np.pad(ref, 1, mode="edge")
        neigh = sum(pad[1 + dy: 1 + dy + ref.shape[0], 1 + dx: 1 + dx + ref.shape[1]].astype(int) for dy in (-1, 0, 1) for dx in (-1, 0, 1))
        assert np.all((neigh[ys, xs] > 0) & (neigh[ys, xs] < 9))
    both = cov & ref
    assert np.abs(d_clip[both] - d_fine[both]).max() < 2e-4   # same plane, different interpolation paths


def test_clipped_pieces_are_watertight(orc):
    """the two triangles of the big quad share the diagonal and are clipped independently: no pixel is drawn by both, and
    together they leave no hole along the diagonal"""
    sc = ground_scene(1, split=True)
    hs = orc.HostScene(sc)
    cam = sc.camera()
    mi, vis, _ = orc.cull_meshes(hs, cam, abi.CULL_TEST_ALL)
    assert int(vis["total"][0]) == 2
    covers = []
    for k in range(2):
        img = orc.clear_visbuffer(sc.width, sc.height)
        assert orc.raster_clip(hs, mi, np.array([k], dtype=np.uint32), 0, 1, cam, img) == (1, 1)
        covers.append(orc.resolve(img)[0] != 0xFFFFFFFF)
    assert covers[0].sum() > 100 and covers[1].sum() > 100 and (covers[0] | covers[1]).sum() > 20000
    assert not (covers[0] & covers[1]).any()
    union = covers[0] | covers[1]
    whole, _, _, _ = render(orc, ground_scene(1), clip=True)
    assert np.array_equal(union, whole != 0xFFFFFFFF)
    # no holes strictly inside: every interior pixel of the union's bounding rows is covered between the row's extremes
    for y in np.nonzero(union.any(axis=1))[0]:
        xs = np.nonzero(union[y])[0]
        assert union[y, xs[0]: xs[-1] + 1].all(), f"hole in row {y}"


def test_unclipped_triangles_are_untouched(orc, small_scene):
    sc = small_scene
    hs = orc.HostScene(sc)
    cam = sc.camera()
    mi, vis, _ = orc.cull_meshes(hs, cam, abi.CULL_TEST_ALL)
    total = int(vis["total"][0])
    ids = np.arange(total, dtype=np.uint32)
    a, b = orc.clear_visbuffer(sc.width, sc.height), orc.clear_visbuffer(sc.width, sc.height)
    na = orc.raster(hs, mi, ids, 0, total, cam, a)
    nb, nclip = orc.raster_clip(hs, mi, ids, 0, total, cam, b)
    assert na == nb
    if nclip == 0:
        assert np.array_equal(a, b)
    else:  # clipping may only ADD coverage
        assert np.all(b >= a)
