"""Specification tests of the alpha-tested discard (visbuffer_encode.slang:54-66; spec: oracle/oxc_oracle.c above raster_triangle).
Oracle only: hand-computed samples, an independent binary64 perspective-correct interpolation, and pass-level properties."""
import numpy as np
import pytest

from oxylus_b200 import abi, capi, synth


def half(x):
    return int(np.float16(x).view(np.uint16))


def material(image=None, cutoff=0.5, albedo_a=1.0, sampler=0, flags=None):
    m = np.zeros(1, dtype=abi.MATERIAL_DT)
    m["albedo_color"][0] = [half(1.0), half(1.0), half(1.0), half(albedo_a)]
    m["alpha_cutoff"] = half(cutoff)
    m["sampler_index"] = sampler
    if image is not None:
        m["flags"] = abi.MATERIAL_HAS_ALBEDO_IMAGE | abi.MATERIAL_ALPHA_MASK
        m["albedo_image_index"] = image
    if flags is not None:
        m["flags"] = flags
    return m[0]


def checker(n, cell, lo=0, hi=255, rgba=True):
    y, x = np.mgrid[0:n, 0:n]
    a = np.where(((x // cell) + (y // cell)) % 2 == 0, hi, lo).astype(np.uint8)
    if not rgba:
        return a
    t = np.full((n, n, 4), 200, dtype=np.uint8)
    t[:, :, 3] = a
    return t


def textured_ground(cells, width=320, height=180):
    """tests/test_oracle_clip.py's ground plane (crosses the near and both side planes) with uv = (x, z) / 10"""
    xs = np.linspace(-50.0, 50.0, cells + 1)
    zs = np.linspace(10.0, -100.0, cells + 1)
    gx, gz = np.meshgrid(xs, zs, indexing="ij")
    pos = np.stack([gx, np.full_like(gx, -1.0), gz], axis=2).reshape(-1, 3).astype(np.float32)
    uv = np.stack([gx / 10.0, gz / 10.0], axis=2).reshape(-1, 2).astype(np.float32)
    i, j = np.meshgrid(np.arange(cells), np.arange(cells), indexing="ij")
    a, b = i * (cells + 1) + j, (i + 1) * (cells + 1) + j
    c, d = i * (cells + 1) + j + 1, (i + 1) * (cells + 1) + j + 1
    tris = np.stack([a, d, c, a, b, d], axis=2).reshape(-1, 3).astype(np.uint32)
    built = [capi.BuiltMesh(pos, [(tris.reshape(-1), 0.0)], texcoords=uv)]
    xf = np.eye(4, dtype=np.float32).reshape(1, 16)
    return capi.assemble_scene(built, np.arange(1), xf, width, height)


def test_alpha_sample_hand_cases(orc):
    img = np.array([[0, 255], [255, 0]], dtype=np.uint8)  # R8, texel (x, y) = img[y, x]
    s = orc.alpha_sample
    # texel centres (linear filter returns the texel itself)
    assert s(img, abi.IMAGE_R8_UNORM, 0.25, 0.25) == 0.0
    assert s(img, abi.IMAGE_R8_UNORM, 0.75, 0.25) == 1.0
    assert s(img, abi.IMAGE_R8_UNORM, 0.25, 0.75) == 1.0
    # half way between two texel centres: 0.5; the centre of the image: mean of all four
    assert s(img, abi.IMAGE_R8_UNORM, 0.5, 0.25) == 0.5
    assert s(img, abi.IMAGE_R8_UNORM, 0.5, 0.5) == 0.5
    # repeat: u = 0 lies half way between texel 1 (wrapped) and texel 0
    assert s(img, abi.IMAGE_R8_UNORM, 0.0, 0.25) == 0.5
    assert s(img, abi.IMAGE_R8_UNORM, 1.25, -0.75) == s(img, abi.IMAGE_R8_UNORM, 0.25, 0.25)
    # clamp to edge: left of the first texel centre is the first texel
    clamp = abi.sampler(u=abi.ADDRESS_CLAMP_TO_EDGE, v=abi.ADDRESS_CLAMP_TO_EDGE)
    assert s(img, abi.IMAGE_R8_UNORM, 0.0, 0.25, clamp) == 0.0
    assert s(img, abi.IMAGE_R8_UNORM, -3.0, 0.25, clamp) == 0.0
    assert s(img, abi.IMAGE_R8_UNORM, 7.0, 0.25, clamp) == 1.0
    # mirrored repeat: [0,1) forward, [1,2) backward
    mirror = abi.sampler(abi.FILTER_NEAREST, abi.FILTER_NEAREST, u=abi.ADDRESS_MIRRORED_REPEAT, v=abi.ADDRESS_MIRRORED_REPEAT)
    assert s(img, abi.IMAGE_R8_UNORM, 0.25, 0.25, mirror) == 0.0
    assert s(img, abi.IMAGE_R8_UNORM, 1.25, 0.25, mirror) == 1.0   # texel 1 mirrored
    assert s(img, abi.IMAGE_R8_UNORM, 1.75, 0.25, mirror) == 0.0
    assert s(img, abi.IMAGE_R8_UNORM, -0.25, 0.25, mirror) == 0.0  # texel -1 -> 0
    # nearest, repeat; RGBA8 reads byte 3
    near = abi.sampler(abi.FILTER_NEAREST, abi.FILTER_NEAREST)
    assert s(img, abi.IMAGE_R8_UNORM, 0.49, 0.0, near) == 0.0 and s(img, abi.IMAGE_R8_UNORM, 0.51, 0.0, near) == 1.0
    rgba = np.zeros((1, 1, 4), dtype=np.uint8)
    rgba[0, 0] = [9, 9, 9, 51]
    assert s(rgba, abi.IMAGE_RGBA8_UNORM, 0.3, 0.9) == np.float32(51) / np.float32(255)
    # special values: NaN / Inf coordinates are defined (texel 0 / saturated index), never a crash
    for u in (np.nan, np.inf, -np.inf, 1e30, -1e30):
        v = s(img, abi.IMAGE_R8_UNORM, u, 0.25)
        assert np.isnan(v) or 0.0 <= v <= 1.0
        assert 0.0 <= s(img, abi.IMAGE_R8_UNORM, u, 0.25, near) <= 1.0


def _front_facing(clip):
    """orient the triangle so that the raster keeps it (negative clip-space determinant, cull.slang:169-171)"""
    m = np.stack([clip[:, 0], clip[:, 1], clip[:, 3]], axis=0).astype(np.float64)
    return np.linalg.det(m) < 0


def test_alpha_matches_binary64_perspective_interpolation(orc):
    """the kept / discarded pattern of single triangles with very different w per vertex equals the textbook
    perspective-correct interpolation (homogeneous barycentrics solved in binary64 from the UNSNAPPED clip coordinates) of a
    per-texel checkerboard — away from texel boundaries, where f32 rounding and the 1/256-pixel snapping cannot matter"""
    rng = np.random.default_rng(11)
    W, H = 96, 64
    tex = checker(8, 1, rgba=False)  # 8x8 texels, every texel its own cell; nearest filter => alpha = cell parity
    near = np.array([abi.sampler(abi.FILTER_NEAREST, abi.FILTER_NEAREST)], dtype=abi.SAMPLER_DT)
    tab = orc.MaterialTable([material(image=0, cutoff=0.5)], [(tex, abi.IMAGE_R8_UNORM)], near)
    plain_tab = orc.MaterialTable([material()])
    checked = kept = dropped = 0
    for _ in range(200):
        w = rng.uniform(0.5, 20.0, 3)
        ndc = rng.uniform(-0.9, 0.9, (3, 2))
        clip = np.zeros((3, 4), dtype=np.float32)
        clip[:, 0:2] = (ndc * w[:, None]).astype(np.float32)
        clip[:, 2] = (0.5 * w).astype(np.float32)
        clip[:, 3] = w.astype(np.float32)
        uv = rng.uniform(-1.0, 2.0, (3, 2)).astype(np.float32)
        if not _front_facing(clip):
            clip[[1, 2]] = clip[[2, 1]]
            uv[[1, 2]] = uv[[2, 1]]
        cover = orc.clear_visbuffer(W, H)
        assert orc.raster_triangle_alpha(plain_tab, 0, clip, uv, 7, cover) == 0
        img = orc.clear_visbuffer(W, H)
        assert orc.raster_triangle_alpha(tab, 0, clip, uv, 7, img) == 0
        covered = (cover & 0xFFFFFFFF) == 7
        got = (img & 0xFFFFFFFF) == 7
        assert not (got & ~covered).any()
        c64, uv64 = clip.astype(np.float64), uv.astype(np.float64)
        M = np.stack([c64[:, 0], c64[:, 1], c64[:, 3]], axis=0)
        for py, px in zip(*np.nonzero(covered)):
            X, Y = (px + 0.5) / W * 2 - 1, (py + 0.5) / H * 2 - 1
            l = np.linalg.solve(M, np.array([X, Y, 1.0]))
            l = l / l.sum()
            if l.min() < 0.05:  # keep to the interior: the snapped triangle differs from the exact one by < 1/256 pixel
                continue
            u, v = l @ uv64[:, 0], l @ uv64[:, 1]
            fu, fv = u * 8 - np.floor(u * 8), v * 8 - np.floor(v * 8)
            if min(fu, 1 - fu, fv, 1 - fv) < 0.05:
                continue
            expect = ((int(np.floor(u * 8)) + int(np.floor(v * 8))) % 2) == 0  # alpha 1 on even cells
            assert bool(got[py, px]) == expect, (clip, uv, px, py)
            checked += 1
            kept += expect
            dropped += not expect
    assert checked > 3000 and kept > 1000 and dropped > 1000


def test_alpha_cutoff_and_flags(orc):
    W, H = 16, 16
    clip = np.array([[-1, -1, 0.5, 1], [-1, 3, 0.5, 1], [3, -1, 0.5, 1]], dtype=np.float32)  # covers the screen
    if not _front_facing(clip):
        clip[[1, 2]] = clip[[2, 1]]
    uv = np.array([[0.5, 0.5], [0.5, 0.5], [0.5, 0.5]], dtype=np.float32)                    # constant uv
    tex = np.full((4, 4), 128, dtype=np.uint8)                                              # alpha 128/255 = 0.50196
    img = [(tex, abi.IMAGE_R8_UNORM)]

    def keep(mat, images=img):
        vis = orc.clear_visbuffer(W, H)
        orc.raster_triangle_alpha(orc.MaterialTable([mat], images), 0, clip, uv, 3, vis)
        n = int(((vis & 0xFFFFFFFF) == 3).sum())
        assert n in (0, W * H)
        return n == W * H

    assert keep(material(image=0, cutoff=0.5))           # 0.50196 >= 0.5
    assert not keep(material(image=0, cutoff=0.51))
    assert not keep(material(image=0, cutoff=0.5, albedo_a=0.9))   # albedo_color.a scales the sample (scene.slang:116-121)
    assert keep(material(image=0, cutoff=0.0))           # clamp(0, 0.001, 1): 0.50196 >= 0.001
    assert not keep(material(image=0, cutoff=5.0))       # clamp(5, 0.001, 1) = 1 > 0.50196
    zero = [(np.zeros((4, 4), dtype=np.uint8), abi.IMAGE_R8_UNORM)]
    assert not keep(material(image=0, cutoff=0.0), zero)  # 0 < 0.001
    # no HasAlbedoImage flag: never tested, whatever the cutoff (visbuffer_encode.slang:55)
    assert keep(material(image=0, cutoff=5.0, flags=abi.MATERIAL_ALPHA_MASK))
    # NaN cutoff keeps the fragment (the comparison is false)
    m = material(image=0)
    m["alpha_cutoff"] = 0x7E00
    assert keep(m)


def _screen_triangle(W, H, texels_per_pixel_x, texels_per_pixel_y, n_tex):
    """one w = 1 triangle covering the whole W x H screen whose uv advances by the given number of level-0 texels per pixel"""
    clip = np.array([[-1, -1, 0.5, 1], [3, -1, 0.5, 1], [-1, 3, 0.5, 1]], dtype=np.float32)
    uv = np.array([[0, 0], [2 * W * texels_per_pixel_x / n_tex, 0], [0, 2 * H * texels_per_pixel_y / n_tex]], dtype=np.float32)
    if not _front_facing(clip):
        clip[[1, 2]] = clip[[2, 1]]
        uv[[1, 2]] = uv[[2, 1]]
    return clip, uv


def _kept_fraction(orc, tab, clip, uv, W, H):
    vis = orc.clear_visbuffer(W, H)
    orc.raster_triangle_alpha(tab, 0, clip, uv, 3, vis)
    return float(((vis & 0xFFFFFFFF) == 3).sum()) / (W * H)


def test_mip_level_selection_hand_cases(orc):
    """SampleGrad's level selection (visbuffer_encode.slang:57-60 + the isotropic LOD rule): an 8 x 8 image whose levels are constant
    (alpha 1, 0, 1, 0 for levels 0..3) tells which level a fragment read"""
    W = H = 32
    levels = [np.full((8 >> l, 8 >> l), 255 if l % 2 == 0 else 0, dtype=np.uint8) for l in range(4)]
    img = [(levels, abi.IMAGE_R8_UNORM)]

    def kept(s_x, s_y, cutoff=0.5, smp=None):
        tab = orc.MaterialTable([material(image=0, cutoff=cutoff)], img, None if smp is None else np.array([smp], dtype=abi.SAMPLER_DT))
        clip, uv = _screen_triangle(W, H, s_x, s_y, 8)
        return _kept_fraction(orc, tab, clip, uv, W, H)

    assert kept(1, 1) == 1.0        # lambda = 0: level 0
    assert kept(0.25, 0.25) == 1.0  # magnified: level 0
    assert kept(2, 2) == 0.0        # lambda = 1: level 1
    assert kept(4, 4) == 1.0        # lambda = 2: level 2
    assert kept(8, 8) == 0.0        # lambda = 3: level 3
    assert kept(64, 64) == 0.0      # clamped to the last level
    assert kept(4, 1) == 1.0 and kept(1, 4) == 1.0  # the larger of the two footprints decides (isotropic rule)
    r2 = float(np.sqrt(2.0))
    assert kept(r2, r2, cutoff=0.4) == 1.0 and kept(r2, r2, cutoff=0.6) == 0.0  # lambda = 0.5: half of level 0, half of level 1
    assert kept(2 ** 1.25, 1, cutoff=0.2) == 1.0 and kept(2 ** 1.25, 1, cutoff=0.3) == 0.0  # 0.75 * level 1 + 0.25 * level 2 = 0.25
    nearest_mip = abi.sampler(mip=abi.MIPMAP_NEAREST)
    assert kept(2 ** 0.4, 1, smp=nearest_mip) == 1.0  # lambda 0.4 -> level 0
    assert kept(2 ** 0.6, 1, smp=nearest_mip) == 0.0  # lambda 0.6 -> level 1
    assert kept(2 ** 1.6, 1, smp=nearest_mip) == 1.0  # lambda 1.6 -> level 2
    # a single-level image never leaves level 0
    one = orc.MaterialTable([material(image=0, cutoff=0.5)], [(levels[0], abi.IMAGE_R8_UNORM)])
    clip, uv = _screen_triangle(W, H, 8, 8, 8)
    assert _kept_fraction(orc, one, clip, uv, W, H) == 1.0


def test_mag_and_min_filter_selection(orc):
    """lambda > 0 uses the min filter, else the mag filter: a sampler with mag = nearest, min = linear equals the all-nearest
    sampler when the image is magnified and the all-linear sampler when it is minified"""
    W = H = 48
    rng = np.random.default_rng(2)
    tex = rng.integers(0, 256, (8, 8), dtype=np.uint8)
    img = [(tex, abi.IMAGE_R8_UNORM)]

    def image(smp, s):
        tab = orc.MaterialTable([material(image=0, cutoff=0.5)], img, np.array([smp], dtype=abi.SAMPLER_DT))
        clip, uv = _screen_triangle(W, H, s, s, 8)
        vis = orc.clear_visbuffer(W, H)
        orc.raster_triangle_alpha(tab, 0, clip, uv, 3, vis)
        return (vis & 0xFFFFFFFF) == 3

    mixed = abi.sampler(mag=abi.FILTER_NEAREST, min=abi.FILTER_LINEAR)
    nearest, linear = abi.sampler(abi.FILTER_NEAREST, abi.FILTER_NEAREST), abi.sampler()
    assert not np.array_equal(image(nearest, 0.13), image(linear, 0.13)) and not np.array_equal(image(nearest, 1.7), image(linear, 1.7))
    np.testing.assert_array_equal(image(mixed, 0.13), image(nearest, 0.13))
    np.testing.assert_array_equal(image(mixed, 1.7), image(linear, 1.7))


def test_mip_levels_follow_the_perspective(orc):
    """the textured ground plane with a full mip chain whose levels are constant (level l: alpha 1 for even l, 0 for odd l): bands
    of kept / discarded rows, coarser levels towards the horizon — the derivative-based selection at work across huge clipped
    triangles; rows closer to the horizon never use a finer level than rows below them"""
    sc = textured_ground(1)
    sc.mesh_instances["material_index"] = 0
    hs = orc.HostScene(sc)
    cam = sc.camera()
    mi, vis, _ = orc.cull_meshes(hs, cam, abi.CULL_TEST_ALL)
    visible, cmd = orc.cull_meshlets(hs, mi, vis, cam)
    visible = visible[: int(cmd["x"][0])]
    n = 256
    levels = [np.full((max(1, n >> l), max(1, n >> l)), 255 if l % 2 == 0 else 0, dtype=np.uint8) for l in range(9)]
    tab = orc.MaterialTable([material(image=0, cutoff=0.5)], [(levels, abi.IMAGE_R8_UNORM)], np.array([abi.sampler(mip=abi.MIPMAP_NEAREST)], dtype=abi.SAMPLER_DT))
    img = orc.clear_visbuffer(sc.width, sc.height)
    orc.raster_alpha(hs, mi, visible, 0, len(visible), cam, img, tab)
    plain = orc.clear_visbuffer(sc.width, sc.height)
    orc.raster_clip(hs, mi, visible, 0, len(visible), cam, plain)
    ground = (plain & 0xFFFFFFFF) != 0xFFFFFFFF
    kept = (img & 0xFFFFFFFF) != 0xFFFFFFFF
    col = sc.width // 2
    rows = np.nonzero(ground[:, col])[0]
    state = kept[rows, col].astype(int)
    flips = int(np.abs(np.diff(state)).sum())
    assert 3 <= flips <= 9, flips        # several level bands between the camera and the horizon in the centre column
    assert 0.2 < kept[ground].mean() < 0.8


@pytest.fixture(scope="module")
def scene_and_tables(orc):
    sc = synth.make_scene(4000, config_index=2, width=480, height=270, n_unique_meshes=12)
    sc.mesh_instances["material_index"] = np.arange(sc.mesh_instance_count) % 4
    return sc


def _survivors(orc, sc):
    hs = orc.HostScene(sc)
    cam = sc.camera()
    mi, vis, _ = orc.cull_meshes(hs, cam, abi.CULL_TEST_ALL)
    visible, cmd = orc.cull_meshlets(hs, mi, vis, cam)
    return hs, cam, mi, visible[: int(cmd["x"][0])]


def test_alpha_pass_properties(orc, scene_and_tables):
    sc = scene_and_tables
    hs, cam, mi, visible = _survivors(orc, sc)
    w, h = sc.width, sc.height
    plain = orc.clear_visbuffer(w, h)
    ntri_plain, _ = orc.raster_clip(hs, mi, visible, 0, len(visible), cam, plain)
    mats = [material(), material(image=0, cutoff=0.5), material(image=1, cutoff=0.3, albedo_a=0.8, sampler=1), material(image=2, cutoff=0.5, sampler=2)]
    smp = np.array([abi.sampler(), abi.sampler(abi.FILTER_NEAREST, abi.FILTER_NEAREST, u=abi.ADDRESS_CLAMP_TO_EDGE, v=abi.ADDRESS_CLAMP_TO_EDGE),
                    abi.sampler(u=abi.ADDRESS_MIRRORED_REPEAT)], dtype=abi.SAMPLER_DT)

    def run(images):
        img = orc.clear_visbuffer(w, h)
        ntri, nalpha = orc.raster_alpha(hs, mi, visible, 0, len(visible), cam, img, orc.MaterialTable(mats, images, smp))
        return img, ntri, nalpha

    opaque = [(np.full((4, 4, 4), 255, np.uint8), abi.IMAGE_RGBA8_UNORM), (np.full((4, 4), 255, np.uint8), abi.IMAGE_R8_UNORM),
              (np.full((2, 2), 255, np.uint8), abi.IMAGE_R8_UNORM)]
    img, ntri, nalpha = run(opaque)
    assert ntri == ntri_plain and 0 < nalpha < ntri
    np.testing.assert_array_equal(img, plain)  # alpha 1 everywhere: nothing is discarded

    clear = [(np.zeros((4, 4, 4), np.uint8), abi.IMAGE_RGBA8_UNORM), (np.zeros((4, 4), np.uint8), abi.IMAGE_R8_UNORM),
             (np.zeros((2, 2), np.uint8), abi.IMAGE_R8_UNORM)]
    img0, ntri0, _ = run(clear)
    assert ntri0 == ntri_plain  # the triangle count is taken before the per-fragment test
    ids = (img0 & 0xFFFFFFFF).astype(np.uint32)
    drawn = ids != 0xFFFFFFFF
    inst_of = mi["mesh_instance_index"][(ids[drawn] >> 8)]
    assert (sc.mesh_instances["material_index"][inst_of] == 0).all()  # only the material without an image is left
    # ... and what is left is exactly the image of the opaque meshlets alone
    keep = sc.mesh_instances["material_index"][mi["mesh_instance_index"][visible]] == 0
    only = orc.clear_visbuffer(w, h)
    orc.raster_clip(hs, mi, np.ascontiguousarray(visible[keep]), 0, int(keep.sum()), cam, only)
    np.testing.assert_array_equal(img0, only)

    rng = np.random.default_rng(5)
    mixed = [(checker(16, 2), abi.IMAGE_RGBA8_UNORM), (rng.integers(0, 256, (8, 8), dtype=np.uint8), abi.IMAGE_R8_UNORM),
             (np.tile(np.linspace(0, 255, 16).astype(np.uint8), (16, 1)), abi.IMAGE_R8_UNORM)]
    img1, _, _ = run(mixed)
    assert (img1 <= plain).all() and (img1 >= img0).all()  # discarding fragments can only lower the packed max
    assert (img1 != plain).any() and (img1 != img0).any()


def test_alpha_through_the_clip_path(orc):
    """a textured ground plane through the camera: every triangle that reaches the screen is clipped; the checker pattern must
    be the same whether the plane is 1 quad (huge clipped triangles) or 40 x 40 quads (mostly unclipped) — the interpolation
    runs over the original triangle, not over the clipped pieces"""
    tab_img = [(checker(4, 1, rgba=False), abi.IMAGE_R8_UNORM)]
    near = np.array([abi.sampler(abi.FILTER_NEAREST, abi.FILTER_NEAREST)], dtype=abi.SAMPLER_DT)
    covers = []
    for cells in (1, 40):
        sc = textured_ground(cells)
        sc.mesh_instances["material_index"] = 0
        hs, cam, mi, visible = _survivors(orc, sc)
        img = orc.clear_visbuffer(sc.width, sc.height)
        tab = orc.MaterialTable([material(image=0, cutoff=0.5)], tab_img, near)
        ntri, nalpha = orc.raster_alpha(hs, mi, visible, 0, len(visible), cam, img, tab)
        assert nalpha == ntri > 0
        covers.append((img & 0xFFFFFFFF) != 0xFFFFFFFF)
    a, b = covers
    assert 0.15 * a.size < a.sum() < 0.35 * a.size   # half of the lower half of the screen
    assert (a ^ b).sum() < 0.01 * a.size            # the two tessellations disagree on cell-boundary pixels only


def test_overdraw_counter_hand_cases(orc):
    """RENDER_OVERDRAW (visbuffer_encode.slang:68-70): every shaded fragment counts, whatever the depth test would say; discarded
    fragments do not.  The ground plane drawn twice (two mesh instances of the same mesh, the second one lower) gives 2 where both
    cover a pixel — although only one of them wins the vis buffer — and the checker material removes its share."""
    sc1 = textured_ground(8)
    xf = np.tile(np.eye(4, dtype=np.float32).reshape(1, 16), (2, 1))
    xf[1, 13] = -0.5  # the second instance half a unit lower (column-major translation y)
    from tests.test_oracle_alpha import textured_ground as _tg  # noqa: F401  (same module: keeps the helper's name searchable)
    xs = np.linspace(-50.0, 50.0, 9)
    zs = np.linspace(10.0, -100.0, 9)
    gx, gz = np.meshgrid(xs, zs, indexing="ij")
    pos = np.stack([gx, np.full_like(gx, -1.0), gz], axis=2).reshape(-1, 3).astype(np.float32)
    uv = np.stack([gx / 10.0, gz / 10.0], axis=2).reshape(-1, 2).astype(np.float32)
    i, j = np.meshgrid(np.arange(8), np.arange(8), indexing="ij")
    a, b, c, d = i * 9 + j, (i + 1) * 9 + j, i * 9 + j + 1, (i + 1) * 9 + j + 1
    tris = np.stack([a, d, c, a, b, d], axis=2).reshape(-1, 3).astype(np.uint32)
    built = [capi.BuiltMesh(pos, [(tris.reshape(-1), 0.0)], texcoords=uv)]
    sc = capi.assemble_scene(built, np.array([0, 0]), xf, sc1.width, sc1.height)
    sc.mesh_instances["material_index"] = [0, 1]
    hs, cam, mi, visible = _survivors(orc, sc)
    w, h = sc.width, sc.height
    plain = orc.clear_visbuffer(w, h)
    orc.raster_clip(hs, mi, visible, 0, len(visible), cam, plain)
    over = np.zeros((h, w), dtype=np.uint32)
    orc.raster_overdraw(hs, mi, visible, 0, len(visible), cam, over)
    covered = (plain & 0xFFFFFFFF) != 0xFFFFFFFF
    assert set(np.unique(over)) == {0, 1, 2}
    np.testing.assert_array_equal(over > 0, covered)       # a pixel is shaded at least once iff something was drawn there
    assert (over == 2).sum() > 0.5 * covered.sum()         # both planes cover most of the lower half of the screen
    # per instance: the counter is the sum of the two single-instance coverages
    total = np.zeros((h, w), dtype=np.uint32)
    for inst in (0, 1):
        keep = mi["mesh_instance_index"][visible] == inst
        img = orc.clear_visbuffer(w, h)
        orc.raster_clip(hs, mi, np.ascontiguousarray(visible[keep]), 0, int(keep.sum()), cam, img)
        total += ((img & 0xFFFFFFFF) != 0xFFFFFFFF).astype(np.uint32)
    np.testing.assert_array_equal(over, total)
    # accumulation: a second call adds the same again
    orc.raster_overdraw(hs, mi, visible, 0, len(visible), cam, over)
    np.testing.assert_array_equal(over, 2 * total)
    # with the checker material on instance 1, its discarded fragments are not counted
    tab = orc.MaterialTable([material(), material(image=0, cutoff=0.5)], [(checker(4, 1, rgba=False), abi.IMAGE_R8_UNORM)],
                            np.array([abi.sampler(abi.FILTER_NEAREST, abi.FILTER_NEAREST)], dtype=abi.SAMPLER_DT))
    over_a = np.zeros((h, w), dtype=np.uint32)
    orc.raster_overdraw(hs, mi, visible, 0, len(visible), cam, over_a, tab)
    assert (over_a <= total).all() and (over_a >= (total > 0) * 0).all()
    keep1 = mi["mesh_instance_index"][visible] == 1
    img1 = orc.clear_visbuffer(w, h)
    orc.raster_alpha(hs, mi, np.ascontiguousarray(visible[keep1]), 0, int(keep1.sum()), cam, img1, tab)
    keep0 = ~keep1
    img0 = orc.clear_visbuffer(w, h)
    orc.raster_clip(hs, mi, np.ascontiguousarray(visible[keep0]), 0, int(keep0.sum()), cam, img0)
    expect = ((img0 & 0xFFFFFFFF) != 0xFFFFFFFF).astype(np.uint32) + ((img1 & 0xFFFFFFFF) != 0xFFFFFFFF).astype(np.uint32)
    np.testing.assert_array_equal(over_a, expect)
    assert (over_a != total).any()


def load_golden_alpha():
    import importlib.util
    import json
    import os

    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_golden_alpha", os.path.join(here, "golden", "make_golden_alpha.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    return mg, json.load(open(os.path.join(here, "golden", "alpha_small.json")))


def test_golden_alpha_fixture(orc):
    """tests/golden/alpha_small.json (written by make_golden_alpha.py from the oracle) still describes what the oracle computes: two
    two-pass frames with a single-level and with a mip-mapped material table (a regression pin — the specification itself is pinned
    by the hand cases above)"""
    mg, want = load_golden_alpha()
    got = mg.generate()
    assert got == want
    a, b = want["tables"]["single_level"], want["tables"]["mip_mapped"]
    assert [f["vis64_sha"] for f in a] != [f["vis64_sha"] for f in b]  # the mip chains are read

