"""Parity of the CUDA closest-hit query (Scene::hit, scene.rs:19-43) against the oracle, through the C-ABI
(rbrt_gpu_trace_rays).  Bar: BIT-EXACT — kind, element, ORIGINAL triangle index, and the f32 bit patterns of
t, dist, point and normal — for both the brute-force kernel (the reference's own loop) and the LBVH kernel."""
import numpy as np
import pytest

import rbrt_b200 as R
from rbrt_b200 import _abi, synth
from rbrt_b200.vec3 import Vec3

from . import golden_util as G
from . import scenes as S

pytestmark = pytest.mark.gpu
MODES = [_abi.TRACE_BRUTE, _abi.TRACE_BVH, _abi.TRACE_WAVEFRONT]   # WAVEFRONT = the renderer's own stage A + k_trace kernels


def assert_same(a, b, what):
    eq = S.hits_equal(a, b)
    if not eq.all():
        i = np.nonzero(~eq)[0]
        raise AssertionError(f"{what}: {len(i)} of {len(a)} hits differ; first: got {a[i[0]]} want {b[i[0]]}")


@pytest.mark.parametrize("name", G.NAMES)
@pytest.mark.parametrize("mode", MODES)
def test_golden_hits(gpu, name, mode):
    z, scene, cam = G.load(name)
    n_px = cam.img_width_pix * cam.img_height_pix
    mine = R.primary_rays(cam, int(z["seed"]), 0)
    assert np.array_equal(mine.view(np.uint32), z["rays"][:n_px].view(np.uint32)), "primary rays (cam.rs:64-82 + Philox)"
    assert_same(scene.hit(z["rays"], mode), z["hits"], f"{name} mode {mode}")


@pytest.mark.parametrize("leaf_size", [1, 2, 4, 8])
@pytest.mark.parametrize("lanes", [8, 4])
def test_bvh_options_do_not_change_results(gpu, oracle, leaf_size, lanes):
    for n_keep in (1275, 1277, 1280, 5, 3, 1):
        scene = S.small_mesh_scene(3, n_keep, simd_lanes=lanes, leaf_size=leaf_size)
        cam = S.example_camera(96, 72)
        rays = np.concatenate([R.primary_rays(cam, 3, 0), S.random_rays(4096, (5.0, 1.4, -12.5), 4.0, 1)], 0)
        ref = oracle.OracleScene.from_scene(scene).hit(rays)
        info = scene.info()
        r = n_keep % lanes
        assert info["num_triangles_tested"] == (n_keep if (r == 0 or 2 * r >= lanes) else n_keep - r)
        for mode in MODES:
            assert_same(scene.hit(rays, mode), ref, f"n={n_keep} lanes={lanes} leaf={leaf_size} mode={mode}")


def test_quirk_cases_on_gpu(gpu, oracle):
    """The hand-derived cases of tests/test_oracle_quirks.py, GPU vs oracle."""
    tri = ((-1, -1, -5), (1, -1, -5), (0, 1, -5))
    cases = []
    def mesh_scene(tris, **kw):
        sc = R.Scene(**kw)
        sc.triangle_meshes.append(R.TriangleMesh.from_triangles(np.array(tris, np.float32), R.Lambertian(Vec3(1, 1, 1))))
        return sc
    cases.append((mesh_scene([tri] * 16), [[0, 0, 0, 0, 0, -1], [0, 0, -10, 0, 0, 1], [0, 0, 0, 0, 0, -2], [0, 0, 0, 0.01, 0.01, -1]]))
    for z in (-999.0, -1000.0, -0.0009, -0.0011):
        cases.append((mesh_scene([tuple((x * 400, y * 400, z) for x, y, _ in tri)] * 8), [[0, 0, 0, 0, 0, -1]]))
    for n in range(1, 18):
        cases.append((mesh_scene([((-1, -1, -50 + i), (1, -1, -50 + i), (0, 1, -50 + i)) for i in range(n)]), [[0, 0, 0, 0, 0, -1]]))
        cases.append((mesh_scene([((-1, -1, -50 + i), (1, -1, -50 + i), (0, 1, -50 + i)) for i in range(n)], simd_lanes=4), [[0, 0, 0, 0, 0, -1]]))
    sc = mesh_scene([tri] * 8)
    sc.elements.append(R.Sphere(Vec3(0, 0, -6.0), 1.0, R.Lambertian(Vec3(1, 1, 1))))       # exact dist tie: sphere wins
    sc.elements.append(R.Sphere(Vec3(0, 0, 0), 1.0, R.Dielectric(1.5)))
    cases.append((sc, [[0, 0, 0, 0, 0, -1], [0, 0, 1.0005, 0, 0, -1], [0, 0, 0.5, 0, 0, 1], [0, 0, 30, 0, 0, 1]]))
    for scene, rays in cases:
        rays = np.array(rays, np.float32)
        ref = oracle.OracleScene.from_scene(scene).hit(rays)
        for mode in MODES:
            assert_same(scene.hit(rays, mode), ref, f"quirk case mode {mode}")


def test_degenerate_rays_and_nan_count(gpu, oracle):
    scene = S.quirk_scene()
    rays = S.random_rays(20000, (0.0, 1.5, -4.0), 3.0, 9)
    rays[:8, 3:] = 0.0                                  # zero direction: a = 0 -> sol = NaN or 0/0: the reference panics
    rays[8:16, 3] = np.nan
    rays[16:24, 3:] = np.float32([0, 0, 1e-30])
    rays[24:32, :3] = np.float32([1e6, 1e6, 1e6])
    ref = oracle.OracleScene.from_scene(scene).hit(rays)
    for mode in MODES:
        st = {}
        got = scene.hit(rays, mode, stats=st)
        assert_same(got, ref, f"degenerate rays mode {mode}")
        assert st["nan_rays"] >= 8


def test_multi_mesh_and_empty_inputs(gpu, oracle):
    scene = S.quirk_scene()
    assert scene.info()["num_meshes"] == 3 and scene.info()["num_triangles"] == 23
    assert len(scene.hit(np.zeros((0, 6), np.float32))) == 0
    empty = R.Scene()
    h = empty.hit(np.float32([[0, 0, 0, 0, 0, -1]]))
    assert h["kind"][0] == -1
    only_mesh = R.Scene()
    only_mesh.triangle_meshes.append(R.TriangleMesh.from_triangles(synth.displaced_icosphere(2, 2.0, (0, 0, -8)), R.Metal(Vec3(1, 1, 1), 0.1)))
    rays = S.random_rays(8192, (0, 0, -8), 3.0, 4)
    assert_same(only_mesh.hit(rays), oracle.OracleScene.from_scene(only_mesh).hit(rays), "mesh-only scene")


def test_medium_mesh_vs_oracle(gpu, oracle):
    """20 480 triangles, camera rays + bounce-like rays starting ON the surface (self-intersection window)."""
    scene = S.small_mesh_scene(5)
    cam = S.example_camera(160, 120)
    prim = R.primary_rays(cam, 21, 0)
    first = scene.hit(prim)
    on = first["kind"] >= 0
    rng = np.random.default_rng(2)
    d2 = rng.normal(size=(on.sum(), 3)).astype(np.float32)
    d2 /= np.linalg.norm(d2, axis=1, keepdims=True)
    bounce = np.concatenate([first["point"][on], d2], 1)
    rays = np.concatenate([prim, bounce], 0)
    ref = oracle.OracleScene.from_scene(scene).hit(rays)
    for mode in MODES:
        assert_same(scene.hit(rays, mode), ref, f"medium mesh mode {mode}")


def c2_scene(tmp=None):
    import os
    d = synth.cache_dir()
    obj = os.path.join(d, "standin6.obj")
    if not os.path.exists(obj):
        synth.write_bunny_standin(obj, 6)
    return R.create_scene_from_scene_blueprint(synth.example_scene_blueprint(obj))


def test_c2_full_size_bvh_equals_brute(gpu):
    """Config C2 at BASELINE.json's full size (1024x768 primary rays of an 81 920-triangle mesh through the OBJ
    loader): the oracle would need minutes, so the size-independent property is used — the BVH kernel must agree
    bit for bit with the GPU brute-force kernel, which is itself pinned to the oracle by the tests above."""
    scene = c2_scene()
    cam = S.example_camera(1024, 768)
    prim = R.primary_rays(cam, 0, 0)
    st = {}
    bvh = scene.hit(prim, _abi.TRACE_BVH, stats=st)
    brute = scene.hit(prim, _abi.TRACE_BRUTE)
    assert_same(bvh, brute, "C2 primary rays")
    assert_same(scene.hit(prim, _abi.TRACE_WAVEFRONT), brute, "C2 primary rays through k_trace")
    assert (bvh["kind"] == 1).sum() > 50000
    on = bvh["kind"] == 1
    n = bvh["normal"][on]
    refl = prim[on, 3:] - 2 * (prim[on, 3:] * n).sum(1, keepdims=True) * n
    sec = np.concatenate([bvh["point"][on], refl.astype(np.float32)], 1)[:200000]
    sec_brute = scene.hit(sec, _abi.TRACE_BRUTE)
    assert_same(scene.hit(sec, _abi.TRACE_BVH), sec_brute, "C2 reflected rays")
    assert_same(scene.hit(sec, _abi.TRACE_WAVEFRONT), sec_brute, "C2 reflected rays through k_trace")
    assert 0 < st["node_visits"] < 60 * len(prim) and st["tri_tests"] < 20 * len(prim)


def test_c3_million_triangles_bvh_equals_brute(gpu):
    """Config C3's mesh (displaced icosphere, subdivision 8 = 1 310 720 triangles, radius 40): BVH == brute on a
    stratified subset of the 1920x1080 primary rays plus reflected rays."""
    scene, cam = S.big_scene(8)
    info = scene.info()
    assert info["num_triangles_tested"] == 1310720
    prim = R.primary_rays(cam, 1, 0)[::23]
    bvh = scene.hit(prim, _abi.TRACE_BVH)
    assert_same(bvh, scene.hit(prim, _abi.TRACE_BRUTE), "C3 primary subset")
    # the hot kernel itself (k_trace: persistent warps, warp-voted traversal, dynamic fetch) on ALL 2 073 600 primary rays
    # against the plain one-lane-one-ray traversal, and on the subset against brute force
    full = R.primary_rays(cam, 1, 0)
    st = {}
    wf = scene.hit(full, _abi.TRACE_WAVEFRONT, stats=st)
    assert_same(wf[::23], bvh, "C3 primary subset through k_trace")
    assert_same(wf, scene.hit(full, _abi.TRACE_BVH), "C3 all primary rays: k_trace vs plain traversal")
    assert st["traversed_rays"] > 500000 and st["node_visits"] > st["traversed_rays"]
    on = bvh["kind"] == 1
    assert on.sum() > 10000
    n = bvh["normal"][on]
    refl = prim[on, 3:] - 2 * (prim[on, 3:] * n).sum(1, keepdims=True) * n
    sec = np.concatenate([bvh["point"][on], refl.astype(np.float32)], 1)[:30000]
    sec_brute = scene.hit(sec, _abi.TRACE_BRUTE)
    assert_same(scene.hit(sec, _abi.TRACE_BVH), sec_brute, "C3 reflected subset")
    assert_same(scene.hit(sec, _abi.TRACE_WAVEFRONT), sec_brute, "C3 reflected subset through k_trace")
    # the same tree without the SAH rotation pass gives the same answers
    plain = S.big_scene(8, sah=False)[0]
    assert_same(plain.hit(sec, _abi.TRACE_WAVEFRONT), sec_brute, "C3 reflected subset, plain LBVH")
    plain.close()


def test_builder_edge_cases(gpu, oracle):
    """BVH builder corner cases, each compared with the oracle bit for bit: all triangles identical (equal Morton codes, a
    binary tree kept together only by the position tie-break), a flat mesh (zero extent in one axis), zero-area and
    needle triangles, coordinates around 1e5, twenty small meshes in one scene, 300 spheres."""
    rng = np.random.default_rng(11)
    cases = []
    tri = np.float32([[-1, -1, -5], [1, -1, -5], [0, 1, -5]])
    cases.append(("identical x 1000", [np.tile(tri, (1000, 1, 1))], (0, 0, -5), 3.0))
    flat = rng.uniform(-3, 3, size=(4096, 3, 3)).astype(np.float32)
    flat[:, :, 2] = -7.0
    cases.append(("flat z = -7", [flat], (0, 0, -7), 4.0))
    deg = synth.displaced_icosphere(3, 2.0, (0, 0, -8)).copy()
    deg[::7, 1] = deg[::7, 0]                                    # zero-area triangles (two equal vertices)
    deg[3::11, 2] = deg[3::11, 0] + np.float32(1e-4)             # needles
    cases.append(("degenerate triangles", [deg], (0, 0, -8), 3.0))
    far = synth.displaced_icosphere(3, 50.0, (1e5, -2e5, 3e5))
    cases.append(("coordinates ~1e5", [far], (1e5, -2e5, 3e5), 80.0))
    many = [synth.displaced_icosphere(1, 0.7, (float(x), float(y), -10.0 - (x + y) % 3)) for x in range(-4, 6, 2) for y in range(-3, 5, 2)]
    cases.append(("20 meshes", many, (0, 0, -10), 6.0))
    for name, meshes, center, spread in cases:
        scene = R.Scene()
        for i, m in enumerate(meshes):
            scene.triangle_meshes.append(R.TriangleMesh.from_triangles(m, [R.Lambertian(Vec3(0.5, 0.5, 0.5)), R.Metal(Vec3(0.9, 0.9, 0.9), 0.1), R.Dielectric(1.5)][i % 3]))
        rays = S.random_rays(30000, center, spread, 5)
        ref = oracle.OracleScene.from_scene(scene).hit(rays)
        assert (ref["kind"] == 1).sum() > 100, name
        for mode in MODES:
            assert_same(scene.hit(rays, mode), ref, f"{name} mode {mode}")
        scene.close()
    scene = R.Scene()
    for i in range(300):
        c = rng.uniform(-20, 20, size=3)
        scene.elements.append(R.Sphere(Vec3(c[0], c[1], c[2] - 40), float(rng.uniform(0.3, 2.0)), R.Lambertian(Vec3(0.5, 0.5, 0.5))))
    rays = S.random_rays(20000, (0, 0, -40), 25.0, 6)
    assert_same(scene.hit(rays), oracle.OracleScene.from_scene(scene).hit(rays), "300 spheres")


def mixed_element_scene():
    """Spheres and BasicTriangles interleaved in Scene.elements, plus a mesh."""
    rng = np.random.default_rng(8)
    els = [R.Sphere(Vec3(0.0, -1000.0, -5.0), 1000.0, R.Lambertian(Vec3(0.02, 0.2, 0.1)))]
    mats = [R.Lambertian(Vec3(0.7, 0.2, 0.2)), R.Metal(Vec3(0.9, 0.9, 0.9), 0.02), R.Dielectric(1.5)]
    for i in range(12):
        c = rng.uniform(-4, 4, size=3) + np.array([0, 4.5, -9])
        if i % 2:
            els.append(R.Sphere(Vec3(*c), float(rng.uniform(0.4, 1.0)), mats[i % 3]))
        else:
            p = c + rng.uniform(-1.5, 1.5, size=(3, 3))
            els.append(R.BasicTriangle.new([tuple(v) for v in p], mats[i % 3]))
    els.append(R.BasicTriangle.new([(-30, 0.5, -30), (30, 0.5, -30), (0, 30, -30)], R.Metal(Vec3(0.8, 0.8, 0.9), 0.0)))   # a big mirror behind
    scene = R.Scene(elements=els)
    scene.triangle_meshes.append(R.TriangleMesh.from_triangles(synth.displaced_icosphere(3, 1.5, (2.5, 1.5, -7.0)), R.Dielectric(1.5)))
    return scene


def test_basic_triangle_elements(gpu, oracle):
    scene = mixed_element_scene()
    cam = S.example_camera(128, 96)
    rays = np.concatenate([R.primary_rays(cam, 4, 0), S.random_rays(20000, (0, 4, -9), 5.0, 3)], 0)
    ref = oracle.OracleScene.from_scene(scene).hit(rays)
    assert (ref["kind"] == 2).sum() > 500 and (ref["kind"] == 0).sum() > 500 and (ref["kind"] == 1).sum() > 100
    for mode in MODES:
        assert_same(scene.hit(rays, mode), ref, f"mixed elements mode {mode}")
    big = R.Scene(elements=[R.BasicTriangle.new(((-900, -900, -1500), (900, -900, -1500), (0, 900, -1500)), R.Lambertian(Vec3(1, 1, 1)))])
    h = big.hit(np.float32([[0, 0, 0, 0, 0, -1]]))
    assert h["kind"][0] == 2 and abs(h["t"][0] - 1500.0) < 1e-2                       # no upper cap on t for BasicTriangle (triangle.rs:118)


def test_tangent_sphere_behind_the_ray_is_hit(gpu, oracle):
    """Quirk Q16 (tests/test_oracle_quirks.py): with a discriminant of exactly 0 the reference reports a sphere BEHIND the ray.  No
    shortcut may treat "points away from the sphere" as a miss, in any of the three trace modes, with few or many spheres."""
    mat = R.Lambertian(Vec3(0.5, 0.5, 0.5))
    rays = np.array([[0, 0, 0, 0, 0, -1], [0, 0, 0, 0, 0, -2.5], [0, 0, 1, 0, 0, -1], [0.5, 0, 0, 0, 0, -1]], np.float32)
    for n_extra in (0, 30):
        sc = R.Scene()
        sc.elements.append(R.Sphere(Vec3(1.0, 0.0, 5.0), 1.0, mat))    # tangent to the z axis, behind the rays
        rng = np.random.default_rng(1)
        for k in range(n_extra):
            c = rng.uniform(-30, 30, size=3)
            sc.elements.append(R.Sphere(Vec3(float(c[0]) + 100.0, float(c[1]), float(c[2])), 0.7, mat))   # far off to the side
        want = oracle.OracleScene.from_scene(sc).hit(rays)
        assert want["kind"][0] == 0 and want["t"][0] == -5.0 and want["kind"][2] == 0 and want["kind"][3] == -1
        for mode in MODES:
            assert_same(sc.hit(rays, mode), want, f"{n_extra} extra spheres, mode {mode}")


def test_small_mesh_near_the_world_origin_far_ray_origins(gpu, oracle):
    """Rays that start up to 990 units away from a small mesh near the world origin, aimed at triangle edges (the reference accepts t < 1000,
    triangle.rs:146): `o - v0` carries ~ulp(|o|) there, more than a box padding relative to the mesh's own coordinates covers.  The builder keeps the
    padding above 2^-21 * (mx + 1000) (bvh_build.cu k_mesh_setup; found and sized with the host build of the traversal,
    tests/test_device_source_on_host.py).  All three trace modes against the oracle, bit for bit."""
    for subdiv, size, centre, dist in ((1, 0.2, (0, 0, 0), 990.0), (0, 0.05, (0, 0, 0), 100.0), (0, 0.05, (0, 0, 0), 990.0), (2, 0.3, (0.05, 0.02, -0.03), 990.0),
                                       (2, 1.0, (0, 0, 0), 990.0)):
        tris = synth.displaced_icosphere(subdiv, size, centre)
        scene = R.Scene()
        scene.triangle_meshes.append(R.TriangleMesh.from_triangles(tris, R.Lambertian(Vec3(0.5, 0.5, 0.5))))
        rays = S.edge_aimed_rays(tris, 60000, dist, seed=subdiv + int(dist))
        want = oracle.OracleScene.from_scene(scene).hit(rays)
        assert (want["kind"] == 1).sum() > 30000
        for mode in MODES:
            assert_same(scene.hit(rays, mode), want, f"size {size} dist {dist} mode {mode}")
