"""RayScattering::scatter (materials.rs:4-12) in isolation, GPU (rbrt_gpu_scatter: the device function the shade kernels call)
against the oracle's rbrt_ref_kat_scatter, bit for bit: Lambertian (lambertian.rs:11-24), Metal incl. absorption and zero /
large roughness (metal.rs:12-25), Dielectric incl. the fixture's ref_idx 0.2, total internal reflection, exit rays with
cosine > 1 and un-normalised normals (dielectric.rs:11-85)."""
import ctypes as C

import numpy as np
import pytest

import rbrt_b200 as R
from rbrt_b200 import _abi
from rbrt_b200.scene import scatter
from rbrt_b200.vec3 import Vec3

pytestmark = pytest.mark.gpu


def oracle_scatter(O, items, seed):
    lib = O.lib()
    sc, att, od = [], [], []
    for mat, d, p, nrm, pixel, sample, bounce in items:
        a, o = _abi.Vec3C(), _abi.RayC()
        ray = _abi.RayC(_abi.Vec3C(0.0, 0.0, 0.0), _abi.Vec3C(*[float(x) for x in d]))
        ok = lib.rbrt_ref_kat_scatter(mat.to_c(), ray, _abi.Vec3C(*[float(x) for x in p]), _abi.Vec3C(*[float(x) for x in nrm]),
                                      seed, pixel, sample, bounce, C.byref(a), C.byref(o))
        sc.append(ok); att.append((a.x, a.y, a.z)); od.append((o.direction.x, o.direction.y, o.direction.z))
    return np.array(sc, np.int32), np.array(att, np.float32), np.array(od, np.float32)


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def test_scatter_matches_oracle(gpu, oracle):
    rng = np.random.default_rng(17)
    mats = [R.Lambertian(Vec3(0.7, 0.3, 0.2)), R.Metal(Vec3(0.8, 0.8, 0.8), 0.005), R.Metal(Vec3(0.9, 0.9, 0.5), 0.0),
            R.Metal(Vec3(0.5, 0.6, 0.7), 0.9), R.Dielectric(1.8), R.Dielectric(1.5), R.Dielectric(0.2), R.Dielectric(1.0), R.Dielectric(3.5)]
    items = []
    for k in range(6000):
        mat = mats[k % len(mats)]
        d = rng.normal(size=3).astype(np.float32)
        n = rng.normal(size=3).astype(np.float32)
        if k % 3 == 0:
            d /= np.float32(np.linalg.norm(d))                          # unit direction (camera / lambertian rays)
        if k % 4 == 0:
            n *= np.float32(rng.uniform(0.01, 1000.0))                  # sphere normals are p - c: length r (sphere.rs:56)
        if k % 7 == 0:                                                  # grazing incidence: total internal reflection on exit rays
            t = np.cross(n, rng.normal(size=3)).astype(np.float32)
            d = (t / np.float32(np.linalg.norm(t)) + np.float32(0.02) * n / np.float32(np.linalg.norm(n)) * np.float32(rng.choice([-1, 1]))).astype(np.float32)
        if k % 11 == 0:
            d = (-n).astype(np.float32)                                 # normal incidence
        p = (rng.normal(size=3) * 20).astype(np.float32)
        items.append((mat, d, p, n, int(rng.integers(0, 2 ** 21)), int(rng.integers(0, 1024)), int(rng.integers(1, 51))))
    seed = 0x5EED0123456789
    g_sc, g_att, g_od = scatter(items, seed)
    o_sc, o_att, o_od = oracle_scatter(oracle, items, seed)
    assert np.array_equal(g_sc, o_sc)
    assert np.array_equal(bits(g_att), bits(o_att))
    ne = (bits(g_od) != bits(o_od)).any(axis=1)
    assert not ne.any(), f"{int(ne.sum())} out directions differ, first item {int(np.argmax(ne))}: {g_od[ne][0]} vs {o_od[ne][0]}"
    # every branch was taken: metal absorbed and not, glass reflected and refracted, total internal reflection present
    kinds = np.array([k % len(mats) for k in range(len(items))])
    metal = np.isin(kinds, [1, 2, 3])
    assert 0 < g_sc[metal].sum() < metal.sum()
    assert g_sc[~metal].all()
    glass = np.nonzero(kinds >= 4)[0]
    def is_reflection(i):                                                # out == d^ - 2 (d^.n^) n^ (materials.rs:32-37), else it was refracted
        d = np.asarray(items[i][1], np.float64); n = np.asarray(items[i][3], np.float64)
        d /= np.linalg.norm(d); n /= np.linalg.norm(n)
        r = d - 2 * d.dot(n) * n
        return np.allclose(g_od[i], r / np.linalg.norm(r), atol=1e-4)
    refl = np.array([is_reflection(i) for i in glass])
    assert 0.05 * len(glass) < refl.sum() < 0.95 * len(glass), "both arms of the dielectric coin / total internal reflection must occur"


def test_scatter_edge_cases(gpu, oracle):
    """Hand-picked: ref_idx 0.2 from inside and outside (example_scene.yaml:27-28), discriminant <= 0 (reflect_prob 1),
    NaN-producing degenerate inputs must agree too (NaN bit patterns included)."""
    g = R.Dielectric(0.2)
    items = [
        (g, (0.0, -1.0, 0.0), (0, 0, 0), (0.0, 1.0, 0.0), 1, 0, 1),            # entering, a < 0: ni_over_nt = 5 -> total internal reflection
        (g, (0.3, -1.0, 0.1), (1, 2, 3), (0.0, 1.0, 0.0), 2, 1, 2),
        (g, (0.0, 1.0, 0.0), (0, 0, 0), (0.0, 1.0, 0.0), 3, 2, 3),             # leaving, a > 0: cosine = 0.2 a
        (g, (0.9, 0.1, 0.0), (0, 0, 0), (0.0, 2.5, 0.0), 4, 3, 4),
        (R.Dielectric(1.8), (1.0, 0.05, 0.0), (0, 0, 0), (0.0, 1.0, 0.0), 5, 4, 5),   # leaving at grazing angle: discriminant < 0
        (R.Dielectric(1.8), (0.0, 0.0, 0.0), (0, 0, 0), (0.0, 1.0, 0.0), 6, 5, 6),    # zero direction -> NaN
        (R.Lambertian(Vec3(1, 1, 1)), (0.0, -1.0, 0.0), (0, 0, 0), (0.0, 0.0, 0.0), 7, 6, 7),   # zero normal -> NaN
        (R.Metal(Vec3(1, 1, 1), 0.0), (0.0, -1.0, 0.0), (0, 0, 0), (0.0, 1.0, 0.0), 8, 7, 8),
        (R.Metal(Vec3(1, 1, 1), 0.0), (0.0, 1.0, 0.0), (0, 0, 0), (0.0, 1.0, 0.0), 9, 8, 9),    # reflected INTO the surface: absorbed
    ]
    for seed in (0, 1, 2 ** 64 - 1):
        g_sc, g_att, g_od = scatter(items, seed)
        o_sc, o_att, o_od = oracle_scatter(oracle, items, seed)
        assert np.array_equal(g_sc, o_sc) and np.array_equal(bits(g_att), bits(o_att))
        # NaN payloads may differ between CPU and GPU; compare NaN-ness, and bits elsewhere
        gn, on = np.isnan(g_od), np.isnan(o_od)
        assert np.array_equal(gn, on)
        assert np.array_equal(bits(g_od)[~gn], bits(o_od)[~on])
    assert g_sc[8] == 0 and g_sc[7] == 1
    assert _abi.lib().rbrt_gpu_scatter(None, 1, 0, None) == _abi.E_INVALID
