"""The CPU oracle against every known-answer test the reference holds for the hot path
(SURVEY.md §4 / §8c).  Each test cites the reference test it restates; exact-equality asserts in the
reference are exact-equality asserts here (bit patterns of f32)."""
import ctypes as C
import math

import numpy as np
import pytest

from rbrt_b200 import _abi

V = _abi.Vec3C


def f32(x):
    return float(np.float32(x))


def tup(v):
    return (v.x, v.y, v.z)


def farr(vals):
    return (C.c_float * len(vals))(*vals)


def test_sphere_intersection(oracle):  # sphere.rs:76-112
    L = oracle.lib()
    h = _abi.HitC()
    r = _abi.RayC(V(0, 0, 0), V(0, 0, -1))
    assert L.rbrt_ref_kat_sphere(V(0, 0, -10), 1.0, r, 0.001, 1000.0, h) == 1
    assert tup(h.point) == (0.0, 0.0, -9.0) and tup(h.normal) == (0.0, 0.0, 1.0)
    r = _abi.RayC(V(0, 0, -15), V(0, 0, 1))
    assert L.rbrt_ref_kat_sphere(V(0, 0, -10), 1.0, r, 0.001, 1000.0, h) == 1
    assert tup(h.point) == (0.0, 0.0, -11.0) and tup(h.normal) == (0.0, 0.0, -1.0)


def test_triangle_normal(oracle):  # triangle.rs:449-475
    L = oracle.lib()
    assert tup(L.rbrt_ref_kat_triangle_normal(V(1, 0, 0), V(1, 1, 0), V(0, 0, 0))) == (0.0, 0.0, 1.0)
    n = L.rbrt_ref_kat_triangle_normal(V(1, 0, 0), V(1, 0, 1), V(0, 1, 0))
    assert tup(n) == tup(L.rbrt_ref_kat_normalize(V(-1, -1, 0)))


def test_mesh_aabbox(oracle):  # aabbox.rs:95-108
    L = oracle.lib()
    lo, hi = V(), V()
    L.rbrt_ref_kat_min_max_3d(farr([1, 0, 0, 1, 0, 1, 0, 1, 0]), 1, lo, hi)
    assert tup(lo) == (0.0, 0.0, 0.0) and tup(hi) == (1.0, 1.0, 1.0)


def test_reflection(oracle):  # materials.rs:49-59
    L = oracle.lib()
    n = L.rbrt_ref_kat_normalize(V(1, 1, 1))
    r = L.rbrt_ref_kat_reflect(V(1, 1, 1), V(1, 1, 1))
    assert tup(r) == tuple(f32(-1.0 * c) for c in tup(n))
    r = L.rbrt_ref_kat_reflect(V(1, 1, 0), V(-1, 0, 0))
    assert tup(r) == (f32(-0.7071068), f32(0.7071068), 0.0)


def test_refraction(oracle):  # dielectric.rs:93-115
    L = oracle.lib()
    out = V()
    d, n = L.rbrt_ref_kat_normalize(V(1, 1, 0)), L.rbrt_ref_kat_normalize(V(-1, 0, 0))
    assert L.rbrt_ref_kat_refract(d, n, 1.4, out) == 1
    assert tup(out) == (f32(0.14142191), f32(0.9899495), 0.0)


def test_random_points_in_unit_sphere(oracle):  # materials.rs:43-47 (via Lambertian scatter: target - point = n + p)
    L = oracle.lib()
    # a metal with roughness 1 and normal == -incoming gives dir = normalize(reflect + p); instead pin the
    # sampler through lambertian: out.direction * |n^ + p| is not recoverable, so sample the Philox stream directly
    key = (C.c_uint32 * 2)(7, 0)
    for pixel in range(20):
        rnd = 0
        while True:
            ctr = (C.c_uint32 * 4)(pixel, 0, 1, rnd)
            o = (C.c_uint32 * 4)()
            L.rbrt_ref_kat_philox(ctr, key, o)
            p = [np.float32(2.0) * (np.float32(o[i] >> 8) * np.float32(1.0 / 16777216.0)) - np.float32(1.0) for i in range(3)]
            ln = np.sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2], dtype=np.float32)
            rnd += 1
            if not ln > 1.0:
                break
        assert ln <= 1.0 and rnd < 64


def test_philox_known_answers(oracle):
    """Philox4x32-10 known-answer vectors of the Random123 distribution (kat_vectors: philox4x32 10)."""
    L = oracle.lib()
    kats = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
            ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
            ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
             (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, exp in kats:
        o = (C.c_uint32 * 4)()
        L.rbrt_ref_kat_philox((C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), o)
        assert tuple(o) == exp


AX, AY, AZ = [1.0, 0.0, 3.0, 2.0], [0.0, 1.0, 4.0, 6.0], [0.0, 0.0, 4.0, 3.0]
BX, BY, BZ = [0.0, 0.0, 1.0, 2.0], [1.0, 0.0, -2.0, 1.0], [0.0, 1.0, 3.0, -2.0]
CX, CY, CZ = [0.0, 1.0, 20.0, -15.0], [0.0, 0.0, -5.0, 10.0], [1.0, 0.0, -10.0, -10.0]
DOT = [0.0, 0.0, 7.0, 4.0]


def test_avx_cross_and_dot(oracle):  # vec3_avx.rs:60-110 (8 lanes = the 4-lane vectors twice)
    L = oracle.lib()
    a, b = farr(AX * 2 + AY * 2 + AZ * 2), farr(BX * 2 + BY * 2 + BZ * 2)
    out = (C.c_float * 24)()
    L.rbrt_ref_kat_avx_cross(a, b, out)
    assert list(out) == CX * 2 + CY * 2 + CZ * 2
    d = (C.c_float * 8)()
    L.rbrt_ref_kat_avx_dot(a, b, d)
    assert list(d) == DOT * 2


def test_sse_cross_and_dot(oracle):  # vec3_sse.rs:59-172
    L = oracle.lib()
    a, b = farr(AX + AY + AZ), farr(BX + BY + BZ)
    out = (C.c_float * 12)()
    L.rbrt_ref_kat_sse_cross(a, b, out)
    assert list(out) == CX + CY + CZ
    d = (C.c_float * 4)()
    L.rbrt_ref_kat_sse_dot(a, b, d)
    assert list(d) == DOT
    # test_trivial_sse_*: unit axes
    a, b = farr([1, 0, 0, 0] + [0, 1, 0, 0] + [0, 0, 1, 0]), farr([0, 0, 1, 0] + [1, 0, 0, 0] + [0, 1, 0, 0])
    L.rbrt_ref_kat_sse_cross(a, b, out)
    assert list(out) == [0, 1, 0, 0] + [0, 0, 1, 0] + [1, 0, 0, 0]


def test_vec3_algebra(oracle):  # vec3.rs:166-215
    L = oracle.lib()
    assert tup(L.rbrt_ref_kat_cross(V(1, 0, 0), V(0, 1, 0))) == (0.0, 0.0, 1.0)
    assert L.rbrt_ref_kat_length(L.rbrt_ref_kat_normalize(V(5, 2, 3))) == 1.0
    assert L.rbrt_ref_kat_dot(V(1, 2, 3), V(1, 2, 3)) == 14.0


def test_rotate_yaw(oracle):  # vec3.rs:217-341: 19 cases, residuum < 1e-6
    L = oracle.lib()
    a = 0.7071067657322372
    r45 = float(np.float32(math.radians(np.float32(45.0))))
    X, Y, Z = (1, 0, 0), (0, 1, 0), (0, 0, 1)
    cases = [(X, (0, 0, r45), (a, a, 0)), (X, (0, 0, -r45), (a, -a, 0)), (Y, (0, 0, r45), (-a, a, 0)),
             (Y, (0, 0, -r45), (a, a, 0)), (Z, (0, 0, r45), Z), (Z, (0, 0, -r45), Z),
             (X, (0, r45, 0), X), (X, (0, -r45, 0), X), (Y, (0, r45, 0), (0, a, a)), (Y, (0, -r45, 0), (0, a, -a)),
             (Z, (0, r45, 0), (0, -a, a)), (Z, (0, -r45, 0), (0, a, a)),
             (X, (r45, 0, 0), (a, a, 0)), (X, (-r45, 0, 0), (a, -a, 0)), (Y, (r45, 0, 0), (-a, a, 0)),
             (Y, (-r45, 0, 0), (a, a, 0)), (Z, (r45, 0, 0), Z), (Z, (r45, 0, 0), Z), (Z, (-r45, 0, 0), Z)]
    assert len(cases) == 19
    for p, rot, exp in cases:
        q = L.rbrt_ref_kat_rotate_point(V(*p), V(*rot))
        res = math.sqrt(sum((e - c) ** 2 for e, c in zip(exp, tup(q))))
        assert res < 1e-6, (p, rot, exp, tup(q))


def test_as_u8_saturates(oracle):  # lib.rs:118-120: Rust `as u8` saturates, NaN -> 0
    L = oracle.lib()
    for v, e in [(256.0, 255), (255.0, 255), (254.99, 254), (0.999, 0), (-3.0, 0), (float("nan"), 0), (float("inf"), 255), (17.7, 17)]:
        assert L.rbrt_ref_kat_as_u8(v) == e


def test_host_math_matches_oracle(oracle):
    """rbrt_camera_new / rbrt_transform_vertices of the product library (host f32 code, no GPU) are
    bit-identical to the oracle's restatement of cam.rs:22-62 and mesh.rs:102-112."""
    G, L = _abi.lib(), oracle.lib()
    rng = np.random.default_rng(5)
    for _ in range(50):
        pos, la, up = (V(*rng.normal(size=3).astype(np.float32)) for _ in range(3))
        h, w, f = int(rng.integers(1, 3000)), int(rng.integers(1, 3000)), float(np.float32(rng.uniform(10, 80)))
        a, b = _abi.CameraC(), _abi.CameraC()
        assert G.rbrt_camera_new(pos, la, up, h, w, f, a) == 0
        assert L.rbrt_ref_camera_new(pos, la, up, h, w, f, b) == 0
        assert bytes(a) == bytes(b)
    pts = rng.normal(size=(1000, 3)).astype(np.float32)
    p1, p2 = pts.copy(), pts.copy()
    rot, tr = V(0.3, -1.2, 2.5), V(5.0, -1.8, -12.5)
    G.rbrt_transform_vertices(p1.ctypes.data_as(_abi.P(C.c_float)), 1000, 45.0, rot, tr)
    L.rbrt_ref_transform_vertices(p2.ctypes.data_as(_abi.P(C.c_float)), 1000, 45.0, rot, tr)
    assert np.array_equal(p1.view(np.uint32), p2.view(np.uint32))
