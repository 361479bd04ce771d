"""ctypes binding of the CPU ORACLE (oracle/rbrt_oracle.cpp).  TEST INFRASTRUCTURE ONLY: imported by
tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs; never by rbrt_b200/."""
import ctypes as C
import os
import subprocess

import numpy as np

from rbrt_b200 import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "build", "librbrt_oracle.so")
P = C.POINTER

_SIG = {
    "rbrt_ref_camera_new": (C.c_int, [_abi.Vec3C, _abi.Vec3C, _abi.Vec3C, C.c_uint32, C.c_uint32, C.c_float, P(_abi.CameraC)]),
    "rbrt_ref_transform_vertices": (C.c_int, [P(C.c_float), C.c_uint64, C.c_float, _abi.Vec3C, _abi.Vec3C]),
    "rbrt_ref_scene_create": (C.c_int, [P(_abi.SphereDescC), C.c_uint32, P(_abi.MeshDescC), C.c_uint32, P(_abi.SceneOptsC), P(C.c_void_p)]),
    "rbrt_ref_scene_create_elements": (C.c_int, [P(_abi.ElementRefC), C.c_uint32, P(_abi.SphereDescC), C.c_uint32, P(_abi.TriangleDescC), C.c_uint32,
                                                P(_abi.MeshDescC), C.c_uint32, P(_abi.SceneOptsC), P(C.c_void_p)]),
    "rbrt_ref_scene_destroy": (C.c_int, [C.c_void_p]),
    "rbrt_ref_render": (C.c_int, [C.c_void_p, P(_abi.CameraC), C.c_uint32, P(_abi.RenderOptsC), C.c_void_p, P(_abi.StatsC)]),
    "rbrt_ref_render_hdr": (C.c_int, [C.c_void_p, P(_abi.CameraC), C.c_uint32, P(_abi.RenderOptsC), C.c_void_p, P(_abi.StatsC)]),
    "rbrt_ref_render_accum": (C.c_int, [C.c_void_p, P(_abi.CameraC), C.c_uint32, P(_abi.RenderOptsC), C.c_void_p, P(_abi.StatsC)]),
    "rbrt_ref_render_subset": (C.c_int, [C.c_void_p, P(_abi.CameraC), C.c_uint32, P(_abi.RenderOptsC), C.c_uint32, C.c_uint32, C.c_void_p, P(_abi.StatsC)]),
    "rbrt_ref_finalize": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]),
    "rbrt_ref_trace_rays": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, P(_abi.StatsC)]),
    "rbrt_ref_primary_rays": (C.c_int, [P(_abi.CameraC), C.c_uint64, C.c_uint32, C.c_void_p]),
    "rbrt_ref_last_error": (C.c_char_p, []),
    "rbrt_ref_version": (C.c_char_p, []),
    "rbrt_ref_set_threads": (C.c_int, [C.c_uint]),
    "rbrt_ref_hardware_threads": (C.c_uint, []),
    "rbrt_ref_kat_philox": (None, [P(C.c_uint32), P(C.c_uint32), P(C.c_uint32)]),
    "rbrt_ref_kat_reflect": (_abi.Vec3C, [_abi.Vec3C, _abi.Vec3C]),
    "rbrt_ref_kat_refract": (C.c_int, [_abi.Vec3C, _abi.Vec3C, C.c_float, P(_abi.Vec3C)]),
    "rbrt_ref_kat_schlick": (C.c_float, [C.c_float, C.c_float]),
    "rbrt_ref_kat_normalize": (_abi.Vec3C, [_abi.Vec3C]),
    "rbrt_ref_kat_length": (C.c_float, [_abi.Vec3C]),
    "rbrt_ref_kat_dot": (C.c_float, [_abi.Vec3C, _abi.Vec3C]),
    "rbrt_ref_kat_cross": (_abi.Vec3C, [_abi.Vec3C, _abi.Vec3C]),
    "rbrt_ref_kat_rotate_point": (_abi.Vec3C, [_abi.Vec3C, _abi.Vec3C]),
    "rbrt_ref_kat_triangle_normal": (_abi.Vec3C, [_abi.Vec3C, _abi.Vec3C, _abi.Vec3C]),
    "rbrt_ref_kat_min_max_3d": (None, [P(C.c_float), C.c_uint64, P(_abi.Vec3C), P(_abi.Vec3C)]),
    "rbrt_ref_kat_bbox_hit": (C.c_int, [_abi.Vec3C, _abi.Vec3C, _abi.RayC]),
    "rbrt_ref_kat_sphere": (C.c_int, [_abi.Vec3C, C.c_float, _abi.RayC, C.c_float, C.c_float, P(_abi.HitC)]),
    "rbrt_ref_kat_avx_dot": (None, [P(C.c_float), P(C.c_float), P(C.c_float)]),
    "rbrt_ref_kat_avx_cross": (None, [P(C.c_float), P(C.c_float), P(C.c_float)]),
    "rbrt_ref_kat_sse_dot": (None, [P(C.c_float), P(C.c_float), P(C.c_float)]),
    "rbrt_ref_kat_sse_cross": (None, [P(C.c_float), P(C.c_float), P(C.c_float)]),
    "rbrt_ref_kat_scatter": (C.c_int, [_abi.MaterialC, _abi.RayC, _abi.Vec3C, _abi.Vec3C, C.c_uint64, C.c_uint32, C.c_uint32,
                                       C.c_uint32, P(_abi.Vec3C), P(_abi.RayC)]),
    "rbrt_ref_kat_as_u8": (C.c_uint8, [C.c_float]),
}

_lib = None


def build():
    subprocess.run(["make", "-C", _HERE], check=True, stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIG.items():
            f = getattr(l, name)
            f.restype, f.argtypes = res, args
        _lib = l
    return _lib


def check(rc):
    if rc != 0:
        raise RuntimeError(f"oracle error {rc}: {lib().rbrt_ref_last_error().decode()}")


def camera_new(position, look_at, up, h, w, focal):
    c = _abi.CameraC()
    check(lib().rbrt_ref_camera_new(position.to_c(), look_at.to_c(), up.to_c(), h, w, focal, c))
    return c


class OracleScene:
    """Same inputs as rbrt_b200.Scene (spheres, meshes), rendered by the CPU restatement."""

    def __init__(self, elements=(), triangle_meshes=(), simd_lanes=8):
        self.elements, self.triangle_meshes = list(elements), list(triangle_meshes)
        from rbrt_b200.scene import element_arrays
        order, spheres, tris, ne, ns, nt = element_arrays(self.elements)
        nm = len(self.triangle_meshes)
        meshes = (_abi.MeshDescC * max(nm, 1))(*[m.to_c() for m in self.triangle_meshes])
        opts = _abi.SceneOptsC(simd_lanes, 0, 0.0, 0)
        self._h = C.c_void_p()
        check(lib().rbrt_ref_scene_create_elements(order, ne, spheres, ns, tris, nt, meshes, nm, opts, C.byref(self._h)))

    @staticmethod
    def from_scene(scene):
        return OracleScene(scene.elements, scene.triangle_meshes, scene.simd_lanes)

    def __del__(self):
        try:
            if self._h:
                lib().rbrt_ref_scene_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def hit(self, rays):
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 6)
        hits = np.zeros(len(rays), dtype=_abi.HIT_DTYPE)
        check(lib().rbrt_ref_trace_rays(self._h, rays.ctypes.data, len(rays), 0, hits.ctypes.data, None))
        return hits

    def render_hdr(self, cam_c, spp, opts=None, stats=None):
        out = np.empty((cam_c.img_height_pix, cam_c.img_width_pix, 3), np.float32)
        st = _abi.StatsC()
        check(lib().rbrt_ref_render_hdr(self._h, cam_c, spp, opts, out.ctypes.data, st))
        if stats is not None:
            stats.update(st.as_dict())
        return out

    def render(self, cam_c, spp, opts=None, stats=None):
        out = np.empty((cam_c.img_height_pix, cam_c.img_width_pix, 3), np.uint8)
        st = _abi.StatsC()
        check(lib().rbrt_ref_render(self._h, cam_c, spp, opts, out.ctypes.data, st))
        if stats is not None:
            stats.update(st.as_dict())
        return out

    def render_accum(self, cam_c, spp, opts=None, stats=None):
        out = np.empty((cam_c.img_height_pix * cam_c.img_width_pix * 4,), np.float32)
        st = _abi.StatsC()
        check(lib().rbrt_ref_render_accum(self._h, cam_c, spp, opts, out.ctypes.data, st))
        if stats is not None:
            stats.update(st.as_dict())
        return out


def render_subset(osc, cam_c, spp, stride_x, stride_y, opts=None, want_image=False):
    """Render only the pixel lattice (row % stride_y == 0, col % stride_x == 0); returns (stats dict, accum or None)."""
    out = np.zeros((cam_c.img_height_pix * cam_c.img_width_pix * 4,), np.float32) if want_image else None
    st = _abi.StatsC()
    check(lib().rbrt_ref_render_subset(osc._h, cam_c, spp, opts, stride_x, stride_y, out.ctypes.data if want_image else None, st))
    return st.as_dict(), out


def finalize(accum, w, h, spp):
    rgb = np.empty((h, w, 3), np.uint8)
    hdr = np.empty((h, w, 3), np.float32)
    accum = np.ascontiguousarray(accum, dtype=np.float32)
    check(lib().rbrt_ref_finalize(accum.ctypes.data, w, h, spp, rgb.ctypes.data, hdr.ctypes.data))
    return rgb, hdr


def primary_rays(cam_c, seed=0, sample=0):
    rays = np.empty((cam_c.img_width_pix * cam_c.img_height_pix, 6), np.float32)
    check(lib().rbrt_ref_primary_rays(cam_c, seed, sample, rays.ctypes.data))
    return rays
