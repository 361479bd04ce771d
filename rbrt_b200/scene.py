"""Scene — host mirror of rbrt_lib::scene::Scene (scene.rs:12-16): `elements` (spheres and BasicTriangles, any `Intersectable` of the reference that has a GPU twin),
`triangle_meshes`, `lights` (dead in the reference, scene.rs:7-10).  The GPU scene (flattened SoA buffers
+ one LBVH per mesh) is created lazily on first use and owned by this object."""
import ctypes as C

import numpy as np

from . import _abi


def element_arrays(elements):
    """Scene.elements (spheres and BasicTriangles, in order) -> the arrays rbrt_gpu_scene_create_elements takes."""
    from .triangle import BasicTriangle
    sph = [e for e in elements if not isinstance(e, BasicTriangle)]
    tri = [e for e in elements if isinstance(e, BasicTriangle)]
    order, si, ti = [], 0, 0
    for e in elements:
        if isinstance(e, BasicTriangle):
            order.append(_abi.ElementRefC(_abi.ELEM_TRIANGLE, ti)); ti += 1
        else:
            order.append(_abi.ElementRefC(_abi.ELEM_SPHERE, si)); si += 1
    return ((_abi.ElementRefC * max(len(order), 1))(*order), (_abi.SphereDescC * max(len(sph), 1))(*[s.to_c() for s in sph]),
            (_abi.TriangleDescC * max(len(tri), 1))(*[t.to_c() for t in tri]), len(order), len(sph), len(tri))


class Scene:
    def __init__(self, elements=None, triangle_meshes=None, lights=None, simd_lanes=8, leaf_size=0, box_pad_rel=0.0, local=False, sah=True, broadcast=False):
        self.elements = list(elements or [])
        self.triangle_meshes = list(triangle_meshes or [])
        self.lights = list(lights or [])
        self.simd_lanes, self.leaf_size, self.box_pad_rel = simd_lanes, leaf_size, box_pad_rel
        self.local, self.sah, self.broadcast = local, sah, broadcast     # RBRT_SCENE_LOCAL / RBRT_SCENE_NO_SAH / RBRT_SCENE_BROADCAST
        self._handle = None

    # ---- GPU handle -----------------------------------------------------------------------
    def handle(self):
        if self._handle is None:
            lib = _abi.lib()
            order, spheres, tris, ne, ns, nt = element_arrays(self.elements)
            nm = len(self.triangle_meshes)
            meshes = (_abi.MeshDescC * max(nm, 1))(*[m.to_c() for m in self.triangle_meshes])
            opts = _abi.SceneOptsC(self.simd_lanes, self.leaf_size, self.box_pad_rel,
                                   (_abi.SCENE_LOCAL if self.local else 0) | (0 if self.sah else _abi.SCENE_NO_SAH)
                                   | (_abi.SCENE_BROADCAST if self.broadcast else 0))
            h = C.c_void_p()
            _abi.check(lib.rbrt_gpu_scene_create_elements(order, ne, spheres, ns, tris, nt, meshes, nm, opts, C.byref(h)))
            self._handle = h
        return self._handle

    def info(self):
        out = _abi.SceneInfoC()
        _abi.check(_abi.lib().rbrt_gpu_scene_info(self.handle(), out))
        return out.as_dict()

    def close(self):
        if self._handle is not None:
            _abi.lib().rbrt_gpu_scene_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- Scene::hit (scene.rs:19-43) for caller-supplied rays ----------------------------------
    def hit(self, rays, trace_mode=_abi.TRACE_BVH, stats=None):
        """rays: [N,6] f32 (origin xyz, direction xyz). Returns a structured array (HIT_DTYPE)."""
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 6)
        hits = np.zeros(len(rays), dtype=_abi.HIT_DTYPE)
        st = _abi.StatsC()
        _abi.check(_abi.lib().rbrt_gpu_trace_rays(self.handle(), rays.ctypes.data, len(rays), trace_mode,
                                                  hits.ctypes.data, st))
        if stats is not None:
            stats.update(st.as_dict())
        return hits


def scatter(items, seed=0):
    """Parity hook for RayScattering::scatter (materials.rs:4-12).  items: iterable of (material, in_ray_direction(3),
    hit_point(3), hit_normal(3), pixel, sample, bounce); returns (scattered [N] i32, attenuation [N,3], out_dir [N,3])."""
    items = list(items)
    n = len(items)
    arr = (_abi.ScatterInC * max(n, 1))()
    for k, (mat, d, p, nrm, pixel, sample, bounce) in enumerate(items):
        a = arr[k]
        a.material = mat.to_c()
        a.in_ray.origin = _abi.Vec3C(0.0, 0.0, 0.0)
        a.in_ray.direction = _abi.Vec3C(*[float(x) for x in d])
        a.hit_point = _abi.Vec3C(*[float(x) for x in p])
        a.hit_normal = _abi.Vec3C(*[float(x) for x in nrm])
        a.pixel, a.sample, a.bounce = int(pixel), int(sample), int(bounce)
    out = (_abi.ScatterOutC * max(n, 1))()
    _abi.check(_abi.lib().rbrt_gpu_scatter(C.cast(arr, C.c_void_p), n, int(seed) & 0xFFFFFFFFFFFFFFFF, C.cast(out, C.c_void_p)))
    sc = np.array([out[k].scattered for k in range(n)], np.int32)
    att = np.array([[out[k].attenuation.x, out[k].attenuation.y, out[k].attenuation.z] for k in range(n)], np.float32).reshape(n, 3)
    od = np.array([[out[k].out_ray.direction.x, out[k].out_ray.direction.y, out[k].out_ray.direction.z] for k in range(n)], np.float32).reshape(n, 3)
    return sc, att, od
