import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rbrt_b200 as R
import bench
R.gpu_init(0)
desc, W, H, spp = bench.WORKLOADS["c4"]
spp = int(os.environ.get("SPP", "8"))
spheres, meshes, camkw = bench.build_workload("c4")
cam = R.Camera.new(camkw["position"], camkw["look_at"], camkw["up"], H, W, camkw["focal_len_mm"])
def scene(groups):
    sc = R.Scene()   # (was: sphere_groups=groups, an experiment that was dropped; the script now just checks run-to-run determinism)
    sc.elements += [R.Sphere(c, r, m) for c, r, m in spheres]
    for t, m in meshes: sc.triangle_meshes.append(R.TriangleMesh.from_triangles(t, m))
    return sc
a, b = scene(True), scene(False)
sa, sb = {}, {}
ia = R.render_scene_hdr(cam, spp, a, seed=1, stats=sa); ib = R.render_scene_hdr(cam, spp, b, seed=1, stats=sb)
ne = (ia.view(np.uint32) != ib.view(np.uint32)).any(axis=2)
print("groups vs plain: differing pixels", int(ne.sum()), "rays", sa["rays"], sb["rays"], "nan", sa["nan_rays"], sb["nan_rays"], "ms", sa["ms_device"], sb["ms_device"])
if ne.any():
    ys, xs = np.nonzero(ne)
    for y, x in list(zip(ys, xs))[:8]: print(y, x, ia[y, x], ib[y, x])
    # first bounce: which primary hits differ?
    rays = R.primary_rays(cam, 1, 0)
    from rbrt_b200 import _abi
    ha, hb = a.hit(rays, _abi.TRACE_WAVEFRONT), b.hit(rays, _abi.TRACE_WAVEFRONT)
    d = (ha["kind"] != hb["kind"]) | (ha["elem_idx"] != hb["elem_idx"]) | (ha["t"].view(np.uint32) != hb["t"].view(np.uint32))
    print("primary hits differing:", int(d.sum()))
    for i in np.nonzero(d)[0][:5]: print(i, ha[i], hb[i])

# bounce-like rays: origins on / near the spheres, random directions of several lengths
from rbrt_b200 import _abi
rng = np.random.default_rng(0)
sph = np.array([[c.x, c.y, c.z, r] for c, r, m in spheres], np.float32)
tot = 0
for rep in range(6):
    N = 4000000
    which = rng.integers(0, len(sph), N)
    dirs = rng.normal(size=(N, 3)); dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    o = (sph[which, :3] + dirs * sph[which, 3:4] * rng.choice([1.0, 1.0, 0.999, 1.001, 3.0], N)[:, None]).astype(np.float32)
    d = rng.normal(size=(N, 3)).astype(np.float32); d /= np.linalg.norm(d, axis=1, keepdims=True); d *= rng.choice(np.float32([1, 1, 0.3, 2.5]), N)[:, None]
    rays = np.concatenate([o, d], 1).astype(np.float32)
    ha, hb = a.hit(rays, _abi.TRACE_WAVEFRONT), b.hit(rays, _abi.TRACE_WAVEFRONT)
    dif = (ha["kind"] != hb["kind"]) | (ha["elem_idx"] != hb["elem_idx"]) | (ha["t"].view(np.uint32) != hb["t"].view(np.uint32))
    tot += int(dif.sum())
    for i in np.nonzero(dif)[0][:4]:
        print("RAY", rays[i].tolist(), "\n  groups:", ha[i], "\n  plain: ", hb[i])
print("bounce-like rays differing:", tot)
