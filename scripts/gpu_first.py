"""First GPU contact: parity of trace (brute + BVH) and render vs the oracle, then a timing of C2."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rbrt_b200 as R
from rbrt_b200 import _abi, synth
from oracle import oracle_ffi as O

R.gpu_init(0)
print(_abi.lib().rbrt_gpu_version())

def cmp_hits(a, b, name):
    same_kind = a["kind"] == b["kind"]
    same = same_kind & (a["elem_idx"] == b["elem_idx"]) & (a["tri_idx"] == b["tri_idx"]) & (a["t"].view(np.uint32) == b["t"].view(np.uint32)) \
        & (a["dist"].view(np.uint32) == b["dist"].view(np.uint32)) & (a["point"].view(np.uint32) == b["point"].view(np.uint32)).all(1) \
        & (a["normal"].view(np.uint32) == b["normal"].view(np.uint32)).all(1)
    print(f"{name}: {len(a)} rays, mismatches {int((~same).sum())}, kinds {np.bincount(a['kind'] + 1, minlength=3)}")
    if (~same).any():
        i = np.nonzero(~same)[0][:5]
        print(a[i]); print(b[i])
    return int((~same).sum())

bp = synth.spheres_only_blueprint()
scene = R.create_scene_from_scene_blueprint(bp)
tris = synth.displaced_icosphere(4, 3.0, (5.0, 1.4, -12.5))[:5115]   # N % 8 == 3: tail rule
scene.triangle_meshes.append(R.TriangleMesh.from_triangles(tris, R.Dielectric(0.2)))
print(scene.info())
cb = bp.camera_blueprint
cam = R.Camera.new(cb.camera_position, cb.camera_look_at, cb.camera_up, 192, 256, cb.camera_focal_length_mm)
osc = O.OracleScene.from_scene(scene)
rays = R.primary_rays(cam, seed=3, sample_idx=0)
orays = O.primary_rays(cam.to_c(), 3, 0)
print("primary rays bit-equal:", np.array_equal(rays.view(np.uint32), orays.view(np.uint32)))
ref = osc.hit(rays)
bad = cmp_hits(scene.hit(rays, _abi.TRACE_BRUTE), ref, "brute vs oracle")
bad += cmp_hits(scene.hit(rays, _abi.TRACE_BVH), ref, "bvh   vs oracle")
# random secondary-like rays from points near the mesh
rng = np.random.default_rng(0)
o = rng.normal(size=(200000, 3)).astype(np.float32) * 4 + np.array([5, 1.4, -12.5], np.float32)
d = rng.normal(size=(200000, 3)).astype(np.float32)
d /= np.linalg.norm(d, axis=1, keepdims=True)
rr = np.concatenate([o, d], 1).astype(np.float32)
ref2 = osc.hit(rr)
bad += cmp_hits(scene.hit(rr, _abi.TRACE_BRUTE), ref2, "brute vs oracle (random)")
bad += cmp_hits(scene.hit(rr, _abi.TRACE_BVH), ref2, "bvh   vs oracle (random)")

st = {}
t0 = time.time(); gpu = R.render_scene_hdr(cam, 8, scene, stats=st, seed=1); t1 = time.time()
ost = {}
refimg = osc.render_hdr(cam.to_c(), 8, _abi.RenderOptsC(seed=1), ost); t2 = time.time()
neq = int((gpu.view(np.uint32) != refimg.view(np.uint32)).any(axis=2).sum())
print(f"render 256x192x8: gpu {t1-t0:.3f}s (device {st['ms_device']:.2f} ms, rays {st['rays']}), oracle {t2-t1:.3f}s (rays {ost['rays']}); differing pixels {neq}")
print("mean", gpu.mean(axis=(0, 1)), refimg.mean(axis=(0, 1)))
if neq:
    ys, xs = np.nonzero((gpu.view(np.uint32) != refimg.view(np.uint32)).any(axis=2))
    for y, x in list(zip(ys, xs))[:5]:
        print(y, x, gpu[y, x], refimg[y, x])

# C2 timing
d = synth.cache_dir(); obj = os.path.join(d, "standin6.obj")
if not os.path.exists(obj): synth.write_bunny_standin(obj, 6)
bp2 = synth.example_scene_blueprint(obj)
t0 = time.time(); scene2 = R.create_scene_from_scene_blueprint(bp2); print("load", time.time() - t0)
print(scene2.info())
cam2 = R.Camera.new(cb.camera_position, cb.camera_look_at, cb.camera_up, 768, 1024, cb.camera_focal_length_mm)
rays2 = R.primary_rays(cam2, seed=0, sample_idx=0)
hb = scene2.hit(rays2, _abi.TRACE_BRUTE); st = {}
hv = scene2.hit(rays2, _abi.TRACE_BVH, st)
bad += cmp_hits(hv, hb, "C2 primary: bvh vs brute (GPU)")
print("trace stats", st)
for spp in (2, 50):
    st = {}
    t0 = time.time(); img = R.render_scene(cam2, spp, scene2, stats=st, seed=0); t1 = time.time()
    print(f"C2 {spp} spp: wall {t1-t0:.3f}s device {st['ms_device']:.1f} ms rays {st['rays']} -> {st['rays']/st['ms_device']/1e3:.1f} Mrays/s, {st['paths']/st['ms_device']/1e3:.1f} Msamples/s, launches {st['launches']}")
img.save(os.path.join(os.path.dirname(d), "c2.png"))
print("TOTAL MISMATCHES", bad)
