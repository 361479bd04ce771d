"""One experiment line per process: python scripts/exp.py <workload> [label]  (library knobs come from the environment / RBRT_GPU_LIB).
Prints: LBVH build ms, node visits + triangle tests per traversed ray, ms/frame (one frame at a time), trace ms, generate+shade-it0 ms,
and the same for a 1/8 tile shard (what one rank of 8 renders)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rbrt_b200 as R
from rbrt_b200 import _abi
import bench
R.gpu_init(0)
wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
label = sys.argv[2] if len(sys.argv) > 2 else ""
desc, W, H, spp = bench.WORKLOADS[wl]
spp = int(os.environ.get("SPP", spp))
spheres, meshes, camkw = bench.build_workload(wl)
cam = R.Camera.new(camkw["position"], camkw["look_at"], camkw["up"], H, W, camkw["focal_len_mm"])
scene = bench.make_scene(spheres, meshes)
scene.sah = os.environ.get("NO_SAH") is None
info = scene.info()
for _ in range(2):                                              # warm build timing
    s2 = bench.make_scene(spheres, meshes); s2.sah = scene.sah; i2 = s2.info(); s2.close()
out = {"wl": wl, "label": label, "build_ms": round(i2["ms_build"], 3), "upload_ms": round(i2["ms_upload"], 3), "nodes": info["num_bvh_nodes"]}
st = {}
img0 = R.render_scene_hdr(cam, spp, scene, stats=st, seed=1, count_visits=True)
tr = max(st["traversed_rays"], 1)
out.update(V=round(st["node_visits"] / tr, 3), T=round(st["tri_tests"] / tr, 3), rays=st["rays"], traversed=st["traversed_rays"])
for shard in (1, 8):
    kw = dict(shard_mode=_abi.SHARD_TILES, shard_rank=3, shard_count=8) if shard > 1 else {}
    ms, trc = [], []
    for rep in range(5):
        st = {}
        R.render_scene_hdr(cam, spp, scene, stats=st, seed=1, time_kernels=True, **kw)
        if rep >= 2:
            ms.append(st["ms_device"]); trc.append(st["ms_trace"])
    out[f"ms_frame_{shard}"] = round(float(np.mean(ms)), 3); out[f"ms_trace_{shard}"] = round(float(np.mean(trc)), 3)
    ms = []
    for rep in range(5):                                       # the same frame on two lanes (RBRT_OPT_SPLIT_BATCHES, what rbrt_gpu_render does)
        st = {}
        img = R.render_scene_hdr(cam, spp, scene, stats=st, seed=1, **kw)
        if rep >= 2:
            ms.append(st["ms_device"])
    out[f"ms_split_{shard}"] = round(float(np.mean(ms)), 3)
    if shard == 1:
        out["split_identical"] = bool(np.array_equal(img.view(np.uint32), img0.view(np.uint32)))
out["launches"] = st["launches"]
out["checksum"] = int(np.ascontiguousarray(img0).view(np.uint32).astype(np.uint64).sum())
print(json.dumps(out))
