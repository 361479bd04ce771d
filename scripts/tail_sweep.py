import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rbrt_b200 as R
from rbrt_b200 import _abi
import bench
R.gpu_init(0)
for wl, n in (("c4", 1), ("c4", 8), ("c3", 1), ("c3", 8), ("c2", 1)):
    desc, W, H, spp = bench.WORKLOADS[wl]
    spheres, meshes, camkw = bench.build_workload(wl)
    cam = R.Camera.new(camkw["position"], camkw["look_at"], camkw["up"], H, W, camkw["focal_len_mm"])
    scene = bench.make_scene(spheres, meshes)
    out = []
    for t in (113000, 262144, 524288, 1048576, 2097152):
        os.environ["RBRT_TAIL_RAYS"] = str(t)
        best = 1e9
        for rep in range(3):
            st = {}
            kw = dict(shard_mode=_abi.SHARD_TILES, shard_rank=0, shard_count=n) if n > 1 else {}
            R.render_scene_hdr(cam, spp, scene, stats=st, seed=1, **kw)
            best = min(best, st["ms_device"])
        out.append(f"{t}: {best:.2f}")
    print(wl, "ranks", n, " | ".join(out), file=sys.stderr, flush=True)
    scene.close()
