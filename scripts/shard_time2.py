import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rbrt_b200 as R
from rbrt_b200 import _abi
import bench
R.gpu_init(0)
desc, W, H, spp = bench.WORKLOADS["c3"]
spheres, meshes, camkw = bench.build_workload("c3")
cam = R.Camera.new(camkw["position"], camkw["look_at"], camkw["up"], H, W, camkw["focal_len_mm"])
scene = bench.make_scene(spheres, meshes)
for n in (3, 4, 5, 6, 8):
    for tail in (False, True):
        ts = []
        for rep in range(4):
            st = {}
            R.render_scene_hdr(cam, spp, scene, stats=st, seed=1, time_kernels=True, no_tail_kernel=tail, shard_mode=_abi.SHARD_TILES, shard_rank=0, shard_count=n)
            ts.append(round(st["ms_device"], 2))
        print(f"ranks {n} no_tail={tail}: device ms {ts} trace {st['ms_trace']:.2f}", file=sys.stderr, flush=True)
