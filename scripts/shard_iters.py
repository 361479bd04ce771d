import sys, os
os.environ["RBRT_DEBUG_ITERS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rbrt_b200 as R
from rbrt_b200 import _abi
import bench
R.gpu_init(0)
n = int(sys.argv[1])
desc, W, H, spp = bench.WORKLOADS["c3"]
spheres, meshes, camkw = bench.build_workload("c3")
cam = R.Camera.new(camkw["position"], camkw["look_at"], camkw["up"], H, W, camkw["focal_len_mm"])
scene = bench.make_scene(spheres, meshes)
for rep in range(3):
    st = {}
    kw = dict(shard_mode=_abi.SHARD_TILES, shard_rank=0, shard_count=n) if n > 1 else {}
    R.render_scene_hdr(cam, spp, scene, stats=st, seed=1, time_kernels=True, **kw)
    print(f"TOTAL device {st['ms_device']:.2f} trace {st['ms_trace']:.2f}", file=sys.stderr)
