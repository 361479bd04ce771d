"""Per-rank render time of C3 when the frame is tile-sharded over N ranks (measured on one GPU, rank 0's shard)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rbrt_b200 as R
from rbrt_b200 import _abi
import bench
R.gpu_init(0)
wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
desc, W, H, spp = bench.WORKLOADS[wl]
spheres, meshes, camkw = bench.build_workload(wl)
cam = R.Camera.new(camkw["position"], camkw["look_at"], camkw["up"], H, W, camkw["focal_len_mm"])
scene = bench.make_scene(spheres, meshes)
base = None
for n in (1, 2, 4, 8):
    for rep in range(3):
        st = {}
        kw = dict(shard_mode=_abi.SHARD_TILES, shard_rank=0, shard_count=n) if n > 1 else {}
        R.render_scene_hdr(cam, spp, scene, stats=st, seed=1, time_kernels=True, **kw)
    base = base or st["ms_device"]
    print(f"ranks {n}: rank-0 device {st['ms_device']:.2f} ms (trace {st['ms_trace']:.2f}) rays {st['rays']}  -> speed-up {base/st['ms_device']:.2f}x", file=sys.stderr, flush=True)
