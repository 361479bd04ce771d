"""C3 render time vs wavefront batch size (paths in flight)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rbrt_b200 as R
from rbrt_b200 import _abi
import bench
R.gpu_init(0)
wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
desc, W, H, spp = bench.WORKLOADS[wl]
spheres, meshes, camkw = bench.build_workload(wl)
cam = R.Camera.new(camkw["position"], camkw["look_at"], camkw["up"], H, W, camkw["focal_len_mm"])
scene = bench.make_scene(spheres, meshes)
print(scene.info())
for bp in [1 << 21, 1 << 23, 1 << 25, 1 << 27]:
    for rep in range(2):
        st = {}
        R.render_scene(cam, spp, scene, stats=st, seed=1, batch_paths=bp, time_kernels=True)
    print(f"batch_paths {bp:>10d}: device {st['ms_device']:8.1f} ms trace {st['ms_trace']:8.1f} ms launches {st['launches']} rays {st['rays']} -> {st['rays']/st['ms_device']/1e3:.1f} Mrays/s", flush=True)
