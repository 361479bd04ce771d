"""Where the e2e step goes: scene_create (upload + build), render (host total vs device), destroy."""
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
pinned, keep = bench.pin_meshes(meshes)
rgb = np.empty((H, W, 3), np.uint8)
for rep in range(4):
    t0 = time.perf_counter()
    sc = bench.make_scene(spheres, meshes, pinned)
    h = sc.handle()
    t1 = time.perf_counter()
    st = _abi.StatsC()
    _abi.check(_abi.lib().rbrt_gpu_render(h, cam.to_c(), spp, R.render.make_opts(seed=1), rgb.ctypes.data, st))
    t2 = time.perf_counter()
    info = sc.info()
    sc.close()
    t3 = time.perf_counter()
    print(f"rep {rep}: create {1e3*(t1-t0):.1f} ms (upload {info['ms_upload']:.1f} build {info['ms_build']:.1f})  render call {1e3*(t2-t1):.1f} ms "
          f"(ms_total {st.ms_total:.1f} device {st.ms_device:.1f} d2h {st.ms_d2h:.1f})  destroy {1e3*(t3-t2):.1f} ms", file=sys.stderr)
