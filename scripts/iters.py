"""Per-iteration queue sizes and trace times of one workload (RBRT_DEBUG_ITERS)."""
import sys, os
os.environ["RBRT_DEBUG_ITERS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rbrt_b200 as R
import bench
R.gpu_init(0)
wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
desc, W, H, spp = bench.WORKLOADS[wl]
spheres, meshes, camkw = bench.build_workload(wl)
cam = R.Camera.new(camkw["position"], camkw["look_at"], camkw["up"], H, W, camkw["focal_len_mm"])
scene = bench.make_scene(spheres, meshes)
for rep in range(2):
    st = {}
    R.render_scene(cam, spp, scene, stats=st, seed=1, time_kernels=True)
print(f"{wl}: device {st['ms_device']:.1f} ms trace {st['ms_trace']:.1f} ms launches {st['launches']} rays {st['rays']} -> {st['rays']/st['ms_device']/1e3:.1f} Mrays/s", file=sys.stderr)
