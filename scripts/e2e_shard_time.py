"""e2e step time (scene upload + LBVH build + render + image to host, frames in flight) for the full frame or one
tile shard of n, on one GPU — to see what an 8-GPU rank's host-bound e2e step costs without 8 GPUs."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rbrt_b200 as R
from rbrt_b200 import _abi
import bench
R.gpu_init(0)
wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
desc, W, H, spp = bench.WORKLOADS[wl]
spheres, meshes, camkw = bench.build_workload(wl)
cam = R.Camera.new(camkw["position"], camkw["look_at"], camkw["up"], H, W, camkw["focal_len_mm"])
pinned, keep = bench.pin_meshes(meshes)
kw = dict(shard_mode=_abi.SHARD_TILES, shard_rank=0, shard_count=n) if n > 1 else {}
depth = int(os.environ.get('DEPTH', '2'))
hostout = os.environ.get('HOSTOUT', '1') == '1'
dep = os.environ.get('DEP', '1') == '1'        # DEP=0: a scene is created per step but the render uses a resident one (cost of the build without the dependency)
pipe = R.FramePipeline(W, H, depth=depth, host_output=hostout)
fixed = bench.make_scene(spheres, meshes, pinned); fixed.handle()

def retire(fin):
    if fin is not None:
        fin[1].close()

def step():
    t0 = time.perf_counter()
    sc = bench.make_scene(spheres, meshes, pinned)
    sc.handle()
    t1 = time.perf_counter()
    for fin in pipe.submit(cam, spp, sc if dep else fixed, tag=sc, seed=1, **kw):
        retire(fin)
    return (t1 - t0) * 1e3

for _ in range(4):
    step()
for f in pipe.drain():
    retire(f)
torch.cuda.synchronize()
t0 = time.perf_counter()
cr = [step() for _ in range(steps)]
for f in pipe.drain():
    retire(f)
dt = (time.perf_counter() - t0) * 1e3 / steps
print(f"{wl} shard 1/{n} depth {depth} hostout {hostout} dep {dep}: e2e {dt:.2f} ms/step, scene_create {sum(cr) / len(cr):.2f} ms (host view)", file=sys.stderr)
