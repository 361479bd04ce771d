"""Frames per second with 1 and 2 frames in flight (FramePipeline), for the full frame or one tile shard of n."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rbrt_b200 as R
from rbrt_b200 import _abi
import bench
R.gpu_init(0)
wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 8
desc, W, H, spp = bench.WORKLOADS[wl]
spp = int(os.environ.get("SPP", spp))
spheres, meshes, camkw = bench.build_workload(wl)
cam = R.Camera.new(camkw["position"], camkw["look_at"], camkw["up"], H, W, camkw["focal_len_mm"])
scene = bench.make_scene(spheres, meshes)
kw = dict(shard_mode=_abi.SHARD_TILES, shard_rank=0, shard_count=n) if n > 1 else {}
ref = None
fpb = int(os.environ.get("FPB", "1"))
for depth in [int(x) for x in os.environ.get("DEPTHS", "1,2,3,4").split(",")]:
    pipe = R.FramePipeline(W, H, depth=depth, host_output=False, frames_per_batch=fpb)
    for _ in range(3 * fpb):
        pipe.submit(cam, spp, scene, seed=1, **kw)
    last = pipe.drain()[-1][0].clone()
    if ref is None: ref = last
    assert torch.equal(ref, last), "pipelined image differs"
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(frames):
        pipe.submit(cam, spp, scene, seed=1, **kw)
    pipe.flush()
    pipe.wait_on()
    e1.record()
    out = pipe.drain()
    torch.cuda.synchronize()
    assert torch.equal(ref, out[-1][0]), "pipelined image differs"
    print(f"{wl} shard 1/{n}: {depth} group(s) of {fpb} frame(s) in flight: {e0.elapsed_time(e1) / frames:.2f} ms/frame", file=sys.stderr)
