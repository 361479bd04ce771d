"""One profiled frame for ncu: python scripts/ncu_frame.py <workload> <N>  renders tile shard 3 of N (N = 1: the whole frame) of the
workload — what ONE rank of N renders — twice untimed, then once between cudaProfilerStart/Stop (ncu --profile-from-start off)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rbrt_b200 as R
from rbrt_b200 import _abi
import bench
R.gpu_init(0)
wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1
desc, W, H, spp = bench.WORKLOADS[wl]
spheres, meshes, camkw = bench.build_workload(wl)
cam = R.Camera.new(camkw["position"], camkw["look_at"], camkw["up"], H, W, camkw["focal_len_mm"])
scene = bench.make_scene(spheres, meshes)
kw = dict(shard_mode=_abi.SHARD_TILES, shard_rank=min(3, n - 1), shard_count=n) if n > 1 else {}
for _ in range(2):
    R.render_scene_hdr(cam, spp, scene, seed=bench.SEED, **kw)
torch.cuda.synchronize()
torch.cuda.profiler.start()
st = {}
R.render_scene_hdr(cam, spp, scene, seed=bench.SEED, stats=st, **kw)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(f"{wl} 1/{n}: {st['ms_device']:.2f} ms, {st['rays']} rays, {st['launches']} launches", file=sys.stderr)
