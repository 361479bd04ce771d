"""One profiled scene build for ncu (--profile-from-start off): python scripts/ncu_build.py <workload>"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rbrt_b200 as R
import bench
R.gpu_init(0)
wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
spheres, meshes, camkw = bench.build_workload(wl)
pinned, keep = bench.pin_meshes(meshes)
for _ in range(3):
    sc = bench.make_scene(spheres, meshes, pinned); sc.info(); sc.close()
torch.cuda.synchronize()
torch.cuda.profiler.start()
sc = bench.make_scene(spheres, meshes, pinned); info = sc.info()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(info, file=sys.stderr)
