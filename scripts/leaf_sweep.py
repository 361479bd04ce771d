"""C3 render time vs BVH leaf size."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rbrt_b200 as R
import bench
R.gpu_init(0)
wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
desc, W, H, spp = bench.WORKLOADS[wl]
spheres, meshes, camkw = bench.build_workload(wl)
cam = R.Camera.new(camkw["position"], camkw["look_at"], camkw["up"], H, W, camkw["focal_len_mm"])
for leaf in (1, 2, 4, 6, 8):
    sc = R.Scene(leaf_size=leaf)
    sc.elements += [R.Sphere(c, r, m) for c, r, m in spheres]
    for tris, mat in meshes:
        sc.triangle_meshes.append(R.TriangleMesh.from_triangles(tris, mat))
    for rep in range(2):
        st = {}
        R.render_scene(cam, spp, sc, stats=st, seed=1, time_kernels=True)
    sc2 = {}
    R.render_scene(cam, spp, sc, stats=sc2, seed=1, count_visits=True)
    print(f"leaf {leaf}: device {st['ms_device']:.1f} ms trace {st['ms_trace']:.1f} ms  nodes/ray {sc2['node_visits']/sc2['traversed_rays']:.1f} tris/ray {sc2['tri_tests']/sc2['traversed_rays']:.1f} bvh_nodes {sc.info()['num_bvh_nodes']}", file=sys.stderr, flush=True)
    sc.close()
