"""Where can a BVH answer differ from the reference's O(N) sweep?  CPU-only exploration with the HOST BUILD of the product's traversal
(tests/host_device: intersect.cuh compiled by g++, trees built on the host by the GPU builder's padding / quantisation rules) against the oracle.
Run from the repo root after `pytest tests/test_device_source_on_host.py` has built tests/host_device/build/libhost_device.so.

Findings of round 2 (300 000 rays per line unless noted):
  * small mesh near the world origin, rays from 100 .. 990 units away aimed at triangle edges: up to 2 % mismatches with a padding that is only relative
    to the mesh's coordinates -> the floor 2^-21 * (mx + 1000) in k_mesh_setup (none left, also with 2 ulp of error on the slab test's reciprocal);
  * bounce rays from the surface, inward rays, direction components down to 1e-38, flat meshes, meshes at coordinates up to 1e6, 20 triangles of
    radius 500 with origins inside, edge-on rays on closed meshes with edges 0.1 .. 105 long from 10 .. 900 away: no mismatch;
  * 2000 free-floating SLIVERS (5 .. 50 long, 0.0003 .. 0.1 of that wide), edge-aimed rays from 20 / 600 away: 1 mismatch in 1.8 M rays.  There the
    reference's own arithmetic is ill-conditioned (a = e1.(d x e2) just above its 1e-3 cull, |o - v0| ~ 600, |e| ~ 50: u and v carry ~eps |s| |e| / a,
    i.e. whole units), it accepts a "hit" whose point is nowhere near the triangle, and no finite box padding contains that.  Known limit; meshes of
    many small triangles (every BASELINE config) are far from it; `box_pad_rel` raises the padding per scene, RBRT_TRACE_BRUTE is the sweep itself."""
import ctypes as C
import sys

import numpy as np

sys.path.insert(0, ".")
import rbrt_b200 as R                                                       # noqa: E402
from oracle import oracle_ffi as O                                          # noqa: E402
from rbrt_b200 import _abi, synth                                           # noqa: E402
from rbrt_b200.vec3 import Vec3                                             # noqa: E402
from tests import scenes as S                                               # noqa: E402
import tests.test_device_source_on_host as T                                # noqa: E402

lib = C.CDLL("tests/host_device/build/libhost_device.so")
P = C.POINTER
sa = [P(_abi.ElementRefC), C.c_uint32, P(_abi.SphereDescC), P(_abi.TriangleDescC), P(_abi.MeshDescC), C.c_uint32, C.c_uint32]
lib.hd_trace_rays.argtypes = sa + [C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p, P(C.c_uint64), P(C.c_uint64)]
lib.hd_set_rcp_error.argtypes = [C.c_float]


def check(name, tris, rays, leafs=(1,17,4)):
    scene=R.Scene(); scene.triangle_meshes.append(R.TriangleMesh.from_triangles(np.asarray(tris,np.float32),R.Lambertian(Vec3(.5,.5,.5))))
    want=O.OracleScene.from_scene(scene).hit(rays)
    res=[]
    for err in (0.0, 2.4e-7):
        lib.hd_set_rcp_error(err)
        for leaf in leafs:
            got=T.hd_hit(lib,scene,rays,leaf); res.append(int((~S.hits_equal(got,want)).sum()))
    lib.hd_set_rcp_error(0.0)
    print(f"{name}: rays {len(rays)}, oracle mesh hits {(want['kind']==1).sum()}, mismatches {res}")
rng = np.random.default_rng(42)
# 1. bounce rays: origins ON the surface (previous hit points), hemisphere + tangent directions
tris=synth.displaced_icosphere(4,3.0,(5.0,1.4,-12.5))
n=200000
ti=rng.integers(0,len(tris),n); b=rng.random((n,3)); b/=b.sum(1,keepdims=True)
p=(tris[ti].astype(np.float64)*b[:,:,None]).sum(1)
nrm=np.cross(tris[ti,1]-tris[ti,0],tris[ti,2]-tris[ti,0]).astype(np.float64); nrm/=np.linalg.norm(nrm,axis=1,keepdims=True)
d=rng.normal(size=(n,3)); d/=np.linalg.norm(d,axis=1,keepdims=True)
tang=d-(d*nrm).sum(1,keepdims=True)*nrm*(1-10.0**rng.uniform(-6,-1,(n,1)))     # nearly tangent
d=np.where(rng.random((n,1))<0.5,d,tang); d/=np.linalg.norm(d,axis=1,keepdims=True)
check("bounce rays from the surface", tris, np.concatenate([p,d],1).astype(np.float32))
# inside (refracted) rays: origin on surface, direction inward
check("inward rays from the surface", tris, np.concatenate([p,-np.abs((d*nrm).sum(1,keepdims=True))*nrm+0.3*d],1).astype(np.float32))
# 2. tiny direction components
d2=d.copy(); k=rng.integers(0,3,n); d2[np.arange(n),k]*=10.0**rng.uniform(-38,-5,n)
o2=np.array((5.0,1.4,-12.5))+rng.normal(0,4,(n,3))
check("tiny direction components", tris, np.concatenate([o2,d2],1).astype(np.float32))
# 3. flat meshes (zero extent in one axis), axis-aligned quads
g=np.linspace(-2,2,21); X,Y=np.meshgrid(g,g)
def grid_tris(z):
    P=np.stack([X,Y,np.full_like(X,z)],-1); t=[]
    for i in range(20):
        for j in range(20):
            t.append([P[i,j],P[i+1,j],P[i,j+1]]); t.append([P[i+1,j],P[i+1,j+1],P[i,j+1]])
    return np.array(t)
flat=grid_tris(-5.0)
rays=S.random_rays(n//2,(0,0,-5),3.0,3)
inplane=np.concatenate([np.stack([rng.uniform(-3,3,n//4),rng.uniform(-3,3,n//4),np.full(n//4,-5.0)],1), np.stack([rng.normal(size=n//4),rng.normal(size=n//4),np.zeros(n//4)],1)],1)
check("flat mesh z=-5", flat, np.concatenate([rays,inplane,S.edge_aimed_rays(flat,n//4,50.0,1)],0).astype(np.float32))
# 4. huge coordinates
for off in (1e3,1e4,1e5,1e6):
    c=(off,0.5*off,-off)
    tr=synth.displaced_icosphere(3,3.0*max(1,off/1e3),c)
    check(f"mesh at {c}", tr, np.concatenate([S.edge_aimed_rays(tr,n//4,10.0*max(1,off/1e3),2),S.edge_aimed_rays(tr,n//4,900.0,3),S.random_rays(n//4,c,4.0*max(1,off/1e3),4)],0))
# 5. big triangles (few, large) and long thin slivers
big=synth.displaced_icosphere(0,500.0,(0,0,0))
check("20 huge triangles r=500, origins inside", big, np.concatenate([S.random_rays(n//2,(0,0,0),100.0,5), S.edge_aimed_rays(big,n//4,300.0,6)],0))
sl=[]
for i in range(2000):
    a=rng.uniform(-5,5,3); dirn=rng.normal(size=3); dirn/=np.linalg.norm(dirn); w=rng.normal(size=3); w-=w.dot(dirn)*dirn; w/=np.linalg.norm(w)
    L=rng.uniform(5,50); wd=10.0**rng.uniform(-3.5,-1)*L      # 2*area = L*wd >= ~1e-3..
    sl.append([a,a+dirn*L,a+dirn*L*0.5+w*wd])
sl=np.array(sl)
check("2000 slivers", sl, np.concatenate([S.edge_aimed_rays(sl,n//2,20.0,7),S.edge_aimed_rays(sl,n//4,600.0,8),S.random_rays(n//4,(0,0,0),30.0,9)],0))
