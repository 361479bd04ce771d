"""Load a fixture of tests/golden/ (see make_golden.py) and rebuild its scene and camera."""
import ctypes as C
import os

import numpy as np

import rbrt_b200 as R
from rbrt_b200 import _abi
from rbrt_b200.vec3 import Vec3

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = ["quirk_48x36", "mesh1275_64x48", "spheres_64x48"]


def _material(kind, r, g, b, param):
    kind = int(kind)
    return [R.Lambertian(Vec3(r, g, b)), R.Metal(Vec3(r, g, b), float(param)), R.Dielectric(float(param))][kind]


def load(name, **scene_opts):
    z = np.load(os.path.join(HERE, name + ".npz"))
    scene = R.Scene(**scene_opts)
    for row in z["spheres"]:
        scene.elements.append(R.Sphere(Vec3(*row[:3]), float(row[3]), _material(*row[4:9])))
    for i in range(int(z["n_meshes"])):
        scene.triangle_meshes.append(R.TriangleMesh.from_triangles(z[f"mesh{i}_tris"], _material(*z[f"mesh{i}_mat"])))
    cam_c = _abi.CameraC.from_buffer_copy(z["camera"].tobytes())
    return z, scene, R.Camera.from_c(cam_c)
