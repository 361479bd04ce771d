"""TriangleMesh — host mirror of rbrt_lib::mesh::TriangleMesh (mesh.rs:12-121).

The reference's constructor loads an .obj with tobj (default LoadOptions: no triangulation, faces read
as consecutive index triples, mesh.rs:96-107), transforms every vertex scale -> rotate_point(Z-X-Z) ->
translate in f32 (mesh.rs:102-112) and converts to padded SoA.  Here the OBJ is parsed on the host, the
transform runs in the C library (rbrt_transform_vertices, bit-identical f32) and the SoA / padding /
normals / AABB / BVH are produced on the GPU at scene upload (csrc/bvh_build.cu).
"""
import numpy as np

from . import _abi
from .vec3 import Vec3


def parse_obj_triangles(filepath):
    """Return (positions [V,3] f32, index triples [F,3] int64) the way rbrt consumes tobj's output
    (mesh.rs:92-107): one model per `o` / `g` record, position indices only (negative = relative), the
    `f` records of a model concatenated and cut into triples (`mesh.indices.len() / 3`, no triangulation:
    tobj's default LoadOptions), models in file order."""
    pos, models, cur = [], [], []
    with open(filepath, "r", errors="replace") as f:
        for line in f:
            parts = line.split()
            if not parts:
                continue
            tag = parts[0]
            if tag == "v":
                pos.append((float(parts[1]), float(parts[2]), float(parts[3])))
            elif tag == "f":
                for tok in parts[1:]:
                    i = int(tok.split("/")[0])
                    cur.append(i - 1 if i > 0 else len(pos) + i)
            elif tag in ("o", "g") and cur:
                models.append(cur)
                cur = []
    if cur:
        models.append(cur)
    idx = []
    for m in models:
        idx.extend(m[:(len(m) // 3) * 3])
    positions = np.asarray(pos, dtype=np.float32).reshape(-1, 3)
    indices = np.asarray(idx, dtype=np.int64).reshape(-1, 3)
    if indices.size and (indices.min() < 0 or indices.max() >= len(positions)):
        raise ValueError(f"{filepath}: face index out of range")
    return positions, indices


def _transform_in_place(ptr, n_vertices, scale, rot_c, tr_c):
    """The library's rbrt_transform_vertices.  (bench.py's reference arm swaps in the oracle's twin so that arm never loads the product.)"""
    _abi.check(_abi.lib().rbrt_transform_vertices(ptr, n_vertices, scale, rot_c, tr_c))


def transform_triangles(tris, translation, rotation, scale):
    """scale -> rotate_point -> translate (mesh.rs:102-112) on an [N,3,3] f32 array, via the C-ABI."""
    tris = np.ascontiguousarray(tris, dtype=np.float32).copy()
    t, r = Vec3.from_any(translation), Vec3.from_any(rotation)
    ptr = tris.ctypes.data_as(_abi.P(_abi.C.c_float))
    _transform_in_place(ptr, tris.size // 3, float(scale), r.to_c(), t.to_c())
    return tris


def load_mesh_vertices_from_file(filepath, translation, rotation, scale):
    """= mesh.rs:78-121. Returns [N,3,3] f32 world-space triangle vertices."""
    positions, indices = parse_obj_triangles(filepath)
    tris = positions[indices] if len(indices) else np.zeros((0, 3, 3), np.float32)
    out = transform_triangles(tris, translation, rotation, scale)
    print(f"Successfully loaded {len(out)} triangles from file {filepath}!")  # mesh.rs:115-119
    return out


class TriangleMesh:
    def __init__(self, triangles, material):
        self.triangles = np.ascontiguousarray(triangles, dtype=np.float32).reshape(-1, 3, 3)
        self.material = material

    @staticmethod
    def new(filepath, translation, rotation, scale, material):  # mesh.rs:41-47
        return TriangleMesh(load_mesh_vertices_from_file(filepath, translation, rotation, scale), material)

    @staticmethod
    def from_triangles(triangles, material):
        """Library-user entry: world-space triangle soup [N,3,3] (what TriangleMesh::new holds after loading)."""
        return TriangleMesh(triangles, material)

    def to_c(self):
        ptr = self.triangles.ctypes.data_as(_abi.P(_abi.C.c_float))
        return _abi.MeshDescC(ptr, len(self.triangles), self.material.to_c())
