"""TriangleMesh — host mirror of rbrt_lib::mesh::TriangleMesh (mesh.rs:12-121).

The reference's constructor loads an .obj with tobj (default LoadOptions: no triangulation, faces read
as consecutive index triples, mesh.rs:96-107), transforms every vertex scale -> rotate_point(Z-X-Z) ->
translate in f32 (mesh.rs:102-112) and converts to padded SoA.  Here the OBJ is parsed and transformed by the
C library (rbrt_mesh_load_obj / rbrt_transform_vertices, bit-identical f32, all host threads) and the SoA / padding /
normals / AABB / BVH are produced on the GPU at scene upload (csrc/bvh_build.cu).
"""
import os
import re
from fractions import Fraction

import numpy as np

from . import _abi
from .vec3 import Vec3


_F32_RE = re.compile(r"[+-]?(?:inf|infinity|nan|(?:\d+\.?\d*|\.\d+)(?:[eE][+-]?\d+)?)\Z", re.I)   # Rust's f32::from_str
_INT_RE = re.compile(r"[+-]?\d+\Z")                                                                 # Rust's isize::from_str
_WS = " \t\n\x0b\x0c\r"


class ObjLoadError(ValueError):
    """The load the reference panics on (`assert!(loaded_mesh.is_ok())`, mesh.rs:89)."""


def _f32(tok):
    """Decimal -> f32 rounded ONCE, as Rust parses it (float() rounds to f64 first: wrong when that lands on an f32 midpoint)."""
    if not _F32_RE.match(tok):
        raise ObjLoadError(f"not a number: {tok!r}")
    d = float(tok)
    with np.errstate(over="ignore"):
        f = np.float32(d)
    if np.isfinite(f) and d != 0.0 and (np.float64(d).view(np.uint64) & 0x1FFFFFFF) == 0x10000000:
        exact = Fraction(tok)                                    # d sits exactly between two f32: decide with the exact value
        if exact != Fraction(d):
            lo, hi = sorted((np.nextafter(np.float32(d), np.float32(-np.inf)), np.nextafter(np.float32(d), np.float32(np.inf))))
            cands = sorted({float(lo), float(f), float(hi)}, key=lambda c: abs(Fraction(c) - exact))
            f = np.float32(cands[0])
    return f


def _words(line):
    return [w for w in re.split("[" + _WS + "]+", line) if w]


def _mtl_names(path):
    """`newmtl` names of a material library, or None where tobj's load_mtl fails (the library then contributes nothing)."""
    try:
        text = open(path, "r", errors="surrogateescape", newline="\n").read()
    except OSError:
        return None
    names = []
    try:
        for line in text.split("\n"):
            w = _words(line)
            if not w:
                continue
            rest = line.strip(_WS)[len(w[0]):].strip(_WS)
            if w[0] == "newmtl":
                if not rest:
                    return None
                names.append(rest)
            elif w[0] in ("Ka", "Kd", "Ks"):
                [_f32(x) for x in (w[1:4] if len(w) >= 4 else [""])]
            elif w[0] in ("Ns", "Ni", "d"):
                _f32(w[1] if len(w) > 1 else "")
            elif w[0] == "illum":
                if len(w) < 2 or not _INT_RE.match(w[1]) or not 0 <= int(w[1]) <= 255:
                    return None
            elif w[0] in ("map_Ka", "map_Kd", "map_Ks", "map_Ns", "map_Bump", "map_bump", "bump", "map_d") and not rest:
                return None
    except ObjLoadError:
        return None
    return names


def parse_obj_triangles(filepath):
    """Return (positions [V,3] f32, index triples [F,3] int64) the way rbrt consumes tobj 4's default output (mesh.rs:84-107).

    A plain, serial restatement of the rules csrc/obj_loader.cpp spells out (the library's rbrt_mesh_load_obj is what the product
    uses; this one is its independent check in tests/ and what bench.py's reference arm loads meshes with): `f` and `l` records
    append their position indices as they stand (negative = relative; no triangulation), a model ends at `o` / `g` when face
    records are pending and at `usemtl` when the material ID changes too, every model is checked against what had been read
    when it ended, and each model's index list is cut into triples (`mesh.indices.len() / 3`)."""
    pos, n_vt, n_vn = [], 0, 0
    models, cur, pending = [], [], 0                              # cur: (v, vt, vn) per corner
    mat_map, n_materials, mat_id = {}, 0, None
    base = os.path.dirname(filepath)

    def close_model():
        nonlocal cur, pending
        for v, vt, vn in cur:
            if not 0 <= v < len(pos) or (n_vt and vt is not None and not 0 <= vt < n_vt) or (n_vn and vn is not None and not 0 <= vn < n_vn):
                raise ObjLoadError(f"{filepath}: face index out of range")
        models.append([c[0] for c in cur])
        cur, pending = [], 0

    with open(filepath, "r", errors="surrogateescape", newline="\n") as f:
        text = f.read()
    for line_no, line in enumerate(text.split("\n"), 1):
        w = _words(line)
        if not w:
            continue
        tag = w[0]
        try:
            if tag == "v":
                if len(w) < 4:
                    raise ObjLoadError("a `v` record needs three numbers")
                pos.append((_f32(w[1]), _f32(w[2]), _f32(w[3])))
            elif tag == "vt":
                if len(w) < 3:
                    raise ObjLoadError("a `vt` record needs two numbers")
                _f32(w[1]), _f32(w[2]); n_vt += 1
            elif tag == "vn":
                if len(w) < 4:
                    raise ObjLoadError("a `vn` record needs three numbers")
                _f32(w[1]), _f32(w[2]), _f32(w[3]); n_vn += 1
            elif tag in ("f", "l"):
                for tok in w[1:]:
                    fields = tok.split("/")
                    corner = [None, None, None]
                    for k, fld in enumerate(fields):
                        if fld == "":
                            continue
                        if k > 2 or not _INT_RE.match(fld) or abs(int(fld)) >= 2 ** 63:
                            raise ObjLoadError(f"bad face corner {tok!r}")
                        x = int(fld)
                        val = x - 1 if x >= 0 else (len(pos), n_vt, n_vn)[k] + x
                        corner[k] = None if (k > 0 and val == -1) else val       # usize::MAX is tobj's "absent"
                    if corner[0] is None:
                        corner[0] = -1                                            # no position index: out of bounds
                    cur.append(tuple(corner))
                pending += 1
            elif tag in ("o", "g"):
                if pending:
                    close_model()
            elif tag in ("usemtl", "mtllib"):
                name = line.strip(_WS)[len(tag):].strip(_WS)
                if tag == "mtllib":
                    names = _mtl_names(os.path.join(base, name)) if name else None
                    if names is not None:
                        for k, nm in enumerate(names):
                            mat_map[nm] = n_materials + k
                        n_materials += len(names)
                else:
                    if not name:
                        raise ObjLoadError("`usemtl` without a name")
                    new_mat = mat_map.get(name)
                    if new_mat != mat_id and pending:
                        close_model()
                    mat_id = new_mat
        except ObjLoadError as e:
            raise ObjLoadError(f"{filepath} line {line_no}: {e}") from None
    close_model()
    idx = []
    for m in models:
        idx.extend(m[:(len(m) // 3) * 3])
    positions = np.asarray(pos, dtype=np.float32).reshape(-1, 3)
    indices = np.asarray(idx, dtype=np.int64).reshape(-1, 3)
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


def load_obj_soup_python(filepath, translation, rotation, scale):
    """parse_obj_triangles + the transform: the slow twin of _load_obj_soup (tests; bench.py's reference arm, which must not load the product)."""
    positions, indices = parse_obj_triangles(filepath)
    tris = positions[indices] if len(indices) else np.zeros((0, 3, 3), np.float32)
    return transform_triangles(tris, translation, rotation, scale)


def _load_obj_soup(filepath, translation, rotation, scale):
    """The library's rbrt_mesh_load_obj (csrc/obj_loader.cpp): the file parsed on all host threads, corners gathered and transformed."""
    lib = _abi.lib()
    ptr, n = _abi.P(_abi.C.c_float)(), _abi.C.c_uint64(0)
    rc = lib.rbrt_mesh_load_obj(os.fsencode(filepath), Vec3.from_any(translation).to_c(), Vec3.from_any(rotation).to_c(), float(scale),
                                _abi.C.byref(ptr), _abi.C.byref(n))
    if rc != 0:
        raise ObjLoadError(lib.rbrt_last_error().decode("utf-8", "replace"))
    try:
        if not n.value:
            return np.zeros((0, 3, 3), np.float32)
        return np.ctypeslib.as_array(ptr, shape=(n.value, 3, 3)).copy()
    finally:
        lib.rbrt_mesh_free(ptr)


def load_mesh_vertices_from_file(filepath, translation, rotation, scale):
    """= mesh.rs:78-121. Returns [N,3,3] f32 world-space triangle vertices."""
    out = _load_obj_soup(filepath, translation, rotation, scale)
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
