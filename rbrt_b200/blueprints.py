"""Scene blueprints — host mirror of rbrt_lib::blueprints (blueprints.rs:15-158): the YAML scene
description, the substring material matching, the silent skip of unknown materials."""
import re
from dataclasses import dataclass
from typing import List, Optional

import numpy as np
import yaml

from .materials import Dielectric, Lambertian, Metal
from .mesh import TriangleMesh
from .scene import Scene
from .sphere import Sphere
from .vec3 import Vec3


@dataclass
class TriangleMeshBlueprint:  # blueprints.rs:15-24
    obj_filepath: str
    scale: float
    translation: Vec3
    rotation_rad: Vec3
    material_type: str
    albedo: Optional[Vec3] = None
    material_param: Optional[float] = None


@dataclass
class SphereBlueprint:  # blueprints.rs:26-33
    radius: float
    center: Vec3
    material_type: str
    albedo: Optional[Vec3] = None
    material_param: Optional[float] = None


@dataclass
class CameraBluePrint:  # blueprints.rs:35-41
    camera_up: Vec3
    camera_look_at: Vec3
    camera_position: Vec3
    camera_focal_length_mm: float


@dataclass
class SceneBlueprint:  # blueprints.rs:43-48
    camera_blueprint: CameraBluePrint
    mesh_blueprints: List[TriangleMeshBlueprint]
    sphere_blueprints: List[SphereBlueprint]


def _v(d):
    return None if d is None else Vec3.from_any(d)


def create_material_from_description(mat_type, albedo, material_param):
    """blueprints.rs:50-74: first match of "metal", "lambert", "dielectric" in the lower-cased type;
    a missing required field is the reference's `expect` panic -> ValueError here."""
    t = mat_type.lower()
    if "metal" in t:
        if albedo is None:
            raise ValueError("you forgot to specify an albedo vector for metal")
        if material_param is None:
            raise ValueError("you forgot to specify a roughness (i.e. material_param: 0.1) for metal")
        return Metal(albedo, material_param)
    if "lambert" in t:
        if albedo is None:
            raise ValueError("you forgot to specify an albedo vector for lambertian")
        return Lambertian(albedo)
    if "dielectric" in t:
        if material_param is None:
            raise ValueError("you forgot to specify a refractory index vector (i.e. material_param: 1.8) dielectric")
        return Dielectric(material_param)
    print(f"Cannot figure out material_type from {mat_type}, material_type must be one of metal, lambertian or dielectric!")
    return None


def blueprint_from_dict(d):
    """A SceneBlueprint from plain Python data (dicts / lists / numbers), e.g. built in code; YAML files go through
    load_blueprints_from_yaml_file, which applies serde's rules to the YAML nodes themselves."""
    cb = d["camera_blueprint"]
    cam = CameraBluePrint(_v(cb["camera_up"]), _v(cb["camera_look_at"]), _v(cb["camera_position"]),
                          float(cb["camera_focal_length_mm"]))
    meshes = [TriangleMeshBlueprint(str(m["obj_filepath"]), float(m["scale"]), _v(m["translation"]), _v(m["rotation_rad"]),
                                    str(m["material_type"]), _v(m.get("albedo")),
                                    None if m.get("material_param") is None else float(m["material_param"]))
              for m in d["mesh_blueprints"]]
    spheres = [SphereBlueprint(float(s["radius"]), _v(s["center"]), str(s["material_type"]), _v(s.get("albedo")),
                               None if s.get("material_param") is None else float(s["material_param"]))
               for s in d["sphere_blueprints"]]
    return SceneBlueprint(cam, meshes, spheres)


# ---- serde_yaml 0.9 + serde derive, restated on PyYAML's node graph (blueprints.rs:15-48,76-92) --------------------------
# PyYAML and serde_yaml both parse with libyaml's grammar (serde_yaml: the unsafe-libyaml port), so `yaml.compose` gives the
# structure serde_yaml sees: mappings, sequences, scalars with their style (plain or quoted), aliases resolved.  What a
# scalar MEANS is serde_yaml's own business (not YAML 1.1's, which safe_load applies): the rules are in
# csrc/host/scene_yaml.hpp, restated here independently; tests/test_scene_yaml.py compares the two hosts on generated files.
_FLOAT_RE = re.compile(r"[+-]?(?:\d+\.?\d*|\.\d+)(?:[eE][+-]?\d+)?\Z")      # Rust's f64::from_str, finite literals
_LEADING_ZEROS_RE = re.compile(r"[+-]?0\d+\Z")                               # "007": a string in YAML 1.2
_NULLS = ("", "~", "null", "Null", "NULL")


class BlueprintError(ValueError):
    pass


def _is_null(node):
    return isinstance(node, yaml.ScalarNode) and node.style is None and node.value in _NULLS


def _kind(node):
    if _is_null(node):
        return "unit value"
    return {yaml.MappingNode: "map", yaml.SequenceNode: "sequence"}.get(type(node), "scalar")


def _f32(node, field):
    if node is None:
        raise BlueprintError(f"missing field `{field}`")
    if not isinstance(node, yaml.ScalarNode) or _is_null(node):
        raise BlueprintError(f"`{field}`: invalid type: {_kind(node)}, expected f32")
    s = node.value
    bad = BlueprintError(f"`{field}`: invalid type: string {s!r}, expected f32")
    if node.style is not None:
        raise bad                                                            # quoted: a string
    if s in ("true", "True", "TRUE", "false", "False", "FALSE"):
        raise BlueprintError(f"`{field}`: invalid type: boolean, expected f32")
    u, neg = s, False
    if u[:1] in ("+", "-"):
        neg, u = u[0] == "-", u[1:]
    radix = {"0x": 16, "0o": 8, "0b": 2}.get(u[:2], 10)
    digits = u if radix == 10 else u[2:]
    alphabet = "0123456789abcdefghijklmnopqrstuvwxyz"[:radix]
    if digits and all(c in alphabet for c in digits.lower()) and not (radix == 10 and _LEADING_ZEROS_RE.match(s)):
        v = int(digits, radix)
        if v < 2 ** 128:
            f = _int_to_f32(v)                                               # `v as f32`: rounded once
            return -f if neg and v else f
        if radix != 10:
            raise bad
    if _LEADING_ZEROS_RE.match(s):
        raise bad
    u = s
    if u[:1] == "+":
        u = u[1:]
        if u[:1] in ("+", "-"):
            raise bad
    if u in (".inf", ".Inf", ".INF"):
        return float("inf")
    if s in ("-.inf", "-.Inf", "-.INF"):
        return float("-inf")
    if s in (".nan", ".NaN", ".NAN"):
        return float("nan")
    if not _FLOAT_RE.match(u):
        raise bad
    d = float(u)                                                             # correctly rounded to f64 ...
    if d in (float("inf"), float("-inf")):
        raise bad                                                            # serde_yaml keeps an overflowing literal a string
    return d                                                                 # ... then `as f32`, where the value crosses the C-ABI (c_float / Vec3)


def _int_to_f32(v):
    """Nearest f32 of a non-negative integer, ties to even (through f64 it would be rounded twice above 2^53)."""
    if v < 2 ** 53:
        with np.errstate(over="ignore"):
            return float(np.float32(v))
    shift = v.bit_length() - 24
    q, r = v >> shift, v & ((1 << shift) - 1)
    half = 1 << (shift - 1)
    if r > half or (r == half and (q & 1)):
        q += 1
    with np.errstate(over="ignore"):
        return float(np.float32(float(q) * 2.0 ** shift))                    # q has <= 25 bits: exact in f64; 2^128 and above -> inf


class _Fields:
    """A struct: a mapping with the named fields (unknown keys ignored, a wanted key given twice is an error) or the sequence of ALL its
    fields in declaration order (serde's derived visit_seq)."""

    def __init__(self, node, names, what):
        if node is None:
            raise BlueprintError(f"missing field `{what}`")
        if not isinstance(node, (yaml.MappingNode, yaml.SequenceNode)):
            raise BlueprintError(f"`{what}`: invalid type: {_kind(node)}, expected struct")
        if isinstance(node, yaml.SequenceNode) and len(node.value) != len(names):
            raise BlueprintError(f"`{what}`: invalid length {len(node.value)}, expected struct with {len(names)} elements")
        self.node, self.names, self.what = node, names, what

    def get(self, key, required=True):
        found = None
        if isinstance(self.node, yaml.MappingNode):
            for k, v in self.node.value:
                if isinstance(k, yaml.ScalarNode) and not _is_null(k) and k.value == key:
                    if found is not None:
                        raise BlueprintError(f"duplicate field `{key}`")
                    found = v
        else:
            found = self.node.value[self.names.index(key)]
        if found is None and required:
            raise BlueprintError(f"`{self.what}`: missing field `{key}`")
        return found


def _vec3(node, what):
    f = _Fields(node, ["x", "y", "z"], what)
    return Vec3(_f32(f.get("x"), "x"), _f32(f.get("y"), "y"), _f32(f.get("z"), "z"))


def _string(node, what):
    if node is None:
        raise BlueprintError(f"missing field `{what}`")
    if not isinstance(node, yaml.ScalarNode):
        raise BlueprintError(f"`{what}`: invalid type: {_kind(node)}, expected a string")
    return node.value


def _seq(node, what):
    if node is None:
        raise BlueprintError(f"missing field `{what}`")
    if not isinstance(node, yaml.SequenceNode):
        raise BlueprintError(f"`{what}`: invalid type: {_kind(node)}, expected a sequence")
    return node.value


def _material(f):
    a, p = f.get("albedo", False), f.get("material_param", False)
    return (_string(f.get("material_type"), "material_type"), None if a is None or _is_null(a) else _vec3(a, "albedo"),
            None if p is None or _is_null(p) else _f32(p, "material_param"))


def blueprint_from_yaml_node(root):
    top = _Fields(root, ["camera_blueprint", "mesh_blueprints", "sphere_blueprints"], "SceneBlueprint")
    c = _Fields(top.get("camera_blueprint"), ["camera_up", "camera_look_at", "camera_position", "camera_focal_length_mm"], "camera_blueprint")
    cam = CameraBluePrint(_vec3(c.get("camera_up"), "camera_up"), _vec3(c.get("camera_look_at"), "camera_look_at"),
                          _vec3(c.get("camera_position"), "camera_position"), _f32(c.get("camera_focal_length_mm"), "camera_focal_length_mm"))
    meshes, spheres = [], []
    for it in _seq(top.get("mesh_blueprints"), "mesh_blueprints"):
        f = _Fields(it, ["obj_filepath", "scale", "translation", "rotation_rad", "material_type", "albedo", "material_param"], "TriangleMeshBlueprint")
        meshes.append(TriangleMeshBlueprint(_string(f.get("obj_filepath"), "obj_filepath"), _f32(f.get("scale"), "scale"),
                                            _vec3(f.get("translation"), "translation"), _vec3(f.get("rotation_rad"), "rotation_rad"), *_material(f)))
    for it in _seq(top.get("sphere_blueprints"), "sphere_blueprints"):
        f = _Fields(it, ["radius", "center", "material_type", "albedo", "material_param"], "SphereBlueprint")
        spheres.append(SphereBlueprint(_f32(f.get("radius"), "radius"), _vec3(f.get("center"), "center"), *_material(f)))
    return SceneBlueprint(cam, meshes, spheres)


def dump_blueprint(bp):
    """The text `rbrt --check --dump` prints: every field as read, f32 as bit patterns (tests compare the two hosts with it)."""
    def bits(x):
        with np.errstate(over="ignore"):
            return f"{int(np.float32(x).view(np.uint32)):08x}"

    def v3(name, v):
        return f"{name} {bits(v.x)} {bits(v.y)} {bits(v.z)}"

    def mat(b):
        return [f"material_type {len(b.material_type.encode())}:{b.material_type}", "albedo None" if b.albedo is None else v3("albedo", b.albedo),
                "material_param None" if b.material_param is None else f"material_param {bits(b.material_param)}"]
    c = bp.camera_blueprint
    out = [v3("camera_up", c.camera_up), v3("camera_look_at", c.camera_look_at), v3("camera_position", c.camera_position),
           f"camera_focal_length_mm {bits(c.camera_focal_length_mm)}"]
    for m in bp.mesh_blueprints:
        out += ["mesh", f"obj_filepath {len(m.obj_filepath.encode())}:{m.obj_filepath}", f"scale {bits(m.scale)}", v3("translation", m.translation),
                v3("rotation_rad", m.rotation_rad)] + mat(m)
    for s in bp.sphere_blueprints:
        out += ["sphere", f"radius {bits(s.radius)}", v3("center", s.center)] + mat(s)
    return "\n".join(out) + "\n"


def load_blueprints_from_yaml_file(filepath):  # blueprints.rs:76-92
    try:
        f = open(filepath, "rb")
    except OSError as e:
        raise RuntimeError(f"Failed to open {e!r} to load content.")
    with f:
        try:
            return blueprint_from_yaml_node(yaml.compose(f, Loader=yaml.SafeLoader))
        except (yaml.YAMLError, BlueprintError) as e:
            raise RuntimeError(f"Unable to parse content of file {filepath!r} to scene blueprint: {e}")


def parse_mesh_bp(mesh_bp):  # blueprints.rs:94-113
    mat = create_material_from_description(mesh_bp.material_type, mesh_bp.albedo, mesh_bp.material_param)
    if mat is None:
        print("Failed to parse material info provided with mesh!")
        return None
    return TriangleMesh.new(mesh_bp.obj_filepath, mesh_bp.translation, mesh_bp.rotation_rad, mesh_bp.scale, mat)


def parse_sphere_bp(sphere_bp):  # blueprints.rs:115-130
    mat = create_material_from_description(sphere_bp.material_type, sphere_bp.albedo, sphere_bp.material_param)
    return None if mat is None else Sphere(sphere_bp.center, sphere_bp.radius, mat)


def create_scene_from_scene_blueprint(scene_bp, **scene_opts):  # blueprints.rs:132-158
    meshes = [m for m in (parse_mesh_bp(bp) for bp in scene_bp.mesh_blueprints) if m is not None]
    spheres = [s for s in (parse_sphere_bp(bp) for bp in scene_bp.sphere_blueprints) if s is not None]
    return Scene(elements=spheres, triangle_meshes=meshes, lights=[], **scene_opts)
