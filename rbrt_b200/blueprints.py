"""Scene blueprints — host mirror of rbrt_lib::blueprints (blueprints.rs:15-158): the YAML scene
description, the substring material matching, the silent skip of unknown materials."""
from dataclasses import dataclass
from typing import List, Optional

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
    cb = d["camera_blueprint"]
    cam = CameraBluePrint(_v(cb["camera_up"]), _v(cb["camera_look_at"]), _v(cb["camera_position"]),
                          float(cb["camera_focal_length_mm"]))
    meshes = [TriangleMeshBlueprint(str(m["obj_filepath"]), float(m["scale"]), _v(m["translation"]), _v(m["rotation_rad"]),
                                    str(m["material_type"]), _v(m.get("albedo")),
                                    None if m.get("material_param") is None else float(m["material_param"]))
              for m in (d.get("mesh_blueprints") or [])]
    spheres = [SphereBlueprint(float(s["radius"]), _v(s["center"]), str(s["material_type"]), _v(s.get("albedo")),
                               None if s.get("material_param") is None else float(s["material_param"]))
               for s in (d.get("sphere_blueprints") or [])]
    return SceneBlueprint(cam, meshes, spheres)


def load_blueprints_from_yaml_file(filepath):  # blueprints.rs:76-92
    try:
        f = open(filepath, "r")
    except OSError as e:
        raise RuntimeError(f"Failed to open {e!r} to load content.")
    with f:
        try:
            return blueprint_from_dict(yaml.safe_load(f))
        except (yaml.YAMLError, KeyError, TypeError, ValueError) as e:
            raise RuntimeError(f"Unable to parse content of file {filepath!r} to scene blueprint: {e!r}")


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
