"""Sphere — host mirror of rbrt_lib::sphere::Sphere (sphere.rs:6-10)."""
from dataclasses import dataclass

from . import _abi
from .vec3 import Vec3


@dataclass
class Sphere:
    center: Vec3
    radius: float
    material: object

    def to_c(self):
        return _abi.SphereDescC(Vec3.from_any(self.center).to_c(), float(self.radius), self.material.to_c())
