"""Vec3 / Ray — host mirrors of rbrt_lib::vec3::Vec3 (vec3.rs:6-10) and rbrt_lib::ray::Ray (ray.rs:4-7).

Values are held as f32 (numpy.float32) so that what crosses the C-ABI is exactly what the caller sees.
Only the data-carrying part is mirrored: arithmetic on the hot path runs on the GPU.
"""
from dataclasses import dataclass

import numpy as np

from ._abi import RayC, Vec3C


@dataclass
class Vec3:
    x: float = 0.0
    y: float = 0.0
    z: float = 0.0

    def __post_init__(self):
        self.x, self.y, self.z = (float(np.float32(v)) for v in (self.x, self.y, self.z))

    @staticmethod
    def new(x, y, z):  # vec3.rs:97-99
        return Vec3(x, y, z)

    @staticmethod
    def zero():  # vec3.rs:101-107
        return Vec3(0.0, 0.0, 0.0)

    @staticmethod
    def from_any(v):
        if isinstance(v, Vec3):
            return v
        if isinstance(v, dict):
            return Vec3(v["x"], v["y"], v["z"])
        x, y, z = v
        return Vec3(x, y, z)

    def to_c(self):
        return Vec3C(self.x, self.y, self.z)

    @staticmethod
    def from_c(c):
        return Vec3(c.x, c.y, c.z)

    def as_tuple(self):
        return (self.x, self.y, self.z)


@dataclass
class Ray:
    origin: Vec3
    direction: Vec3

    def to_c(self):
        return RayC(self.origin.to_c(), self.direction.to_c())
