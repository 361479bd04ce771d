"""Materials — host mirrors of the three RayScattering impls (lambertian.rs:6-9, metal.rs:6-10,
dielectric.rs:6-9).  They only carry parameters; scatter() runs on the GPU (csrc/shade.cuh)."""
from dataclasses import dataclass

from . import _abi
from .vec3 import Vec3


@dataclass
class Lambertian:
    albedo: Vec3

    def to_c(self):
        return _abi.MaterialC(_abi.MAT_LAMBERTIAN, Vec3.from_any(self.albedo).to_c(), 0.0)


@dataclass
class Metal:
    albedo: Vec3
    roughness: float

    def to_c(self):
        return _abi.MaterialC(_abi.MAT_METAL, Vec3.from_any(self.albedo).to_c(), float(self.roughness))


@dataclass
class Dielectric:
    ref_idx: float

    def to_c(self):
        return _abi.MaterialC(_abi.MAT_DIELECTRIC, _abi.Vec3C(0.0, 0.0, 0.0), float(self.ref_idx))
