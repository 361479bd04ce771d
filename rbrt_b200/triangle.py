"""BasicTriangle — host mirror of rbrt_lib::triangle::BasicTriangle (triangle.rs:9-28): a single counter-clockwise
triangle that can be pushed into `Scene.elements` next to the spheres (it is an `Intersectable`, triangle.rs:412-441).
It only carries parameters; normal, edges and the intersection run in the library."""
from dataclasses import dataclass

from . import _abi
from .vec3 import Vec3


@dataclass
class BasicTriangle:
    corners: tuple
    material: object

    @staticmethod
    def new(corners, material):  # triangle.rs:19-27
        return BasicTriangle(tuple(Vec3.from_any(c) for c in corners), material)

    def to_c(self):
        cs = (_abi.Vec3C * 3)(*[Vec3.from_any(c).to_c() for c in self.corners])
        return _abi.TriangleDescC(cs, self.material.to_c())
