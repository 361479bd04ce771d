"""Camera — host mirror of rbrt_lib::cam::Camera (cam.rs:4-62).

`Camera.new` keeps the reference's argument order (height BEFORE width, cam.rs:22-29 / main.rs:71-78)
and is computed by the C library's rbrt_camera_new so that every field is bit-identical f32.
"""
from dataclasses import dataclass

from . import _abi
from .vec3 import Vec3


@dataclass
class Camera:
    hor_fov_rad: float
    img_width_pix: int
    img_height_mm: float
    vert_fov_rad: float
    img_height_pix: int
    img_width_mm: float
    position: Vec3
    focal_len_mm: float
    look_at: Vec3
    up: Vec3
    right: Vec3
    img_center_point: Vec3
    mm_per_pix_hor: float
    mm_per_pix_vert: float

    @staticmethod
    def new(position, look_at, up, img_height_pix, img_width_pix, focal_len_mm):
        position, look_at, up = Vec3.from_any(position), Vec3.from_any(look_at), Vec3.from_any(up)
        c = _abi.CameraC()
        _abi.check(_abi.lib().rbrt_camera_new(position.to_c(), look_at.to_c(), up.to_c(), int(img_height_pix),
                                              int(img_width_pix), float(focal_len_mm), c))
        return Camera.from_c(c)

    @staticmethod
    def from_c(c):
        return Camera(c.hor_fov_rad, c.img_width_pix, c.img_height_mm, c.vert_fov_rad, c.img_height_pix,
                      c.img_width_mm, Vec3.from_c(c.position), c.focal_len_mm, Vec3.from_c(c.look_at),
                      Vec3.from_c(c.up), Vec3.from_c(c.right), Vec3.from_c(c.img_center_point),
                      c.mm_per_pix_hor, c.mm_per_pix_vert)

    def to_c(self):
        return _abi.CameraC(self.hor_fov_rad, self.img_width_pix, self.img_height_mm, self.vert_fov_rad,
                            self.img_height_pix, self.img_width_mm, self.position.to_c(), self.focal_len_mm,
                            self.look_at.to_c(), self.up.to_c(), self.right.to_c(), self.img_center_point.to_c(),
                            self.mm_per_pix_hor, self.mm_per_pix_vert)
