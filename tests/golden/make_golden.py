"""Generates the golden fixtures under tests/golden/ from the CPU oracle (oracle/rbrt_oracle.cpp).

The reference (Rust) cannot be built or run in the build image (no cargo/rustc, no network), and its RNG is
unseedable, so no output of the reference itself exists to record; these vectors pin the ORACLE (they make any
later change to it visible) and give the -m gpu tests committed inputs/outputs that do not depend on
re-running the oracle.  Run from the repo root:   python -m tests.golden.make_golden
"""
import os

import numpy as np

from oracle import oracle_ffi as O
from rbrt_b200 import _abi

from .. import scenes as S

HERE = os.path.dirname(os.path.abspath(__file__))


def scene_arrays(scene):
    """Everything needed to rebuild the scene without synth.py: sphere table, per-mesh triangle soup, materials."""
    sph = np.array([[*s.center.as_tuple(), s.radius, s.material.to_c().kind, s.material.to_c().albedo.x, s.material.to_c().albedo.y,
                     s.material.to_c().albedo.z, s.material.to_c().param] for s in scene.elements], np.float32).reshape(-1, 9)
    out = {"spheres": sph, "n_meshes": np.int32(len(scene.triangle_meshes))}
    for i, m in enumerate(scene.triangle_meshes):
        mc = m.material.to_c()
        out[f"mesh{i}_tris"] = m.triangles
        out[f"mesh{i}_mat"] = np.float32([mc.kind, mc.albedo.x, mc.albedo.y, mc.albedo.z, mc.param])
    return out


def camera_array(cam):
    return np.frombuffer(bytes(cam.to_c()), np.uint8).copy()


def make(name, scene, cam, center, spread, spp, seed):
    osc = O.OracleScene.from_scene(scene)
    rays = np.concatenate([O.primary_rays(cam.to_c(), seed, 0), S.random_rays(2048, center, spread, seed)], 0)
    hits = osc.hit(rays)
    st = {}
    hdr = osc.render_hdr(cam.to_c(), spp, _abi.RenderOptsC(seed=seed), st)
    rgb = osc.render(cam.to_c(), spp, _abi.RenderOptsC(seed=seed))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), camera=camera_array(cam), rays=rays, hits=hits, hdr=hdr, rgb=rgb,
                        spp=np.int32(spp), seed=np.int64(seed), n_rays_rendered=np.int64(st["rays"]), **scene_arrays(scene))
    print(name, "rays", len(rays), "kinds", np.bincount(hits["kind"] + 1, minlength=3), "render rays", st["rays"])


if __name__ == "__main__":
    make("quirk_48x36", S.quirk_scene(), S.quirk_camera(48, 36), (0.0, 1.5, -4.0), 3.0, 4, 7)
    make("mesh1275_64x48", S.small_mesh_scene(3, 1275), S.example_camera(64, 48), (5.0, 1.4, -12.5), 4.0, 4, 11)
    make("spheres_64x48", S.spheres_scene(), S.example_camera(64, 48), (0.0, 2.0, -9.0), 5.0, 8, 13)
