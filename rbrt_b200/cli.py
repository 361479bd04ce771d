"""`python -m rbrt_b200` — the rbrt CLI (src/main.rs:9-92): same flags, defaults and call sequence."""
import argparse

from .blueprints import create_scene_from_scene_blueprint, load_blueprints_from_yaml_file
from .cam import Camera
from .render import render_scene


def build_parser():
    p = argparse.ArgumentParser(prog="rbrt", description="a lighweight raytracer written in rust")  # main.rs:10-13
    p.add_argument("-t", "--target_file", default="dbg_out.png", help="file that will be created witht he rendered output")
    p.add_argument("--height", type=int, default=600, help="target image resolution height")
    p.add_argument("-w", "--width", type=int, default=800, help="target image resolution width")
    p.add_argument("-c", "--config", default="scenes/example_scene.yaml",
                   help="YAML file that specifies the scene layout and camera specification.")
    p.add_argument("-s", "--samples", type=int, default=5, help="number of rays per pixel")
    p.add_argument("--seed", type=int, default=0, help="(extension) Philox seed; the reference is unseeded")
    return p


def main(argv=None):
    a = build_parser().parse_args(argv)
    scene_bp = load_blueprints_from_yaml_file(a.config)
    cb = scene_bp.camera_blueprint
    cam = Camera.new(cb.camera_position, cb.camera_look_at, cb.camera_up, a.height, a.width,
                     cb.camera_focal_length_mm)  # height before width (main.rs:71-78)
    scene = create_scene_from_scene_blueprint(scene_bp)
    img = render_scene(cam, a.samples, scene, seed=a.seed)
    print(f"Saving rendered image to {a.target_file}")  # main.rs:84
    try:
        img.save(a.target_file)
    except OSError:
        raise SystemExit(f"Unable to save target img to {a.target_file}! Maybe the directory does not exist?")
    return 0
