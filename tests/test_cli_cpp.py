"""The compiled host side: rbrt_b200/rbrt, the reference's command line (src/main.rs) in C++ over the C-ABI
(rbrt_b200/csrc/host/rbrt_cli.cpp) — flags/defaults, the YAML scene schema of blueprints.rs, .obj loading with the
reference's transform, material matching, PNG output."""
import os
import subprocess

import numpy as np
import pytest

import rbrt_b200 as R
from rbrt_b200 import png, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "rbrt_b200", "rbrt")

YAML = """---
# scene for the CLI tests (schema of scenes/example_scene.yaml)
camera_blueprint:
  camera_up:
    x: 0.0
    y: 1.0
    z: -0.4
  camera_look_at:
    x: 0.0
    y: -0.1
    z: -1.0
  camera_position: {x: 0.0, y: 5.0, z: 4.0}
  camera_focal_length_mm: 28.0
mesh_blueprints:
  - obj_filepath: %s
    scale: 45.0
    translation:
      x: 5.0
      y: -1.8
      z: -12.5
    rotation_rad:
      x: 0.0
      y: 0.3
      z: 0.0
    material_type: "dielectric"
    material_param: 0.2
    albedo:
      x: 0.8
      y: 0.8
      z: 0.8
sphere_blueprints:
# green earth
  - radius: 1000.0 
    center:
      x: 0.0
      y: -1000.0
      z: -5.0
    material_type: "lambertian"
    albedo:
      x: 0.02
      y: 0.2
      z: 0.1
  - radius: 3.0
    center:
      x: -2.5
      y: 2.9
      z: -15.0
    material_type: "Polished Metal"   # substring match on the lower-cased type
    albedo:
      x: 0.8
      y: 0.8
      z: 0.8
    material_param: 0.005
  - radius: 1.0
    center:
      x: 0.0
      y: 1.0
      z: -3.0
    material_type: "plastic"
"""


@pytest.fixture(scope="module")
def cli():
    if not os.path.exists(CLI):
        subprocess.run(["make", "-C", os.path.join(ROOT, "rbrt_b200", "csrc")], check=True)
    return CLI


def write_scene(tmp_path, flow_map=False):
    obj = tmp_path / "m.obj"
    n = synth.write_bunny_standin(str(obj), subdiv=2)
    y = tmp_path / "scene.yaml"
    text = YAML % obj
    if not flow_map:                                       # the reference's files use block style only
        text = text.replace("camera_position: {x: 0.0, y: 5.0, z: 4.0}", "camera_position:\n    x: 0.0\n    y: 5.0\n    z: 4.0")
    y.write_text(text)
    return str(y), n


def test_cli_help_and_flag_errors(cli):
    out = subprocess.run([cli, "--help"], capture_output=True, text=True)
    assert out.returncode == 0
    for flag, default in [("--target_file", "dbg_out.png"), ("--height", "600"), ("--width", "800"), ("--config", "scenes/example_scene.yaml"), ("--samples", "5")]:
        assert flag in out.stdout and f"[default: {default}]" in out.stdout          # main.rs:14-50
    assert subprocess.run([cli, "--bogus"], capture_output=True).returncode == 2
    assert subprocess.run([cli, "-c", "/nonexistent.yaml", "--check"], capture_output=True).returncode == 101   # panic exit code


def test_cli_parses_scene_like_the_python_mirror(cli, tmp_path):
    y, n = write_scene(tmp_path)
    out = subprocess.run([cli, "-c", y, "--height", "48", "-w", "64", "--check"], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert "Cannot figure out material_type from plastic" in out.stdout            # blueprints.rs:70-73: skipped, not fatal
    assert f"Successfully loaded {n} triangles" in out.stdout                      # mesh.rs:115-119
    assert f"scene ok: 2 spheres, 1 meshes, {n} triangles, camera 64x48" in out.stdout
    bp = R.load_blueprints_from_yaml_file(y)
    assert len(bp.sphere_blueprints) == 3 and len(bp.mesh_blueprints) == 1


def test_cli_image_writers_hold_the_pixels(cli, tmp_path):
    """The CLI's writers for every format it offers (main.rs:86: `image::save` picks by extension), without a GPU: `--from-ppm` feeds them an
    image; PNG, BMP, TGA, TIFF and QOI files decode (PIL) to the same pixels, and the Python mirror writes byte-identical BMP / TGA / TIFF / QOI."""
    from PIL import Image
    rng = np.random.default_rng(3)
    smooth = (np.cumsum(rng.integers(-3, 4, size=(40, 31, 3)), axis=1) % 256).astype(np.uint8)
    smooth[5:9] = 0; smooth[20:23, 4:20] = smooth[20, 3]
    palette = rng.integers(0, 256, size=(5, 3), dtype=np.uint8)[rng.integers(0, 5, size=(16, 16))]
    for k, img in enumerate((rng.integers(0, 256, size=(10, 13, 3), dtype=np.uint8), smooth, palette, np.zeros((3, 70, 3), np.uint8))):
        src = str(tmp_path / f"in{k}.ppm")
        R.ImageBuffer(img).save(src)
        for ext in ("png", "ppm", "bmp", "tga", "tif", "tiff", "qoi"):
            out = str(tmp_path / f"o{k}.{ext}")
            r = subprocess.run([cli, "--from-ppm", src, "-t", out], capture_output=True, text=True)
            assert r.returncode == 0, r.stderr
            assert np.array_equal(np.asarray(Image.open(out).convert("RGB")), img), (k, ext)
            if ext not in ("png",):                                                  # (zlib streams may differ; the pixels may not)
                mine = str(tmp_path / f"p{k}.{ext}")
                R.ImageBuffer(img).save(mine)
                assert open(mine, "rb").read() == open(out, "rb").read(), (k, ext)
    lossy = subprocess.run([cli, "--from-ppm", src, "-t", str(tmp_path / "o.jpg")], capture_output=True, text=True)
    assert lossy.returncode == 101 and "Unable to save target img" in lossy.stderr   # main.rs:86-91


def test_cli_missing_fields_are_fatal(cli, tmp_path):
    y, _ = write_scene(tmp_path)
    txt = open(y).read().replace("    material_param: 0.005\n", "")
    open(y, "w").write(txt)
    out = subprocess.run([cli, "-c", y, "--check"], capture_output=True, text=True)
    assert out.returncode == 101 and "you forgot to specify a roughness" in out.stderr   # blueprints.rs:58-59 `expect`


@pytest.mark.gpu
def test_cli_render_matches_python_host(cli, tmp_path, gpu):
    """Same scene through the C++ CLI and through the Python mirror: the two hosts must hand the library identical
    inputs, so the PNG holds the identical pixels."""
    y, _ = write_scene(tmp_path)
    target = str(tmp_path / "out.png")
    out = subprocess.run([cli, "-c", y, "--height", "96", "-w", "128", "-s", "3", "-t", target, "--seed", "5"], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert "Starting rendering..." in out.stdout and f"Saving rendered image to {target}" in out.stdout   # lib.rs:80, main.rs:84
    got = png.decode_png(open(target, "rb").read())
    bp = R.load_blueprints_from_yaml_file(y)
    cb = bp.camera_blueprint
    cam = R.Camera.new(cb.camera_position, cb.camera_look_at, cb.camera_up, 96, 128, cb.camera_focal_length_mm)
    ref = R.render_scene(cam, 3, R.create_scene_from_scene_blueprint(bp), seed=5).pixels
    assert np.array_equal(got, ref)
    ppm = str(tmp_path / "out.ppm")
    assert subprocess.run([cli, "-c", y, "--height", "8", "-w", "8", "-s", "1", "-t", ppm], capture_output=True).returncode == 0
    assert open(ppm, "rb").read().startswith(b"P6\n8 8\n255\n")
    # the other lossless formats the `image` crate picks by extension (main.rs:86): 24-bit BMP (bottom-up BGR, rows padded
    # to 4 bytes) and uncompressed TGA (top-down BGR) hold the PNG's pixels
    small = [cli, "-c", y, "--height", "10", "-w", "13", "-s", "2", "--seed", "5"]
    files = {e: str(tmp_path / f"s.{e}") for e in ("png", "bmp", "tga")}
    for f in files.values():
        assert subprocess.run(small + ["-t", f], capture_output=True).returncode == 0
    want = png.decode_png(open(files["png"], "rb").read())
    bmp = open(files["bmp"], "rb").read()
    stride = (3 * 13 + 3) & ~3
    assert bmp[:2] == b"BM" and len(bmp) == 54 + stride * 10 and int.from_bytes(bmp[18:22], "little") == 13 and int.from_bytes(bmp[22:26], "little") == 10
    rows = np.frombuffer(bmp[54:], np.uint8).reshape(10, stride)[::-1, :39].reshape(10, 13, 3)[:, :, ::-1]
    assert np.array_equal(rows, want)
    tga = open(files["tga"], "rb").read()
    assert tga[2] == 2 and tga[12:16] == bytes([13, 0, 10, 0]) and tga[16] == 24 and tga[17] == 0x20 and len(tga) == 18 + 390
    assert np.array_equal(np.frombuffer(tga[18:], np.uint8).reshape(10, 13, 3)[:, :, ::-1], want)
    lossy = subprocess.run(small + ["-t", str(tmp_path / "s.jpg")], capture_output=True, text=True)
    assert lossy.returncode == 101 and "Unable to save target img" in lossy.stderr
    bad = subprocess.run([cli, "-c", y, "--height", "8", "-w", "8", "-s", "1", "-t", str(tmp_path / "nodir" / "o.png")], capture_output=True, text=True)
    assert bad.returncode == 101 and "Unable to save target img" in bad.stderr      # main.rs:86-91
