"""Host side of the drop-in: YAML blueprints, material matching, OBJ loading with the
reference's transform order, the CLI's flags/defaults, PNG save.  No GPU needed."""
import os

import numpy as np
import pytest

import rbrt_b200 as R
from rbrt_b200 import blueprints as B
from rbrt_b200 import cli, mesh, png, synth
from rbrt_b200.vec3 import Vec3

YAML = """---
camera_blueprint:
  camera_up: {x: 0.0, y: 1.0, z: -0.4}
  camera_look_at: {x: 0.0, y: -0.1, z: -1.0}
  camera_position: {x: 0.0, y: 5.0, z: 4.0}
  camera_focal_length_mm: 28.0
mesh_blueprints:
  - obj_filepath: %s
    scale: 45.0
    translation: {x: 5.0, y: -1.8, z: -12.5}
    rotation_rad: {x: 0.0, y: 0.0, z: 0.0}
    material_type: "dielectric"
    material_param: 0.2
    albedo: {x: 0.8, y: 0.8, z: 0.8}
sphere_blueprints:
  - radius: 1000.0
    center: {x: 0.0, y: -1000.0, z: -5.0}
    material_type: "lambertian"
    albedo: {x: 0.02, y: 0.2, z: 0.1}
  - radius: 3.0
    center: {x: -2.5, y: 2.9, z: -15.0}
    material_type: "Shiny METAL"
    albedo: {x: 0.8, y: 0.8, z: 0.8}
    material_param: 0.005
  - radius: 1.0
    center: {x: 0.0, y: 0.0, z: 0.0}
    material_type: "plastic"
"""


def test_yaml_blueprint_and_scene_creation(tmp_path, capsys):
    obj = tmp_path / "m.obj"
    n = synth.write_bunny_standin(str(obj), subdiv=1)
    y = tmp_path / "scene.yaml"
    y.write_text(YAML % obj)
    bp = R.load_blueprints_from_yaml_file(str(y))
    assert bp.camera_blueprint.camera_up.as_tuple() == Vec3(0.0, 1.0, -0.4).as_tuple()
    assert len(bp.mesh_blueprints) == 1 and len(bp.sphere_blueprints) == 3
    scene = R.create_scene_from_scene_blueprint(bp)
    out = capsys.readouterr().out
    assert "Cannot figure out material_type from plastic" in out          # blueprints.rs:70-73: skipped, not fatal
    assert len(scene.elements) == 2 and len(scene.triangle_meshes) == 1
    assert isinstance(scene.elements[1].material, R.Metal)                 # substring match on the lower-cased type
    assert isinstance(scene.triangle_meshes[0].material, R.Dielectric) and scene.triangle_meshes[0].material.ref_idx == pytest.approx(0.2)
    assert scene.triangle_meshes[0].triangles.shape == (n, 3, 3)
    assert f"Successfully loaded {n} triangles" in out                      # mesh.rs:115-119


def test_reference_scene_files_parse():
    """The reference's own fixtures (scenes/*.yaml) parse when the reference tree is mounted."""
    d = "/root/reference/scenes"
    if not os.path.isdir(d):
        pytest.skip("reference tree not mounted on this box")
    ex = R.load_blueprints_from_yaml_file(os.path.join(d, "example_scene.yaml"))
    assert len(ex.sphere_blueprints) == 4 and ex.mesh_blueprints[0].obj_filepath == "bunny.obj"
    mine = synth.spheres_only_blueprint()
    assert [(s.radius, s.center.as_tuple(), s.material_type) for s in ex.sphere_blueprints] == \
           [(s.radius, s.center.as_tuple(), s.material_type) for s in mine.sphere_blueprints]
    assert ex.camera_blueprint.camera_position.as_tuple() == mine.camera_blueprint.camera_position.as_tuple()
    hc = R.load_blueprints_from_yaml_file(os.path.join(d, "header_card.yaml"))
    assert len(hc.sphere_blueprints) == 7 and len(hc.mesh_blueprints) == 1
    # synth.header_card_blueprint (what the GPU parity test renders, the reference tree being absent on the GPU box) restates it exactly
    mine = synth.header_card_blueprint("bunny.obj")
    key = lambda s: (s.radius, s.center.as_tuple(), s.material_type, None if s.albedo is None else s.albedo.as_tuple(), s.material_param)
    assert [key(s) for s in hc.sphere_blueprints] == [key(s) for s in mine.sphere_blueprints]
    m, n = hc.mesh_blueprints[0], mine.mesh_blueprints[0]
    assert (m.obj_filepath, m.scale, m.translation.as_tuple(), m.rotation_rad.as_tuple(), m.material_type, m.albedo.as_tuple(), m.material_param) == \
           (n.obj_filepath, n.scale, n.translation.as_tuple(), n.rotation_rad.as_tuple(), n.material_type, n.albedo.as_tuple(), n.material_param)
    cb, cm = hc.camera_blueprint, mine.camera_blueprint
    assert (cb.camera_up.as_tuple(), cb.camera_look_at.as_tuple(), cb.camera_position.as_tuple(), cb.camera_focal_length_mm) == \
           (cm.camera_up.as_tuple(), cm.camera_look_at.as_tuple(), cm.camera_position.as_tuple(), cm.camera_focal_length_mm)


def test_blueprint_yaml_round_trip(tmp_path):
    bp = synth.header_card_blueprint("some.obj")
    p = tmp_path / "scene.yaml"
    p.write_text(synth.blueprint_to_yaml(bp))
    back = R.load_blueprints_from_yaml_file(str(p))
    assert back == bp


def test_material_description_errors():
    with pytest.raises(ValueError):
        B.create_material_from_description("metal", Vec3(1, 1, 1), None)   # blueprints.rs: expect(roughness)
    with pytest.raises(ValueError):
        B.create_material_from_description("lambertian", None, None)
    with pytest.raises(ValueError):
        B.create_material_from_description("dielectric", None, None)
    with pytest.raises(RuntimeError):
        R.load_blueprints_from_yaml_file("/nonexistent/scene.yaml")


def test_obj_loader_transform_order(tmp_path):
    """scale -> rotate_point (Z-X-Z) -> translate, in f32 (mesh.rs:102-112); faces cut into index triples per
    model, `v/vt/vn` tokens and negative indices accepted, non-face records ignored."""
    obj = tmp_path / "t.obj"
    obj.write_text("# c\nv 1 0 0\nv 0 1 0\nv 0 0 1\nvn 0 0 1\nvt 0 0\no a\nf 1/1/1 2/1/1 3/1/1\ng b\nf -3 -1 -2\n")
    pos, idx = mesh.parse_obj_triangles(str(obj))
    assert pos.shape == (3, 3) and idx.tolist() == [[0, 1, 2], [0, 2, 1]]
    tris = R.load_mesh_vertices_from_file(str(obj), Vec3(1, 2, 3), Vec3(0, 0, float(np.float32(np.pi / 2))), 2.0)
    assert tris.shape == (2, 3, 3)
    assert np.allclose(tris[0], [[1, 4, 3], [-1, 2, 3], [1, 2, 5]], atol=1e-6)   # (2,0,0) rotated about z by 90 deg -> (0,2,0), + t
    empty = tmp_path / "e.obj"
    empty.write_text("v 0 0 0\n")
    assert R.load_mesh_vertices_from_file(str(empty), Vec3(0, 0, 0), Vec3(0, 0, 0), 1.0).shape == (0, 3, 3)


def test_cli_flags_and_defaults():  # main.rs:14-50
    a = cli.build_parser().parse_args([])
    assert (a.target_file, a.height, a.width, a.config, a.samples) == ("dbg_out.png", 600, 800, "scenes/example_scene.yaml", 5)
    a = cli.build_parser().parse_args(["-t", "o.png", "--height", "768", "-w", "1024", "-c", "s.yaml", "-s", "50"])
    assert (a.target_file, a.height, a.width, a.config, a.samples) == ("o.png", 768, 1024, "s.yaml", 50)
    a = cli.build_parser().parse_args(["--target_file", "x.png", "--width", "10", "--config", "c.yaml", "--samples", "2"])
    assert (a.target_file, a.width, a.config, a.samples) == ("x.png", 10, "c.yaml", 2)


def test_png_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, size=(37, 53, 3), dtype=np.uint8)
    p = tmp_path / "o.png"
    R.ImageBuffer(img).save(str(p))
    assert np.array_equal(png.decode_png(p.read_bytes()), img)
    try:
        from PIL import Image
        assert np.array_equal(np.asarray(Image.open(str(p)).convert("RGB")), img)
    except ImportError:
        pass
    with pytest.raises(ValueError):
        R.ImageBuffer(img).save(str(tmp_path / "o.xyz"))
    assert R.ImageBuffer(img).get_pixel(5, 7) == tuple(int(v) for v in img[7, 5])   # (x, y) like image::ImageBuffer


def test_bmp_and_tga_hold_the_same_pixels(tmp_path):
    """The other lossless formats `image::save` picks by extension (main.rs:86); same byte layout as the C++ CLI's writers
    (tests/test_cli_cpp.py decodes those the same way)."""
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, size=(10, 13, 3), dtype=np.uint8)       # 13 px: BMP rows need 1 byte of padding
    b, t = tmp_path / "o.bmp", tmp_path / "o.tga"
    R.ImageBuffer(img).save(str(b)); R.ImageBuffer(img).save(str(t))
    bmp, tga = b.read_bytes(), t.read_bytes()
    stride = (3 * 13 + 3) & ~3
    assert bmp[:2] == b"BM" and len(bmp) == 54 + stride * 10 and int.from_bytes(bmp[2:6], "little") == len(bmp) and int.from_bytes(bmp[10:14], "little") == 54
    assert int.from_bytes(bmp[18:22], "little") == 13 and int.from_bytes(bmp[22:26], "little") == 10 and int.from_bytes(bmp[28:30], "little") == 24
    assert np.array_equal(np.frombuffer(bmp[54:], np.uint8).reshape(10, stride)[::-1, :39].reshape(10, 13, 3)[:, :, ::-1], img)
    assert len(tga) == 18 + 390 and tga[2] == 2 and tga[12:16] == bytes([13, 0, 10, 0]) and tga[16] == 24 and tga[17] == 0x20
    assert np.array_equal(np.frombuffer(tga[18:], np.uint8).reshape(10, 13, 3)[:, :, ::-1], img)
    try:
        from PIL import Image
        for f in (b, t):
            assert np.array_equal(np.asarray(Image.open(str(f)).convert("RGB")), img)
    except ImportError:
        pass


def test_tiff_and_qoi_hold_the_same_pixels(tmp_path):
    """Two more lossless formats of `image::save`; decoded here with PIL.  The images have runs, repeats, small and large steps and black
    pixels so that every QOI op is written (run, index, diff, luma, rgb)."""
    from PIL import Image
    rng = np.random.default_rng(2)
    noise = rng.integers(0, 256, size=(9, 11, 3), dtype=np.uint8)
    smooth = (np.cumsum(rng.integers(-3, 4, size=(40, 31, 3)), axis=1) % 256).astype(np.uint8)       # diff / luma ops
    smooth[5:9] = 0; smooth[20:23, 4:20] = smooth[20, 3]                                             # runs (also > 62 long), black after colour
    palette = rng.integers(0, 256, size=(5, 3), dtype=np.uint8)[rng.integers(0, 5, size=(16, 16))]   # index ops
    for k, img in enumerate((noise, smooth, palette, np.zeros((3, 70, 3), np.uint8), np.full((1, 1, 3), 7, np.uint8))):
        for ext in ("tif", "tiff", "qoi"):
            f = tmp_path / f"o{k}.{ext}"
            R.ImageBuffer(img).save(str(f))
            assert np.array_equal(np.asarray(Image.open(str(f)).convert("RGB")), img), (k, ext)
    ops = (tmp_path / "o1.qoi").read_bytes()[14:-8]
    assert any(b >> 6 == 3 and b < 0xFE for b in ops) and any(b >> 6 == 1 for b in ops) and any(b >> 6 == 2 for b in ops)


def test_camera_new_argument_order():
    cam = R.Camera.new(Vec3(0, 5, 4), Vec3(0, -0.1, -1), Vec3(0, 1, -0.4), 600, 800, 28.0)   # height BEFORE width
    assert (cam.img_height_pix, cam.img_width_pix) == (600, 800) and cam.img_width_mm == 35.0
    assert cam.img_height_mm == pytest.approx(35.0 * 600 / 800)


def test_synthetic_meshes_respect_the_determinant_cull():
    """SURVEY.md hard part: 2*area of the synthetic triangles must sit well above the 1e-3 absolute cull."""
    for subdiv, radius, floor in [(6, 3.15, 2e-3), (8, 40.0, 2e-2)]:   # C2 is bunny-sized like the fixture (grazing rays ARE culled there)
        v, f = synth.icosphere(min(subdiv, 5))
        scale = 4.0 ** (subdiv - min(subdiv, 5))
        t = (v * radius)[f]
        area2 = np.linalg.norm(np.cross(t[:, 1] - t[:, 0], t[:, 2] - t[:, 0]), axis=1) / scale
        assert np.median(area2) > floor, (subdiv, np.median(area2))
