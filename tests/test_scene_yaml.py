"""The scene file (blueprints.rs:15-48, read with serde_yaml at blueprints.rs:76-92) as the two hosts read it: the C++ `rbrt` command line
(csrc/host/scene_yaml.hpp, its own YAML reader) and the Python mirror (PyYAML's node graph + the same serde rules, written separately).
serde_yaml is absent from /root/reference and pinned only by the two scene files, so the checks are: the reference's
files, hand-made files for every rule scene_yaml.hpp's header lists, and generated files in mixed block / flow style on which both
hosts must read the same bits or both refuse."""
import os
import subprocess

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

import rbrt_b200 as R
from rbrt_b200 import blueprints as B

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "rbrt_b200", "rbrt")
REF_SCENES = "/root/reference/scenes"

CAMERA = """camera_blueprint:
  camera_up: {x: 0.0, y: 1.0, z: -0.4}
  camera_look_at: {x: 0.0, y: -0.1, z: -1.0}
  camera_position: {x: 0.0, y: 5.0, z: 4.0}
  camera_focal_length_mm: 28.0
"""


def cli_dump(path):
    out = subprocess.run([CLI, "-c", str(path), "--check", "--dump"], capture_output=True, text=True)
    assert out.returncode in (0, 101), (out.returncode, out.stderr)
    return (out.stdout if out.returncode == 0 else None), out.stderr


def py_dump(path):
    try:
        return B.dump_blueprint(R.load_blueprints_from_yaml_file(str(path))), ""
    except RuntimeError as e:
        return None, str(e)


def both(tmp_path, text, name="s.yaml"):
    p = tmp_path / name
    p.write_bytes(text.encode())
    c, c_err = cli_dump(p)
    y, y_err = py_dump(p)
    assert (c is None) == (y is None), f"C++: {c_err or 'ok'}\nPython: {y_err or 'ok'}\n{text}"
    assert c == y, text
    return c


def bits(x):
    return f"{int(np.float32(x).view(np.uint32)):08x}"


@pytest.mark.skipif(not os.path.isdir(REF_SCENES), reason="the reference tree is not on this machine")
def test_reference_scene_files_read_the_same():
    for name in ("example_scene.yaml", "header_card.yaml"):
        c, err = cli_dump(os.path.join(REF_SCENES, name))
        y, _ = py_dump(os.path.join(REF_SCENES, name))
        assert c is not None and c == y, err
        assert c.count("sphere\n") == (4 if name == "example_scene.yaml" else 7) and c.count("mesh\n") == 1


def test_styles_a_scene_author_may_use(tmp_path):
    block = CAMERA + """mesh_blueprints: []
sphere_blueprints:
  - radius: 1.5
    center:
      x: 1
      y: 2
      z: 3
    material_type: metal
    albedo: {x: 0.8, y: 0.8, z: 0.8}
    material_param: 0.005
"""
    want = both(tmp_path, block)
    assert want is not None and "radius " + bits(1.5) in want and "center " + " ".join(bits(v) for v in (1, 2, 3)) in want
    same = [
        # flow everything, on several lines, with comments, a trailing comma, quoted keys and a document marker
        """--- # a scene
{camera_blueprint: {camera_up: [0.0, 1.0, -0.4], camera_look_at: {x: 0.0, y: -0.1, z: -1.0},   # a struct may be the sequence of its fields
   camera_position: {"x": 0.0, 'y': 5.0, z: 4.0,}, camera_focal_length_mm: 28.0},
 mesh_blueprints: [],
 sphere_blueprints: [{radius: 1.5, center: [1, 2, 3], material_type: "metal", albedo: {x: 0.8, y: 0.8, z: 0.8}, material_param: 0.005}]}
...
""",
        # a sequence at its key's indentation, keys in another order, unknown keys, anchors
        CAMERA.replace("camera_up: {x: 0.0, y: 1.0, z: -0.4}", "camera_up: &up {z: -0.4, y: 1.0, x: 0.0, w: ignored}") + """comment: this key is not in the struct
mesh_blueprints: [   ]
sphere_blueprints:
- material_param: 5.0e-3
  material_type: 'metal'
  albedo: &grey
    x: 0.8
    y: 0.8
    z: 0.8
  center: {x: 1.0, y: +2, z: 0x3}
  radius: 1.5
  lights: [1, 2, {a: b}]
""",
    ]
    for text in same:
        assert both(tmp_path, text) == want, text
    # aliases
    text = block.replace("albedo: {x: 0.8, y: 0.8, z: 0.8}", "albedo: &a {x: 0.8, y: 0.8, z: 0.8}") + """  - radius: 2
    center: *a
    material_type: lambertian
    albedo: *a
"""
    got = both(tmp_path, text)
    assert got.count("albedo " + " ".join([bits(0.8)] * 3)) == 2 and "center " + " ".join([bits(0.8)] * 3) in got
    # Options: absent, null, ~ and empty all mean None; the material check comes later (blueprints.rs:50-74), not while parsing
    for none in ("", "    albedo: null\n    material_param: ~\n", "    albedo:\n    material_param: NULL\n"):
        text = CAMERA + "mesh_blueprints: []\nsphere_blueprints:\n  - radius: 1\n    center: [0, 0, 0]\n    material_type: lambertian\n" + none
        assert "albedo None\nmaterial_param None" in both(tmp_path, text)


def test_numbers_as_serde_yaml_reads_them(tmp_path):
    def focal(tok):
        return both(tmp_path, CAMERA.replace("camera_focal_length_mm: 28.0", f"camera_focal_length_mm: {tok}") + "mesh_blueprints: []\nsphere_blueprints: []\n")
    for tok, val in (("28", 28.0), ("+28", 28.0), ("-28", -28.0), ("2.8e1", 28.0), ("28.", 28.0), (".5", 0.5), ("+.5", 0.5), ("-.5e-1", -0.05), ("1E3", 1000.0),
                     ("0x1C", 28.0), ("-0x1c", -28.0), ("0o34", 28.0), ("0b11100", 28.0), ("0", 0.0), ("-0", 0.0), ("-0.0", -0.0), ("00.5", 0.5),
                     (".inf", np.inf), ("-.INF", -np.inf), ("+.Inf", np.inf), ("1e-60", 0.0), ("16777217", 16777216.0), ("16777219", 16777220.0),
                     ("18446744073709551615", 2.0 ** 64), ("340282366920938463463374607431768211455", np.inf), ("1e39", np.inf),
                     # through f64 first: 16777217.0000000001 is 16777217 as f64, a tie that goes to the even f32 (Rust parsing straight to f32 would say 16777218)
                     ("16777217.0000000001", 16777216.0)):
        got = focal(tok)
        assert got is not None and f"camera_focal_length_mm {bits(val)}" in got, (tok, got)
    assert "camera_focal_length_mm 7fc00000" in focal(".nan") or "camera_focal_length_mm ffc00000" in focal(".nan")
    for tok in ("'28.0'", '"28"', "true", "False", "~", "null", "", "028", "-007", "0x", "0xZZ", "1_000", "28mm", "inf", "nan", ".Nan", "++5", "+-5", "1e", "e5", ".", "1e400",
                "[28]", "{x: 28}", "0x-1C", "2 8"):
        assert focal(tok) is None, tok


def test_what_serde_refuses(tmp_path):
    ok = CAMERA + "mesh_blueprints: []\nsphere_blueprints: []\n"
    assert both(tmp_path, ok) is not None
    for text in (CAMERA + "sphere_blueprints: []\n",                                   # `mesh_blueprints` is a Vec, not an Option: it must be there
                 CAMERA + "mesh_blueprints: []\n",
                 CAMERA + "mesh_blueprints:\nsphere_blueprints: []\n",                  # null is not a sequence
                 CAMERA + "mesh_blueprints: {}\nsphere_blueprints: []\n",
                 "mesh_blueprints: []\nsphere_blueprints: []\n",
                 ok.replace("z: 4.0}", "z: 4.0, z: 5.0}"),                              # duplicate field
                 ok.replace("camera_up: {x: 0.0, y: 1.0, z: -0.4}", "camera_up: {x: 0.0, y: 1.0}"),
                 ok.replace("camera_up: {x: 0.0, y: 1.0, z: -0.4}", "camera_up: [0.0, 1.0]"),
                 ok.replace("camera_up: {x: 0.0, y: 1.0, z: -0.4}", "camera_up: [0.0, 1.0, 2.0, 3.0]"),
                 ok.replace("camera_up: {x: 0.0, y: 1.0, z: -0.4}", "camera_up: 1.0"),
                 ok + "---\n" + ok,                                                     # two documents
                 ok.replace("sphere_blueprints: []", "sphere_blueprints: [{radius: 1, center: [0, 0, 0]}]"),          # material_type missing
                 ok.replace("sphere_blueprints: []", "sphere_blueprints: [{radius: 1, center: [0, 0, 0], material_type: [metal]}]"),
                 ok.replace("sphere_blueprints: []", "sphere_blueprints: [[1, [0, 0, 0], metal]]"),                   # sequence form needs all five
                 ok.replace("camera_focal_length_mm: 28.0", "camera_focal_length_mm:\t28.0"),                             # libyaml: a tab cannot start a token
                 ok.replace("sphere_blueprints: []", "sphere_blueprints:\n-\tradius: 1"),
                 "", "# nothing\n", "just a scalar\n", "[1, 2\n", "{a: 1\n", "a: 'unterminated\n", "a: b: c\n", "a: *nowhere\n"):
        assert both(tmp_path, text) is None, text
    for text in (ok.replace("\n", "\r\n"), ok.replace("\n", "   \n"), "\ufeff" + ok):
        assert both(tmp_path, text) is not None, text                                   # CRLF, trailing blanks, a BOM
    assert both(tmp_path, ok.replace("camera_blueprint:\n", "camera_blueprint:\t\n")) is None       # a tab where the value would start
    # One known difference between the hosts: a tab AFTER a value (`28.0<tab>`, `}<tab># comment`).  libyaml — hence serde_yaml — skips it, and so does
    # the CLI; PyYAML's pure-Python scanner never skips tabs and refuses the file.
    p = tmp_path / "tab.yaml"
    p.write_text(ok.replace("28.0\n", "28.0\t\n").replace("z: 4.0}", "z: 4.0}\t# tab before a comment"))
    assert cli_dump(p)[0] is not None and py_dump(p)[0] is None
    # ... and what it takes that one might not expect: the sequence form of a whole blueprint, a numeric-looking material_type
    text = ok.replace("sphere_blueprints: []", "sphere_blueprints: [[1, [0, 0, 0], 42, ~, 0.5]]")
    assert "material_type 2:42\nalbedo None\nmaterial_param " + bits(0.5) in both(tmp_path, text)


# ---- generated files ---------------------------------------------------------------------------------------------------------
GOOD_NUMBERS = ["0", "1", "-1", "28.0", "-0.4", "1e3", "2.5E-2", ".5", "+3", "7.", "0x10", "0o17", "0b101", "1000.0", "-1000", "16777217.0000000001", "0.1", "3.4e38",
                "1e-46", ".inf", "-.inf", "123456789", "9007199254740993"]
BAD_NUMBERS = ["'1.0'", '"2"', "true", "~", "", "007", "1_0", "abc", "1e", "0x", "inf", "[1]", "1e999"]
STRINGS = ["metal", "Lambertian", "my dielectric glass", "plastic", "'metal'", '"dielectric"', "bunny.obj", "/tmp/a b.obj", "'quoted: colon'", '"esc\\t\\"x\\""', "42", "~"]


@st.composite
def scene_text(draw):
    dirty = draw(st.integers(0, 3)) == 0
    num = st.sampled_from(GOOD_NUMBERS + BAD_NUMBERS) if dirty else st.sampled_from(GOOD_NUMBERS)

    def emit(value, indent, flow):
        """YAML text of `value` (dict / list / token string) as the value of a key written at `indent`; returns what follows `key:`."""
        if isinstance(value, str):
            return " " + value if value else ""
        flow = flow or draw(st.integers(0, 2)) == 0
        pad = " " * (indent + 2)
        if isinstance(value, dict):
            items = list(value.items())
            if draw(st.booleans()):
                items = draw(st.permutations(items))
            if flow:
                inner = ", ".join(f"{k}:{emit(v, 0, True) or ' '}" for k, v in items)
                return " {" + inner + ("," if items and draw(st.integers(0, 4)) == 0 else "") + "}"
            return "".join(f"\n{pad}{k}:{emit(v, indent + 2, False)}" for k, v in items) if items else " {}"
        if flow or not value:
            return " [" + ", ".join(emit(v, 0, True).strip() for v in value) + "]"
        out = ""
        same_level = draw(st.booleans())                                     # "key:\n- a" is allowed
        ipad = " " * indent if same_level else pad
        for v in value:
            if isinstance(v, dict) and not draw(st.integers(0, 3)) == 0:
                items = list(v.items())
                first, rest = items[0], items[1:]
                out += f"\n{ipad}- {first[0]}:{emit(first[1], len(ipad) + 2, False)}"
                out += "".join(f"\n{ipad}  {k}:{emit(x, len(ipad) + 2, False)}" for k, x in rest)
            else:
                out += f"\n{ipad}-{emit(v, len(ipad), True)}"
        return out

    def vec3():
        comps = {"x": draw(num), "y": draw(num), "z": draw(num)}
        form = draw(st.integers(0, 5))
        if form == 0:
            return list(comps.values())
        if dirty and form == 1:
            comps.pop(draw(st.sampled_from("xyz")))
        if form == 2:
            comps["extra"] = "1"
        return comps

    def material():
        m = {"material_type": draw(st.sampled_from(STRINGS))}
        if draw(st.booleans()):
            m["albedo"] = draw(st.sampled_from(["~", "null", ""])) if draw(st.integers(0, 4)) == 0 else vec3()
        if draw(st.booleans()):
            m["material_param"] = draw(st.sampled_from(["~", ""])) if draw(st.integers(0, 4)) == 0 else draw(num)
        return m

    cam = {"camera_up": vec3(), "camera_look_at": vec3(), "camera_position": vec3(), "camera_focal_length_mm": draw(num)}
    meshes = [dict({"obj_filepath": draw(st.sampled_from(STRINGS)), "scale": draw(num), "translation": vec3(), "rotation_rad": vec3()}, **material())
              for _ in range(draw(st.integers(0, 2)))]
    spheres = [dict({"radius": draw(num), "center": vec3()}, **material()) for _ in range(draw(st.integers(0, 3)))]
    top = {"camera_blueprint": cam, "mesh_blueprints": meshes, "sphere_blueprints": spheres}
    if dirty and draw(st.integers(0, 3)) == 0:
        top.pop(draw(st.sampled_from(list(top))))
    items = draw(st.permutations(list(top.items())))
    text = "".join(f"{k}:{emit(v, 0, False)}\n" for k, v in items)
    if draw(st.integers(0, 3)) == 0:
        text = "---\n" + text
    if draw(st.integers(0, 3)) == 0:
        text = text.replace("\n", "   # note\n", 1)
    return text


@settings(max_examples=250, deadline=None, derandomize=True, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
@given(text=scene_text())
def test_generated_scene_files_read_the_same(tmp_path, text):
    both(tmp_path, text)
