"""rbrt_mesh_load_obj (csrc/obj_loader.cpp, CPU code inside librbrt_gpu.so) = load_mesh_vertices_from_file (mesh.rs:78-121).

tobj is absent from /root/reference and unpinned by the reference's tests, so the checks are: (1) hand-made files for
every rule the loader's header spells out, (2) the multi-threaded C++ loader against the serial Python restatement
(rbrt_b200.mesh.parse_obj_triangles) on generated files, bit for bit, for every way of cutting the file into pieces,
(3) the f32 parse against correctly rounded values computed with exact rationals."""
import os
from fractions import Fraction

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

from rbrt_b200 import _abi, mesh

ZERO = (0.0, 0.0, 0.0)


def load_c(path, translation=ZERO, rotation=ZERO, scale=1.0, piece_bytes=None, threads=None):
    old = {k: os.environ.get(k) for k in ("RBRT_OBJ_PIECE_BYTES", "RBRT_HOST_THREADS")}
    try:
        if piece_bytes is not None:
            os.environ["RBRT_OBJ_PIECE_BYTES"] = str(piece_bytes)
        if threads is not None:
            os.environ["RBRT_HOST_THREADS"] = str(threads)
        return mesh._load_obj_soup(str(path), translation, rotation, scale)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def load_py(path, translation=ZERO, rotation=ZERO, scale=1.0):
    return mesh.load_obj_soup_python(str(path), translation, rotation, scale)


def same_bits(a, b):
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


def both(path, **kw):
    """Both loaders on one file: the same soup, or the same refusal."""
    try:
        ref = load_py(path, **kw)
    except mesh.ObjLoadError:
        for pb in (None, 16):
            with pytest.raises(mesh.ObjLoadError):
                load_c(path, piece_bytes=pb, **kw)
        return None
    for pb, th in ((None, None), (16, 7), (64, 3), (1, 64)):
        got = load_c(path, piece_bytes=pb, threads=th, **kw)
        assert same_bits(got, ref), (pb, th)
    return ref


def test_records_and_model_boundaries(tmp_path):
    p = tmp_path / "a.obj"
    # quads are NOT triangulated: the reference walks the index list in triples (mesh.rs:96), so 4 + 4 indices = 2 "triangles" + 2 dropped
    p.write_text("v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nf 1 2 3 4\nf 4 3 2 1\n")
    soup = both(p)
    assert soup.shape == (2, 3, 3)
    assert soup[0].tolist() == [[0, 0, 0], [1, 0, 0], [1, 1, 0]] and soup[1].tolist() == [[0, 1, 0], [0, 1, 0], [1, 1, 0]]
    # an `o` between them makes two models: each is cut on its own -> 1 + 1 triangles, both in phase
    p.write_text("v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nf 1 2 3 4\no second\nf 4 3 2 1\n")
    soup = both(p)
    assert soup[1].tolist() == [[0, 1, 0], [1, 1, 0], [1, 0, 0]]
    # `l` records add their two indices (tobj's default keeps lines), so the triangle after one is out of phase
    p.write_text("v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nl 1 2\nf 2 3 4\nf 1 2 3\n")
    soup = both(p)
    assert soup.shape == (2, 3, 3) and soup[0].tolist() == [[0, 0, 0], [1, 0, 0], [1, 0, 0]]
    # `o` / `g` with nothing pending do not split; unknown records and comments are ignored; "\r\n" line ends
    p.write_bytes(b"# head\r\no a\r\ng b\r\nv 0 0 0\r\nv 1 0 0\r\nv 0 1 0\r\ns off\r\nvp 1 2\r\nf 1 2 3\r\n#f 1 1 1\r\n")
    assert both(p).shape == (1, 3, 3)
    # v/vt/vn forms, relative indices, vertex colours after the position, extra white space, no newline at the end
    p.write_text("v 0 0 0 1 0 0\nv\t1   0 0\nv 0 1 0\nvt 0.5 0.5\nvn 0 0 1\nf 1/1/1 2//1 3/1\nf -3/-1/-1 -2 -1")
    soup = both(p)
    assert soup.shape == (2, 3, 3) and same_bits(soup[0], soup[1])
    # empty file, vertices only, faces only in a later model
    p.write_text("")
    assert both(p).shape == (0, 3, 3)
    p.write_text("v 0 0 0\nv 1 1 1\n")
    assert both(p).shape == (0, 3, 3)


def test_usemtl_splits_only_when_the_material_id_changes(tmp_path):
    (tmp_path / "m.mtl").write_text("newmtl red\nKd 1 0 0\nnewmtl blue\nKd 0 0 1\nillum 2\n")
    (tmp_path / "bad.mtl").write_text("newmtl green\nKd 1 oops 0\n")
    verts = "v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\n"
    quads = "f 1 2 3 4\n{}f 4 3 2 1\n"
    p = tmp_path / "a.obj"
    in_phase, out_of_phase = [[0, 1, 0], [1, 1, 0], [1, 0, 0]], [[0, 1, 0], [0, 1, 0], [1, 1, 0]]
    for middle, want in (("usemtl red\n", out_of_phase),                               # no library loaded: every name is "no material"
                         ("mtllib m.mtl\nusemtl red\n", in_phase),                     # None -> red: split
                         ("mtllib nope.mtl\nusemtl red\n", out_of_phase),              # unreadable library
                         ("mtllib bad.mtl\nusemtl green\n", out_of_phase),             # a library tobj fails to parse contributes nothing
                         ("mtllib m.mtl\nusemtl purple\n", out_of_phase)):             # unknown name
        p.write_text(verts + quads.format(middle))
        assert both(p)[1].tolist() == want, middle
    p.write_text("mtllib m.mtl\n" + verts + "usemtl red\nf 1 2 3 4\nusemtl red\nf 4 3 2 1\nusemtl blue\nf 1 2 3 4\n")
    soup = both(p)                                                                      # red+red = one model of 8, blue its own
    assert soup.shape == (3, 3, 3) and soup[1].tolist() == out_of_phase and soup[2].tolist() == [[0, 0, 0], [1, 0, 0], [1, 1, 0]]
    p.write_text(verts + "usemtl\nf 1 2 3\n")
    assert both(p) is None


def test_loads_the_reference_refuses(tmp_path):
    p = tmp_path / "a.obj"
    tri = "v 0 0 0\nv 1 0 0\nv 0 1 0\n"
    for text in (tri + "f 1 2 4\n",                  # beyond the positions
                 tri + "f 0 1 2\n",                  # index 0 wraps
                 tri + "f -4 1 2\n",                 # relative, before the first
                 tri + "f 1 2 x\n", tri + "f 1/1/1/1 2 3\n", tri + "f 1.0 2 3\n",
                 "v 0 0\n", "v 0 zero 0\n", "v 0x10 0 0\n", "v 1_000 0 0\n", "vt 1\n", "vn 0 0\n",
                 tri + "vt 0 0\nf 1/2 2/1 3/1\n",    # texcoord index beyond the texcoords (only checked when the file has any)
                 tri + "vn 0 0 1\nf 1//1 2//1 3//3\n",
                 "f 1 2 3\no late\n" + tri,          # the model is exported at `o`, before its vertices were read
                 tri + "f /1 2 3\n"):
        p.write_text(text)
        assert both(p) is None, text
    # ... and what it accepts: texcoord / normal indices of a file without such records are not looked at, vt/vn index 0 is "absent",
    # a forward reference inside the LAST model is fine (it is exported at the end of the file)
    for text in (tri + "f 1/9 2/9 3//9\n", tri + "vt 0 0\nf 1/0 2/0 3/1\n", "f 1 2 3\n" + tri, tri + "f 1/ 2// 3///\n"):
        p.write_text(text)
        assert both(p).shape == (1, 3, 3), text
    with pytest.raises(mesh.ObjLoadError):
        load_c(tmp_path / "missing.obj")
    with pytest.raises(mesh.ObjLoadError):
        load_c(tmp_path)                              # a directory
    lib = _abi.lib()
    assert lib.rbrt_mesh_load_obj(None, _abi.Vec3C(), _abi.Vec3C(), 1.0, None, None) == _abi.E_INVALID
    lib.rbrt_mesh_free(None)


def correctly_rounded(tok):
    exact = Fraction(tok)
    with np.errstate(over="ignore"):
        f = np.float32(float(tok))
    if not np.isfinite(f):
        return f
    with np.errstate(over="ignore"):
        cands = [np.nextafter(f, np.float32(-np.inf)), f, np.nextafter(f, np.float32(np.inf))]
    cands = [c for c in cands if np.isfinite(c)]
    best = min(abs(Fraction(float(c)) - exact) for c in cands)
    ties = [c for c in cands if abs(Fraction(float(c)) - exact) == best]
    return ties[0] if len(ties) == 1 else [c for c in ties if not (c.view(np.uint32) & 1)][0]


def test_numbers_are_rounded_once(tmp_path):
    """Rust parses the decimal straight to f32.  Going through f64 (Python's float, C's atof) is wrong when the f64 lands on the midpoint of
    two f32: 16777217.0000000001 must give 16777218, not 16777216."""
    toks = ["16777217.0000000001", "16777217", "16777216.9999999999", "0.1", "-0.3", "1e-45", "7e-46", "1.17549435e-38", "3.4028235e38", "3.4028236e38",
            "1e39", "-1e39", "1e-60", "+5", "5.", ".5", "5.e1", "1E2", "1e+2", "inf", "-inf", "+Infinity", "nan", "0.50000002980232238769531250001",
            "0.5000000298023223876953125", "1.00000005960464477539062500000001", "33554434.0000001", "8388609.5", "8388610.5"]
    rng = np.random.default_rng(5)
    for _ in range(400):                                   # decimals hugging f32 midpoints from either side
        f = np.float32(rng.uniform(-1, 1) * 10.0 ** rng.integers(-20, 20))
        mid = Fraction(float(f)) + (Fraction(float(np.nextafter(f, np.float32(np.inf)))) - Fraction(float(f))) / 2
        eps = Fraction(1, 10 ** 60) * int(rng.integers(-1, 2))
        x = mid + eps
        digits = 70
        scaled = x * 10 ** digits
        assert scaled.denominator == 1 or True
        n = int(scaled) if scaled.denominator == 1 else None
        if n is None:
            continue
        s = ("-" if n < 0 else "") + f"{abs(n) // 10 ** digits}.{abs(n) % 10 ** digits:0{digits}d}"
        toks.append(s)
    p = tmp_path / "n.obj"
    p.write_text("".join(f"v {t} 0 0\n" for t in toks) + "".join(f"f {i + 1} {i + 1} {i + 1}\n" for i in range(len(toks))))
    got = load_c(p)[:, 0, 0]
    py = load_py(p)[:, 0, 0]
    for t, g, q in zip(toks, got, py):
        if "nan" in t.lower():
            assert np.isnan(g) and np.isnan(q)
            continue
        want = np.float32(float(t)) if "inf" in t.lower() else correctly_rounded(t)
        assert g.view(np.uint32) == want.view(np.uint32), (t, g, want)
        assert q.view(np.uint32) == want.view(np.uint32), (t, q, want)


good_number = st.one_of(st.floats(-1e6, 1e6, width=32).map(lambda x: repr(float(x))), st.integers(-10 ** 9, 10 ** 9).map(str),
                        st.sampled_from(["1e-3", "-2.5E+2", ".25", "7.", "+3", "16777217.0000000001", "1e50", "-1e-50"]))
bad_number = st.sampled_from(["inf", "bad", "0x1p3", "", "1e", "--1", "nan(1)"])
ws = st.sampled_from([" ", "  ", "\t", " \t "])


@st.composite
def obj_text(draw):
    """An .obj with every record kind the loader knows.  Clean files (4 of 5) only hold loadable records — faces of 0..5 corners, lines, relative
    indices, groups, materials — so that the soups are compared; the others mix in what the reference refuses, so that the refusals are."""
    dirty = draw(st.integers(0, 4)) == 0
    number = st.one_of(good_number, bad_number) if dirty else good_number
    lines = []
    n_v = n_t = n_n = 0
    for _ in range(draw(st.integers(0, 40))):
        kind = draw(st.sampled_from(["v"] * 6 + ["f"] * 8 + ["l", "o", "g", "usemtl", "mtllib", "vt", "vn", "#", "s", "", "junk"]))
        sep = draw(ws)
        if kind == "v":
            lines.append("v" + sep + sep.join(draw(number) for _ in range(draw(st.sampled_from([3, 3, 3, 3, 4, 6] + [2] * dirty)))))
            n_v += 1
        elif kind == "vt":
            lines.append("vt" + sep + sep.join(draw(number) for _ in range(draw(st.sampled_from([2, 2, 3] + [1] * dirty)))))
            n_t += 1
        elif kind == "vn":
            lines.append("vn" + sep + sep.join(draw(number) for _ in range(draw(st.sampled_from([3, 3, 4] + [2] * dirty)))))
            n_n += 1
        elif kind in ("f", "l"):
            if not n_v and not dirty:
                continue
            corners = []
            for _ in range(draw(st.sampled_from([3, 3, 3, 3, 4, 5, 2, 1, 0]))):
                def index(count):
                    if dirty:
                        hi = max(count, 1) + 1
                        return draw(st.one_of(st.integers(1, hi), st.integers(-hi, -1), st.sampled_from([0, 10 ** 19])))
                    return draw(st.one_of(st.integers(1, count), st.integers(-count, -1))) if count else draw(st.integers(-3, 3))
                forms = ["v", "v", "v", "v/t", "v//n", "v/t/n", "v/", "v//", "v/t/"] + ["/t", "v/t/n/x", "v.0"] * dirty
                form = draw(st.sampled_from(forms))
                corners.append(form.replace("v", str(index(n_v))).replace("t", str(index(n_t))).replace("n", str(index(n_n))))
            lines.append(kind + sep + sep.join(corners))
        elif kind in ("o", "g"):
            lines.append(kind + sep + draw(st.sampled_from(["name", "", "two words"])))
        elif kind == "usemtl":
            lines.append("usemtl" + sep + draw(st.sampled_from(["red", "blue", "none", "red"] + [""] * dirty)))
        elif kind == "mtllib":
            lines.append("mtllib" + sep + draw(st.sampled_from(["m.mtl", "bad.mtl", "nope.mtl"])))
        elif kind == "#":
            lines.append("# f 1 2 3")
        elif kind == "s":
            lines.append("s 1")
        elif kind == "junk":
            lines.append(draw(st.sampled_from(["vp 1 2 3", "#v 1 2 3", "fv 1 2 3", "\t", "vv"])))
        else:
            lines.append("")
    eol = draw(st.sampled_from(["\n", "\n", "\r\n"]))
    return eol.join(lines) + draw(st.sampled_from(["", eol]))


@settings(max_examples=300, deadline=None, derandomize=True, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
@given(text=obj_text(), scale=st.sampled_from([1.0, 0.5, 60.0]))
def test_generated_files_agree_with_the_python_restatement(tmp_path, text, scale):
    (tmp_path / "m.mtl").write_text("newmtl red\nKd 1 0 0\nnewmtl blue\n")
    (tmp_path / "bad.mtl").write_text("newmtl none\nNs x\n")
    p = tmp_path / "g.obj"
    p.write_bytes(text.encode())
    both(p, translation=(3.5, -1.8, -14.0), rotation=(0.3, -1.1, 2.0), scale=scale)


def test_a_large_file_in_many_pieces(tmp_path):
    """81 920 triangles written the way synth writes C2's stand-in (≈ 4 MB: really cut into pieces at the default size), plus relative
    indices that reach back across piece boundaries and a group in the middle."""
    from rbrt_b200 import synth
    p = tmp_path / "big.obj"
    synth.write_bunny_standin(str(p), 6)
    text = p.read_text().splitlines()
    n_v = sum(1 for l in text if l.startswith("v "))
    faces = [l for l in text if l.startswith("f ")]
    rel = []
    for i, l in enumerate(faces):                                # every third face as relative indices (all vertices precede the faces)
        if i % 3 == 0:
            a, b, c = (int(t) for t in l.split()[1:])
            l = f"f {a - n_v - 1} {b - n_v - 1} {c - n_v - 1}"
        rel.append(l)
    rel.insert(len(rel) // 2, "g second half")
    p.write_text("\n".join([l for l in text if not l.startswith("f ")] + rel) + "\n")
    assert p.stat().st_size > 3 << 20
    kw = dict(translation=(5.0, -1.8, -12.5), rotation=(0.1, 0.2, 0.3), scale=60.0)
    ref = load_py(p, **kw)
    assert ref.shape == (81920, 3, 3)
    for pb, th in ((None, None), (None, 3), (1 << 16, 16), (1 << 30, 1)):
        assert same_bits(load_c(p, piece_bytes=pb, threads=th, **kw), ref), (pb, th)
    # the transform entry on its own (threads above 65 536 vertices) agrees with the loader's
    raw = load_c(p)
    assert same_bits(mesh.transform_triangles(raw, **kw), ref)
