"""The drop-in boundary: rbrt_b200/librbrt_gpu.so loads and exports every function include/rbrt_gpu.h
declares, struct layouts of the ctypes mirror match the header, and without a GPU the entry points fail
loudly (no CPU fallback) instead of computing anything."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from rbrt_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rbrt_gpu.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rbrt_[a-z0-9_]+)\s*\(", src)))


def test_header_functions_are_exported():
    names = declared_functions()
    assert len(names) >= 14
    lib = C.CDLL(_abi.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/rbrt_gpu.h but not exported"
    assert set(names) == set(_abi.GPU_SIGNATURES), set(names) ^ set(_abi.GPU_SIGNATURES)


def test_struct_layouts_match_header(tmp_path):
    """Compile a C program against the header and compare sizeof/offsetof with the ctypes mirror."""
    fields = {"rbrt_camera": _abi.CameraC, "rbrt_material": _abi.MaterialC, "rbrt_sphere_desc": _abi.SphereDescC,
              "rbrt_mesh_desc": _abi.MeshDescC, "rbrt_hit": _abi.HitC, "rbrt_scene_opts": _abi.SceneOptsC,
              "rbrt_render_opts": _abi.RenderOptsC, "rbrt_stats": _abi.StatsC, "rbrt_scene_info": _abi.SceneInfoC,
              "rbrt_ray": _abi.RayC, "rbrt_vec3": _abi.Vec3C}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void){"]
    for cname, ct in fields.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in ct._fields_:
            lines.append(f'printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines.append("return 0;}")
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-o", str(exe), str(src)], check=True)     # the header is plain C
    got = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for cname, ct in fields.items():
        assert int(got[cname]) == C.sizeof(ct), cname
        for fname, _ in ct._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(ct, fname).offset, (cname, fname)
    assert np.dtype(_abi.HIT_DTYPE).itemsize == C.sizeof(_abi.HitC)


def test_version_and_argument_errors():
    lib = _abi.lib()
    assert b"sm_100a" in lib.rbrt_gpu_version()
    assert lib.rbrt_gpu_scene_create(None, 1, None, 0, None, C.byref(C.c_void_p())) == _abi.E_INVALID
    assert b"null" in lib.rbrt_last_error()
    bad = (_abi.SphereDescC * 1)(_abi.SphereDescC(_abi.Vec3C(0, 0, 0), 1.0, _abi.MaterialC(9, _abi.Vec3C(0, 0, 0), 0.0)))
    assert lib.rbrt_gpu_scene_create(bad, 1, None, 0, None, C.byref(C.c_void_p())) == _abi.E_INVALID
    assert lib.rbrt_gpu_scene_create(None, 0, None, 0, _abi.SceneOptsC(3, 0, 0.0, 0), C.byref(C.c_void_p())) == _abi.E_INVALID
    assert lib.rbrt_gpu_scene_destroy(None) == 0
    assert lib.rbrt_camera_new(_abi.Vec3C(), _abi.Vec3C(), _abi.Vec3C(), 1, 1, 1.0, None) == _abi.E_INVALID


def test_no_cpu_fallback_without_gpu():
    """In a container without a GPU every compute entry point must refuse: RBRT_E_NODEVICE / RBRT_E_CUDA."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the refusal path is exercised on the CPU-only build box")
    lib = _abi.lib()
    assert lib.rbrt_gpu_init(0) in (_abi.E_NODEVICE, _abi.E_CUDA)
    h = C.c_void_p()
    assert lib.rbrt_gpu_scene_create(None, 0, None, 0, None, C.byref(h)) in (_abi.E_NODEVICE, _abi.E_CUDA)
    assert not h.value
    import rbrt_b200 as R
    with pytest.raises(_abi.RbrtGpuError):
        R.render_scene(R.Camera.new((0, 0, 0), (0, 0, -1), (0, 1, 0), 4, 4, 28.0), 1, R.Scene())


def test_product_never_imports_oracle():
    """Nothing under rbrt_b200/ or include/ may reference oracle/ (the oracle is test infrastructure)."""
    for base in ("rbrt_b200", "include"):
        for d, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                    txt = open(os.path.join(d, f), errors="replace").read()
                    assert "oracle_ffi" not in txt and "rbrt_ref_" not in txt and "librbrt_oracle" not in txt, os.path.join(d, f)


def test_plain_c_host_compiles_and_fails_loudly_without_a_gpu(tmp_path):
    """integration/c/render_scene.c — main.rs:70-91 from C99 over the header, -Wall -Wextra -Werror -pedantic.  It loads a mesh through
    rbrt_mesh_load_obj (CPU) and then needs the GPU: on a machine without one it must stop with the library's message and exit code 3."""
    import torch
    from rbrt_b200 import synth
    src = os.path.join(ROOT, "integration", "c", "render_scene.c")
    exe = tmp_path / "render_scene"
    libdir = os.path.join(ROOT, "rbrt_b200")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), src, "-L", libdir, "-lrbrt_gpu",
                    f"-Wl,-rpath,{libdir}", "-o", str(exe)], check=True)
    obj = tmp_path / "m.obj"
    n = synth.write_bunny_standin(str(obj), subdiv=2)
    out = subprocess.run([str(exe), str(obj), str(tmp_path / "o.ppm")], capture_output=True, text=True)
    assert f"Successfully loaded {n} triangles" in out.stdout
    if torch.cuda.is_available():
        assert out.returncode == 0 and (tmp_path / "o.ppm").read_bytes().startswith(b"P6\n256 192\n255\n")
    else:
        assert out.returncode == 3 and "no CPU fallback" in out.stderr
    bad = subprocess.run([str(exe), str(tmp_path / "missing.obj"), str(tmp_path / "o.ppm")], capture_output=True, text=True)
    assert bad.returncode == 1 and "cannot open" in bad.stderr
