"""integration/rust/rbrt_gpu_sys (SOURCE ONLY: no rustc in the image) must stay in step with the C header: every #[repr(C)] mirror has the
field names and order of its ctypes twin in rbrt_b200/_abi.py (which tests/test_abi.py compares with the compiled header), and every
function the crate declares exists in the header with the same number of parameters."""
import os
import re

from rbrt_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RS = open(os.path.join(ROOT, "integration", "rust", "rbrt_gpu_sys", "src", "lib.rs")).read()
HDR = open(os.path.join(ROOT, "include", "rbrt_gpu.h")).read()

PAIRS = {"RbrtVec3": _abi.Vec3C, "RbrtRay": _abi.RayC, "RbrtCamera": _abi.CameraC, "RbrtMaterial": _abi.MaterialC, "RbrtSphereDesc": _abi.SphereDescC,
         "RbrtTriangleDesc": _abi.TriangleDescC, "RbrtElementRef": _abi.ElementRefC, "RbrtMeshDesc": _abi.MeshDescC, "RbrtRenderOpts": _abi.RenderOptsC,
         "RbrtStats": _abi.StatsC, "RbrtSceneOpts": _abi.SceneOptsC, "RbrtCommInfo": _abi.CommInfoC}


def rust_fields(name):
    m = re.search(r"pub struct %s\s*\{(.*?)\}" % name, RS, re.S)
    assert m, f"struct {name} missing from the Rust crate"
    return re.findall(r"pub (\w+)\s*:", m.group(1))


def test_struct_fields_match_the_ctypes_mirror():
    for name, cls in PAIRS.items():
        assert rust_fields(name) == [f for f, _ in cls._fields_], name


def test_every_rust_function_is_in_the_header_with_the_same_arity():
    block = re.search(r'extern "C" \{(.*?)\n\}', RS, re.S).group(1)
    fns = re.findall(r"pub fn (\w+)\s*\((.*?)\)\s*(?:->|;)", block, re.S)
    assert len(fns) >= 20
    for name, args in fns:
        m = re.search(r"\b%s\s*\(([^()]*?)\)\s*;" % name, HDR, re.S)
        assert m, f"{name} is not declared in include/rbrt_gpu.h"
        n_rs = len([a for a in args.split(",") if a.strip()])
        c_args = m.group(1).strip()
        n_c = 0 if c_args in ("", "void") else len([a for a in re.sub(r"/\*.*?\*/", "", c_args, flags=re.S).split(",") if a.strip()])
        assert n_rs == n_c, f"{name}: {n_rs} parameters in Rust, {n_c} in the header"
    for const, val in re.findall(r"pub const (RBRT_\w+): \w+ = (\d+);", RS):
        m = re.search(r"\b%s\s*=?\s*(\d+)" % const, HDR)
        assert m and int(m.group(1)) == int(val), f"{const} differs from the header"
