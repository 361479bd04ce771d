"""The committed measurement evidence is self-consistent (CPU-only): profiles/r2_ncu_summary.json is what scripts/ncu_summary.py
makes of the committed raw ncu launch lists, and the committed bench lines of the final code quote exactly those figures in their
`roofline` record (bench.py's ncu_evidence) and satisfy the bench contract's keys."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROF = os.path.join(ROOT, "profiles")


def bench_line(name):
    with open(os.path.join(PROF, name)) as f:
        lines = [ln for ln in f.read().splitlines() if ln.startswith("{")]
    assert len(lines) == 1, f"{name}: one JSON line expected"
    return json.loads(lines[0])


def test_ncu_summary_is_reproducible_from_the_raw_launch_lists():
    csvs = [os.path.join("profiles", f"r2_ncu_metrics_c3_n{n}.csv") for n in (1, 2, 4, 8)]
    out = subprocess.run([sys.executable, os.path.join("scripts", "ncu_summary.py")] + csvs, cwd=ROOT, capture_output=True, text=True, check=True)
    made = json.loads(out.stdout)
    with open(os.path.join(PROF, "r2_ncu_summary.json")) as f:
        kept = json.load(f)
    assert made == kept
    for n in ("1", "2", "4", "8"):
        e = kept["c3"][n]
        assert e["k_trace_launches"] > 0 and 0.5 < e["k_trace_share_of_frame"] < 0.75      # the kernel the roofline record is about dominates the frame
        assert abs(e["dram_bytes_per_step"] - 1e9 * (e["dram_read_gb"] + e["dram_write_gb"])) < 2e6
        assert abs(sum(k["share"] for k in e["per_kernel"].values()) - 1.0) < 1e-3


@pytest.mark.parametrize("n", [1, 8])
def test_final_bench_lines_quote_the_committed_ncu_figures(n):
    b = bench_line(f"r2_bench_c3_n{n}.json")
    with open(os.path.join(PROF, "r2_ncu_summary.json")) as f:
        ev = json.load(f)["c3"][str(n)]
    assert b["metric"] == "Mrays/s" and b["n_gpus"] == n and b["scaling"] == "strong" and b["higher_is_better"] is True
    assert b["steps"] == 20 and b["warmup"] == 5 and b["gpu_launches"] > 0 and b["vs_baseline"] is None
    r = b["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["traffic"] == ev["dram_bytes_per_step"]
    assert abs(r["dram_frac"] - r["traffic"] / (r["trace_ms_per_step"] * 1e-3) / (r["peak"] * 1e9)) < 1e-6
    assert r["dram_frac"] < r["frac"]                                   # requested bytes are mostly served by L1/L2
    # whole-job value and ms/step agree: rays per frame x frames per second
    assert abs(b["value"] - b["rays_per_step"] / b["ms_per_step"] / 1e3) / b["value"] < 1e-6
    assert abs(b["samples_per_s"] - 1920 * 1080 * 64 / (b["ms_per_step"] * 1e-3)) / b["samples_per_s"] < 1e-6
    e = b["e2e"]
    assert e["h2d_bytes_per_step"] > 40e6 and e["d2h_bytes_per_step"] == 1920 * 1080 * 3 and e["value"] < b["value"]
    assert b["image_check"]["bit_identical"] is True and b["image_check"]["n_gpus"] == n
    c = b["clocks"]
    assert c["sm_mhz"] >= 0.95 * c["sm_max_mhz"] and not any("slowdown" in x for x in c["reasons"])
    if n == 1:
        cb = b["cpu_baseline"]
        assert cb["kind"] == "port" and cb["cores"] >= 1 and 0 < cb["value"] < 1.0


def test_reference_arm_line_did_not_load_the_product_library():
    r = bench_line("r2_bench_c3_reference_arm.json")
    assert r["impl"] == "reference" and r["product_lib_loaded"] is False and r["metric"] == "Mrays/s"
    assert r["e2e"]["h2d_bytes_per_step"] == 0 and r["e2e"]["d2h_bytes_per_step"] == 0 and r["e2e"]["value"] == r["value"]
    assert r["cpu_baseline"]["value"] == r["value"]


def test_device_code_is_the_gpu_verified_build():
    """Commits after the round's last GPU run touched host code, plus ONE kernel (k_mesh_setup: a floor under the leaf-box padding, see
    profiles/r2_device_code.md): every other kernel of the library in the tree has the SASS of the build that passed the GPU tests and produced
    the committed bench lines."""
    import hashlib
    import re
    import shutil
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    lib = os.path.join(ROOT, "rbrt_b200", "librbrt_gpu.so")
    if not os.path.exists(cuobjdump) or not os.path.exists(lib):
        pytest.skip("cuobjdump or the built library is not here")
    want = {}
    with open(os.path.join(PROF, "r2_device_code_kernels.txt")) as f:
        for ln in f:
            if not ln.startswith("#"):
                h, name = ln.split()
                want[name] = h
    sass = subprocess.run([cuobjdump, "-sass", lib], capture_output=True, text=True, check=True).stdout
    got, name, buf = {}, None, []
    for ln in sass.splitlines(True):
        m = re.match(r"\s+Function : (\S+)", ln)
        if m:
            if name:
                got[name] = hashlib.md5("".join(buf).encode()).hexdigest()
            name, buf = m.group(1), []
        elif name:
            buf.append(ln)
    got[name] = hashlib.md5("".join(buf).encode()).hexdigest()
    assert set(got) == set(want) and len(want) == 51
    changed = sorted(k for k in want if got[k] != want[k])
    assert changed == ["_ZN4rbrt12k_mesh_setupEPKiNS_7MeshDevEfPS2_PNS_11BuildParamsE"], (
        "device code differs from the GPU-verified build in more than k_mesh_setup: re-run pytest -m gpu and the bench, then regenerate "
        "profiles/r2_device_code_kernels.txt")
