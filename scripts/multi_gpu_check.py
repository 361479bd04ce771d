"""Multi-GPU inside the C-ABI on N REAL GPUs, one process (rbrt_gpu_init_multi): for the NCCL and the PEER transport, the collective
render of C3 (tile shards) must be bit-identical to the one-GPU render, sample shards within f32 re-association; times one frame at a time
and the scene replication; then the compiled CLI: `rbrt --gpus N` must write the same PNG as `rbrt --gpus 1`.
    python scripts/multi_gpu_check.py N [workload]        ->  one JSON report on stdout"""
import ctypes as C, json, os, subprocess, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
_OUT = os.fdopen(os.dup(1), "w"); os.dup2(2, 1); sys.stdout = sys.stderr      # the JSON alone on stdout; host-mirror prints go to stderr
import numpy as np
import rbrt_b200 as R
from rbrt_b200 import _abi, synth
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
wl = sys.argv[2] if len(sys.argv) > 2 else "c3"
lib = _abi.lib()
desc, W, H, spp = bench.WORKLOADS[wl]
spheres, meshes, camkw = bench.build_workload(wl)
R.gpu_init(0)
cam = R.Camera.new(camkw["position"], camkw["look_at"], camkw["up"], H, W, camkw["focal_len_mm"])
report = {"n_gpus": n, "workload": desc}


def timed(scene, reps=5, **kw):
    ms = []
    for k in range(reps + 2):
        st = {}
        t0 = time.perf_counter()
        img = R.render_scene(cam, spp, scene, stats=st, seed=bench.SEED, **kw)
        if k >= 2:
            ms.append((time.perf_counter() - t0) * 1e3)
    return img, float(np.median(ms)), st


one = bench.make_scene(spheres, meshes)
ref, ms1, st1 = timed(one)
ref_hdr = R.render_scene_hdr(cam, 4, one, seed=3)
report["one_gpu"] = {"ms_per_frame_host_clock": ms1, "rays": st1["rays"], "ms_device": st1["ms_device"]}
one.close()
for name, tr in (("nccl", _abi.TRANSPORT_NCCL), ("peer", _abi.TRANSPORT_PEER)):
    devs = (C.c_int * n)(*range(n))
    rc = lib.rbrt_gpu_init_multi(devs, n, tr)
    if rc:
        report[name] = {"error": lib.rbrt_last_error().decode()}
        continue
    info = _abi.CommInfoC(); lib.rbrt_gpu_comm_info(info)
    t0 = time.perf_counter()
    sc = bench.make_scene(spheres, meshes)
    sc.handle()
    t_create = (time.perf_counter() - t0) * 1e3
    inf = sc.info()
    img, ms, st = timed(sc)
    os.environ["RBRT_DEBUG_MULTI"] = "1"
    R.render_scene(cam, spp, sc, stats={}, seed=bench.SEED)
    del os.environ["RBRT_DEBUG_MULTI"]
    hdr = R.render_scene_hdr(cam, 4, sc, seed=3)
    hs = R.render_scene_hdr(cam, 4, sc, seed=3, shard_mode=_abi.SHARD_SAMPLES)
    report[name] = {"comm": info.as_dict(), "scene_create_ms": t_create, "ms_upload_incl_replication": inf["ms_upload"], "ms_build": inf["ms_build"],
                    "ms_per_frame_host_clock": ms, "speedup_vs_one_gpu": ms1 / ms, "rays": st["rays"], "paths": st["paths"],
                    "u8_identical_to_one_gpu": bool(np.array_equal(img.pixels, ref.pixels)),
                    "hdr_identical_to_one_gpu": bool(np.array_equal(hdr.view(np.uint32), ref_hdr.view(np.uint32))),
                    "sample_shards_max_rel_err": float(np.max(np.abs(hs - ref_hdr) / np.maximum(np.abs(ref_hdr), 1e-6)))}
    sc.close()
    lib.rbrt_gpu_comm_destroy()

# the compiled host: rbrt --gpus N
cli = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "rbrt_b200", "rbrt")
with tempfile.TemporaryDirectory() as d:
    obj = os.path.join(d, "bunny.obj"); synth.write_bunny_standin(obj, 6)
    yml = os.path.join(d, "scene.yaml"); open(yml, "w").write(synth.blueprint_to_yaml(synth.example_scene_blueprint(obj)))
    outs = []
    for g in (1, n):
        png = os.path.join(d, f"out{g}.png")
        r = subprocess.run([cli, "-c", yml, "-t", png, "-w", "1024", "--height", "768", "-s", "50", "--seed", "7", "--gpus", str(g)], capture_output=True, text=True)
        outs.append((r.returncode, open(png, "rb").read() if os.path.exists(png) else b"", r.stderr.strip().splitlines()[-1:] ))
    report["cli"] = {"returncodes": [o[0] for o in outs], "png_identical": outs[0][1] == outs[1][1] and len(outs[0][1]) > 1000, "stderr_tail": [o[2] for o in outs]}
print(json.dumps(report), file=_OUT, flush=True)
