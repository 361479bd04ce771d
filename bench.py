#!/usr/bin/env python
"""bench.py — the headline benchmark of BASELINE.json: Mrays/s (and samples/s) of the path-tracing hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c2|c1|c4] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" is ONE render of the whole frame of the workload (all pixels x all samples, up to 51 ray segments per
path) — the reference's `render_scene(cam, spp, scene)` (lib.rs:75-124).  Rays = closest-hit queries
(`Scene::hit` calls, primary + bounces), counted on the device.

  value      whole-job Mrays/s with the scene (SoA buffers + LBVH) already resident in HBM; the timed region is
             K x (render -> [NCCL reduce to rank 0] -> finalize (1/spp, sqrt, x256, saturating u8) on the device), with
             --frames-in-flight frames (default 2) in flight on their own streams (rbrt_b200.FramePipeline): the sparse,
             latency-bound last bounces of one frame overlap the dense first bounces of the next; on 4 / 8 GPUs 2 / 4
             consecutive frames are also rendered in the same wavefront batches (--frames-per-batch), which gives a rank's
             launches the size they have on fewer GPUs.  `single_frame` is the same measurement one frame at a time.  The
             last pipelined image is checked against the single-frame one.
  e2e        the same metric through the reference-facing call with HOST buffers: every step uploads the
             triangle soup from pinned host memory (rbrt_gpu_scene_create: H2D + LBVH build), renders
             (rbrt_gpu_render / the multi-rank building blocks) and copies the RGB8 image back to the host.
  roofline   the trace kernel (BVH traversal + intersection tests): algorithmic bytes per step from an
             instrumented counting pass (64 B per node visit, 48 B per triangle test, 64 B of queue / path-record traffic per
             traversed ray) / the summed CUDA-event durations of that kernel's
             launches inside the timed region, against the measured HBM copy bandwidth (MEASURED_PEAKS.json).
  cpu_baseline / --impl reference
             the CPU oracle (C++/AVX restatement of the reference; the Rust reference cannot be built here) on all
             host threads, on a bounded stratified pixel sample of the same workload.

Multi-GPU: one process per GPU, the image sharded by interleaved 8x4-pixel tiles (fixed total work => "strong"
scaling), scene + LBVH replicated, ONE NCCL reduce of the f32 accumulation buffer to rank 0 per step.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (description, width, height, spp)
    "c1": ("C1: spheres-only scene of scenes/example_scene.yaml (mesh removed), 256x192, 8 spp", 256, 192, 8),
    "c2": ("C2: scenes/example_scene.yaml with an 81920-triangle displaced icosphere (.obj) for bunny.obj, 1024x768, 50 spp", 1024, 768, 50),
    "c3": ("C3: 1310720-triangle displaced icosphere (radius 40) + 4 spheres, 1920x1080, 64 spp, tile-sharded", 1920, 1080, 64),
    "c4": ("C4: dielectric/metal-heavy scene (36 glass/metal spheres + 20480-triangle glass mesh), depth 50, 1024x768, 256 spp", 1024, 768, 256),
    "c5": ("C5: 5242880-triangle displaced icosphere (radius 100) + 4 spheres, 3840x2160, 1024 spp, sample-range sharded + NCCL reduce", 3840, 2160, 1024),
}
SAMPLE_SHARDED = {"c5"}
SEED = 0x5EED
_STDOUT = sys.stdout


def build_workload(name, pinned=False):
    """Returns (spheres [(center, radius, material)], [(triangles [N,3,3] f32, material)], camera kwargs)."""
    import rbrt_b200 as R
    from rbrt_b200 import synth
    from rbrt_b200.vec3 import Vec3

    _, W, H, spp = WORKLOADS[name]
    excam = synth.EXAMPLE_CAMERA
    cam_ex = dict(position=Vec3(*excam["camera_position"]), look_at=Vec3(*excam["camera_look_at"]), up=Vec3(*excam["camera_up"]),
                  focal_len_mm=excam["camera_focal_length_mm"])
    if name == "c1":
        bp = synth.spheres_only_blueprint()
        sc = R.create_scene_from_scene_blueprint(bp)
        return [(s.center, s.radius, s.material) for s in sc.elements], [], cam_ex
    if name == "c2":
        obj = os.path.join(synth.cache_dir(), "standin6.obj")
        if not os.path.exists(obj):
            synth.write_bunny_standin(obj, 6)
        sc = R.create_scene_from_scene_blueprint(synth.example_scene_blueprint(obj))
        return ([(s.center, s.radius, s.material) for s in sc.elements],
                [(m.triangles, m.material) for m in sc.triangle_meshes], cam_ex)
    if name == "c3":
        cam, spheres, tris, mat = synth.big_mesh_config(8, 40.0)
        return spheres, [(tris, mat)], cam
    if name == "c5":
        cam, spheres, tris, mat = synth.big_mesh_config(9, 100.0)
        return spheres, [(tris, mat)], cam
    if name == "c4":
        spheres, tris, mat = synth.stress_config()
        return spheres, [(tris, mat)], cam_ex
    raise SystemExit(f"unknown workload {name}")


def make_scene(spheres, meshes, pinned_cache=None):
    import rbrt_b200 as R
    sc = R.Scene()
    sc.elements += [R.Sphere(c, r, m) for c, r, m in spheres]
    for i, (tris, mat) in enumerate(meshes):
        if pinned_cache is not None:
            tris = pinned_cache[i]
        sc.triangle_meshes.append(R.TriangleMesh.from_triangles(tris, mat))
    return sc


def pin_meshes(meshes):
    """Copies of the triangle arrays in page-locked host memory (the e2e leg uploads from these)."""
    import torch
    out, keep = [], []
    for tris, _ in meshes:
        t = torch.empty(tris.shape, dtype=torch.float32, pin_memory=True)
        t.numpy()[...] = tris
        keep.append(t)
        out.append(t.numpy())
    return out, keep


class ClockSampler:
    """nvidia-smi clocks / throttle reasons of one GPU (B200_PROFILING.md).  nvidia-smi needs a few hundred ms before its first
    sample, longer than a whole timed region of this benchmark, so the sampler is started BEFORE the warm-up steps and the
    samples are filtered by their timestamps: those inside [mark_begin, mark_end] (the timed region, padded by one sampling
    period) are used; if the region was too short to catch any, the samples of the warm-up steps (same load) are used and
    `window` says so."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    PERIOD_MS = 20

    def __init__(self, index):
        self.index, self.proc, self.lines, self.t_begin, self.t_end, self.t_start = index, None, [], None, None, time.time()

    def __enter__(self):
        if os.environ.get("RBRT_BENCH_NO_CLOCKS"):            # debugging aid: measure without the nvidia-smi side process
            return self
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", str(self.PERIOD_MS)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def mark_begin(self):
        self.t_begin = time.time()

    def mark_end(self):
        self.t_end = time.time()

    def __exit__(self, *a):
        if self.proc:
            time.sleep(2.5 * self.PERIOD_MS / 1e3)
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        import datetime
        rows = []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                vals = (float(f[1]), float(f[2]), float(f[3]))
            except ValueError:
                continue
            ts = None
            for fmt in ("%Y/%m/%d %H:%M:%S.%f", "%Y-%m-%d %H:%M:%S.%f", "%Y/%m/%d %H:%M:%S"):
                try:
                    ts = datetime.datetime.strptime(f[0], fmt).timestamp()
                    break
                except ValueError:
                    pass
            rows.append((ts if ts is not None else (self.t_begin or 0.0), *vals, f[4:8]))
        pad = self.PERIOD_MS / 1e3
        inside = [r for r in rows if self.t_begin is not None and self.t_begin - pad <= r[0] <= (self.t_end or time.time()) + pad]
        window = "timed region"
        if not inside:
            inside = [r for r in rows if r[0] <= (self.t_end or time.time()) + pad]
            window = "warm-up + timed region (timed region shorter than the sampling period)"
        if not inside:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        reasons = set()
        for r in inside:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(r[1] for r in inside), "sm_max_mhz": max(r[2] for r in inside), "reasons": sorted(reasons),
                "samples": len(inside), "power_w_max": max(r[3] for r in inside), "window": window}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def algorithmic_bytes(st, n_spheres, n_meshes):
    """Trace kernel (stage B: LBVH traversal of the rays that entered a mesh AABB), DESIGN.md section 4:
    bytes/step = 64 V + 48 T + 64 C   (V node visits x 64-B node, T triangle tests x 48-B record, C traversed rays x
    64 B of queue / path-record traffic: 4 B queue index + 40 B of the path record (16 B sphere pre-result + 24 B ray) read,
    16 B hit record + 4 B material-queue index written).  The sphere and mesh-AABB tests of SURVEY.md section 8(d)
    (16 S + 24 M per ray) run in the producing kernels (k_generate / k_shade) and are not charged to this kernel."""
    return 64 * st["node_visits"] + 48 * st["tri_tests"] + 64 * st["traversed_rays"]      # counts of k_trace only (tail kernel subtracted by the caller)


def algorithmic_flops(st, n_spheres, n_meshes):
    """SURVEY.md section 8(d) restated for the 4-wide node: four slab tests of 12 flops (6 FMA) + 12 min/max each = 96 per node
    visit; 46 per Moeller-Trumbore test (triangle.rs:189-234)."""
    return 96 * st["node_visits"] + 46 * st["tri_tests"]


# ------------------------------------------------------------------------------------------- CPU legs
def oracle_scene(spheres, meshes):
    from oracle import oracle_ffi as O
    import rbrt_b200 as R
    els = [R.Sphere(c, r, m) for c, r, m in spheres]
    ms = [R.TriangleMesh.from_triangles(t, m) for t, m in meshes]
    return O, O.OracleScene(els, ms, 8)


def choose_stride(O, osc, cam_c, spp_full, target_s):
    """Pick a pixel lattice (stride, stride) and a sample count so that one oracle pass takes about target_s."""
    from rbrt_b200 import _abi
    W, H = cam_c.img_width_pix, cam_c.img_height_pix
    stride = max(1, int(max(W, H) // 24))
    st, _ = O.render_subset(osc, cam_c, 1, stride, stride, _abi.RenderOptsC(seed=SEED))
    per_path = max(st["ms_total"], 1e-3) / 1e3 / max(st["paths"], 1)
    paths = max(1.0, target_s / per_path)
    full_px = W * H
    if paths >= full_px * spp_full:
        return 1, spp_full
    if paths >= full_px:
        return 1, max(1, int(paths // full_px))
    s = int(np.ceil(np.sqrt(full_px / paths)))
    return max(1, s), 1


def cpu_leg(O, osc, cam_c, stride, spp, want_image=False):
    from rbrt_b200 import _abi
    st, acc = O.render_subset(osc, cam_c, spp, stride, stride, _abi.RenderOptsC(seed=SEED), want_image=want_image)
    return (st, acc) if want_image else st


def run_reference(args):
    """--impl reference: the oracle on all host threads, each step a bounded stratified sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import rbrt_b200 as R
    desc, W, H, spp = WORKLOADS[args.workload]
    spheres, meshes, camkw = build_workload(args.workload)
    O, osc = oracle_scene(spheres, meshes)
    cam = R.Camera.new(camkw["position"], camkw["look_at"], camkw["up"], H, W, camkw["focal_len_mm"])
    cores = O.lib().rbrt_ref_hardware_threads()
    budget = 150.0 / max(1, args.steps + args.warmup)
    stride, s_spp = choose_stride(O, osc, cam.to_c(), spp, min(8.0, budget))
    for _ in range(args.warmup):
        cpu_leg(O, osc, cam.to_c(), stride, s_spp)
    t0 = time.perf_counter()
    rays = paths = 0
    for _ in range(args.steps):
        st = cpu_leg(O, osc, cam.to_c(), stride, s_spp)
        rays += st["rays"]; paths += st["paths"]
    dt = time.perf_counter() - t0
    val = rays / dt / 1e6
    sample = f"pixel lattice stride {stride}x{stride} of the {W}x{H} frame, {s_spp} of {spp} spp: {paths // max(1, args.steps)} paths/step"
    line = {"impl": "reference", "metric": "Mrays/s", "value": val, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "samples_per_s": paths / dt,
            "config": workload_config(args.workload, spheres, meshes, args.gpus),
            "cpu_baseline": {"value": val, "unit": "Mrays/s", "cores": int(cores), "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "CPU oracle (C++/AVX restatement; rustc/cargo absent so the Rust reference cannot be built); full-frame time is "
                    "extrapolated only in DESIGN.md, this value is measured rays / measured seconds on the sample"}
    print(json.dumps(line), file=_STDOUT, flush=True)
    return 0


def workload_config(name, spheres, meshes, n_gpus):
    desc, W, H, spp = WORKLOADS[name]
    return {"workload": desc, "width": W, "height": H, "spp": spp, "max_depth": 50, "spheres": len(spheres),
            "triangles": int(sum(len(t) for t, _ in meshes)), "seed": SEED,
            "sharding": "none" if n_gpus == 1 else (f"sample ranges over {n_gpus} ranks + one NCCL reduce (sum) of the f32 accumulator" if name in SAMPLE_SHARDED
                                                    else f"interleaved 8x4-pixel tiles over {n_gpus} ranks + one NCCL reduce of the f32 accumulator"),
            "l2_policy": "no explicit flush: per step the wavefront streams >400 MB of ray/hit queues and the scene (nodes+triangles+normals) "
                         "is larger than or comparable to L2; inputs larger than L2"}


# ------------------------------------------------------------------------------------------- GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist

    import rbrt_b200 as R
    from rbrt_b200 import _abi
    from rbrt_b200.render import make_opts

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch multi-GPU runs with torch.distributed.run --nproc-per-node N (one process per GPU)")
        args.gpus = world
    torch.cuda.set_device(local)
    R.gpu_init(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _abi.lib()
    desc, W, H, spp = WORKLOADS[args.workload]
    spheres, meshes, camkw = build_workload(args.workload)
    cam = R.Camera.new(camkw["position"], camkw["look_at"], camkw["up"], H, W, camkw["focal_len_mm"])
    cam_c = cam.to_c()
    scene = make_scene(spheres, meshes)
    info = scene.info()
    stream = torch.cuda.current_stream()
    accum = torch.empty(H * W * 4, dtype=torch.float32, device="cuda")
    rgb = torch.empty(H * W * 3, dtype=torch.uint8, device="cuda")
    shard_mode = _abi.SHARD_SAMPLES if args.workload in SAMPLE_SHARDED else _abi.SHARD_TILES
    shard = dict(shard_mode=shard_mode, shard_rank=rank, shard_count=world) if world > 1 else {}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident(handle, flags_kw, st):
        _abi.check(lib.rbrt_gpu_render_accum_device(handle, cam_c, spp, make_opts(seed=SEED, **shard, **flags_kw), accum.data_ptr(),
                                                    stream.cuda_stream, st))
        if world > 1:
            dist.reduce(accum, dst=0, op=dist.ReduceOp.SUM)
        if rank == 0:
            _abi.check(lib.rbrt_gpu_finalize_device(accum.data_ptr(), W, H, spp, rgb.data_ptr(), None, stream.cuda_stream))

    # ---- counting pass (untimed): node visits / triangle tests per step, identical every step (fixed seed)
    cst = _abi.StatsC()
    step_resident(scene.handle(), dict(count_visits=True), cst)
    barrier()
    counts = cst.as_dict()

    # ---- single-frame pass: one frame at a time (host waits for each), every trace launch bracketed by CUDA events.
    #      Gives the per-frame counters (identical every frame: fixed seed), the un-overlapped kernel times the roofline
    #      is computed from, and the reference image the pipelined frames are checked against.
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stats = []
    for _ in range(max(args.warmup, 3)):
        step_resident(scene.handle(), dict(time_kernels=True), _abi.StatsC())
    barrier()
    n_single = 3
    e0.record(stream)
    for _ in range(n_single):
        st = _abi.StatsC()
        step_resident(scene.handle(), dict(time_kernels=True), st)
        stats.append(st.as_dict())
    e1.record(stream)
    barrier()
    ms_single = e0.elapsed_time(e1) / n_single
    rgb_ref = rgb.clone() if rank == 0 else None
    frame = stats[-1]
    assert all(s_["rays"] == frame["rays"] and s_["paths"] == frame["paths"] for s_ in stats), "frames of one seed differ"

    # ---- resident arm (timed region): `steps` frames through the FramePipeline, `frames_in_flight` of them in flight on
    #      their own streams and wavefront pools, so the sparse last bounces of one frame overlap the next frame's first
    fpb = args.frames_per_batch or (1 if world == 1 else 2 if world < 8 else 4)     # measured: 2 GPUs 17.10 -> 16.55 ms/frame, 4 GPUs 9.22 -> 8.60 with 2;
                                                                                    # 1/8 shard on one GPU 5.15 -> 4.29 with 4 (profiles/r1_summary.md)
    if args.workload in SAMPLE_SHARDED:
        fpb = args.frames_per_batch or 1
    pipe = R.FramePipeline(W, H, depth=args.frames_in_flight, host_output=False, shard_mode=shard_mode, frames_per_batch=fpb)
    with ClockSampler(local) as clk:
        time.sleep(0.3)                                       # let nvidia-smi deliver its first samples
        for _ in range(max(args.warmup, 3) * fpb):
            pipe.submit(cam_c, spp, scene, seed=SEED)
        pipe.drain()
        barrier()
        clk.mark_begin()
        e0.record(stream)
        for _ in range(args.steps):
            pipe.submit(cam_c, spp, scene, seed=SEED)
        pipe.flush()                                          # a last, partially filled group of frames
        pipe.wait_on(stream)
        e1.record(stream)
        barrier()
        clk.mark_end()
    last = pipe.drain()[-1][0]
    if rank == 0:
        assert torch.equal(last, rgb_ref), "pipelined frame differs from the single-frame render"
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms, ms_single], dtype=torch.float64, device="cuda")
    agg = torch.tensor([frame["rays"] * args.steps, frame["paths"] * args.steps, (frame["launches"] + (1 if rank == 0 else 0)) * args.steps,
                        counts["node_visits"] - counts["tail_node_visits"], counts["tri_tests"] - counts["tail_tri_tests"], counts["rays"],
                        counts["traversed_rays"] - counts["tail_traversed_rays"]], dtype=torch.float64, device="cuda")
    trace_ms = torch.tensor([sum(s_["ms_trace"] for s_ in stats) / n_single * args.steps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(agg, op=dist.ReduceOp.SUM)
        dist.all_reduce(trace_ms, op=dist.ReduceOp.MAX)
    ms, ms_single = (float(x) for x in t.tolist())
    rays, paths, launches, V, T, Rc, Cc = (float(x) for x in agg.tolist())
    trace_ms = float(trace_ms.item())
    clocks = clk.summary()

    # ---- e2e arm: host buffers in, host image out, every step (scene upload + LBVH build inside the timed region)
    pinned, keep = pin_meshes(meshes)
    h2d = sum(t.nbytes for t in pinned) + 36 * len(spheres) + 20 * (len(spheres) + len(meshes)) + 64
    d2h = H * W * 3

    pipe_e = R.FramePipeline(W, H, depth=args.frames_in_flight, host_output=True, shard_mode=shard_mode)
    e2e_last = [None]

    def retire(fin):
        if fin is not None:
            img, old_scene = fin
            old_scene.close()
            e2e_last[0] = img

    def step_e2e():
        t_a = time.perf_counter()
        sc = make_scene(spheres, meshes, pinned)
        sc.handle()                                           # rbrt_gpu_scene_create: H2D of the triangle soup + LBVH build
        t_b = time.perf_counter()
        # render -> [reduce] -> finalize -> RGB8 image to pinned host memory, enqueued on the frame's stream; the image of
        # the frame submitted `frames_in_flight` steps ago is collected (host buffer ready) and its scene destroyed
        for fin in pipe_e.submit(cam_c, spp, sc, tag=sc, seed=SEED):
            retire(fin)
        t_c = time.perf_counter()
        return (t_b - t_a) * 1e3, (t_c - t_b) * 1e3

    for _ in range(max(2, args.frames_in_flight + 1)):
        step_e2e()
    for fin in pipe_e.drain():
        retire(fin)
    barrier()
    e_steps = max(2, min(args.steps, 5))
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(e_steps):
        ms_create, ms_render = step_e2e()
        print(f"[e2e rank {rank}] scene_create {ms_create:.1f} ms, submit (+ wait for the frame {args.frames_in_flight} steps back) {ms_render:.1f} ms", file=sys.stderr)
    for fin in pipe_e.drain():                                # every image of the timed steps is on the host when the clock stops
        retire(fin)
    e1.record(stream)
    barrier()
    e_ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)   # host-side work (malloc, sync copies) counts too
    if rank == 0:
        assert np.array_equal(e2e_last[0].pixels.reshape(-1), rgb_ref.cpu().numpy()), "e2e image differs from the single-frame render"
    e_rays = frame["rays"] * e_steps
    te = torch.tensor([e_ms], dtype=torch.float64, device="cuda")
    re = torch.tensor([float(e_rays)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dist.all_reduce(re, op=dist.ReduceOp.SUM)
    e_val = float(re.item()) / (float(te.item()) / 1e3) / 1e6

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peak, peak_src = peaks()
    ns, nm = len(spheres), len(meshes)
    cstep = {"rays": Rc, "node_visits": V, "tri_tests": T, "traversed_rays": Cc}
    bytes_step = algorithmic_bytes(cstep, ns, nm)
    flops_step = algorithmic_flops(cstep, ns, nm)
    trace_ms_step = trace_ms / args.steps          # max over ranks of the per-rank sum; ranks run concurrently
    achieved = bytes_step / max(world, 1) / (trace_ms_step / 1e3) / 1e9 if trace_ms_step > 0 else None
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get(args.workload)
    line = {
        "metric": "Mrays/s", "value": rays / (ms / 1e3) / 1e6, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(workload_config(args.workload, spheres, meshes, world), frames_in_flight=args.frames_in_flight, frames_per_batch=fpb),
        "single_frame": {"ms_per_step": ms_single, "value": rays / args.steps / (ms_single / 1e3) / 1e6, "unit": "Mrays/s",
                         "note": "one frame at a time, host waits for each (latency of a lone render_scene call); `value` keeps "
                                 "`frames_in_flight` groups of `frames_per_batch` frames in flight on separate streams"},
        "samples_per_s": paths / (ms / 1e3), "rays_per_step": rays / args.steps, "rays_per_sample": rays / max(paths, 1),
        "clocks": clocks,
        "e2e": {"value": e_val, "unit": "Mrays/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": e_steps,
                "ms_per_step": float(te.item()) / e_steps, "includes": "scene upload from pinned host memory + LBVH build + render + RGB8 image to pinned host memory, per step; "
                            "frames_in_flight frames overlap (FramePipeline), all images on the host when the clock stops"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "k_trace (stage B: persistent LBVH traversal with dynamic fetch + Moeller-Trumbore tests)",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                     "peak_source": peak_src, "algorithmic_bytes_per_traversed_ray": bytes_step / max(Cc, 1),
                     "traversed_rays_per_step": Cc, "traversed_share_of_rays": Cc / max(Rc, 1),
                     "node_visits_per_traversed_ray": V / max(Cc, 1), "tri_tests_per_traversed_ray": T / max(Cc, 1),
                     "trace_ms_per_step": trace_ms_step, "trace_share_of_step": trace_ms_step / ms_single,
                     "fp32_achieved_tflops": flops_step / max(world, 1) / (trace_ms_step / 1e3) / 1e12 if trace_ms_step > 0 else None,
                     "fp32_peak_tflops": 37.2, "fp32_frac": (flops_step / max(world, 1) / (trace_ms_step / 1e3) / 1e12 / 37.2) if trace_ms_step > 0 else None,
                     "fp32_note": "secondary bound of SURVEY.md 8(d): 148 SMs x 128 lanes x 1.965 GHz = 37.2 T lane-ops/s without FMA contraction "
                                  "(parity forbids it); the bandwidth fraction above is the larger one and is reported as `frac`",
                     "note": "achieved = per-GPU algorithmic bytes of all trace launches of a step / their summed CUDA-event time (per-launch "
                             "average x launches), events taken in this run's single-frame pass (frames of the timed region overlap, which "
                             "would smear per-kernel times); nodes+triangles mostly hit in L2, so this is requested bandwidth against the HBM copy peak"},
        "scene": {"bvh_nodes": info["num_bvh_nodes"], "device_bytes": info["device_bytes"], "ms_upload": info["ms_upload"], "ms_build": info["ms_build"]},
    }
    if world == 1 and not args.no_cpu:
        O, osc = oracle_scene(spheres, meshes)
        cores = int(O.lib().rbrt_ref_hardware_threads())
        stride, s_spp = choose_stride(O, osc, cam_c, spp, 22.0)
        st, acc_cpu = cpu_leg(O, osc, cam_c, stride, s_spp, want_image=True)
        if st["ms_total"] < 10e3 and stride > 1 and s_spp == 1:       # the probe under-estimated the cost per path: one denser pass (~18 s)
            stride2 = max(1, int(stride * (st["ms_total"] / 18e3) ** 0.5))
            if stride2 < stride:
                stride = stride2
                st, acc_cpu = cpu_leg(O, osc, cam_c, stride, s_spp, want_image=True)
        # image check (BASELINE.json's metric names the image RMSE): the pixels the oracle just rendered against the GPU's
        # render of the same frame at the same seed and sample count — same Philox streams, so the sums must agree bit for bit
        _abi.check(lib.rbrt_gpu_render_accum_device(scene.handle(), cam_c, s_spp, make_opts(seed=SEED), accum.data_ptr(), stream.cuda_stream, _abi.StatsC()))
        acc_gpu = accum.cpu().numpy().reshape(H, W, 4)[::stride, ::stride, :3]
        acc_ref = np.asarray(acc_cpu, np.float32).reshape(H, W, 4)[::stride, ::stride, :3]
        hdr_gpu, hdr_ref = acc_gpu * np.float32(1.0 / s_spp), acc_ref * np.float32(1.0 / s_spp)
        to_u8 = lambda h_: np.clip(np.nan_to_num(np.sqrt(np.maximum(h_, 0)) * 256.0), 0, 255).astype(np.uint8).astype(np.float64)
        line["image_check"] = {"pixels": int(acc_ref.shape[0] * acc_ref.shape[1]), "spp": int(s_spp),
                               "bit_identical": bool(np.array_equal(acc_gpu.view(np.uint32), acc_ref.view(np.uint32))),
                               "hdr_rmse": float(np.sqrt(np.mean((hdr_gpu.astype(np.float64) - hdr_ref) ** 2))),
                               "u8_rmse": float(np.sqrt(np.mean((to_u8(hdr_gpu) - to_u8(hdr_ref)) ** 2))),
                               "note": "GPU vs CPU oracle on the cpu_baseline sample's pixels, same seed and spp (equal Philox streams => RMSE 0 "
                                       "expected); the criterion for independent seeds, RMSE(gpu, cpu_a) <= 1.10 RMSE(cpu_b, cpu_a), is in tests/"}
        line["cpu_baseline"] = {"value": st["rays"] / (st["ms_total"] / 1e3) / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port",
                                "sample": f"pixel lattice stride {stride}x{stride} of the {W}x{H} frame, {s_spp} of {spp} spp: {st['paths']} paths, "
                                          f"{st['rays']} rays in {st['ms_total'] / 1e3:.1f} s"}
    print(json.dumps(line), file=_STDOUT, flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--frames-per-batch", type=int, default=0, choices=[0, 1, 2, 3, 4],
                    help="frames rendered together in the same wavefront batches in the timed region (0 = 1 on one GPU, 2 on 2-7, 4 on 8: "
                         "a rank's launches then have about the size they have on fewer GPUs)")
    ap.add_argument("--frames-in-flight", type=int, default=2, choices=[1, 2, 3, 4],
                    help="frames kept in flight on separate streams in the timed region (1 = one frame at a time)")
    args = ap.parse_args()
    # exactly ONE line goes to stdout (the JSON); the host mirror's progress prints (lib.rs:80,114, mesh.rs:115) go to stderr
    # (NCCL and the host mirror also write to fd 1 from C: redirect the descriptor itself, keep a private duplicate for the JSON)
    global _STDOUT
    sys.stdout.flush()
    _STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = sys.stderr
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
