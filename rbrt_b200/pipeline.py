"""FramePipeline — several render_scene calls in flight on one GPU (one per CUDA stream).

A frame of the wavefront tracer ends with a sparse phase: the last bounce iterations and the tail kernel
are bounded by the latency of their longest path (C3: ~2 ms of a 38 ms frame on one GPU, ~1.7 ms of the
6.4 ms a rank spends on its eighth of the frame), during which most of the GPU idles.  The reference
renders one image per process, so nothing can hide that there; a host that renders a SEQUENCE of images
(an animation, a camera sweep, progressive refinement) can: frame k+1 is enqueued on a second stream with
its own pool of wavefront state (RBRT_OPT_POOL_*) while frame k is still finishing, and its dense first
bounces fill the SMs frame k's sparse phase leaves empty.  Every frame is computed by exactly the same
kernels on the same inputs as a lone render_scene call, so images stay bit-identical (tests).

    pipe = FramePipeline(width, height, depth=2)
    for cam, scene in frames:
        done = pipe.submit(cam, spp, scene)      # returns the frame submitted `depth` calls ago (or None)
    rest = pipe.drain()

With torch.distributed initialised every rank calls submit() in the same order; the per-frame reduce of the
f32 accumulation buffers (dist.py) is enqueued on the frame's stream and rank 0 gets the images.
"""
import numpy as np

from . import _abi
from .render import ImageBuffer, make_opts


class _Slot:
    def __init__(self, torch, n_px, host_output, hdr):
        self.stream = torch.cuda.Stream()
        self.accum = torch.empty(n_px * 4, dtype=torch.float32, device="cuda")
        self.out = torch.empty(n_px * 3, dtype=torch.float32 if hdr else torch.uint8, device="cuda")
        self.host = torch.empty(n_px * 3, dtype=self.out.dtype, pin_memory=True) if host_output else None
        self.done = torch.cuda.Event()
        self.busy = False
        self.keep = None       # the scene (and anything else) that must outlive the frame in flight
        self.tag = None


class FramePipeline:
    def __init__(self, width, height, depth=2, host_output=True, hdr=False, shard_mode=_abi.SHARD_TILES):
        import torch
        import torch.distributed as dist
        if depth not in (1, 2, 3, 4):
            raise ValueError("depth must be 1..4 (the library keeps four pools of wavefront state per device)")
        self._torch, self._dist = torch, dist
        self.width, self.height, self.depth, self.hdr = int(width), int(height), depth, hdr
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.shard_mode = shard_mode
        self._slots = [_Slot(torch, self.width * self.height, host_output and self.rank == 0, hdr) for _ in range(depth)]
        self._n = 0
        self._lib = _abi.lib()

    # ------------------------------------------------------------------ internals
    def _collect(self, slot):
        """Wait for the frame in `slot` and hand out its result (rank 0: image; other ranks: None)."""
        slot.done.synchronize()
        slot.busy = False
        keep, tag = slot.keep, slot.tag
        slot.keep = slot.tag = None
        res = None
        if self.rank == 0:
            src = slot.host if slot.host is not None else slot.out
            if slot.host is not None:
                arr = slot.host.numpy().reshape(self.height, self.width, 3).copy()
                res = arr if self.hdr else ImageBuffer(arr)
            else:
                res = src          # device tensor, valid until the slot is reused
        return res, tag, keep

    # ------------------------------------------------------------------ API
    def submit(self, cam, num_samples, scene, tag=None, keep=None, **opts):
        """Enqueue render_scene(cam, num_samples, scene) and return at once.  Returns (image, tag) of the frame whose
        slot is being reused — the one submitted `depth` calls earlier — or None while the pipeline fills.  `scene`
        must stay alive (not closed) until its frame has been returned; it is held here until then."""
        torch, dist = self._torch, self._dist
        slot = self._slots[self._n % self.depth]
        finished = None
        if slot.busy:
            img, t, _ = self._collect(slot)
            finished = (img, t)
        cam_c = cam.to_c() if hasattr(cam, "to_c") else cam
        handle = scene.handle() if hasattr(scene, "handle") else scene
        W, H = int(cam_c.img_width_pix), int(cam_c.img_height_pix)
        if (W, H) != (self.width, self.height):
            raise ValueError("camera size differs from the pipeline's")
        sm, sr, sc = opts.pop("shard_mode", _abi.SHARD_NONE), opts.pop("shard_rank", 0), opts.pop("shard_count", 1)
        if self.world > 1:                                    # one process per GPU: the process group decides the shard
            sm, sr, sc = self.shard_mode, self.rank, self.world
        o = make_opts(shard_mode=sm, shard_rank=sr, shard_count=sc, pool=self._n % self.depth, **opts)
        s = slot.stream
        with torch.cuda.stream(s):
            # stats = NULL: the call only enqueues (no event synchronisation inside the library)
            _abi.check(self._lib.rbrt_gpu_render_accum_device(handle, cam_c, int(num_samples), o, slot.accum.data_ptr(), s.cuda_stream, None))
            if self.world > 1:
                dist.reduce(slot.accum, dst=0, op=dist.ReduceOp.SUM)
            if self.rank == 0:
                rgb, hdr = (None, slot.out.data_ptr()) if self.hdr else (slot.out.data_ptr(), None)
                _abi.check(self._lib.rbrt_gpu_finalize_device(slot.accum.data_ptr(), W, H, int(num_samples), rgb, hdr, s.cuda_stream))
                if slot.host is not None:
                    slot.host.copy_(slot.out, non_blocking=True)
            slot.done.record(s)
        slot.busy, slot.keep, slot.tag = True, (scene, keep), tag
        self._n += 1
        return finished

    def drain(self):
        """Wait for every frame still in flight; returns their (image, tag) in submission order."""
        out = []
        for k in range(self.depth):
            slot = self._slots[(self._n + k) % self.depth]
            if slot.busy:
                img, t, _ = self._collect(slot)
                out.append((img, t))
        return out

    def wait_on(self, stream=None):
        """Make `stream` (default: the current stream) wait for everything enqueued so far — for device-side timing."""
        torch = self._torch
        stream = stream or torch.cuda.current_stream()
        for slot in self._slots:
            if slot.busy:
                stream.wait_event(slot.done)
