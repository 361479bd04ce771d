"""Multi-GPU rendering: one process per GPU (torch.distributed), the image sharded by interleaved
8x4-pixel tiles or by sample range, no data-path collective except ONE reduce of the per-pixel f32
accumulation buffers to rank 0 (NCCL over NVLink on GPUs; gloo on CPU in tests).

The reference's only parallelism is rayon over image columns (lib.rs:84-86): pixels and samples are
independent, so the path shards without any exchange step.  Tile sharding gives an image bit-identical
to the 1-GPU render (every pixel is summed by exactly one rank, the others add +0.0); sample sharding
changes the f32 summation order across ranks only.
"""
import numpy as np

from . import _abi
from .render import ImageBuffer, make_opts


def shard_sample_range(spp, rank, count):
    return (spp * rank) // count, (spp * (rank + 1)) // count


def tile_owner(row, col, width, count):
    """Rank that owns pixel (row, col) under tile sharding (csrc/common.cuh shard_pixel)."""
    tiles_x = (width + 7) // 8
    return ((row // 4) * tiles_x + (col // 8)) % count


def reduce_and_finalize(accum, width, height, num_samples, finalize_fn, dist=None, dst=0):
    """Sum the ranks' accumulation buffers on `dst` and apply lib.rs:101,116-122 there.
    accum: torch tensor [H*W*4] f32 on the backend's device.  Returns finalize_fn(accum) on dst, None elsewhere."""
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(accum, dst=dst, op=dist.ReduceOp.SUM)
        if dist.get_rank() != dst:
            return None
    return finalize_fn(accum, width, height, num_samples)


def render_scene_distributed(cam, num_samples, scene, shard_mode=_abi.SHARD_TILES, stats=None, hdr=False, **opts):
    """render_scene across all ranks of the default process group; rank 0 returns the ImageBuffer
    (or the HDR array), other ranks return None.  Call after torch.cuda.set_device(local_rank) and
    rbrt_gpu_init(local_rank)."""
    import torch
    import torch.distributed as dist

    lib = _abi.lib()
    W, H = cam.img_width_pix, cam.img_height_pix
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    o = make_opts(shard_mode=shard_mode if world > 1 else _abi.SHARD_NONE, shard_rank=rank, shard_count=world, **opts)
    accum = torch.empty(H * W * 4, dtype=torch.float32, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    st = _abi.StatsC()
    _abi.check(lib.rbrt_gpu_render_accum_device(scene.handle(), cam.to_c(), int(num_samples), o, accum.data_ptr(), stream, st))
    if stats is not None:
        stats.update(st.as_dict())

    def fin(acc, w, h, spp):
        if hdr:
            out = torch.empty(h * w * 3, dtype=torch.float32, device="cuda")
            _abi.check(lib.rbrt_gpu_finalize_device(acc.data_ptr(), w, h, spp, None, out.data_ptr(), stream))
            return out.cpu().numpy().reshape(h, w, 3)
        out = torch.empty(h * w * 3, dtype=torch.uint8, device="cuda")
        _abi.check(lib.rbrt_gpu_finalize_device(acc.data_ptr(), w, h, spp, out.data_ptr(), None, stream))
        return ImageBuffer(out.cpu().numpy().reshape(h, w, 3))

    return reduce_and_finalize(accum, W, H, int(num_samples), fin, dist if world > 1 else None)
