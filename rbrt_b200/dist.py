"""Multi-GPU rendering, one process per GPU.  The sharding, the per-GPU finalise and the gather / reduce on rank 0 live INSIDE
the C library (csrc/multi.cu: NCCL over NVLink); this module only brings the library's communicator up from an existing
torch.distributed process group (torch is the plumbing that carries the 128-byte NCCL unique id between the ranks).

The reference's only parallelism is rayon over image columns (lib.rs:84-86): pixels and samples are independent, so the path
shards without any exchange step.  Tile sharding gives an image bit-identical to the 1-GPU render (every pixel is summed and
finalised by exactly one rank); sample sharding changes the f32 summation order across ranks only.
"""
import ctypes as C

import numpy as np

from . import _abi
from .render import ImageBuffer, make_opts


def shard_sample_range(spp, rank, count):
    return (spp * rank) // count, (spp * (rank + 1)) // count


def tile_owner(row, col, width, count):
    """Rank that owns pixel (row, col) under tile sharding (csrc/common.cuh shard_pixel)."""
    tiles_x = (width + 7) // 8
    return ((row // 4) * tiles_x + (col // 8)) % count


def exchange_unique_id(dist, make_id, nbytes=_abi.COMM_ID_BYTES):
    """Rank 0 calls make_id() -> bytes; every rank returns those bytes (one broadcast over the torch process group, any backend)."""
    import torch
    rank = dist.get_rank()
    buf = torch.zeros(nbytes, dtype=torch.uint8)
    if rank == 0:
        raw = make_id()
        if len(raw) != nbytes:
            raise ValueError(f"unique id must be {nbytes} bytes")
        buf = torch.tensor(list(raw), dtype=torch.uint8)
    if dist.get_backend() == "nccl":
        buf = buf.cuda()
    dist.broadcast(buf, src=0)
    return bytes(buf.cpu().tolist())


def init_comm(dist=None):
    """Bring up the library's communicator over the ranks of the default torch.distributed process group.  Call after
    torch.cuda.set_device(local_rank) and rbrt_b200.gpu_init(local_rank).  Returns the rbrt_comm_info dict."""
    if dist is None:
        import torch.distributed as dist
    lib = _abi.lib()
    if dist.is_initialized() and dist.get_world_size() > 1:
        def make_id():
            raw = (C.c_uint8 * _abi.COMM_ID_BYTES)()
            _abi.check(lib.rbrt_gpu_comm_unique_id(raw))
            return bytes(raw)
        uid = exchange_unique_id(dist, make_id)
        raw = (C.c_uint8 * _abi.COMM_ID_BYTES)(*uid)
        _abi.check(lib.rbrt_gpu_comm_init_rank(raw, dist.get_rank(), dist.get_world_size()))
    info = _abi.CommInfoC()
    _abi.check(lib.rbrt_gpu_comm_info(info))
    return info.as_dict()


def render_scene_distributed(cam, num_samples, scene, shard_mode=_abi.SHARD_TILES, stats=None, hdr=False, **opts):
    """render_scene across all ranks of the library's communicator; rank 0 returns the ImageBuffer (or the HDR array), other
    ranks return None.  Without a communicator this is the 1-GPU render."""
    lib = _abi.lib()
    info = _abi.CommInfoC()
    _abi.check(lib.rbrt_gpu_comm_info(info))
    root = not info.active or info.rank == 0
    W, H = cam.img_width_pix, cam.img_height_pix
    o = make_opts(shard_mode=shard_mode, **opts)
    st = _abi.StatsC()
    out = np.empty((H, W, 3), dtype=np.float32 if hdr else np.uint8) if root else None
    fn = lib.rbrt_gpu_render_hdr if hdr else lib.rbrt_gpu_render
    _abi.check(fn(scene.handle(), cam.to_c(), int(num_samples), o, out.ctypes.data if root else None, st))
    if stats is not None:
        stats.update(st.as_dict())
    if not root:
        return None
    return out if hdr else ImageBuffer(out)
