"""Multi-rank host logic on CPU (world_size 2, gloo).  On the GPU the sharding, the per-rank finalise and the gather / reduce on
rank 0 run inside the C library over NCCL (csrc/multi.cu; tests/test_gpu_multi.py, scripts/multi_gpu_check.py); what stays on the
host is (1) carrying the 128-byte NCCL unique id from rank 0 to the other ranks over the torch process group
(rbrt_b200.dist.exchange_unique_id) and (2) the shard arithmetic.  Both are exercised here with two real processes: each rank
renders ITS shard with the oracle, finalises its own pixels, rank 0 gathers — the same plan the library executes — and the
composed image must equal the one-rank render."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, mode, out_path):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from oracle import oracle_ffi as O
    from rbrt_b200 import _abi
    from rbrt_b200 import dist as D
    from tests import golden_util as G

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # (1) the unique id made on rank 0 reaches every rank unchanged
        made = bytes((7 * k + 3) % 256 for k in range(_abi.COMM_ID_BYTES))
        uid = D.exchange_unique_id(dist, lambda: made)
        assert uid == made and len(uid) == 128
        # (2) the library's plan with the oracle as the renderer: shard -> per-rank finalise -> gather (tiles) / sum (samples)
        z, scene, cam = G.load("mesh1275_64x48")
        osc = O.OracleScene.from_scene(scene)
        spp = 6
        opts = _abi.RenderOptsC(seed=2, shard_mode=mode, shard_rank=rank, shard_count=world)
        accum = osc.render_accum(cam.to_c(), spp, opts)
        W, H = cam.img_width_pix, cam.img_height_pix
        if mode == _abi.SHARD_TILES:
            rgb, hdr = O.finalize(accum, W, H, spp)                          # every rank finalises its own pixels
            parts_rgb = [torch.zeros(H, W, 3, dtype=torch.uint8) for _ in range(world)] if rank == 0 else None
            parts_hdr = [torch.zeros(H, W, 3) for _ in range(world)] if rank == 0 else None
            dist.gather(torch.from_numpy(rgb), parts_rgb, dst=0)
            dist.gather(torch.from_numpy(hdr), parts_hdr, dst=0)
            if rank == 0:
                owner = np.array([[D.tile_owner(r, c, W, world) for c in range(W)] for r in range(H)])
                rgb_out, hdr_out = np.zeros((H, W, 3), np.uint8), np.zeros((H, W, 3), np.float32)
                for q in range(world):
                    rgb_out[owner == q] = parts_rgb[q].numpy()[owner == q]
                    hdr_out[owner == q] = parts_hdr[q].numpy()[owner == q]
                np.savez(out_path, rgb=rgb_out, hdr=hdr_out)
        else:
            s0, s1 = D.shard_sample_range(spp, rank, world)
            assert s1 > s0
            t = torch.from_numpy(accum)
            dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)
            if rank == 0:
                rgb, hdr = O.finalize(t.numpy(), W, H, spp)
                np.savez(out_path, rgb=rgb, hdr=hdr)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode_name", ["tiles", "samples"])
def test_two_rank_plan_matches_single_rank(tmp_path, oracle, mode_name):
    import torch.multiprocessing as mp

    from rbrt_b200 import _abi

    from . import golden_util as G

    mode = _abi.SHARD_TILES if mode_name == "tiles" else _abi.SHARD_SAMPLES
    out = str(tmp_path / "r0.npz")
    mp.spawn(_worker, args=(2, _free_port(), mode, out), nprocs=2, join=True)
    got = np.load(out)
    z, scene, cam = G.load("mesh1275_64x48")
    osc = oracle.OracleScene.from_scene(scene)
    ref_hdr = osc.render_hdr(cam.to_c(), 6, _abi.RenderOptsC(seed=2))
    ref_rgb = osc.render(cam.to_c(), 6, _abi.RenderOptsC(seed=2))
    if mode_name == "tiles":       # every pixel is summed and finalised by exactly one rank: bit-identical to one rank
        assert np.array_equal(got["hdr"].view(np.uint32), ref_hdr.view(np.uint32))
        assert np.array_equal(got["rgb"], ref_rgb)
    else:                          # per-rank partial sums are re-associated by the reduce
        assert np.allclose(got["hdr"], ref_hdr, rtol=1e-5, atol=1e-6)
        assert np.abs(got["rgb"].astype(int) - ref_rgb.astype(int)).max() <= 1


def test_shard_plans():
    from rbrt_b200 import dist as D
    for spp, n in [(64, 8), (50, 8), (5, 8), (1024, 3)]:
        rs = [D.shard_sample_range(spp, r, n) for r in range(n)]
        assert rs[0][0] == 0 and rs[-1][1] == spp and all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
    owners = np.array([[D.tile_owner(r, c, 70, 4) for c in range(70)] for r in range(30)])
    assert set(np.unique(owners)) == {0, 1, 2, 3}
    assert (owners[:4, :8] == owners[0, 0]).all() and owners[0, 8] == (owners[0, 0] + 1) % 4   # 8x4 tiles, round-robin
