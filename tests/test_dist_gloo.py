"""Multi-rank host logic on CPU (world_size 2, gloo): shard planning, the one reduce of the per-pixel f32
accumulation buffers to rank 0 and the finalize step of rbrt_b200.dist.  The per-rank renders are produced
by the oracle here (no GPU in this container); on the GPU box the same code path runs with NCCL and the
CUDA library (tests/test_gpu_render.py covers the shard arithmetic of the kernels)."""
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
        z, scene, cam = G.load("mesh1275_64x48")
        osc = O.OracleScene.from_scene(scene)
        spp = 6
        opts = _abi.RenderOptsC(seed=2, shard_mode=mode, shard_rank=rank, shard_count=world)
        accum = torch.from_numpy(osc.render_accum(cam.to_c(), spp, opts))
        W, H = cam.img_width_pix, cam.img_height_pix

        def fin(acc, w, h, n):
            return O.finalize(acc.numpy(), w, h, n)

        res = D.reduce_and_finalize(accum, W, H, spp, fin, dist)
        if rank == 0:
            rgb, hdr = res
            np.savez(out_path, rgb=rgb, hdr=hdr)
        else:
            assert res is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode_name", ["tiles", "samples"])
def test_two_rank_reduce_matches_single_rank(tmp_path, oracle, mode_name):
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
    if mode_name == "tiles":       # every pixel is summed by exactly one rank: bit-identical to one rank
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
