"""Life-cycle and threading of the library state (round-1 VERDICT weak #7, ADVICE): scene create is asynchronous (build on the device's
build stream, two scratch slots, pooled scene blocks), every entry point may be called from any thread (one process-wide lock)."""
import threading

import numpy as np
import pytest

import rbrt_b200 as R
from rbrt_b200 import _abi, synth

from . import scenes as S

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def test_create_destroy_churn_with_changing_sizes(gpu, oracle):
    """Scenes of very different sizes created back to back (the scratch grows, the two slots alternate, blocks are pooled and
    re-used), some destroyed before their build can have finished, some rendered at once: every render must equal the oracle's."""
    cam = S.example_camera(64, 48)
    want = {}
    for k in range(14):
        sub = (1, 4, 2, 5, 0, 3, 3)[k % 7]
        n_keep = None if k % 3 else 7 + k
        sc = S.small_mesh_scene(sub, n_keep)
        if k % 4 == 1:
            sc.handle(); sc.close()                                    # destroyed while its build may still be running
            continue
        got = R.render_scene_hdr(cam, 2, sc, seed=k)                   # the render's stream waits for the scene's `ready` event
        key = (sub, n_keep)
        if key not in want:
            want[key] = oracle.OracleScene.from_scene(sc).render_hdr(cam.to_c(), 2, _abi.RenderOptsC(seed=k)), k
        ref, k0 = want[key]
        if k0 == k:
            assert np.array_equal(bits(got), bits(ref)), f"scene {k} (subdiv {sub}, keep {n_keep})"
        info = sc.info()
        assert info["num_triangles"] == len(sc.triangle_meshes[0].triangles) and info["ms_build"] >= 0.0
        if k % 2:
            sc.close()


def test_entry_points_from_several_threads(gpu):
    """render_scene from four threads at once (different scenes, cameras and seeds) gives what the same calls give one after another."""
    jobs = [(S.small_mesh_scene(3), S.example_camera(96, 64), 5, 11), (S.spheres_scene(), S.example_camera(80, 60), 7, 12),
            (S.quirk_scene(), S.quirk_camera(64, 48), 4, 13), (S.small_mesh_scene(2, 37), S.example_camera(50, 40), 6, 14)]
    want = [R.render_scene_hdr(cam, spp, sc, seed=seed) for sc, cam, spp, seed in jobs]
    got = [None] * len(jobs)
    errs = []

    def work(i):
        try:
            sc, cam, spp, seed = jobs[i]
            for _ in range(3):
                got[i] = R.render_scene_hdr(cam, spp, sc, seed=seed)
            fresh = S.small_mesh_scene(1 + i % 3)                      # scene create / info / destroy race with the other threads' renders
            assert fresh.info()["num_meshes"] == 1
            fresh.close()
        except Exception as e:                                         # noqa: BLE001
            errs.append((i, repr(e)))

    ts = [threading.Thread(target=work, args=(i,)) for i in range(len(jobs))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs
    for i, (g, w) in enumerate(zip(got, want)):
        assert np.array_equal(bits(g), bits(w)), f"thread {i}"


def test_release_cache_and_render_again(gpu):
    scene, cam = S.small_mesh_scene(3), S.example_camera(64, 48)
    a = R.render_scene_hdr(cam, 3, scene, seed=5)
    assert _abi.lib().rbrt_gpu_release_cache() == 0                    # wavefront pools, build scratch, pooled scene blocks
    b = R.render_scene_hdr(cam, 3, scene, seed=5)
    assert np.array_equal(bits(a), bits(b))
    fresh = S.small_mesh_scene(3)
    assert np.array_equal(bits(R.render_scene_hdr(cam, 3, fresh, seed=5)), bits(a))
