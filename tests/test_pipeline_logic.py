"""FramePipeline's host-side logic — grouping of frames into batches, slot reuse, the order in which frames leave the
pipeline — with the GPU work replaced by a recorder (no CUDA needed)."""
import pytest

from rbrt_b200 import _abi
from rbrt_b200.pipeline import FramePipeline


class _FakeEvent:
    def synchronize(self):
        pass


class _FakeSlot:
    def __init__(self):
        self.stream, self.host, self.out, self.done = None, None, None, _FakeEvent()
        self.busy, self.keep, self.tags = False, None, []


class _Cam:
    img_width_pix, img_height_pix = 16, 8


class _Scene:
    def __init__(self, h):
        self._h = h

    def handle(self):
        return self._h


class RecordingPipeline(FramePipeline):
    """Same bookkeeping, no device: every launched group is recorded as (slot index, pool, [tags], scene handle, spp, opts)."""

    def __init__(self, depth, fpb, rank=1):
        self.width, self.height, self.depth, self.hdr, self.fpb = 16, 8, depth, False, fpb
        self.world, self.rank, self.shard_mode = 1, rank, _abi.SHARD_TILES      # rank != 0: _collect hands out None images
        self._torch = None
        self._slots = [_FakeSlot() for _ in range(depth)]
        self._n, self._pending, self._pending_key, self._pending_hv = 0, [], None, None
        self.groups = []

    def _enqueue(self, slot, frames, scene, spp, opts):
        self.groups.append((self._slots.index(slot), self._n % self.depth, [f[2] for f in frames], scene.handle(), spp, opts))


@pytest.mark.parametrize("depth,fpb", [(1, 1), (2, 1), (2, 2), (3, 4), (4, 3)])
def test_frames_leave_in_submission_order_and_groups_are_homogeneous(depth, fpb):
    a, b = _Scene(11), _Scene(22)
    jobs = [(a, 4, {}), (a, 4, {}), (a, 4, {}), (a, 4, {}), (a, 4, {}), (b, 4, {}), (b, 4, {}), (b, 2, {}), (a, 2, {"max_depth": 3}),
            (a, 2, {"max_depth": 3}), (a, 2, {}), (a, 2, {})]
    pipe = RecordingPipeline(depth, fpb)
    out = []
    for k, (sc, spp, kw) in enumerate(jobs):
        got = pipe.submit(_Cam(), spp, sc, tag=k, seed=k, **kw)
        assert all(img is None for img, _ in got)
        out += got
        assert len(pipe._pending) < fpb                                      # a full group is launched at once
    out += pipe.drain()
    assert [t for _, t in out] == list(range(len(jobs)))                      # every frame exactly once, in order
    assert not pipe._pending and not any(s.busy for s in pipe._slots)
    seen = []
    for n, (slot, pool, tags, handle, spp, opts) in enumerate(pipe.groups):
        assert slot == pool == n % depth                                      # slots (= wavefront pools) are used round-robin
        assert 1 <= len(tags) <= fpb
        for t in tags:                                                        # a group never mixes scenes, sample counts or options
            assert (jobs[t][0].handle(), jobs[t][1], jobs[t][2]) == (handle, spp, opts)
        seen += tags
    assert seen == list(range(len(jobs)))
    # consecutive equal jobs are packed: 5 x (a, 4) need ceil(5 / fpb) groups
    assert sum(1 for g in pipe.groups if g[3] == 11 and g[4] == 4) == -(-5 // fpb)


def test_a_frame_is_returned_only_when_its_slot_is_reused_or_drained():
    sc = _Scene(5)
    pipe = RecordingPipeline(depth=2, fpb=1)
    assert pipe.submit(_Cam(), 1, sc, tag="f0") == []
    assert pipe.submit(_Cam(), 1, sc, tag="f1") == []
    assert [t for _, t in pipe.submit(_Cam(), 1, sc, tag="f2")] == ["f0"]     # slot 0 reused: its frame is collected first
    assert [t for _, t in pipe.flush()] == []                                 # nothing pending
    assert [t for _, t in pipe.drain()] == ["f1", "f2"]
    pipe = RecordingPipeline(depth=1, fpb=3)
    assert pipe.submit(_Cam(), 1, sc, tag=0) == [] and pipe.submit(_Cam(), 1, sc, tag=1) == []
    assert [t for _, t in pipe.flush()] == [] and len(pipe.groups) == 1 and pipe.groups[0][2] == [0, 1]   # partial group launched
    assert [t for _, t in pipe.drain()] == [0, 1]


def test_wrong_camera_size_and_bad_parameters():
    class Big:
        img_width_pix, img_height_pix = 17, 8
    pipe = RecordingPipeline(2, 2)
    with pytest.raises(ValueError):
        pipe.submit(Big(), 1, _Scene(1))
