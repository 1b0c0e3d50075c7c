"""CPU: the host-side logic of the torch-facing layer (fovgs/ops.py, fovgs/surface.py) with the C library stubbed out — the
capacity / overflow protocol, the deferred statistics check, the packed-model cache's state machine and the no-grad call path.
The kernels behind them are covered by the `-m gpu` tests; nothing here computes an image."""
import gc
import types

import numpy as np
import pytest
import torch

from fovgs import ops, surface


class _Stream:
    cuda_stream = 0

    def synchronize(self):
        pass


class _Event:
    """Stand-in for torch.cuda.Event: `landed` says whether the frame statistics guarded by it have reached the host buffer."""
    auto_land = True
    waits = 0

    def __init__(self):
        self.landed = False
        self.cuda_event = id(self)      # the handle the library gets (it records the event behind the colour stage)

    def record(self, stream=None):
        self.landed = _Event.auto_land

    def query(self):
        return self.landed

    def synchronize(self):
        _Event.waits += 1
        self.landed = True


def _library_records_slot_event(item):
    """What libfovgs does in a deferred frame: the scan kernel writes the statistics into the ring slot it was handed
    (item["early"][0]) and the library records that slot's event on the stream behind the colour stage."""
    item["ring_events"][item["ring_pos"]].record()


@pytest.fixture
def stubbed(monkeypatch):
    """ops with a fake library: workspaces are tiny CPU tensors, `launch` is the test's own function."""
    fake = types.SimpleNamespace(fovgs_workspace_bytes=lambda *a: 1024, fovgs_read_stats_async=lambda *a: 0)
    monkeypatch.setattr(ops, "lib", lambda: fake)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda device=None: _Stream())
    monkeypatch.setattr(torch.cuda, "Event", _Event)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda device=None: None)
    monkeypatch.setattr(_Event, "auto_land", True)
    monkeypatch.setattr(_Event, "waits", 0)
    monkeypatch.setattr(ops, "_pool", ops._Pool())
    monkeypatch.setattr(ops, "_train_capacity_hint", {})
    monkeypatch.setattr(ops, "_DEFERRED", False)
    return ops


def test_capacity_rounding_and_environment_override(monkeypatch):
    assert ops._round_capacity(1) == 1 << 22 and ops._round_capacity((1 << 22) + 1) == 2 << 22
    assert ops._round_capacity(1 << 40) == 0xFFFFFFF0                       # instance ids and counts travel in 32 bits
    monkeypatch.delenv("FOVGS_INSTANCE_CAPACITY", raising=False)
    assert ops._initial_capacity(10) == 1 << 20 and ops._initial_capacity(1 << 20) == 8 << 20
    monkeypatch.setenv("FOVGS_INSTANCE_CAPACITY", "12345")
    assert ops._initial_capacity(10) == 12345


def test_overflow_reruns_with_a_larger_workspace_and_never_truncates(stubbed):
    dev = torch.device("cpu")
    caps = []

    def launch(item, stream):
        caps.append(item["cap"])
        need = 5_000_000
        item["stats_np"][:] = 0
        item["stats_np"][0, 0] = need
        item["stats_np"][0, 1] = 1 if item["cap"] < need else 0
        item["stats_np"][0, 2] = 77

    item, st = stubbed._run_with_capacity(launch, dev, ops.MODE_FOV, 1000, 64, 64, fresh_workspace=False)
    assert len(caps) == 2 and caps[0] == 1 << 20 and caps[1] >= 5_000_000 * 1.25 and caps[1] % (1 << 22) == 0
    assert st["num_rendered"] == 5_000_000 and st["overflow"] == 0 and st["num_visible"] == 77
    assert stubbed.last_stats == st
    # the pooled workspace kept its grown capacity: the next frame runs once
    caps.clear()
    stubbed._run_with_capacity(launch, dev, ops.MODE_FOV, 1000, 64, 64, fresh_workspace=False)
    assert caps == [item["cap"]]


def test_frames_on_different_streams_do_not_share_a_workspace(stubbed, monkeypatch):
    dev = torch.device("cpu")
    used = []

    def launch(item, stream):
        used.append((stream, item["ws_ptr"]))
        item["stats_np"][:] = 0

    class S1(_Stream):
        cuda_stream = 111

    class S2(_Stream):
        cuda_stream = 222

    for S in (S1, S2, S1):
        monkeypatch.setattr(torch.cuda, "current_stream", lambda device=None, S=S: S())
        stubbed._run_with_capacity(launch, dev, ops.MODE_FOV, 100, 32, 32, fresh_workspace=False)
    assert [s for s, _ in used] == [111, 222, 111]
    assert used[0][1] != used[1][1] and used[0][1] == used[2][1]


def test_training_workspace_capacity_only_grows(stubbed):
    dev = torch.device("cpu")
    seen = []

    def launch_n(n):
        def launch(item, stream):
            seen.append(item["cap"])
            item["stats_np"][:] = 0
            item["stats_np"][0, 0] = n
            item["stats_np"][0, 1] = 1 if item["cap"] < n else 0
        return launch

    stubbed._run_with_capacity(launch_n(9_000_000), dev, ops.MODE_SUM, 1000, 64, 64, fresh_workspace=True)
    big = seen[-1]
    stubbed._run_with_capacity(launch_n(100), dev, ops.MODE_SUM, 1000, 64, 64, fresh_workspace=True)
    assert seen[-1] == big                                                  # same size again: the allocator's block is reused


def test_prefiltered_violation_and_deferred_overflow_raise(stubbed):
    dev = torch.device("cpu")

    def bad_prefilter(item, stream):
        item["stats_np"][:] = 0
        item["stats_np"][0, 8] = 3

    with pytest.raises(RuntimeError, match="Point is filtered although prefiltered is set"):
        stubbed._run_with_capacity(bad_prefilter, dev, ops.MODE_OBB, 10, 32, 32, fresh_workspace=False)

    stubbed.set_deferred_check(True)
    try:
        def overflowing(item, stream):
            slot = item["ring_pos"]
            item["stats_np"][slot] = 0
            item["stats_np"][slot, 0] = 30_000_000
            item["stats_np"][slot, 1] = 1 if item["cap"] < 30_000_000 else 0
            assert item["early"][0] == item["stats_ptr"] + 64 * slot        # the library is told to write THIS slot
            _library_records_slot_event(item)

        item, st = stubbed._run_with_capacity(overflowing, dev, ops.MODE_FOV, 20, 32, 32, fresh_workspace=False)
        assert st is None and item["ring_pending"][0]                       # not inspected yet
        with pytest.raises(RuntimeError, match="deferred frame overflowed"):
            stubbed._run_with_capacity(overflowing, dev, ops.MODE_FOV, 20, 32, 32, fresh_workspace=False)
        # the overflow was learnt: the pooled workspace has grown, the retry and every later frame fit and nothing raises
        item2, _ = stubbed._run_with_capacity(overflowing, dev, ops.MODE_FOV, 20, 32, 32, fresh_workspace=False)
        assert item2 is not item and item2["cap"] >= 30_000_000
        stubbed._run_with_capacity(overflowing, dev, ops.MODE_FOV, 20, 32, 32, fresh_workspace=False)
        stubbed.check_pending()
    finally:
        stubbed.set_deferred_check(False)


def test_deferred_statistics_are_only_read_after_their_copy_landed(stubbed, monkeypatch):
    """the host may run ahead of the GPU.  A frame's slot is inspected only once the event the library records
    behind its statistics says they have landed, every frame has its own slot (a ring), and an overflow in ANY frame is seen — not only the last one."""
    dev = torch.device("cpu")
    frame = [0]

    def launch(item, stream):
        slot = item["ring_pos"]
        item["stats_np"][slot] = 0
        item["stats_np"][slot, 0] = 30_000_000 if frame[0] == 2 else 1000
        item["stats_np"][slot, 1] = 1 if (frame[0] == 2 and item["cap"] < 30_000_000) else 0
        frame[0] += 1
        _library_records_slot_event(item)

    stubbed.set_deferred_check(True)
    try:
        monkeypatch.setattr(_Event, "auto_land", False)                     # the GPU is behind: no copy has landed
        items = [stubbed._run_with_capacity(launch, dev, ops.MODE_FOV, 20, 32, 32, fresh_workspace=False)[0] for _ in range(ops._RING)]
        item = items[0]
        assert all(it is item for it in items) and all(item["ring_pending"]) and _Event.waits == 0
        # frame 2 overflowed, but nothing may be concluded yet; the ring is full, so the next call waits for the OLDEST slot only
        stubbed._run_with_capacity(launch, dev, ops.MODE_FOV, 20, 32, 32, fresh_workspace=False)
        assert _Event.waits == 1
        stubbed._run_with_capacity(launch, dev, ops.MODE_FOV, 20, 32, 32, fresh_workspace=False)
        assert _Event.waits == 2
        with pytest.raises(RuntimeError, match="deferred frame overflowed"):  # slot 2 comes up: frame 2's overflow is seen
            stubbed._run_with_capacity(launch, dev, ops.MODE_FOV, 20, 32, 32, fresh_workspace=False)
        assert ops._pool.min_caps[item["key"]] >= 30_000_000
        # copies land: later frames are checked without blocking, on the grown workspace
        monkeypatch.setattr(_Event, "auto_land", True)
        w = _Event.waits
        it2, _ = stubbed._run_with_capacity(launch, dev, ops.MODE_FOV, 20, 32, 32, fresh_workspace=False)
        assert it2["cap"] >= 30_000_000
        stubbed._run_with_capacity(launch, dev, ops.MODE_FOV, 20, 32, 32, fresh_workspace=False)
        assert _Event.waits <= w + ops._RING        # only the old workspace's ring was drained (blocking) when it was replaced
    finally:
        stubbed.set_deferred_check(False)


def test_pool_evicts_least_recently_used_workspaces(stubbed, monkeypatch):
    dev = torch.device("cpu")
    monkeypatch.setattr(ops._pool, "budget", 3 * 1024)          # the fake library sizes every workspace at 1024 bytes

    def launch(item, stream):
        item["stats_np"][:] = 0

    for P in (10, 20, 30, 40, 50):
        stubbed._run_with_capacity(launch, dev, ops.MODE_FOV, P, 32, 32, fresh_workspace=False)
    assert [k[2] for k in ops._pool.items] == [30, 40, 50]
    stubbed._run_with_capacity(launch, dev, ops.MODE_FOV, 30, 32, 32, fresh_workspace=False)      # touch: 30 becomes most recent
    stubbed._run_with_capacity(launch, dev, ops.MODE_FOV, 60, 32, 32, fresh_workspace=False)
    assert [k[2] for k in ops._pool.items] == [50, 30, 60]
    ops.release_workspaces()
    assert len(ops._pool.items) == 0


def test_packed_model_cache_state_machine(monkeypatch):
    packs = []

    def fake_pack(P, M_rest, means3D, shs_rest, shs_dcs, opacities, device):
        packs.append(P)
        return torch.zeros(P, 64)

    monkeypatch.setattr(ops, "_pack_rows", fake_pack)
    monkeypatch.setattr(ops, "_MODEL_CACHE", True)
    ops.invalidate_model_cache()
    P = 8
    model = (torch.zeros(P, 3), torch.zeros(P, 15, 3), torch.zeros(P, 4, 3), torch.zeros(P, 4))
    dev = torch.device("cpu")
    assert ops._packed_rows(model, model, P, 15, dev, False) is None and packs == []      # first sighting: un-packed gather
    rows = ops._packed_rows(model, model, P, 15, dev, False)                              # second: packed once
    assert rows is not None and packs == [P]
    assert ops._packed_rows(model, model, P, 15, dev, False) is rows and packs == [P]     # then served from the cache
    model[3].add_(1.0)                                                                    # in-place update bumps _version
    assert ops._packed_rows(model, model, P, 15, dev, False) is None
    assert ops._packed_rows(model, model, P, 15, dev, False) is not None and packs == [P, P]
    fresh = tuple(t.clone() for t in model)                                               # one-shot tensors never pay
    assert ops._packed_rows(fresh, fresh, P, 15, dev, False) is None and packs == [P, P]
    n = len(ops._packed_cache)
    del model, rows
    gc.collect()
    assert len(ops._packed_cache) < n                                                     # entry dies with its tensors
    ops.set_model_cache(False)
    try:
        assert ops._packed_rows(fresh, fresh, P, 15, dev, False) is None
        assert ops._packed_rows(fresh, fresh, P, 15, dev, False) is None and packs == [P, P]
    finally:
        ops.set_model_cache(True)
        ops.invalidate_model_cache()


def test_no_grad_calls_skip_function_apply_and_grad_calls_do_not():
    calls = []

    class Fn(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, tag):
            calls.append(type(ctx).__name__)
            ctx.tag = tag
            ctx.save_for_backward(x)
            y = x * 2
            ctx.mark_non_differentiable(y)
            return y

        @staticmethod
        def backward(ctx, g):
            return None, None

    x = torch.ones(3, requires_grad=True)
    with torch.no_grad():
        y = surface._apply(Fn, x, "a")
    assert calls == ["_NoGradCtx"] and not y.requires_grad and torch.equal(y, torch.full((3,), 2.0))
    surface._apply(Fn, x, "b")
    assert len(calls) == 2 and calls[1] != "_NoGradCtx"
    e = surface._empty_if_none(None)
    assert e.numel() == 0 and surface._empty_if_none(None) is e and surface._empty_if_none(x) is x


def test_stats_dict_reads_unsigned_counters():
    item = {"stats_np": np.zeros((1, 16), np.int32)}
    item["stats_np"][0, 0] = -1                                             # 0xFFFFFFFF instances as int32
    item["stats_np"][0, 2] = 5
    st = ops._stats_dict(item)
    assert st["num_rendered"] == 0xFFFFFFFF and st["num_visible"] == 5 and st["overflow"] == 0
