"""Torch-facing operators on top of the C-ABI: tensor checks, output/workspace allocation (torch's caching
allocator owns all memory, like the reference's resizeFunctional callbacks, FOV/rasterize_points.cu:27-33) and
the overflow protocol.  These replace the reference's C++ glue functions:

  forward_fov   <- RasterizeGaussiansCUDA          FOV/rasterize_points.cu:35-152
  forward_ps1   <- RasterizeGaussiansCUDA          OBB|SUM/rasterize_points.cu:35-135
  backward_ps1  <- RasterizeGaussiansBackwardCUDA  SUM/rasterize_points.cu:137-216
  mark_visible  <- markVisible                     FOV/rasterize_points.cu:236-253

Host synchronisation: the reference blocks four times per frame (2x .item(), 2x cudaMemcpy).  Here nothing in
the kernel pipeline synchronises; the only host wait is ONE read of the 64-byte frame statistics after all
launches are queued (to detect instance-capacity overflow and to return `num_rendered`).  Set
FOVGS_DEFERRED_CHECK=1 to skip even that (the previous frame's statistics are then checked at the next call).
"""
import ctypes as C
import os
import weakref

import torch

from . import _lib
from ._lib import Camera, FovFwdArgs, FrameStats, MmfrFwdArgs, Ps1BwdArgs, Ps1FwdArgs, SmfrFwdArgs, check, lib, new_args

MODE_OBB, MODE_SUM, MODE_FOV = 0, 1, 2
# pruning-metric variants of the training rasterizer: SUM's workspace layout and backward, other statistics
MODE_MAX, MODE_LWMC = 3, 4
MODE_SMFR = 5   # shared-model foveation baseline (naive_pcheck_obb)
MODE_MMFR = 6   # multi-model foveation baseline (mmfr_pcheck_obb), one call per level
MODE_VANILLA = 7   # the stock diff_gaussian_rasterization: SUM's state and backward; no OBB test, no -4.5 cut, no statistics
_TRAIN_MODES = (MODE_SUM, MODE_MAX, MODE_LWMC, MODE_VANILLA)      # SUM's workspace layout (final_T, n_contrib, cov3D, point_list)
_STAT_MODES = (MODE_SUM, MODE_MAX, MODE_LWMC)                     # ... that return gaussians_count / contributions
_PS1_ABI_MODE = {MODE_OBB: _lib.FOVGS_PS1_OBB, MODE_SUM: _lib.FOVGS_PS1_SUM, MODE_MAX: _lib.FOVGS_PS1_MAX,
                 MODE_LWMC: _lib.FOVGS_PS1_LWMC, MODE_VANILLA: _lib.FOVGS_PS1_VANILLA}
_DEFERRED = os.environ.get("FOVGS_DEFERRED_CHECK", "0") == "1"
# Blocking (default) mode waits for the statistics the library copies out right after the colour stage, which hosts the tile scan (instance count,
# overflow flag: final under half of the way into the frame) — not for the end of the frame: the call returns with a validated
# instance count while colour / scatter / blend are still running, and an overflow is repaired (the frame re-runs into the same
# outputs, stream-ordered) before anybody can have consumed it.  The blend stage's own counters (`blend_consumed`,
# `blend_block_pairs`) are only final at the end of the frame: FOVGS_FULL_STATS=1 / set_full_stats(True) waits for them.
_FULL_STATS = os.environ.get("FOVGS_FULL_STATS", "0") == "1"


def set_full_sort(on):
    """Force the complete per-tile depth sort in the inference variants (default: lazy, consumption-driven)."""
    check(lib().fovgs_set_option(1, 1 if on else 0), "fovgs_set_option")


def set_no_tma(on):
    """Colour stage: use register-staged cooperative loads instead of TMA bulk copies (debug / comparison)."""
    check(lib().fovgs_set_option(2, 1 if on else 0), "fovgs_set_option")


def set_full_stats(on):
    """Blocking mode: wait for the end of the frame so that `last_stats` also carries the blend stage's counters."""
    global _FULL_STATS
    _FULL_STATS = bool(on)


def set_deferred_check(on):
    """Pipelined mode for the inference paths: skip the per-frame host read of the frame statistics; the previous
    frame's overflow flag is checked at the next call (and by `check_pending()`)."""
    global _DEFERRED
    _DEFERRED = bool(on)


def check_pending(device=None):
    """Synchronise and validate every frame whose statistics were not yet inspected (deferred mode)."""
    torch.cuda.synchronize(device)
    for key, item in list(_pool.items.items()):
        _drain_ring(item, key, block=True)


def _ptr(t):
    if t is None:
        return None
    if t.numel() == 0:
        return None
    return t.data_ptr()


def _prep(t, name, device, dtype=torch.float32, optional=False):
    """Contiguous fp32 CUDA view of an input (reference: `.contiguous().data<float>()`); empty -> None (Q13)."""
    if t is None or (isinstance(t, torch.Tensor) and t.numel() == 0):
        if optional:
            return None
        raise RuntimeError(f"{name} must be a non-empty tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if t.device != device:
        raise RuntimeError(f"{name} is on {t.device} but means3D is on {device}: all inputs must share one device")
    if t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def _camera(rs, device, keep):
    bg = _prep(rs.bg, "bg", device)
    view = _prep(rs.viewmatrix, "viewmatrix", device)
    proj = _prep(rs.projmatrix, "projmatrix", device)
    campos = _prep(rs.campos, "campos", device)
    keep += [bg, view, proj, campos]
    cam = Camera()
    cam.image_height = int(rs.image_height)
    cam.image_width = int(rs.image_width)
    cam.tanfovx = float(rs.tanfovx)
    cam.tanfovy = float(rs.tanfovy)
    cam.scale_modifier = float(rs.scale_modifier)
    cam.sh_degree = int(rs.sh_degree)
    cam.prefiltered = int(bool(rs.prefiltered))
    cam.debug = int(bool(rs.debug))
    cam.bg = bg.data_ptr()
    cam.viewmatrix = view.data_ptr()
    cam.projmatrix = proj.data_ptr()
    cam.campos = campos.data_ptr()
    return cam


_RING = 8   # per-workspace ring of pinned statistics slots: frames whose statistics may be un-inspected in deferred mode


class _Pool:
    """Reusable workspaces for the inference paths, one per (device, mode, P, W, H, stream): one workspace is one in-flight
    frame, frames queued on one stream are ordered, frames on different streams must not share scratch memory.
    Least-recently-used entries are dropped once the pool holds more than FOVGS_POOL_BYTES (default 24 GiB; the entry in use
    is never dropped), so resolution sweeps, MMFR's four level models or evaluation during densification do not pin one
    multi-GB workspace per configuration for the life of the process.  `release_workspaces()` empties it."""

    def __init__(self):
        import collections
        self.items = collections.OrderedDict()
        self.min_caps = {}          # capacities learnt from overflows (kept across evictions)
        self.budget = int(float(os.environ.get("FOVGS_POOL_BYTES", 24 * 2 ** 30)))

    def get(self, device, mode, P, W, H, min_cap=0, stream=0):
        key = (device.index, mode, P, W, H, stream)
        min_cap = max(min_cap, self.min_caps.get(key, 0))
        it = self.items.get(key)
        if it is None or it["cap"] < min_cap:
            if it is not None:
                _drain_ring(it, key, block=True, raise_on_overflow=False)
            cap = max(min_cap, _initial_capacity(P))
            it = _new_workspace(device, mode, P, W, H, cap)
            it["key"] = key
            self.items[key] = it
            total = sum(x["bytes"] for x in self.items.values())
            for k in list(self.items):
                if total <= self.budget or k == key:
                    continue
                old = self.items[k]
                if any(old["ring_pending"]):
                    continue                   # frames with un-inspected statistics: keep until they are checked
                total -= old["bytes"]
                del self.items[k]
        self.items.move_to_end(key)
        return it

    def require(self, key, cap):
        self.min_caps[key] = max(self.min_caps.get(key, 0), int(cap))

    def clear(self):
        self.items.clear()


def release_workspaces():
    """Drops every pooled inference workspace (their memory returns to torch's caching allocator)."""
    _pool.clear()


def _initial_capacity(P):
    env = os.environ.get("FOVGS_INSTANCE_CAPACITY")
    if env:
        return int(env)
    return int(min(max(1 << 20, 8 * P), 0xFFFFFFF0))


def _new_workspace(device, mode, P, W, H, cap):
    nbytes = lib().fovgs_workspace_bytes(P, W, H, cap, {MODE_FOV: 1, MODE_SMFR: 2, MODE_MMFR: 3}.get(mode, 0),
                                         1 if mode in _TRAIN_MODES else 0)
    if nbytes == 0:
        raise RuntimeError("fovgs_workspace_bytes rejected the frame configuration")
    ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
    # _RING slots of 16 ints: slot 0 serves the blocking mode; deferred frames cycle through all of them, each guarded by a
    # CUDA event recorded behind its 64-byte copy, so a slot is only read once its copy has landed and only reused once read
    stats = torch.zeros((_RING, 16), dtype=torch.int32)
    if torch.cuda.is_available():
        stats = stats.pin_memory()
    # `stats_np` aliases the (pinned) host tensor: reading 16 ints through it costs ~1 us, indexing the tensor ~3 us each
    return {"ws": ws, "cap": cap, "bytes": nbytes, "stats": stats, "stats_np": stats.numpy(), "stats_ptr": stats.data_ptr(),
            "ws_ptr": ws.data_ptr(), "ring_events": [None] * _RING, "ring_pending": [False] * _RING, "ring_pos": 0, "key": None,
            "early_event": None, "early_handle": None}


def _early_event(item, stream_obj):
    """The CUDA event the library records behind the early statistics copy (created on first use: torch creates the underlying
    cudaEvent_t lazily, at its first record)."""
    ev = item["early_event"]
    if ev is None:
        ev = item["early_event"] = torch.cuda.Event()
        ev.record(stream_obj)
        item["early_handle"] = int(ev.cuda_event)
    return ev


def _ring_event_handle(item, slot, stream_obj):
    """cudaEvent_t of ring slot `slot` (the library records it; torch creates the underlying event at its first record)."""
    if item["ring_events"][slot] is None:
        ev = item["ring_events"][slot] = torch.cuda.Event()
        ev.record(stream_obj)
        item.setdefault("ring_handles", {})[slot] = int(ev.cuda_event)
    return item["ring_handles"][slot]


_pool = _Pool()


def _read_stats(item, stream, slot=0):
    check(lib().fovgs_read_stats_async(item["ws_ptr"], item["stats_ptr"] + 64 * slot, stream), "fovgs_read_stats_async")


def _stats_dict(item, slot=0):
    s = item["stats_np"][slot].tolist()
    return {
        "num_rendered": s[0] & 0xFFFFFFFF,
        "overflow": s[1],
        "num_visible": s[2] & 0xFFFFFFFF,
        "num_blend_tiles": s[3] & 0xFFFFFFFF,
        "max_tile_instances": s[4] & 0xFFFFFFFF,
        "blend_consumed": s[5] & 0xFFFFFFFF,
        "blend_block_pairs": s[6] & 0xFFFFFFFF,
        "candidates": s[7] & 0xFFFFFFFF,
        "prefiltered_violations": s[8] & 0xFFFFFFFF,
    }


last_stats = {}


def _check_slot(item, key, slot, raise_on_overflow=True):
    """Inspects one landed statistics slot of a deferred frame."""
    item["ring_pending"][slot] = False
    st = _stats_dict(item, slot)
    if st["overflow"]:
        # the frame that overflowed has already been handed out truncated — that cannot be undone, so it is reported; the
        # pooled workspace grows to what the frame needed, so the caller's retry (and every later frame) fits
        need = _round_capacity(int(st["num_rendered"] * 1.25) + 1024)
        if key is not None:
            _pool.require(key, need)
        if raise_on_overflow:
            raise RuntimeError(f"fovgs: a deferred frame overflowed its instance capacity ({st['num_rendered']} > {item['cap']}); "
                               f"its image is incomplete — the workspace now grows to {need} instances, render the frame again")
    if st["prefiltered_violations"] and raise_on_overflow:
        raise RuntimeError("Point is filtered although prefiltered is set. This shouldn't happen! "
                           f"({st['prefiltered_violations']} Gaussians behind the near plane)")


def _drain_ring(item, key, block, raise_on_overflow=True, only_slot=None):
    """Checks pending deferred slots, oldest first: all of them when `block` (waiting for their events), otherwise only those
    whose 64-byte copy has already landed (event query)."""
    n = _RING
    start = item["ring_pos"]
    for i in range(n):
        slot = (start + i) % n            # ring_pos is the NEXT slot to be written = the oldest one still held
        if only_slot is not None and slot != only_slot:
            continue
        if not item["ring_pending"][slot]:
            continue
        ev = item["ring_events"][slot]
        if block:
            ev.synchronize()
        elif not ev.query():
            break                          # younger frames cannot have landed either
        _check_slot(item, key, slot, raise_on_overflow)


def _run_with_capacity(launch, device, mode, P, W, H, fresh_workspace):
    """Runs `launch(item)` with the tensor's device current (a tensor on another device than the current one would otherwise
    get its kernels launched on a stream of the wrong device)."""
    if device.type == "cuda" and torch.cuda.current_device() != device.index:
        with torch.cuda.device(device):
            return _run_with_capacity_impl(launch, device, mode, P, W, H, fresh_workspace)
    return _run_with_capacity_impl(launch, device, mode, P, W, H, fresh_workspace)


def _run_with_capacity_impl(launch, device, mode, P, W, H, fresh_workspace):
    """Runs `launch(item)`; on instance overflow grows the workspace and re-runs (no silent truncation)."""
    global last_stats
    cur_stream = torch.cuda.current_stream(device)
    stream = cur_stream.cuda_stream
    min_cap = 0
    while True:
        if fresh_workspace:
            cap = max(min_cap, _train_capacity_hint.get((device.index, P, W, H), _initial_capacity(P)))
            item = _new_workspace(device, mode, P, W, H, cap)
        else:
            item = _pool.get(device, mode, P, W, H, min_cap, stream)
        if _DEFERRED and not fresh_workspace:
            key = item["key"]
            # statistics of earlier frames whose copies have landed; then make sure the slot this frame will use is free
            _drain_ring(item, key, block=False)
            slot = item["ring_pos"]
            if item["ring_pending"][slot]:
                _drain_ring(item, key, block=True, only_slot=slot)     # the host ran _RING frames ahead: wait for the oldest
            if item is not _pool.items.get(key) or item["cap"] < _pool.min_caps.get(key, 0):
                continue                                                # an overflow was learnt: take the grown workspace
            # the library writes this frame's statistics into ring slot `slot` (the scan kernel stores them straight into the
            # pinned buffer) and records the slot's event behind the colour stage: no copy operation enters the stream
            item["early"] = (item["stats_ptr"] + 64 * slot, _ring_event_handle(item, slot, cur_stream))
            launch(item, stream)
            item["ring_pending"][slot] = True
            item["ring_pos"] = (slot + 1) % _RING
            return item, None
        if not fresh_workspace:
            _drain_ring(item, item["key"], block=True)                  # frames left over from a deferred phase
        if _FULL_STATS or not hasattr(cur_stream, "cuda_stream") or not torch.cuda.is_available():
            item["early"] = None
            launch(item, stream)
            _read_stats(item, stream)
            cur_stream.synchronize()
        else:
            ev = _early_event(item, cur_stream)
            item["early"] = (item["stats_ptr"], item["early_handle"])   # launch() passes these to the library
            launch(item, stream)
            ev.synchronize()                                            # binning done, statistics landed; the frame runs on
        st = _stats_dict(item)
        last_stats = st
        if st["prefiltered_violations"]:
            raise RuntimeError("Point is filtered although prefiltered is set. This shouldn't happen! "
                               f"({st['prefiltered_violations']} Gaussians behind the near plane)")
        if not st["overflow"]:
            if fresh_workspace:
                # The training workspace is allocated per forward (it is the saved state of backward).  Its capacity only
                # ever grows, in 4 Mi-instance steps: every iteration (a new camera, a new instance count) then asks torch's
                # caching allocator for the SAME size and gets the block the previous iteration released — a capacity that
                # followed num_rendered made every step a fresh multi-GB cudaMalloc (measured: 73 ms per forward).
                key = (device.index, P, W, H)
                need = int(st["num_rendered"] * 1.25) + 1024
                _train_capacity_hint[key] = max(_train_capacity_hint.get(key, 1 << 20), _round_capacity(need))
            return item, st
        min_cap = _round_capacity(int(st["num_rendered"] * 1.25) + 1024)


_train_capacity_hint = {}


def _round_capacity(n, step=1 << 22):
    return int(min(((int(n) + step - 1) // step) * step, 0xFFFFFFF0))


# ---- packed model rows for the foveated colour stage --------------------------------------------------------------------
# The colour stage gathers, per visible Gaussian, 180 B of SH rest (4-byte aligned), 48 B of dc, 16 B of opacity and 12 B of
# xyz from four tensors: ~11.5 DRAM sectors in 6-7 lines for 256 useful bytes.  Those four tensors are the static part of a
# foveated model (render_compose_gazes_fps.py passes the same tensor objects every frame: pc.get_xyz, pc.get_rest_features,
# shs_dcs.pt, opacities.pt), so they are re-laid ONCE into one aligned 256-byte row per Gaussian and the gather becomes one
# bulk copy of 8 full sectors.  The cache entry is bound to the caller's tensor OBJECTS: it is keyed by (id, data_ptr,
# _version, shape) of all four — any in-place update bumps `_version` and re-packs — and dropped by weak-reference callbacks
# when one of them dies (so a recycled id / address cannot alias).  Writes that bypass autograd's version counter
# (`tensor.data[...] = ...`, foreign kernels on `data_ptr()`) are invisible to it: call `invalidate_model_cache()` after
# such writes or set FOVGS_MODEL_CACHE=0.  `raster_settings.debug=True` re-packs and compares on every frame.
# Packing costs about one frame (1 ms for 6 M Gaussians), so it only happens at the SECOND consecutive call that brings the same
# four tensor objects: a caller that builds fresh tensors every frame never pays for it (and gets the un-packed gather).
_MODEL_CACHE = os.environ.get("FOVGS_MODEL_CACHE", "1") != "0"
_packed_cache = {}
_packed_candidate = [None]


def set_model_cache(on):
    global _MODEL_CACHE
    _MODEL_CACHE = bool(on)
    if not on:
        _packed_cache.clear()


def invalidate_model_cache():
    _packed_cache.clear()
    _packed_candidate[0] = None


def _pack_rows(P, M_rest, means3D, shs_rest, shs_dcs, opacities, device):
    rows = torch.empty((P, 64), dtype=torch.float32, device=device)
    stream = torch.cuda.current_stream(device).cuda_stream
    check(lib().fovgs_pack_color_rows(P, M_rest, means3D.data_ptr(), _ptr(shs_rest), shs_dcs.data_ptr(), opacities.data_ptr(),
                                      rows.data_ptr(), stream), "fovgs_pack_color_rows")
    return rows


def _packed_rows(user, prepared, P, M_rest, device, verify):
    """user: the caller's four tensor objects (means3D, shs_rest, shs_dcs, opacities); prepared: their contiguous fp32 forms."""
    if not _MODEL_CACHE or M_rest > 15 or any(not isinstance(t, torch.Tensor) for t in user if t is not None):
        return None
    key = tuple(None if t is None else (id(t), t.data_ptr(), t._version, tuple(t.shape), t.dtype) for t in user)
    ent = _packed_cache.get(key)
    if ent is None:
        if _packed_candidate[0] != key:      # first sighting of these tensors: remember them, gather un-packed this frame
            _packed_candidate[0] = key
            return None
        _packed_candidate[0] = None
        rows = _pack_rows(P, M_rest, *prepared, device)
        # an in-place update bumped `_version`: the entry of the same tensor objects at the older version is dead weight
        # (P * 256 bytes, 1.5 GB at 6 M Gaussians) — drop it now instead of waiting for four newer entries to push it out
        ident = tuple(None if k is None else k[:2] for k in key)
        for old in [k for k in _packed_cache if tuple(None if x is None else x[:2] for x in k) == ident]:
            for r in _packed_cache.pop(old)["refs"]:
                r.detach()
        if len(_packed_cache) >= 4:
            for r in _packed_cache.pop(next(iter(_packed_cache)))["refs"]:
                r.detach()
        refs = []
        for t in user:
            if t is not None:
                try:
                    refs.append(weakref.finalize(t, _packed_cache.pop, key, None))
                except TypeError:
                    return rows          # not weak-referenceable: use the rows for this call only
        ent = {"rows": rows, "refs": refs, "stream": None, "event": None}
        if rows.is_cuda:
            # the packing pass ran on the current stream; a frame queued later on ANOTHER stream must wait for it
            cur = torch.cuda.current_stream(device)
            ent["stream"] = cur.cuda_stream
            ent["event"] = torch.cuda.Event()
            ent["event"].record(cur)
        _packed_cache[key] = ent
        return rows
    if verify and not torch.equal(ent["rows"], _pack_rows(P, M_rest, *prepared, device)):
        raise RuntimeError("fovgs: the packed model cache is stale (a model tensor was modified without bumping its version "
                           "counter); call fovgs.ops.invalidate_model_cache() after such writes")
    if ent.get("event") is not None:
        cur = torch.cuda.current_stream(device)
        if cur.cuda_stream != ent["stream"]:
            cur.wait_event(ent["event"])
            # the rows were allocated on another stream: tell the caching allocator this stream reads them too, so an evicted
            # or finalized entry is not handed out again while a frame queued here is still gathering from it
            ent["rows"].record_stream(cur)
    return ent["rows"]


def _gaze_tensor(gazeArray, device):
    gaze = gazeArray
    if not isinstance(gaze, torch.Tensor):
        gaze = torch.tensor([float(gaze[0]), float(gaze[1])], dtype=torch.float32)
    if not gaze.is_cuda:
        gaze = gaze.to(device, non_blocking=True)
    return _prep(gaze, "gazeArray", device)


# entry point, args struct and pool mode of the three foveated forwards (they share everything but a few fields)
_FOV_KINDS = {
    "fov": ("fovgs_forward_fov", FovFwdArgs, MODE_FOV),
    "smfr": ("fovgs_forward_smfr", SmfrFwdArgs, MODE_SMFR),
    "mmfr": ("fovgs_forward_mmfr", MmfrFwdArgs, MODE_MMFR),
}


def _forward_foveated(kind, fields, means3D, scales, rotations, gazeArray, alpha, blending, raster_settings, want_lists, keep,
                      out_uint8=False):
    """The marshalling the three foveated forwards share: camera block, outputs, workspace, optional sorted lists, launch through
    the capacity protocol.  `fields`: the variant's own struct fields (name -> int / float / device pointer)."""
    entry, Args, mode = _FOV_KINDS[kind]
    device = means3D.device
    rs = raster_settings
    H, W = int(rs.image_height), int(rs.image_width)
    P = means3D.size(0)
    gaze = _gaze_tensor(gazeArray, device)
    cam = _camera(rs, device, keep)
    # out_uint8 (foveated rasterizer only): the blend epilogue writes the 8-bit image the reference's scripts store
    # (torchvision.utils.save_image's quantisation) INSTEAD of the fp32 one: 6.2 MB per 1080p frame instead of 24.9 MB
    color = torch.empty((3, H, W), dtype=torch.uint8 if out_uint8 else torch.float32, device=device)
    radii = torch.empty((P,), dtype=torch.int32, device=device)
    T = ((W + 15) // 16) * ((H + 15) // 16)
    lists = {}
    fn = getattr(lib(), entry)
    a = new_args(Args)
    a.cam = cam
    a.P = P
    a.means3D = means3D.data_ptr()
    a.scales = scales.data_ptr()
    a.rotations = rotations.data_ptr()
    a.gaze = gaze.data_ptr()
    a.alpha = float(alpha) if alpha is not None else 0.0
    a.blending = int(bool(blending))
    if out_uint8:
        a.out_color, a.out_color_u8 = None, color.data_ptr()
    else:
        a.out_color = color.data_ptr()
    a.radii = radii.data_ptr()
    for k, v in fields.items():
        setattr(a, k, v)

    def launch(item, stream):
        a.workspace = item["ws_ptr"]
        a.workspace_bytes = item["bytes"]
        a.max_instances = item["cap"]
        early = item.get("early")
        a.early_stats_host, a.early_stats_event = early if early else (None, None)
        if want_lists:
            lists["point_list"] = torch.zeros((item["cap"],), dtype=torch.int32, device=device)
            lists["ranges"] = torch.zeros((T, 2), dtype=torch.int32, device=device)
            a.out_point_list = lists["point_list"].data_ptr()
            a.out_ranges = lists["ranges"].data_ptr()
        check(fn(C.byref(a), stream), entry)

    item, st = _run_with_capacity(launch, device, mode, P, W, H, fresh_workspace=False)
    n = st["num_rendered"] if st is not None else -1
    if want_lists:
        return n, color, radii, lists["point_list"][: max(n, 0)], lists["ranges"], item
    return n, color, radii


def _empty_frame(means3D, raster_settings):
    device = means3D.device
    H, W = int(raster_settings.image_height), int(raster_settings.image_width)
    return 0, torch.zeros((3, H, W), dtype=torch.float32, device=device), torch.zeros((0,), dtype=torch.int32, device=device)


def forward_fov(means3D, opacities, scales, rotations, shs_rest, shs_dcs, highest_levels, gazeArray, alpha, blending,
                raster_settings, want_lists=False, out_uint8=False):
    """Foveated forward.  Returns (num_rendered, color[3,H,W], radii[P]) (+ point_list, ranges when want_lists).
    out_uint8: `color` is the uint8 image clamp(v * 255 + 0.5, 0, 255) (what torchvision.utils.save_image would store of the
    fp32 image, fov3dgs/render.py:52), written by the blend epilogue in place of the fp32 image."""
    if means3D.dim() != 2 or means3D.size(1) != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")
    device = means3D.device
    P = means3D.size(0)
    if P == 0:
        n0, c0, r0 = _empty_frame(means3D, raster_settings)
        return (n0, c0.to(torch.uint8), r0) if out_uint8 else (n0, c0, r0)
    user_model = (means3D, shs_rest if (shs_rest is not None and shs_rest.numel()) else None, shs_dcs, opacities)
    means3D = _prep(means3D, "means3D", device)
    opacities = _prep(opacities, "opacities", device)
    if opacities.numel() != P * 4:
        raise RuntimeError("opacities must have dimensions (num_points, 4) for the foveated rasterizer")
    scales = _prep(scales, "scales", device)
    rotations = _prep(rotations, "rotations", device)
    shs_rest = _prep(shs_rest, "shs_rest", device, optional=True)
    shs_dcs = _prep(shs_dcs, "shs_dcs", device)
    highest_levels = _prep(highest_levels, "highest_levels", device)
    if shs_dcs.numel() != P * 12 or highest_levels.numel() != P:
        raise RuntimeError("shs_dcs must be (P,4,3) and highest_levels (P,1)")
    M_rest = 0 if shs_rest is None else int(shs_rest.size(1))
    packed = _packed_rows(user_model, (means3D, shs_rest, shs_dcs, opacities), P, M_rest, device, bool(raster_settings.debug))
    fields = {"M_rest": M_rest, "opacities": opacities.data_ptr(), "shs_rest": _ptr(shs_rest), "shs_dcs": shs_dcs.data_ptr(),
              "highest_levels": highest_levels.data_ptr(), "packed_color_rows": None if packed is None else packed.data_ptr()}
    return _forward_foveated("fov", fields, means3D, scales, rotations, gazeArray, alpha, blending, raster_settings, want_lists,
                             [opacities, shs_rest, shs_dcs, highest_levels, packed], out_uint8=out_uint8)


def forward_smfr(means3D, opacities, scales, rotations, shs, highest_levels, gazeArray, alpha, blending, raster_settings,
                 want_lists=False):
    """SMFR baseline forward (replaces RasterizeGaussiansCUDA of naive_pcheck_obb/rasterize_points.cu:35-150).
    Returns (num_rendered, color[3,H,W], radii[P]) (+ point_list, ranges, workspace item when want_lists)."""
    if means3D.dim() != 2 or means3D.size(1) != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")
    device = means3D.device
    P = means3D.size(0)
    if P == 0:
        return _empty_frame(means3D, raster_settings)
    means3D = _prep(means3D, "means3D", device)
    opacities = _prep(opacities, "opacities", device)
    if opacities.numel() != P:
        raise RuntimeError("opacities must have dimensions (num_points, 1) for the shared-model foveated rasterizer")
    scales = _prep(scales, "scales", device)
    rotations = _prep(rotations, "rotations", device)
    shs = _prep(shs, "shs", device)
    highest_levels = _prep(highest_levels, "highest_levels", device)
    if highest_levels.numel() != P:
        raise RuntimeError("highest_levels must be (P,1)")
    fields = {"M": int(shs.size(1)), "opacities": opacities.data_ptr(), "shs": shs.data_ptr(), "highest_levels": highest_levels.data_ptr()}
    return _forward_foveated("smfr", fields, means3D, scales, rotations, gazeArray, alpha, blending, raster_settings, want_lists,
                             [opacities, shs, highest_levels])


def forward_mmfr(means3D, opacities, scales, rotations, shs, cur_level, gazeArray, alpha, blending, raster_settings,
                 want_lists=False):
    """MMFR baseline forward for ONE level model (replaces RasterizeGaussiansCUDA of mmfr_pcheck_obb/rasterize_points.cu).
    Returns (num_rendered, color[3,H,W], radii[P]) (+ point_list, ranges, workspace item when want_lists)."""
    if means3D.dim() != 2 or means3D.size(1) != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")
    device = means3D.device
    P = means3D.size(0)
    if P == 0:
        return _empty_frame(means3D, raster_settings)
    if cur_level is None:
        raise RuntimeError("cur_level must be given (0..3)")
    means3D = _prep(means3D, "means3D", device)
    opacities = _prep(opacities, "opacities", device)
    if opacities.numel() != P:
        raise RuntimeError("opacities must have dimensions (num_points, 1)")
    scales = _prep(scales, "scales", device)
    rotations = _prep(rotations, "rotations", device)
    shs = _prep(shs, "shs", device)
    fields = {"M": int(shs.size(1)), "opacities": opacities.data_ptr(), "shs": shs.data_ptr(), "cur_level": float(cur_level)}
    return _forward_foveated("mmfr", fields, means3D, scales, rotations, gazeArray, alpha, blending, raster_settings, want_lists,
                             [opacities, shs])


def forward_ps1(mode, means3D, opacities, scales, rotations, cov3D_precomp, shs, colors_precomp, raster_settings,
                want_lists=False, loss_map=None):
    """PS=1 forward (mode = MODE_OBB | MODE_SUM | MODE_MAX | MODE_LWMC | MODE_VANILLA; `loss_map` [H,W] for MODE_LWMC).
    Returns (num_rendered, color, radii, workspace_item[, gaussians_count, contributions][, point_list, ranges]); the
    statistics pair is returned by SUM / MAX / LWMC only."""
    if means3D.dim() != 2 or means3D.size(1) != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")
    device = means3D.device
    rs = raster_settings
    H, W = int(rs.image_height), int(rs.image_width)
    P = means3D.size(0)
    sum_mode = mode in _TRAIN_MODES
    stat_mode = mode in _STAT_MODES
    if P == 0:
        z = torch.zeros((3, H, W), dtype=torch.float32, device=device)
        r = torch.zeros((0,), dtype=torch.int32, device=device)
        out = [0, z, r, None]
        if stat_mode:
            out += [torch.zeros((0,), dtype=torch.int32, device=device), torch.zeros((0,), dtype=torch.float32, device=device)]
        return tuple(out)
    keep = []
    means3D = _prep(means3D, "means3D", device)
    opacities = _prep(opacities, "opacities", device)
    scales = _prep(scales, "scales", device, optional=True)
    rotations = _prep(rotations, "rotations", device, optional=True)
    cov3D_precomp = _prep(cov3D_precomp, "cov3D_precomp", device, optional=True)
    shs = _prep(shs, "shs", device, optional=True)
    colors_precomp = _prep(colors_precomp, "colors_precomp", device, optional=True)
    M = 0 if shs is None else int(shs.size(1))
    if mode == MODE_LWMC:
        loss_map = _prep(loss_map, "loss_map", device)
        if loss_map.numel() != H * W:
            raise RuntimeError("loss_map must have image_height * image_width elements")
    cam = _camera(rs, device, keep)
    color = torch.empty((3, H, W), dtype=torch.float32, device=device)
    radii = torch.empty((P,), dtype=torch.int32, device=device)
    gcount = contrib = None
    T = ((W + 15) // 16) * ((H + 15) // 16)
    lists = {}
    extra = {}

    def launch(item, stream):
        a = new_args(Ps1FwdArgs)
        a.cam = cam
        a.mode = _PS1_ABI_MODE[mode]
        a.loss_map = loss_map.data_ptr() if mode == MODE_LWMC else None
        a.P = P
        a.M = M
        a.means3D = means3D.data_ptr()
        a.opacities = opacities.data_ptr()
        a.scales = _ptr(scales)
        a.rotations = _ptr(rotations)
        a.cov3D_precomp = _ptr(cov3D_precomp)
        a.shs = _ptr(shs)
        a.colors_precomp = _ptr(colors_precomp)
        a.out_color = color.data_ptr()
        a.radii = radii.data_ptr()
        if sum_mode:   # (vanilla: scratch the C-ABI still wants)
            extra["gcount"] = torch.zeros((P,), dtype=torch.int32, device=device)
            extra["contrib"] = torch.zeros((P,), dtype=torch.float32, device=device)
            a.gaussians_count = extra["gcount"].data_ptr()
            a.contributions = extra["contrib"].data_ptr()
        a.workspace = item["ws_ptr"]
        a.workspace_bytes = item["bytes"]
        a.max_instances = item["cap"]
        early = item.get("early")
        a.early_stats_host, a.early_stats_event = early if early else (None, None)
        if want_lists:
            lists["point_list"] = torch.zeros((item["cap"],), dtype=torch.int32, device=device)
            lists["ranges"] = torch.zeros((T, 2), dtype=torch.int32, device=device)
            a.out_point_list = lists["point_list"].data_ptr()
            a.out_ranges = lists["ranges"].data_ptr()
        check(lib().fovgs_forward_ps1(C.byref(a), stream), "fovgs_forward_ps1")

    item, st = _run_with_capacity(launch, device, mode, P, W, H, fresh_workspace=sum_mode)
    n = st["num_rendered"] if st is not None else -1
    out = [n, color, radii, item]
    if stat_mode:
        out += [extra["gcount"], extra["contrib"]]
    if want_lists:
        out += [lists["point_list"][: max(n, 0)], lists["ranges"]]
    return tuple(out)


def backward_ps1(workspace_item, means3D, radii, scales, rotations, cov3D_precomp, shs, colors_precomp, raster_settings,
                 grad_out_color):
    """Gradient of the SUM-family forward (SUM / MAX / LWMC / VANILLA: the falloff cut of the matching forward travels in the
    workspace).  Returns the reference's 8-tuple
    (dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations)."""
    device = means3D.device
    rs = raster_settings
    P = means3D.size(0)
    keep = []
    means3D = _prep(means3D, "means3D", device)
    scales = _prep(scales, "scales", device, optional=True)
    rotations = _prep(rotations, "rotations", device, optional=True)
    cov3D_precomp = _prep(cov3D_precomp, "cov3D_precomp", device, optional=True)
    shs = _prep(shs, "shs", device, optional=True)
    colors_precomp = _prep(colors_precomp, "colors_precomp", device, optional=True)
    M = 0 if shs is None else int(shs.size(1))
    # The reference zero-fills nine gradient tensors with nine launches (SUM/rasterize_points.cu:171-179; 300 B per Gaussian).
    # Here they are nine views of ONE zero-filled slab: one allocation, one memset (blocks start on 256-byte boundaries: the
    # per-Gaussian kernel writes dL/dsh rows with 16-byte bulk stores).
    widths = (3, 3, 3, 4, 1, 6, 3 * M, 3, 4)      # means3D, means2D, colors, conic, opacity, cov3D, sh, scales, rotations
    offs, total = [], 0
    for w in widths:
        offs.append(total)
        total += (P * w + 63) // 64 * 64
    slab = torch.zeros((max(total, 1),), dtype=torch.float32, device=device)
    view = lambda i, shape: slab[offs[i]:offs[i] + P * widths[i]].view(shape)
    dL_dmeans3D, dL_dmeans2D, dL_dcolors = view(0, (P, 3)), view(1, (P, 3)), view(2, (P, 3))
    dL_dconic, dL_dopacity, dL_dcov3D = view(3, (P, 2, 2)), view(4, (P, 1)), view(5, (P, 6))
    dL_dsh, dL_dscales, dL_drotations = view(6, (P, M, 3)), view(7, (P, 3)), view(8, (P, 4))
    if P != 0:
        if workspace_item is None:
            raise RuntimeError("backward_ps1 needs the workspace of the matching forward call")
        g = _prep(grad_out_color, "grad_out_color", device)
        a = new_args(Ps1BwdArgs)
        a.cam = _camera(rs, device, keep)
        a.P = P
        a.M = M
        a.means3D = means3D.data_ptr()
        a.scales = _ptr(scales)
        a.rotations = _ptr(rotations)
        a.cov3D_precomp = _ptr(cov3D_precomp)
        a.shs = _ptr(shs)
        a.colors_precomp = _ptr(colors_precomp)
        a.radii = radii.contiguous().data_ptr()
        a.dL_dout_color = g.data_ptr()
        a.workspace = workspace_item["ws_ptr"]
        a.workspace_bytes = workspace_item["bytes"]
        a.max_instances = workspace_item["cap"]
        a.dL_dmeans2D = dL_dmeans2D.data_ptr()
        a.dL_dconic = dL_dconic.data_ptr()
        a.dL_dopacity = dL_dopacity.data_ptr()
        a.dL_dcolors = dL_dcolors.data_ptr()
        a.dL_dmeans3D = dL_dmeans3D.data_ptr()
        a.dL_dcov3D = dL_dcov3D.data_ptr()
        a.dL_dsh = _ptr(dL_dsh)
        a.dL_dscales = dL_dscales.data_ptr()
        a.dL_drotations = dL_drotations.data_ptr()
        stream = torch.cuda.current_stream(device).cuda_stream
        check(lib().fovgs_backward_ps1(C.byref(a), stream), "fovgs_backward_ps1")
    return dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations


def mark_visible(positions, viewmatrix, projmatrix):
    device = positions.device
    P = positions.size(0)
    present = torch.zeros((P,), dtype=torch.bool, device=device)
    if P != 0:
        pos = _prep(positions, "positions", device)
        view = _prep(viewmatrix, "viewmatrix", device)
        proj = _prep(projmatrix, "projmatrix", device)
        stream = torch.cuda.current_stream(device).cuda_stream
        check(lib().fovgs_mark_visible(P, pos.data_ptr(), view.data_ptr(), proj.data_ptr(), present.data_ptr(), stream),
              "fovgs_mark_visible")
    return present


def knn_mean_dist2(points):
    """simple_knn.distCUDA2 (fov3dgs/submodules/simple-knn/spatial.cu:15-26): mean squared distance of every point to its
    three nearest neighbours, float32 [P]."""
    if points.dim() != 2 or points.size(1) != 3:
        raise RuntimeError("points must have dimensions (num_points, 3)")
    device = points.device
    P = points.size(0)
    out = torch.zeros((P,), dtype=torch.float32, device=device)
    if P != 0:
        pts = _prep(points, "points", device)
        nbytes = lib().fovgs_knn_workspace_bytes(P)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        stream = torch.cuda.current_stream(device).cuda_stream
        check(lib().fovgs_knn_mean_dist2(P, pts.data_ptr(), out.data_ptr(), ws.data_ptr(), nbytes, stream), "fovgs_knn_mean_dist2")
    return out


def fov_tile_tables(item, W, H):
    """Parity helper: (tile_level, tile_min, grad_x, grad_y, blending) of the last FOV frame in `item`."""
    device = item["ws"].device
    T = ((W + 15) // 16) * ((H + 15) // 16)
    f = lambda: torch.empty((T,), dtype=torch.float32, device=device)
    lvl, mn, gx, gy = f(), f(), f(), f()
    bl = torch.empty((T,), dtype=torch.uint8, device=device)
    stream = torch.cuda.current_stream(device).cuda_stream
    check(lib().fovgs_fov_tile_tables(item["ws"].data_ptr(), W, H, lvl.data_ptr(), mn.data_ptr(), gx.data_ptr(),
                                      gy.data_ptr(), bl.data_ptr(), stream), "fovgs_fov_tile_tables")
    return lvl, mn, gx, gy, bl


def geometry(item, mode, P, W, H):
    """Parity helper: dense per-Gaussian projection results of the last frame (valid where radii > 0)."""
    device = item["ws"].device
    opts = dict(dtype=torch.float32, device=device)
    means2D = torch.zeros((P, 2), **opts)
    depths = torch.zeros((P,), **opts)
    conic = torch.zeros((P, 3), **opts)
    stream = torch.cuda.current_stream(device).cuda_stream
    if mode == MODE_FOV:
        lc = torch.zeros((P, 4, 3), **opts)
        check(lib().fovgs_fov_geometry(item["ws"].data_ptr(), P, W, H, means2D.data_ptr(), depths.data_ptr(),
                                       conic.data_ptr(), lc.data_ptr(), stream), "fovgs_fov_geometry")
        return {"means2D": means2D, "depths": depths, "conic": conic, "level_colors": lc}
    cov3D = torch.zeros((P, 6), **opts)
    rgb = torch.zeros((P, 3), **opts)
    check(lib().fovgs_ps1_geometry(item["ws"].data_ptr(), P, W, H, 1 if mode in _TRAIN_MODES else 0, means2D.data_ptr(),
                                   depths.data_ptr(), conic.data_ptr(), cov3D.data_ptr() if mode in _TRAIN_MODES else None,
                                   rgb.data_ptr(), stream), "fovgs_ps1_geometry")
    return {"means2D": means2D, "depths": depths, "conic": conic, "cov3D": cov3D, "rgb": rgb}


STAGE_NAMES = ("setup", "preprocess", "color", "scatter", "tile_sort", "blend")


def profile_enable(on=True, blend_only=False):
    """Stage events on the launch stream.  blend_only: bracket just the blend stage (2 events per frame instead of 7; the
    other durations read 0) — what bench.py's timed loop uses so the headline carries the dominant kernel's live time."""
    check(lib().fovgs_profile_enable((2 if blend_only else 1) if on else 0), "fovgs_profile_enable")


def profile_read():
    """Stage durations (ms) of the last profiled forward: dict keyed by STAGE_NAMES."""
    buf = (C.c_float * 6)()
    check(lib().fovgs_profile_read(buf, 6), "fovgs_profile_read")
    return dict(zip(STAGE_NAMES, [float(x) for x in buf]))


def profile_read_all():
    """Stage durations (ms) of every profiled frame still held (<= 256, oldest first): list of dicts."""
    out = []
    buf = (C.c_float * 6)()
    for k in range(lib().fovgs_profile_count()):
        check(lib().fovgs_profile_read_frame(k, buf, 6), "fovgs_profile_read_frame")
        out.append(dict(zip(STAGE_NAMES, [float(x) for x in buf])))
    return out


# ---------------------------------------------------------------- training-step elementwise work
class _Activate(torch.autograd.Function):
    """exp / normalize / sigmoid of the raw scaling, rotation and opacity parameters in one pass each way
    (fov3dgs/scene/gaussian_model.py:40-60,200-237)."""

    @staticmethod
    def forward(ctx, raw_scale, raw_rot, raw_opacity):
        device = raw_scale.device
        P = raw_scale.size(0)
        rs, rr, ro = _prep(raw_scale, "scaling", device), _prep(raw_rot, "rotation", device), _prep(raw_opacity, "opacity", device)
        scale, rot, opac = torch.empty_like(rs), torch.empty_like(rr), torch.empty_like(ro)
        stream = torch.cuda.current_stream(device).cuda_stream
        check(lib().fovgs_activate_forward(P, rs.data_ptr(), rr.data_ptr(), ro.data_ptr(), scale.data_ptr(), rot.data_ptr(),
                                           opac.data_ptr(), stream), "fovgs_activate_forward")
        ctx.save_for_backward(rr, scale, opac)
        return scale, rot, opac

    @staticmethod
    def backward(ctx, d_scale, d_rot, d_opac):
        rr, scale, opac = ctx.saved_tensors
        device = rr.device
        P = rr.size(0)
        need = ctx.needs_input_grad
        ds = _prep(d_scale, "d_scale", device) if need[0] else None
        dr = _prep(d_rot, "d_rot", device) if need[1] else None
        do = _prep(d_opac, "d_opacity", device) if need[2] else None
        g_s = torch.empty_like(scale) if need[0] else None
        g_r = torch.empty_like(rr) if need[1] else None
        g_o = torch.empty_like(opac) if need[2] else None
        stream = torch.cuda.current_stream(device).cuda_stream
        check(lib().fovgs_activate_backward(P, rr.data_ptr(), scale.data_ptr(), opac.data_ptr(), _ptr(ds), _ptr(dr), _ptr(do),
                                            _ptr(g_s), _ptr(g_r), _ptr(g_o), stream), "fovgs_activate_backward")
        return g_s, g_r, g_o


def activate(raw_scale, raw_rot, raw_opacity):
    """(get_scaling, get_rotation, get_opacity) of GaussianModel from (_scaling [P,3], _rotation [P,4], _opacity [P,1]),
    differentiable; one launch forward, one backward."""
    if raw_scale.dim() != 2 or raw_scale.size(1) != 3 or raw_rot.dim() != 2 or raw_rot.size(1) != 4:
        raise RuntimeError("scaling must have dimensions (num_points, 3) and rotation (num_points, 4)")
    if raw_rot.size(0) != raw_scale.size(0) or raw_opacity.numel() != raw_scale.size(0):
        raise RuntimeError("scaling, rotation and opacity must describe the same number of points")
    if raw_scale.size(0) == 0:
        return torch.empty_like(raw_scale), torch.empty_like(raw_rot), torch.empty_like(raw_opacity)
    return _Activate.apply(raw_scale, raw_rot, raw_opacity)
