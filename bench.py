#!/usr/bin/env python
"""bench.py — headline benchmark of the hot path: frames/s of the foveated rasterizer call at 1920x1080 on a
~6 M-Gaussian "bicycle-shaped" synthetic scene (BASELINE.json metric; config 3 "4-level foveated, moving gaze").

  python bench.py --gpus 1 --steps K --warmup W            # our sm_100a library through the drop-in package
  python bench.py --impl reference ...                     # the UNMODIFIED reference CUDA extension (oracle/_ref)
  python -m torch.distributed.run --nproc-per-node N bench.py --gpus N ...   # frames sharded by (camera, gaze)

A "step" is one rendered frame: camera = ring[f % 30], gaze = the reference's 9-gaze grid cycled per frame
(render_compose_gazes_fps.py:26) so the tile levels change every frame.  Per rank the K frames are
f = rank + i * world (weak scaling: the model is replicated, frames are independent; NCCL only gathers timings).

value  : whole-job frames/s with every input resident in HBM, timed on the device (CUDA events around the K steps,
         max over ranks), library called in its pipelined mode (no per-frame host read-back).
e2e    : the same frames/s through the public API with HOST inputs: per frame the camera matrices + gaze are copied
         from pinned host memory and the [3,H,W] image is copied back to pinned host memory (copy of frame i overlaps
         the rendering of frame i+1 on a copy stream, a ring of host buffers; all copies finish inside the timed region),
         wall clock with a synchronize on both sides.  `e2e.value` is the library's DEFAULT drop-in behaviour (one blocking
         64-byte statistics read per frame, like the reference's own host syncs) for both arms; `e2e.pipelined_value` is
         the same loop in the opt-in serving mode (ops.set_deferred_check: the statistics of frame i are inspected when
         frame i+1 is queued) and is reported as an extra, never as the headline.
roofline: ONE byte model — SURVEY.md §8(d)'s ALGORITHMIC bytes per frame, split by stage (DESIGN.md §5 states the same
         split) — over the stage durations from CUDA events recorded by the library on its launch stream over the timed
         region; `roofline.stages` lists bytes / ms / GB/s / frac for every stage, the headline entry is the dominant
         kernel (largest mean time).  `roofline.alu` is the blend's pixel-Gaussian pairs/s against the fp32 issue peak.
extra   : BASELINE.json configs 2 and 5 on the same scene and box: `ps1_fwd` (pcheck_obb full-quality forward,
         render.py:47) and `train_step` (pcheck_obb_sum forward + backward with a fixed dL/dpixel, eff_finetune.py:107,127),
         device-timed over the same K steps; the reference arm prints the same keys for the reference binaries.
cpu_baseline: the CPU oracle (oracle/fovgs_oracle.c, a port: the reference ships no CPU path) on ONE frame of the same
         workload, all host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "fov-3dgs_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

METRIC = "frames_per_second_foveated_1080p_6M"
UNIT = "frames/s"


# ----------------------------------------------------------------------------------------------------------------
class Readback:
    """Every frame's image is copied to pinned host memory; the copy of frame i runs on a copy stream while frame i+1
    renders (two host buffers).  Both bench arms use the same protocol.  `drain()` waits for the last copies, so all
    images of the timed frames are on the host when the clock stops."""

    def __init__(self, dev, shape, depth=4, dtype=torch.float32):
        self.host = [torch.empty(shape, dtype=dtype).pin_memory() for _ in range(depth)]
        self.done = [None] * depth
        self.stream = torch.cuda.Stream(dev)
        self.n = 0

    def push(self, img):
        slot = self.n % len(self.host)
        self.n += 1
        if self.done[slot] is not None:
            self.done[slot].synchronize()          # the host buffer is free again (a consumer would read it here)
        ready = torch.cuda.Event()
        ready.record()
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ready)
            self.host[slot].copy_(img, non_blocking=True)
            img.record_stream(self.stream)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.done[slot] = ev

    def drain(self):
        for ev in self.done:
            if ev is not None:
                ev.synchronize()


# ----------------------------------------------------------------------------------------------------------------
def sample_clocks(stop_evt, out, gpu_index):
    """nvidia-smi clocks line of /opt/skills/guides/B200_PROFILING.md, sampled during the timed region."""
    q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    try:
        p = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    except Exception:
        return
    def reader():
        for line in p.stdout:
            out.append(line.strip())
    t = threading.Thread(target=reader, daemon=True)
    t.start()
    stop_evt.wait()
    p.terminate()
    t.join(timeout=1.0)


def clocks_summary(lines):
    sm, mx, reasons = [], 0.0, set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    for ln in lines:
        f = [x.strip() for x in ln.split(",")]
        if len(f) < 7:
            continue
        try:
            sm.append(float(f[0])); mx = max(mx, float(f[1]))
        except ValueError:
            continue
        for nm, v in zip(names, f[3:7]):
            if v.lower().startswith("active"):
                reasons.add(nm)
    if not sm:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
    return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


class CameraUpload:
    """Per-frame host inputs of the e2e loops: viewmatrix (16) + projmatrix (16) + campos (3) + gaze (2) floats are packed
    into ONE pinned staging buffer and reach the device with ONE 148-byte copy per frame (four separate copies cost ~90 us
    of stream time per frame); the rasterizer gets views of the device buffer.  Both bench arms use it."""

    def __init__(self, cams, gazes, dev, depth=8):
        self.dev = dev
        self.meta = cams
        self.host = []
        for c in cams:
            per_gaze = []
            for g in gazes:
                buf = np.concatenate([np.asarray(c["viewmatrix"], np.float32).ravel(), np.asarray(c["projmatrix"], np.float32).ravel(),
                                      np.asarray(c["campos"], np.float32).ravel(), np.asarray(g, np.float32).ravel()])
                per_gaze.append(torch.from_numpy(buf).pin_memory())
            self.host.append(per_gaze)
        self.bytes = 37 * 4

    def upload(self, ci, gi):
        d = self.host[ci][gi].to(self.dev, non_blocking=True)
        c = dict(self.meta[ci])
        c["viewmatrix"] = d[0:16].view(4, 4)
        c["projmatrix"] = d[16:32].view(4, 4)
        c["campos"] = d[32:35]
        return c, d[35:37]


def bind_to_gpu_numa_node(local):
    """Multi-GPU runs: pin this rank to the CPU cores (NUMA node) next to its GPU BEFORE any pinned host buffer is allocated,
    so the per-frame 24.9 MB image read-backs of eight ranks land in local memory instead of crossing the socket link
    (8 x 830 frames/s x 24.9 MB = 165 GB/s of device-to-host traffic).  Best effort; both bench arms use it."""
    try:
        import pynvml
        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(local)
        try:
            h = pynvml.nvmlDeviceGetHandleByPciBusId(f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0")
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(local)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return sorted(os.sched_getaffinity(0))
    except Exception:
        return None


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


def to_dev(d, dev):
    return {k: (torch.from_numpy(v).to(dev) if isinstance(v, np.ndarray) else v) for k, v in d.items()}


class Workload:
    def __init__(self, size):
        from fovgs import synth
        self.synth = synth
        if size == "big":
            self.P, self.W, self.H = 6_000_000, 1920, 1080
            scene = synth.make_scene_bicycle(self.P, 1)
        elif size == "mid":
            self.P, self.W, self.H = 300_000, 800, 600
            scene = synth.make_scene_bicycle(self.P, 1, log_scale_mu=-3.6)
        elif size == "ref":
            # the size the reference's published FPS is quoted on (pruned 1.16 M-Gaussian model, Mip360 images_4
            # ~1237x822; pnum/ours-Q/bicycle.txt:1); fewer, larger Gaussians covering the same synthetic scene
            self.P, self.W, self.H = 1_161_358, 1237, 822
            scene = synth.make_scene_bicycle(self.P, 1, log_scale_mu=-3.8)
        else:
            raise SystemExit("unknown --size")
        self.scene = synth.add_foveation(scene)
        self.cams = synth.ring_cameras(30, self.W, self.H)
        self.gazes = synth.GAZES_9
        self.name = f"fov_{self.P // 1000}k_{self.W}x{self.H}_4level_moving_gaze_ring30"

    def frame(self, f):
        return self.cams[f % len(self.cams)], self.gazes[f % len(self.gazes)]


# ----------------------------------------------------------------------------------------------------------------
def run_ours(args, wl, rank, world, dev):
    from fovgs import ops
    import diff_gaussian_rasterization_fov_pcheck_obb as fovpkg

    sc = to_dev(wl.scene, dev)
    bg = torch.zeros(3, device=dev)
    cams_dev = [to_dev(c, dev) for c in wl.cams]
    gazes_dev = [torch.tensor(g, dtype=torch.float32, device=dev) for g in wl.gazes]

    def settings(c):
        return fovpkg.GaussianRasterizationSettings(c["image_height"], c["image_width"], c["tanfovx"], c["tanfovy"], bg, 1.0,
                                                    c["viewmatrix"], c["projmatrix"], wl.scene["sh_degree"], c["campos"], False, False)

    rs_dev = [settings(c) for c in cams_dev]

    def render(rs, gaze):
        r = fovpkg.GaussianRasterizer(raster_settings=rs)
        return r(means3D=sc["means3D"], means2D=None, opacities=sc["opacities4"], shs_rest=sc["shs_rest"], scales=sc["scales"],
                 rotations=sc["rotations"], shs_dcs=sc["shs_dcs"], highest_levels=sc["highest_levels"], gazeArray=gaze,
                 alpha=0.05, blending=True)

    if os.environ.get("FOVGS_NO_DIRECT_STATS", "0") == "1":      # A/B switch (tools): statistics by cudaMemcpyAsync
        from fovgs._lib import lib as _rawlib
        assert _rawlib().fovgs_set_option(4, 1) == 0
    frames = [rank + i * world for i in range(args.steps + args.warmup)]
    with torch.no_grad():
        # ---------------- value: device-resident inputs, pipelined ----------------
        ops.set_deferred_check(False)
        for f in frames[: args.warmup]:
            render(rs_dev[f % 30], gazes_dev[f % 9])
        torch.cuda.synchronize(dev)
        ops.set_deferred_check(True)
        # live timing of the dominant kernel only (two events per frame around the blend stage): the full stage table is taken
        # by a separate pass over the same frames below — seven events per frame cost the frame ~2.5 %
        ops.profile_enable(True, blend_only=True)
        stop_evt, clk = threading.Event(), []
        th = threading.Thread(target=sample_clocks, args=(stop_evt, clk, torch.cuda.current_device()), daemon=True)
        th.start()
        time.sleep(0.25)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        stats = []
        for f in frames[args.warmup:]:
            render(rs_dev[f % 30], gazes_dev[f % 9])
        e1.record()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        ms = e0.elapsed_time(e1)
        stop_evt.set()
        th.join(timeout=2.0)
        ops.check_pending(dev)
        blend_live = [x["blend"] for x in ops.profile_read_all()[-args.steps:]]
        # the same frames again with an event between every two stages: the per-stage table (not part of `value`)
        ops.profile_enable(True)
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for f in frames[args.warmup:]:
            render(rs_dev[f % 30], gazes_dev[f % 9])
        s1.record()
        torch.cuda.synchronize(dev)
        ops.check_pending(dev)
        ms_staged = s0.elapsed_time(s1)
        stage_frames = ops.profile_read_all()[-args.steps:]
        ops.profile_enable(False)

        # the same loop without the packed-model cache (FOVGS_MODEL_CACHE=0): the un-cached number beside the headline
        ops.set_model_cache(False)
        try:
            for f in frames[: 3]:
                render(rs_dev[f % 30], gazes_dev[f % 9])
            torch.cuda.synchronize(dev)
            ops.profile_enable(True, blend_only=True)   # same conditions as the headline loop
            u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            u0.record()
            for f in frames[args.warmup:]:
                render(rs_dev[f % 30], gazes_dev[f % 9])
            u1.record()
            torch.cuda.synchronize(dev)
            ops.check_pending(dev)
            ms_uncached = u0.elapsed_time(u1)
            ops.profile_enable(True)                    # and its stage table
            for f in frames[args.warmup:]:
                render(rs_dev[f % 30], gazes_dev[f % 9])
            torch.cuda.synchronize(dev)
            ops.check_pending(dev)
            unc = ops.profile_read_all()[-args.steps:]
            stage_uncached = {k: float(np.mean([x[k] for x in unc])) for k in unc[0]}
            ops.profile_enable(False)
        finally:
            ops.set_model_cache(True)
        ops.set_deferred_check(False)
        # re-prime the packed rows (two sightings of the same tensors) before the loops below
        for f in frames[: 2]:
            render(rs_dev[f % 30], gazes_dev[f % 9])

        # per-frame statistics (N, V, blending tiles) for the roofline byte model: re-render synchronously, untimed
        ops.set_full_stats(True)             # the blend stage's counters are only final at the end of a frame
        try:
            for f in frames[args.warmup: args.warmup + min(args.steps, 18)]:
                render(rs_dev[f % 30], gazes_dev[f % 9])
                stats.append(dict(ops.last_stats))
        finally:
            ops.set_full_stats(False)

        # ---------------- e2e: host inputs, image back to the host ----------------
        upl = CameraUpload(wl.cams, wl.gazes, dev)
        h2d = upl.bytes
        d2h = 3 * wl.H * wl.W * 4
        readback = Readback(dev, (3, wl.H, wl.W))

        def e2e_frame(f):
            cd, g = upl.upload(f % 30, f % 9)     # one 148-byte H2D copy from pinned memory
            img, _ = render(settings(cd), g)      # public API
            readback.push(img)

        def e2e_loop():
            for f in frames[: args.warmup]:
                e2e_frame(f)
            readback.drain()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            for f in frames[args.warmup:]:
                e2e_frame(f)
            readback.drain()
            torch.cuda.synchronize(dev)
            dt = time.perf_counter() - t0
            if world > 1:
                dist.barrier()
            return dt

        e2e_sync_s = e2e_loop()                 # default API behaviour: one blocking 64-byte read per frame (the headline)
        ops.set_deferred_check(True)            # opt-in pipelined serving mode (reported as an extra)
        try:
            e2e_s = e2e_loop()
            ops.check_pending(dev)
        finally:
            ops.set_deferred_check(False)
        del readback
        # extra: the same end-to-end loop (default blocking mode) with the opt-in 8-bit image of the blend epilogue — what the
        # reference's scripts store (torchvision.utils.save_image) — 6.2 MB per frame back to the host instead of 24.9 MB
        readback = Readback(dev, (3, wl.H, wl.W), dtype=torch.uint8)

        def render_u8(rs, gaze):
            r = fovpkg.GaussianRasterizer(raster_settings=rs)
            r.output_uint8 = True
            return r(means3D=sc["means3D"], means2D=None, opacities=sc["opacities4"], shs_rest=sc["shs_rest"], scales=sc["scales"],
                     rotations=sc["rotations"], shs_dcs=sc["shs_dcs"], highest_levels=sc["highest_levels"], gazeArray=gaze,
                     alpha=0.05, blending=True)

        render_f32, render = render, render_u8
        try:
            e2e_u8_s = e2e_loop()
        finally:
            render = render_f32
        del readback

    extra = extra_ours(args, wl, sc, cams_dev, bg, frames, dev) if not args.no_extra else None
    # kernels per frame: k_setup, k_tile_levels, k_tile_infos, k_pre, k_tile_scan, k_color_tma, k_scatter, k_lazy_blend x 2
    # (blending tiles, plain tiles)
    with torch.no_grad():
        last_image = render(rs_dev[frames[-1] % 30], gazes_dev[frames[-1] % 9])[0]

    def render_frame(f):
        with torch.no_grad():
            return render(rs_dev[f % 30], gazes_dev[f % 9])[0]

    return {"last_image": last_image, "last_frame": frames[-1], "render_frame": render_frame, "n_frames": args.steps,
            "ms": ms, "ms_staged": ms_staged, "blend_live": blend_live, "ms_uncached": ms_uncached, "stage_uncached": stage_uncached, "e2e_s": e2e_sync_s, "e2e_pipelined_s": e2e_s, "e2e_u8_s": e2e_u8_s, "stages": stage_frames, "stats": stats,
            "clocks": clocks_summary(clk), "h2d": h2d, "d2h": d2h, "launches_per_frame": FOV_LAUNCHES_PER_FRAME, "extra": extra}


FOV_LAUNCHES_PER_FRAME = 9
PS1_LAUNCHES_PER_FRAME = 6      # k_setup, k_pre, k_tile_scan, k_color_tma, k_scatter, k_lazy_blend
BWD_LAUNCHES = 3                # slab memset, k_bwd_render, k_bwd_preprocess


def _event_loop(step, frames, warmup, dev, marks):
    """Runs step(f, mark) for every frame; `mark()` records the next of `marks` CUDA events of the frame on the current stream.
    Returns per-phase milliseconds summed over the timed frames (list of marks - 1 floats) after the warm-up frames."""
    for f in frames[:warmup]:
        step(f, lambda: None)
    torch.cuda.synchronize(dev)
    evs = []
    for f in frames[warmup:]:
        row = []
        def mark(row=row):
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            row.append(e)
        step(f, mark)
        assert len(row) == marks
        evs.append(row)
    torch.cuda.synchronize(dev)
    total = evs[0][0].elapsed_time(evs[-1][-1])
    phases = [sum(r[i].elapsed_time(r[i + 1]) for r in evs) for i in range(marks - 1)]
    return total, phases


def extra_ours(args, wl, sc, cams_dev, bg, frames, dev):
    """BASELINE.json configs 2 and 5 through the drop-in packages (our library)."""
    import diff_gaussian_rasterization_pcheck_obb as obbpkg
    import diff_gaussian_rasterization_pcheck_obb_sum as sumpkg
    K = args.steps
    out = {}

    def settings(pkg, c):
        return pkg.GaussianRasterizationSettings(c["image_height"], c["image_width"], c["tanfovx"], c["tanfovy"], bg, 1.0,
                                                 c["viewmatrix"], c["projmatrix"], wl.scene["sh_degree"], c["campos"], False, False)

    rs_obb = [settings(obbpkg, c) for c in cams_dev]
    rs_sum = [settings(sumpkg, c) for c in cams_dev]
    # ---- config 2: PS=1 full-quality forward (render.py:47, cuda_type pcheck_obb)
    with torch.no_grad():
        def ps1(f, mark):
            mark()
            r = obbpkg.GaussianRasterizer(raster_settings=rs_obb[f % 30])
            r(means3D=sc["means3D"], means2D=None, opacities=sc["opacity"], shs=sc["shs"], scales=sc["scales"], rotations=sc["rotations"])
            mark()
        total, _ = _event_loop(ps1, frames, args.warmup, dev, 2)
    out["ps1_fwd"] = {"config": "BASELINE configs[1]: PS=1 forward, pcheck_obb, same 6M scene, ring cameras", "steps": K,
                      "ms_per_step": total / K, "frames_per_s": K / (total * 1e-3), "gpu_launches": PS1_LAUNCHES_PER_FRAME * K}
    # ---- config 5: training step = pcheck_obb_sum forward + backward, fixed dL/dpixel (eff_finetune.py:107,127)
    grad = torch.from_numpy(np.random.default_rng(3).standard_normal((3, wl.H, wl.W)).astype(np.float32)).to(dev)
    params = {k: sc[k].clone().requires_grad_(True) for k in ("means3D", "opacity", "shs", "scales", "rotations")}
    m2d = torch.zeros_like(params["means3D"], requires_grad=True)

    def train(f, mark):
        for p in list(params.values()) + [m2d]:
            p.grad = None
        mark()
        r = sumpkg.GaussianRasterizer(raster_settings=rs_sum[f % 30])
        color, radii, cnt, contrib = r(means3D=params["means3D"], means2D=m2d, opacities=params["opacity"], shs=params["shs"],
                                       scales=params["scales"], rotations=params["rotations"])
        mark()
        color.backward(grad)
        mark()
    total, (fwd, bwd) = _event_loop(train, frames, args.warmup, dev, 3)
    out["train_step"] = {"config": "BASELINE configs[4]: pcheck_obb_sum forward + backward, fixed dL/dpixel (seed 3), same 6M scene", "steps": K,
                         "ms_per_step": total / K, "fwd_ms": fwd / K, "bwd_ms": bwd / K, "steps_per_s": K / (total * 1e-3),
                         "gpu_launches": (PS1_LAUNCHES_PER_FRAME + BWD_LAUNCHES) * K}
    return out


def extra_reference(args, wl, sc, cams_dev, bg, frames, dev):
    """The same two configs on the unmodified reference binaries (oracle/_ref/ref_obb_C, ref_sum_C) through their pybind entries."""
    import ref_api
    K = args.steps
    out = {}
    obb, sm = ref_api.ref_module("ref_obb_C"), ref_api.ref_module("ref_sum_C")
    if obb is not None:
        with torch.no_grad():
            def ps1(f, mark):
                mark()
                ref_api.ps1_forward(obb, sc, cams_dev[f % 30], bg)
                mark()
            total, _ = _event_loop(ps1, frames, args.warmup, dev, 2)
        out["ps1_fwd"] = {"config": "BASELINE configs[1]: PS=1 forward, pcheck_obb, same 6M scene, ring cameras", "steps": K,
                          "ms_per_step": total / K, "frames_per_s": K / (total * 1e-3)}
    else:
        out["ps1_fwd"] = {"unavailable": "oracle/_ref/ref_obb_C not built"}
    if sm is not None:
        grad = torch.from_numpy(np.random.default_rng(3).standard_normal((3, wl.H, wl.W)).astype(np.float32)).to(dev)
        with torch.no_grad():
            def train(f, mark):
                c = cams_dev[f % 30]
                mark()
                res = ref_api.ps1_forward(sm, sc, c, bg)
                n, color, radii, geom, binning, img = res[:6]
                mark()
                ref_api.ps1_backward(sm, sc, c, radii, grad, geom, n, binning, img, bg)
                mark()
            total, (fwd, bwd) = _event_loop(train, frames, args.warmup, dev, 3)
        out["train_step"] = {"config": "BASELINE configs[4]: pcheck_obb_sum forward + backward, fixed dL/dpixel (seed 3), same 6M scene",
                             "steps": K, "ms_per_step": total / K, "fwd_ms": fwd / K, "bwd_ms": bwd / K, "steps_per_s": K / (total * 1e-3)}
    else:
        out["train_step"] = {"unavailable": "oracle/_ref/ref_sum_C not built"}
    return out


def run_reference(args, wl, rank, world, dev):
    """The unmodified reference extension (oracle/_ref/ref_fov_C.so) through its pybind entry point, same frames."""
    import ref_api
    mod = ref_api.ref_module("ref_fov_C")
    if mod is None:
        return None
    sc = to_dev(wl.scene, dev)
    bg = torch.zeros(3, device=dev)
    cams_dev = [to_dev(c, dev) for c in wl.cams]
    gazes_dev = [torch.tensor(g, dtype=torch.float32, device=dev) for g in wl.gazes]
    frames = [rank + i * world for i in range(args.steps + args.warmup)]

    def render(c, g):
        return ref_api.fov_forward(mod, sc, c, g, 0.05, True, bg)

    with torch.no_grad():
        for f in frames[: args.warmup]:
            render(cams_dev[f % 30], gazes_dev[f % 9])
        torch.cuda.synchronize(dev)
        stop_evt, clk = threading.Event(), []
        th = threading.Thread(target=sample_clocks, args=(stop_evt, clk, torch.cuda.current_device()), daemon=True)
        th.start()
        time.sleep(0.25)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for f in frames[args.warmup:]:
            render(cams_dev[f % 30], gazes_dev[f % 9])
        e1.record()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        ms = e0.elapsed_time(e1)
        stop_evt.set()
        th.join(timeout=2.0)
        # e2e with host camera / gaze and image read-back, same protocol as our arm
        upl = CameraUpload(wl.cams, wl.gazes, dev)
        readback = Readback(dev, (3, wl.H, wl.W))

        def e2e_frame(f):
            cd, g = upl.upload(f % 30, f % 9)
            out = render(cd, g)
            readback.push(out[1])

        for f in frames[: args.warmup]:
            e2e_frame(f)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for f in frames[args.warmup:]:
            e2e_frame(f)
        readback.drain()
        torch.cuda.synchronize(dev)
        e2e_s = time.perf_counter() - t0
        del readback
    extra = extra_reference(args, wl, sc, cams_dev, bg, frames, dev) if not args.no_extra else None
    with torch.no_grad():
        last_image = render(cams_dev[frames[-1] % 30], gazes_dev[frames[-1] % 9])[1]
    return {"last_image": last_image, "last_frame": frames[-1], "n_frames": args.steps,
            "ms": ms, "e2e_s": e2e_s, "clocks": clocks_summary(clk), "h2d": (16 + 16 + 3 + 2) * 4, "d2h": 3 * wl.H * wl.W * 4,
            "extra": extra}


def cpu_baseline(wl, frames=3):
    """A bounded sample (3 frames, about 12 s) of the same workload on the host cores with the CPU oracle (a port: the
    reference ships no CPU path), all host threads."""
    import oracle
    nthreads = oracle.lib().orc_num_threads()
    t0 = time.perf_counter()
    inst = []
    for f in range(frames):
        cam, gaze = wl.frame(f)
        o = oracle.forward_fov(wl.scene, cam, gaze, list_cap=1 << 27)
        inst.append(int(o["num_rendered"]))
    dt = time.perf_counter() - t0
    return {"value": frames / dt, "unit": UNIT, "cores": int(nthreads), "kind": "port",
            "sample": f"{frames} frames (cameras 0-{frames - 1}, gazes of the 9-gaze cycle) of {wl.name}; {inst} instances; {dt:.2f} s"}


FP32_INST_PER_PAIR = 9   # gauss_power (fovgs_math.cuh): 2 subtractions + 7 multiplies / FMAs per (pixel, Gaussian) evaluation


def roofline(res, wl, steps):
    """SURVEY.md §8(d)'s algorithmic bytes per foveated frame, split by stage (the ONE byte model; DESIGN.md §5 repeats it):

        B_fwd = P*(12+12+16+4) + P*4 + V*(180 + Lv*16) + N*24 + N*R + pixels*12
        preprocess = P*44 (xyz, scale, rot, highest_level) + P*4 (radii) + N*12 (first half of the binning floor: key+id written once)
        color      = V*(180 + Lv*16), Lv = 4 (this library colours all four levels of a visible Gaussian)
        scatter    = N*12 (second half of the binning floor: key+id read once)
        blend      = N*R + pixels*12, R = 36 B per instance on plain tiles, 56 B on blending tiles (share beta by tile count)

    P, V, N, beta are printed so the figures can be recomputed.  achieved = bytes / mean stage duration: the
    dominant stage (blend) from CUDA events the library records around it inside the timed loop, the others from a second pass
    over the same frames with an event between every two stages."""
    stages = res["stages"]
    names = list(stages[0].keys())
    mean = {k: float(np.mean([s[k] for s in stages])) for k in names}
    # the dominant stage's duration comes from the TIMED loop itself (events around that stage only); the other stages from
    # the stage-table pass over the same frames right after it
    live = res.get("blend_live")
    if live:
        mean["blend"] = float(np.mean(live))
    st = res["stats"]
    N = float(np.mean([s["num_rendered"] for s in st]))
    V = float(np.mean([s["num_visible"] for s in st]))
    Tb = float(np.mean([s["num_blend_tiles"] for s in st]))
    C = float(np.mean([s["blend_consumed"] for s in st]))
    pairs = float(np.mean([s["blend_block_pairs"] for s in st])) * 32.0     # (8x4 pixel block, instance) pairs x 32 pixels
    T = ((wl.W + 15) // 16) * ((wl.H + 15) // 16)
    P, pix = wl.P, wl.W * wl.H
    beta = Tb / T
    R = 36.0 * (1.0 - beta) + 56.0 * beta
    bytes_model = {
        "setup": 0.0,
        "preprocess": P * 44.0 + P * 4.0 + N * 12.0,
        "color": V * (180.0 + 4 * 16.0),
        "scatter": N * 12.0,
        "tile_sort": 0.0,                      # inference frames sort inside the blend kernel
        "blend": N * R + pix * 12.0,
    }
    peak, src = measured_peak_gbs()
    per_stage = {}
    for k in names:
        gbs = bytes_model[k] / (mean[k] * 1e-3) / 1e9 if mean[k] > 0 else 0.0
        per_stage[k] = {"bytes": bytes_model[k], "ms": mean[k], "gbs": gbs, "frac": gbs / peak}
    dom = max(mean, key=mean.get)
    frame_bytes = float(sum(bytes_model.values()))
    frame_ms = res["ms"] / steps
    traffic, traffic_note = None, None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            traffic, traffic_note = tj.get(dom), tj.get("note")
        except Exception:
            traffic = None
    # the blend is fp32-issue-bound, not HBM-bound: pixel-Gaussian falloff evaluations per second against the fp32 pipe
    clk = (res["clocks"].get("sm_mhz") or 1965.0) * 1e6
    fp32_peak = 148 * 128 * clk                               # fp32 instructions/s (an FMA is one instruction)
    blend_s = mean["blend"] * 1e-3
    alu = {"pairs_per_frame": pairs, "pairs_per_s": pairs / blend_s, "fp32_inst_per_pair": FP32_INST_PER_PAIR,
           "fp32_inst_per_s": pairs * FP32_INST_PER_PAIR / blend_s, "fp32_peak_inst_per_s": fp32_peak,
           "frac": pairs * FP32_INST_PER_PAIR / blend_s / fp32_peak,
           "note": "counts only the falloff exponent of pairs that pass the block footprint test; sort, staging, exp and compositing of hits come on top"}
    d = per_stage[dom]
    return {"bound": "hbm", "kernel": dom, "achieved": d["gbs"], "peak": peak, "peak_source": src, "unit": "GB/s",
            "frac": d["frac"], "traffic": traffic, "traffic_note": traffic_note, "algorithmic_bytes_per_launch": d["bytes"],
            "byte_model": "SURVEY.md 8(d), split by stage (bench.py roofline() docstring, DESIGN.md 5)",
            "timing": "blend: CUDA events around the blend stage inside the timed loop; other stages: the same frames re-run with an "
                      "event between every two stages (%.4f ms per frame in that pass)" % (res.get("ms_staged", 0.0) / steps),
            "stages": per_stage, "frame": {"bytes": frame_bytes, "ms": frame_ms, "gbs": frame_bytes / (frame_ms * 1e-3) / 1e9,
                                           "frac": frame_bytes / (frame_ms * 1e-3) / 1e9 / peak},
            "P": P, "N_mean": N, "V_mean": V, "Lv": 4, "composited_mean": C, "blend_tile_share": beta, "alu": alu}


def gather_frames(res, wl, world, rank, dev):
    """N > 1: the one communication step of the sharded forward path (north_star: NCCL only to gather timings / images).  Every
    rank contributes the image of its last timed frame and its timing row; rank 0 receives them with ONE NCCL gather each
    (fovgs/shard.py), device-timed (max over ranks is rank 0's view: it is the receiver).  Outside the timed region of `value`."""
    from fovgs import shard
    img = res["last_image"]
    table = shard.gather_timings(res["ms"], res["n_frames"], device=dev)
    torch.cuda.synchronize(dev)
    dist.barrier()
    times = []
    out = None
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = shard.gather_images(img)
        e1.record()
        torch.cuda.synchronize(dev)
        times.append(e0.elapsed_time(e1))
    dist.barrier()
    if rank != 0:
        return None
    nbytes = (world - 1) * img.numel() * 4
    ms = min(times[1:])
    info = {"collective": "nccl gather -> rank 0", "images": world, "bytes_received": nbytes, "ms": ms, "gbs": nbytes / (ms * 1e-3) / 1e9,
            "rank_ms": [float(x) for x in table[:, 0]], "rank_frames": [int(x) for x in table[:, 1]],
            "frames_per_s_from_table": shard.aggregate_fps(table)}
    check = res.get("render_frame")
    if check is not None:
        # rank 0 re-renders the frames the other ranks sent: a replicated model must give bit-identical images on every GPU
        last = res["last_frame"]
        info["images_equal_rank0_rerender"] = all(bool(torch.equal(out[r], check(last - rank + r))) for r in range(world))
    return info


def config_of(wl, args):
    """The workload description both arms print (identical keys and values, so the driver's same_config check holds)."""
    return {"workload": wl.name, "gaussians": wl.P, "width": wl.W, "height": wl.H, "levels": 4, "alpha": 0.05,
            "frames_per_rank": args.steps, "sharding": "frame (camera,gaze) round-robin, model replicated",
            "l2_policy": "inputs larger than L2 (model 1.7 GB + 0.6 GB of per-frame records vs 126 MB L2)",
            "model_cache": "on (library default): packed colour rows of the static model tensors, built once"}


def cpu_reference_line(wl, args):
    """The reference arm when the reference's CUDA extension cannot run: the CPU oracle port on a bounded sample."""
    cb = cpu_baseline(wl)
    return {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": 1, "warmup": 0, "ms_per_step": 1000.0 / cb["value"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_of(wl, args), "reference_class": "cpu: oracle port (no CUDA device or reference binary here)", "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=90)
    ap.add_argument("--warmup", type=int, default=9)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", default="big", choices=["big", "mid", "ref"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the PS=1 forward and training-step lines (configs 2 and 5)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        if args.impl == "reference":
            # no GPU: the reference's CUDA cannot run either; its CPU restatement (the oracle port) is the only arm left
            if rank == 0:
                print(json.dumps(cpu_reference_line(Workload(args.size), args)))
            return 0
        print(json.dumps({"impl": args.impl, "error": "no CUDA device: the product path has no CPU fallback"}))
        return 1
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    cpus = bind_to_gpu_numa_node(local) if world > 1 else None
    if world > 1 and os.environ.get("FOVGS_BENCH_VERBOSE"):
        print(f"rank {rank}: cuda:{local} bound to {len(cpus) if cpus else 'all'} cpus", file=sys.stderr)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner to STDOUT when the first communicator is created; the contract is ONE JSON line on
        # stdout, so file descriptor 1 points at stderr while the communicator comes up (init + one warm-up all-reduce)
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
            warm = torch.zeros(1, device=dev)
            dist.all_reduce(warm)
            torch.cuda.synchronize(dev)
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    wl = Workload(args.size)
    if args.impl == "reference":
        res = run_reference(args, wl, rank, world, dev)
        if res is None:
            # reference extension not built here: fall back to the CPU oracle port on rank 0 (bounded sample)
            if rank == 0:
                print(json.dumps(cpu_reference_line(wl, args)))
            return 0
    else:
        res = run_ours(args, wl, rank, world, dev)

    # max over ranks (device time and wall time)
    t = torch.tensor([res["ms"], res["e2e_s"], res.get("e2e_pipelined_s", 0.0), res.get("ms_uncached", 0.0), res.get("e2e_u8_s", 0.0)],
                     dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max, e2e_max, e2e_piped_max, ms_unc_max, e2e_u8_max = float(t[0]), float(t[1]), float(t[2]), float(t[3]), float(t[4])
    total_frames = args.steps * world
    value = total_frames / (ms_max * 1e-3)
    e2e_v = total_frames / e2e_max
    gathered = gather_frames(res, wl, world, rank, dev) if world > 1 else None

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": config_of(wl, args),
            "clocks": res["clocks"],
            "e2e": {"value": e2e_v, "unit": UNIT, "h2d_bytes_per_step": res["h2d"], "d2h_bytes_per_step": res["d2h"],
                    "mode": "default drop-in call: every call returns with its validated instance count (host waits for the event behind "
                            "the colour stage; the tile-scan kernel has stored the statistics in its pinned buffer by then)"},
        }
        if res.get("extra") is not None:
            line["extra"] = res["extra"]
        if gathered is not None:
            line["gather"] = gathered
        if args.impl == "reference":
            line["impl"] = "reference"
            line["e2e"]["mode"] = "default drop-in call: the reference's own blocking call (it synchronises with the host inside every frame)"
            line["reference_class"] = "gpu: the unmodified reference CUDA extension (oracle/_ref, built by oracle/build_ref.py) on the same B200"
            line["cpu_baseline"] = {"value": value, "unit": UNIT, "cores": 0, "kind": "reference",
                                    "sample": "reference CUDA rasterizer on the same B200 (reference_class)"}
            line["gpu_launches"] = 0
        else:
            line["gpu_launches"] = res["launches_per_frame"] * args.steps
            line["e2e"]["pipelined_value"] = total_frames / e2e_piped_max      # opt-in serving mode, not the headline
            if e2e_u8_max > 0:                                                  # opt-in 8-bit image (extra, not the headline)
                line["e2e"]["uint8_output"] = {"value": total_frames / e2e_u8_max, "d2h_bytes_per_step": res["d2h"] // 4,
                                               "note": "rasterizer.output_uint8 = True: the blend epilogue writes the image the "
                                                       "reference's scripts store (torchvision save_image quantisation)"}
            line["value_uncached"] = total_frames / (ms_unc_max * 1e-3)       # FOVGS_MODEL_CACHE=0
            line["stage_ms_uncached"] = res["stage_uncached"]
            line["roofline"] = roofline(res, wl, args.steps)
            if not args.no_cpu_baseline and world == 1:
                try:
                    line["cpu_baseline"] = cpu_baseline(wl)
                except Exception as ex:  # never lose the GPU numbers to a host-side problem
                    line["cpu_baseline"] = {"error": repr(ex)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
