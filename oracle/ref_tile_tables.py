"""Test infrastructure: the reference's OWN tile-level kernels, run from the unmodified reference binary.

The foveated reference keeps its per-tile tables (tile_levels, gradients, tile_level_min, tile_blendings) in
function-static cudaMallocs (FOV/cuda_rasterizer/rasterizer_impl.cu:716-753) that no pybind entry exposes.  To pin
our `k_tile_levels` / `k_tile_infos` against the reference and not against ourselves, this module takes the two
kernels that fill those tables —
    compute_tile_levels_cuda        FOV/cuda_rasterizer/rasterizer_impl.cu:120-177
    compute_tile_level_infos_cuda   FOV/cuda_rasterizer/rasterizer_impl.cu:182-260
— as compiled machine code out of oracle/_ref/ref_fov_C/ref_fov_C.so (`cuobjdump -xelf`: the sm_100 cubin the reference
build produced, nothing recompiled, no reference source touched), loads it with the CUDA driver API and launches them with
the reference's own launch configuration (:755-775) on caller-owned buffers.

Only tests/ and tools/parity_gpu.py import this.  Nothing here is on the product path.
"""
import glob
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
K_LEVELS = b"_Z24compute_tile_levels_cudaiPf6float2iiif"
K_INFOS = b"_Z29compute_tile_level_infos_cudaiPfiiS_S_S_Pb"


def cubin_path(name="ref_fov_C"):
    return os.path.join(HERE, "_ref", name, name + "_tiles.cubin")


def extract_cubin(name="ref_fov_C", force=False):
    """Writes the cubin of the reference .so that holds the tile-level kernels next to it (oracle/_ref/<name>/<name>_tiles.cubin).
    Returns the path, or None when the reference .so (or cuobjdump) is missing."""
    out = cubin_path(name)
    if os.path.exists(out) and not force:
        return out
    sos = glob.glob(os.path.join(HERE, "_ref", name, name + "*.so"))
    if not sos:
        return None
    import shutil
    import tempfile
    tmp = tempfile.mkdtemp(prefix="fovgs_cubin_")
    try:
        subprocess.check_call(["cuobjdump", "-xelf", "all", sos[0]], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        for f in sorted(glob.glob(os.path.join(tmp, "*.cubin"))):
            if K_LEVELS in open(f, "rb").read():
                shutil.copyfile(f, out)
                return out
    except (OSError, subprocess.CalledProcessError):
        return None
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return None


_module = {}


def _functions(name="ref_fov_C"):
    if name in _module:
        return _module[name]
    p = cubin_path(name) if os.path.exists(cubin_path(name)) else extract_cubin(name)
    if p is None:
        return None
    import torch
    from cuda.bindings import driver as cu
    torch.cuda.init()
    torch.zeros(1, device="cuda")          # makes torch's primary context current on this thread
    data = open(p, "rb").read()
    err, mod = cu.cuModuleLoadData(data)
    if int(err) != 0:
        raise RuntimeError(f"cuModuleLoadData({p}) failed: {err}")
    fns = []
    for k in (K_LEVELS, K_INFOS):
        err, fn = cu.cuModuleGetFunction(mod, k)
        if int(err) != 0:
            raise RuntimeError(f"cuModuleGetFunction({k!r}) failed: {err}")
        fns.append(fn)
    _module[name] = (mod, fns[0], fns[1])
    return _module[name]


def available(name="ref_fov_C"):
    return os.path.exists(cubin_path(name)) or extract_cubin(name) is not None


def reference_tile_tables(W, H, gaze, alpha):
    """Runs the reference binary's two tile kernels.  gaze: (x, y) floats.  Returns dict of numpy arrays of T = tiles elements:
    tile_level, grad_y, grad_x, tile_min (f32), blending (bool), in the reference's tile order (row-major, 16x16 tiles)."""
    import ctypes as C
    import torch
    from cuda.bindings import driver as cu
    f = _functions()
    if f is None:
        raise RuntimeError("reference tile kernels unavailable (oracle/_ref/ref_fov_C missing)")
    _, k_levels, k_infos = f
    gx, gy = (W + 15) // 16, (H + 15) // 16
    T = gx * gy
    dev = torch.device("cuda", torch.cuda.current_device())
    lvl = torch.full((T,), -7.0, dtype=torch.float32, device=dev)
    g_y, g_x, mn = torch.zeros_like(lvl), torch.zeros_like(lvl), torch.zeros_like(lvl)
    bl = torch.zeros((T,), dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream

    def launch(fn, args):
        # args: list of ctypes scalars / structs, passed by address (kernelParams)
        ptrs = (C.c_void_p * len(args))(*[C.addressof(a) for a in args])
        err, = cu.cuLaunchKernel(fn, (T + 255) // 256, 1, 1, 256, 1, 1, 0, stream, C.addressof(ptrs), 0)
        if int(err) != 0:
            raise RuntimeError(f"cuLaunchKernel failed: {err}")

    class Float2(C.Structure):
        _fields_ = [("x", C.c_float), ("y", C.c_float)]

    launch(k_levels, [C.c_int(T), C.c_void_p(lvl.data_ptr()), Float2(float(gaze[0]), float(gaze[1])), C.c_int(W), C.c_int(H),
                      C.c_int(gx), C.c_float(float(alpha))])
    launch(k_infos, [C.c_int(T), C.c_void_p(lvl.data_ptr()), C.c_int(gx), C.c_int(gy), C.c_void_p(g_y.data_ptr()),
                     C.c_void_p(g_x.data_ptr()), C.c_void_p(mn.data_ptr()), C.c_void_p(bl.data_ptr())])
    torch.cuda.synchronize(dev)
    return {"tile_level": lvl.cpu().numpy(), "grad_y": g_y.cpu().numpy(), "grad_x": g_x.cpu().numpy(),
            "tile_min": mn.cpu().numpy(), "blending": bl.cpu().numpy().astype(np.bool_)}


if __name__ == "__main__":
    print(extract_cubin(force=True))
