"""The C-ABI shared library loads and exports every symbol include/fovgs.h declares (no compute calls: no GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "fovgs.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fovgs_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_expected_entry_points():
    syms = _header_symbols()
    for s in ("fovgs_forward_fov", "fovgs_forward_ps1", "fovgs_backward_ps1", "fovgs_mark_visible", "fovgs_workspace_bytes",
              "fovgs_last_error"):
        assert s in syms


def test_library_exports_every_declared_symbol():
    from fovgs import _lib
    L = _lib.lib()
    for s in _header_symbols():
        assert hasattr(L, s), f"libfovgs.so does not export {s}"
    assert sorted(_lib.EXPORTS) == _header_symbols()
    assert L.fovgs_version() == _lib.FOVGS_VERSION == 202


def test_python_constants_mirror_the_header():
    """Enumerators and option numbers of include/fovgs.h against their mirrors in fovgs/_lib.py (a drifted constant would
    select another variant silently: the PS=1 mode picks OBB test / falloff cut / statistics)."""
    import re
    from fovgs import _lib
    text = open(os.path.join(ROOT, "include", "fovgs.h")).read()
    enums = dict((m.group(1), int(m.group(2))) for m in re.finditer(r"\b(FOVGS_PS1_[A-Z]+)\s*=\s*(\d+)", text))
    assert enums == {"FOVGS_PS1_OBB": 0, "FOVGS_PS1_SUM": 1, "FOVGS_PS1_MAX": 2, "FOVGS_PS1_LWMC": 3, "FOVGS_PS1_VANILLA": 4}
    for name, value in enums.items():
        assert getattr(_lib, name) == value, name
    opts = dict((m.group(1), int(m.group(2))) for m in re.finditer(r"#define\s+(FOVGS_OPT_[A-Z_]+)\s+(\d+)", text))
    assert opts == {"FOVGS_OPT_FULL_SORT": 1, "FOVGS_OPT_NO_TMA": 2, "FOVGS_OPT_NO_PDL": 3, "FOVGS_OPT_NO_DIRECT_STATS": 4}
    assert int(re.search(r"#define\s+FOVGS_VERSION\s+(\d+)", text).group(1)) == _lib.FOVGS_VERSION
    # unknown modes / options are rejected before any CUDA call
    L = _lib.lib()
    assert L.fovgs_set_option(99, 1) != 0 and b"unknown option" in L.fovgs_last_error()


def test_workspace_bytes_is_monotone_and_mode_dependent():
    from fovgs import _lib
    L = _lib.lib()
    a = L.fovgs_workspace_bytes(10000, 256, 256, 1 << 20, 0, 0)
    b = L.fovgs_workspace_bytes(10000, 256, 256, 1 << 21, 0, 0)
    c = L.fovgs_workspace_bytes(20000, 256, 256, 1 << 20, 0, 0)
    f = L.fovgs_workspace_bytes(10000, 256, 256, 1 << 20, 1, 0)
    s = L.fovgs_workspace_bytes(10000, 256, 256, 1 << 20, 0, 1)
    assert 0 < a < b and a < c and f > a and s > a
    assert L.fovgs_workspace_bytes(-1, 256, 256, 10, 0, 0) == 0
    assert L.fovgs_workspace_bytes(10, 0, 256, 10, 0, 0) == 0


def test_argument_errors_do_not_need_a_gpu():
    """Null / malformed arguments are rejected before any CUDA call, with a message (reference: AT_ERROR)."""
    from fovgs import _lib
    L = _lib.lib()
    assert L.fovgs_forward_fov(None, None) == -1
    assert b"null args" in L.fovgs_last_error()
    a = _lib.new_args(_lib.FovFwdArgs)
    a.cam.image_width = 0
    a.cam.image_height = 16
    assert L.fovgs_forward_fov(ctypes.byref(a), None) == -1
    assert b"image size" in L.fovgs_last_error()
    p = _lib.new_args(_lib.Ps1FwdArgs)
    p.cam.image_width = 16
    p.cam.image_height = 16
    assert L.fovgs_forward_ps1(ctypes.byref(p), None) == -1  # camera pointers null
    assert L.fovgs_mark_visible(5, None, None, None, None, None) == -1
    assert L.fovgs_mark_visible(0, None, None, None, None, None) == 0


def test_missing_library_fails_loudly(monkeypatch):
    from fovgs import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libfovgs.so")
    with pytest.raises(RuntimeError, match="no CPU/PyTorch fallback|not found"):
        _lib.lib()


def test_step_entry_points_validate_before_any_cuda_call():
    from fovgs import _lib
    L = _lib.lib()
    g = _lib.AdamGroup()
    g.n, g.step = 4, 0                                                    # step must be >= 1, pointers null
    assert L.fovgs_adam_step((_lib.AdamGroup * 1)(g), 1, None) == -1
    assert b"group 0" in L.fovgs_last_error()
    assert L.fovgs_adam_step(None, 9, None) == -1
    assert L.fovgs_adam_step(None, 0, None) == 0
    assert L.fovgs_activate_forward(0, None, None, None, None, None, None, None) == 0
    assert L.fovgs_activate_forward(4, None, None, None, 16, None, None, None) == -1    # output without its input
    assert L.fovgs_activate_backward(-1, None, None, None, None, None, None, None, None, None, None) == -1


def test_args_header_is_checked_before_anything_else():
    """A caller built against another header (wrong struct_size / abi_version) is rejected, whatever else the struct holds."""
    from fovgs import _lib
    L = _lib.lib()
    for cls, fn in ((_lib.FovFwdArgs, L.fovgs_forward_fov), (_lib.SmfrFwdArgs, L.fovgs_forward_smfr), (_lib.MmfrFwdArgs, L.fovgs_forward_mmfr),
                    (_lib.Ps1FwdArgs, L.fovgs_forward_ps1), (_lib.Ps1BwdArgs, L.fovgs_backward_ps1)):
        a = cls()                                   # zeroed: struct_size 0
        assert fn(ctypes.byref(a), None) == -1 and b"args header mismatch" in L.fovgs_last_error()
        a = _lib.new_args(cls)
        a.struct_size -= 8                          # the round-1 INTEGRATION.md stub: one trailing pointer short
        assert fn(ctypes.byref(a), None) == -1 and b"args header mismatch" in L.fovgs_last_error()
        a = _lib.new_args(cls)
        a.abi_version = 102
        assert fn(ctypes.byref(a), None) == -1 and b"args header mismatch" in L.fovgs_last_error()


def test_struct_mirrors_match_the_library():
    from fovgs import _lib
    L = _lib.lib()
    for sid, cls in _lib.STRUCT_IDS.items():
        assert L.fovgs_struct_size(sid) == ctypes.sizeof(cls), cls.__name__
    assert L.fovgs_struct_size(99) == 0
    # field order / names of the Python mirrors against the header text
    src = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "fovgs.h")).read(), flags=re.S)
    for cname, cls in (("fovgs_fov_fwd_args", _lib.FovFwdArgs), ("fovgs_smfr_fwd_args", _lib.SmfrFwdArgs), ("fovgs_mmfr_fwd_args", _lib.MmfrFwdArgs),
                       ("fovgs_ps1_fwd_args", _lib.Ps1FwdArgs), ("fovgs_ps1_bwd_args", _lib.Ps1BwdArgs), ("fovgs_camera", _lib.Camera)):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (cname, cname), src, re.S).group(1)
        body = body.replace("FOVGS_ARGS_HEADER", "uint32_t struct_size; uint32_t abi_version")
        names = [re.search(r"(\w+)\s*$", d.strip()).group(1) for d in body.split(";") if d.strip()]
        assert names == [f[0] for f in cls._fields_], cname


def test_integration_doc_quotes_the_binding_stub_and_its_structs_match():
    """INTEGRATION.md section 2 is examples/binding_stub.py verbatim; the stub's struct mirrors have the library's sizes
    (round 1's stub lacked `packed_color_rows`: the library would have read 8 bytes past the caller's struct)."""
    from fovgs import _lib
    L = _lib.lib()
    stub = open(os.path.join(ROOT, "examples", "binding_stub.py")).read()
    quoted = stub[stub.index("# --- binding stub (begin)\n") + len("# --- binding stub (begin)\n"):stub.index("# --- binding stub (end)")]
    assert quoted in open(os.path.join(ROOT, "INTEGRATION.md")).read()
    ns = {}
    classes = re.findall(r"(class (fovgs_\w+)\(C\.Structure\):.*?\n)(?=\n)", quoted, re.S)
    exec("import ctypes as C\n" + "\n".join(c[0] for c in classes), ns)
    ids = {"fovgs_camera": 0, "fovgs_fov_fwd_args": 2, "fovgs_adam_group": 7}
    assert {c[1] for c in classes} == set(ids)
    for name, sid in ids.items():
        assert ctypes.sizeof(ns[name]) == L.fovgs_struct_size(sid), name
    assert [f[0] for f in ns["fovgs_fov_fwd_args"]._fields_] == [f[0] for f in _lib.FovFwdArgs._fields_]
