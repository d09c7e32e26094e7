"""CPU: the C-ABI library builds, loads, and exports every symbol include/ndcn_b200.h declares.
No compute calls (there is no GPU here)."""
import ctypes
import os
import re

from ndcn_b200 import _build, _ffi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "ndcn_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(ndcn_[a-z0-9_]+)\s*\(", src)
    return sorted(set(n for n in names if not n.endswith("_t")))


def test_library_builds_and_loads():
    path = _build.build()
    assert os.path.exists(path)
    lib = _ffi.lib()
    assert lib.ndcn_sm_arch() == 100
    assert b"sm_100a" in lib.ndcn_version()


def test_every_declared_symbol_is_exported():
    lib = ctypes.CDLL(_build.build())
    declared = _declared_functions()
    assert len(declared) >= 12
    for name in declared:
        assert hasattr(lib, name), "missing export: " + name


def test_prototype_table_matches_header():
    assert sorted(_ffi.PROTOTYPES) == _declared_functions()


def test_workspace_size_formula():
    lib = _ffi.lib()
    n, H = 1000, 256
    b = lib.ndcn_solver_workspace_bytes(n, n, H, _ffi.DOPRI5)
    assert b >= 11 * n * H * 4 + H * H * 4
    assert lib.ndcn_solver_workspace_bytes(n, n + 100, H, _ffi.DOPRI5) > b  # halo rows cost space


def test_argument_validation_without_gpu():
    lib = _ffi.lib()
    out = ctypes.c_void_p()
    assert lib.ndcn_graph_create(-1, 0, 0, None, None, None, ctypes.byref(out)) == _ffi.E_ARG
    assert lib.ndcn_graph_create(4, 2, 0, None, None, None, ctypes.byref(out)) == _ffi.E_ARG
    assert lib.ndcn_odeint_f32(None, None, None, 0, None, None, None, None) == _ffi.E_ARG


def test_peer_api_argument_validation_without_gpu():
    """multi-GPU peer-memory entry points reject null / malformed arguments before touching the device"""
    lib = _ffi.lib()
    assert lib.ndcn_solver_set_peers(None, None) == _ffi.E_ARG
    assert lib.ndcn_solver_set_feature_peers(None, None, None) == _ffi.E_ARG
    ptr = ctypes.c_void_p()
    assert lib.ndcn_peer_alloc(0, ctypes.byref(ptr), ctypes.create_string_buffer(64)) == _ffi.E_ARG
    assert lib.ndcn_peer_alloc(4096, None, ctypes.create_string_buffer(64)) == _ffi.E_ARG
    assert lib.ndcn_peer_open(None, ctypes.byref(ptr)) == _ffi.E_ARG
    assert lib.ndcn_peer_close(None) == _ffi.OK and lib.ndcn_peer_free(None) == _ffi.OK
    # struct layouts the library reads: sizes must match the C side (8-byte aligned fields)
    assert ctypes.sizeof(_ffi.PeerConfig) == 8 + 8 * 8 + 8 * 8
    assert ctypes.sizeof(_ffi.FeaturePeerConfig) == 8 + 3 * 8 * 8 + 9 * 8


def test_sass_has_bulk_copy_and_no_legacy_paths():
    """The W^T chunks move with cp.async.bulk (SASS UBLKCP); built for sm_100a only."""
    import shutil
    import subprocess

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        return
    out = subprocess.run([cuobjdump, "-lelf", _build.build()], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sass = subprocess.run([cuobjdump, "-sass", _build.build()], capture_output=True, text=True).stdout
    assert "UBLKCP" in sass and "k_stage_ndcn_gemm" in sass
