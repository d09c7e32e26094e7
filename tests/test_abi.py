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


def test_struct_sizes_match_the_library():
    """every ctypes struct of the binding has the size the C side compiled with (ndcn_sizeof)"""
    lib = _ffi.lib()
    for which, cls in enumerate([_ffi.RhsDesc, _ffi.SolveOpts, _ffi.SolveStats, _ffi.PeerConfig,
                                 _ffi.FeaturePeerConfig, _ffi.GatherRequest]):
        assert ctypes.sizeof(cls) == lib.ndcn_sizeof(which), cls.__name__
    assert lib.ndcn_sizeof(99) == -1


def test_integration_md_stub_structs_match_the_library():
    """INTEGRATION.md's binding stub declares the same struct layouts as include/ndcn_b200.h (a stale stub would
    make the library read past the caller's struct)."""
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = re.search(r"```python\n(import ctypes as C, torch.*?)```", text, flags=re.S).group(1)
    lib = _ffi.lib()
    ns = {}
    cwd = os.getcwd()
    os.chdir(ROOT)  # the stub loads the library by its in-tree relative path
    try:
        exec(compile(block, "INTEGRATION.md", "exec"), ns)  # its own asserts compare against ndcn_sizeof
    finally:
        os.chdir(cwd)
    assert ctypes.sizeof(ns["RhsDesc"]) == lib.ndcn_sizeof(0)
    assert ctypes.sizeof(ns["SolveOpts"]) == lib.ndcn_sizeof(1)
    assert ctypes.sizeof(ns["SolveStats"]) == lib.ndcn_sizeof(2)
    fields = [f[0] for f in ns["SolveOpts"]._fields_]
    assert fields == [f[0] for f in _ffi.SolveOpts._fields_]


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
