"""The C-ABI library loads without a GPU and exports every symbol include/rqae_b200.h declares."""
import ctypes
import os
import re

from rqae_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "rqae_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rqae_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_list_agree():
    assert header_symbols() == sorted(_lib.SYMBOLS)


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in header_symbols():
        assert hasattr(lib, name), f"{name} missing from librqae_b200.so"


def test_host_only_entry_points():
    lib = _lib.load()
    assert b"sm_100a" in lib.rqae_version()
    assert lib.rqae_strerror(0) == b"ok" and lib.rqae_strerror(2).startswith(b"unsupported")
    # packed size model: header + biases + search table + (nq+1) stages of E*256*36 bytes
    n = lib.rqae_packed_bytes(1024, 2304, 4, 625)
    assert n >= 1025 * 9 * 256 * 36 and n < 1025 * 9 * 256 * 36 + (1 << 16)
    assert lib.rqae_packed_bytes(2048, 3584, 4, 625) >= 2049 * 14 * 256 * 36
    assert lib.rqae_packed_bytes(8, 2304, 8, 625) == 0      # codebook_dim != 4
    assert lib.rqae_packed_bytes(8, 5000, 4, 625) == 0      # dim beyond the compiled shapes
    # argument validation happens before any CUDA call
    assert lib.rqae_forward_f32(None, None, 1, 8, 8, 256, 4, 625, None, 4, None, 2, 8, None, None, None, None) == 1
    assert lib.rqae_decode_f32(None, None, 8, 8, 256, 4, 625, None, 2, 8, None, None, 4, None, None) == 1
