"""CPU: the C-ABI library loads and exports exactly the symbols include/destr_b200.h declares
(no compute calls -- there is no GPU here), and the ctypes table mirrors the header."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "destr_b200.h")
LIB = os.path.join(ROOT, "object_detection_destr_b200", "libdestr_b200.so")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(?:int|int64_t|const char\*)\s+(destr_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = [a.strip() for a in m.group(2).split(",") if a.strip() and a.strip() != "void"]
        out[m.group(1)] = args
    return out


def test_library_exports_every_declared_symbol():
    assert os.path.exists(LIB), "run __graft_entry__.build() first"
    lib = ctypes.CDLL(LIB)
    funcs = header_functions()
    assert len(funcs) >= 20
    for name in funcs:
        assert hasattr(lib, name), f"{name} declared in destr_b200.h but not exported by libdestr_b200.so"
    lib.destr_version.restype = ctypes.c_int
    assert lib.destr_version() >= 100
    lib.destr_last_error.restype = ctypes.c_char_p
    assert isinstance(lib.destr_last_error(), bytes)


def test_ctypes_table_matches_header_arity():
    from object_detection_destr_b200 import _lib
    funcs = header_functions()
    for name, argtypes in _lib.SIGNATURES.items():
        assert name in funcs, f"{name} bound in _lib.py but not declared in destr_b200.h"
        assert len(argtypes) == len(funcs[name]), f"{name}: {len(argtypes)} ctypes args vs {len(funcs[name])} in header"
    bound = set(_lib.SIGNATURES) | {"destr_last_error", "destr_split_cross_attn_ws_floats",
                                       "destr_enc_attn_bwd_stats_floats"}
    assert set(funcs) <= bound, f"declared but not bound: {set(funcs) - bound}"


def test_bad_arguments_fail_loudly_without_gpu():
    """argument validation happens before any launch, so it is testable on CPU."""
    from object_detection_destr_b200 import _lib
    rc = _lib.lib.destr_add_layernorm_fwd(None, 256, None, 0, None, None, None, 256, None, None, 4, 256, None, 0, 0, None)
    assert rc != 0 and b"null pointer" in _lib.lib.destr_last_error()
    rc = _lib.lib.destr_enc_attn_fwd(1, 1, 1, 256, 256, 256, 1, 4, 1, None, 1, 128, 9, 0.1, None, 0, 0, None)
    assert rc != 0 and b"shape" in _lib.lib.destr_last_error()


def test_ops_refuse_cpu_tensors():
    import pytest
    import torch
    from object_detection_destr_b200 import ops
    with pytest.raises(RuntimeError, match="no CPU path"):
        ops.add_layernorm(torch.zeros(4, 256, dtype=torch.bfloat16), None, torch.ones(256), torch.zeros(256))
