"""CPU: the C-ABI library loads and exports exactly what include/deco_b200.h declares; argument validation works
without a GPU (checks run before any launch)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "deco_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(deco_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    from deco_b200 import _lib
    names = declared_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/deco_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes table and header disagree"


def test_header_arity_matches_ctypes_table():
    from deco_b200 import _lib
    src = open(os.path.join(ROOT, "include", "deco_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    for name, (_, args) in _lib.SIGNATURES.items():
        m = re.search(r"\b" + name + r"\s*\((.*?)\)\s*;", src, flags=re.S)
        assert m, name
        params = [p for p in m.group(1).split(",") if p.strip() and p.strip() != "void"]
        assert len(params) == len(args), (name, len(params), len(args))


def test_host_only_entry_points(lib):
    assert lib.deco_abi_version() == 1
    assert lib.deco_decoder_blob_bytes(3) == 36512
    assert lib.deco_decoder_blob_bytes(4) > lib.deco_decoder_blob_bytes(3)


def test_argument_errors_do_not_need_a_gpu(lib):
    rc = lib.deco_cfg_step(None, None, 1, None, None, None, 1.0, 0.1, 1.0, 0, 0, 0, None, None, None, None, 8, None)
    assert rc == -1 and b"null" in lib.deco_last_error()
    rc = lib.deco_gemm_bf16(ctypes.c_void_p(16), 8, ctypes.c_void_p(16), 8, ctypes.c_void_p(16), 8, 4, 8, 7, 0,
                            None, None, 0, None, 0, 1, 0, None)
    assert rc == -1 and b"multiples of 8" in lib.deco_last_error()
    rc = lib.deco_qknorm_rope(ctypes.c_void_p(16), ctypes.c_void_p(16), ctypes.c_void_p(16), ctypes.c_void_p(16),
                              4, 2, 48, 4, 1e-6, None)
    assert rc == -2 and b"head_dim" in lib.deco_last_error()


def test_missing_library_fails_loudly(monkeypatch):
    from deco_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libdeco_b200.so")
    with pytest.raises(_lib.DecoLibraryError, match="no CPU or PyTorch fallback"):
        _lib.load()


def test_backward_entry_points_validate_arguments_without_a_gpu(lib):
    """Training-path entry points: host-only size queries and argument checks that run before any launch."""
    assert lib.deco_decoder_train_blob_floats(3) == 1152 + 3 * 5344 + 132
    assert lib.deco_decoder_bwd_blob_bytes(3) == 4 * (512 + 3 * 2560 + 96)
    p = ctypes.c_void_p(16)
    rc = lib.deco_gemm_bf16_tn(p, 8, p, 8, p, 8, 12, 8, 64, 0, 0, None)       # M not a multiple of 8
    assert rc == -1 and b"multiples of 8" in lib.deco_last_error()
    rc = lib.deco_pixel_decoder_bwd_tc(p, p, p, p, p, p, p, p, 1, 64, 64, 16, 32, 4, None)   # 4 res-blocks: not built
    assert rc == -1 and b"R <= 3" in lib.deco_last_error()
    rc = lib.deco_attention_bwd(p, 8, p, p, 8, p, 8, p, 8, p, 8, p, p, 8, p, p, 0, 1, 2, 16, 16, 48, 0.1, None)
    assert rc == -2 and b"head_dim" in lib.deco_last_error()
    rc = lib.deco_transpose_cast(p, 0, 8, p, 8, 4, 3, 4, None)               # odd column count
    assert rc == -1 and b"even" in lib.deco_last_error()


def test_gemm_tile_plan_is_balanced_and_complete():
    """Host-side walk of the GEMM kernels' static tile schedule (deco_gemm_tile_plan, no GPU): every tile is visited exactly
    once, and at the training shapes (8192 tokens, 74 CTA pairs) the 1152-wide GEMMs take 256-wide tiles whose ragged last
    column is dealt out so that every pair ends with exactly two tile-times of work (192-wide tiles: three rounds)."""
    import ctypes
    from deco_b200 import _lib
    lib = _lib.load()

    def plan(M, N, K, ctas=148):
        a, b, c = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        assert lib.deco_gemm_tile_plan(M, N, K, ctas, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)) == 0
        return a.value, b.value / 256.0, c.value

    for shape in [(8192, 1152, 1152), (8192, 3456, 1152), (8192, 6144, 1152), (8192, 1152, 6144), (131072, 8192, 1152),
                  (16384, 1152, 3072), (32, 1152, 256), (300, 688, 144), (77, 96, 64), (1024, 5472, 1024)]:
        bn, load, bad = plan(*shape)
        assert bn in (128, 192, 256) and bad == 0, (shape, bn, bad)
    assert plan(8192, 1152, 1152) == (256, 2.0, 0)           # 128 full + 32 half tiles on 74 pairs
    assert plan(8192, 1152, 3072) == (256, 2.0, 0)
    assert plan(8192, 1152, 6144) == (256, 3.0, 0)           # A = 100 MB: row-major order (ragged tiles not deferred)
    bn, load, _ = plan(8192, 3456, 1152)                     # 13.5 column tiles x 32 row tiles = 432 units on 74 pairs
    assert bn == 256 and load == 6.0
    assert lib.deco_gemm_tile_plan(0, 8, 8, 148, None, None, None) == -1
