"""CPU: the C-ABI library builds for sm_100a, loads without a GPU and exports every symbol include/bts_b200.h
declares; the ctypes mirror of the structs has the sizes the header implies.  No compute calls here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "bts_b200.h")


@pytest.fixture(scope="module")
def lib():
    from behindthescenes_b200 import build, _lib
    build.build()
    return _lib.load()


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bts_[a-z_0-9]+)\s*\(", src)))


def test_header_declares_the_expected_entry_points():
    names = _declared_functions()
    for n in ("bts_abi_version", "bts_last_error", "bts_fold_features", "bts_pack_images", "bts_render_fwd",
              "bts_query_field", "bts_last_kernel", "bts_render_bwd", "bts_render_views_fwd", "bts_pack_g16", "bts_g16_bytes"):
        assert n in names


def test_every_declared_symbol_is_exported(lib):
    from behindthescenes_b200 import _lib
    for n in _declared_functions():
        assert hasattr(lib, n), f"{n} declared in include/bts_b200.h but not exported"
        assert n in _lib.EXPORTS, f"{n} has no ctypes prototype"
    assert lib.bts_abi_version() == 11
    assert lib.bts_last_error() is not None


def test_struct_mirrors_match_header_layout():
    from behindthescenes_b200 import _lib
    # sizes under the LP64 C ABI: int32 fields, 8-byte pointers, natural alignment
    assert ctypes.sizeof(_lib.BtsScene) == 16 + 4 * 8 + 16 + 3 * 8 + 8 + 2 * 8   # + same_view_mask (padded), G16, g16_flags
    assert ctypes.sizeof(_lib.BtsMlp) == 16 + 2 * 8 + 4 * 8 * 8 + 2 * 8
    assert ctypes.sizeof(_lib.BtsRenderOut) == 10 * 8 + 8 + 8              # + scratch, scratch_ints
    assert ctypes.sizeof(_lib.BtsRayGen) == 2 * 8 + 4 * 4 + 2 * 4
    assert ctypes.sizeof(_lib.BtsRenderCfg) == 9 * 4 + 6 * 4 + 4 + 8 + 8 + 8 + 8  # 9 ints, 6 floats, pad, z_steps, step+band_mode, seed_dev, precision+pad
    assert lib_g16_bytes() == 3 * 192 * 81 * 2048        # three copies, ceil((640 + 6) / 8) blocks per row


def lib_g16_bytes():
    from behindthescenes_b200 import _lib
    return _lib.load().bts_g16_bytes(1, 192, 640, 64)


def test_argument_validation_without_gpu(lib):
    """Validation paths return before any CUDA call, so they can be exercised on the CPU box."""
    from behindthescenes_b200 import _lib
    assert lib.bts_fold_features(None, 1, 64, 4, 4, None, 103, None, 64, None, None, None, None) == -1
    assert b"null" in lib.bts_last_error()
    rc = lib.bts_render_fwd(None, None, None, None, 1, 1, None, 0, None, None)
    assert rc == -1
    with pytest.raises(ValueError):
        _lib.check(rc, "bts_render_fwd")


def test_built_for_sm100a():
    import subprocess
    from behindthescenes_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_sparse_fold_entry_points_validate_without_gpu(lib):
    """The sparse-fold predicate is pure host logic; the entry points reject null pointers before any CUDA call."""
    assert lib.bts_sparse_fold_supported(64, 192, 640, 64) == 1          # KITTI field: tensor-core fold kernels take it
    assert lib.bts_sparse_fold_supported(64, 370, 1226, 64) == 1         # full-resolution KITTI-Raw frame (H*W % 4 == 0)
    assert lib.bts_sparse_fold_supported(48, 16, 16, 64) == 0            # C outside {32, 64}: dense CUDA-core fold only
    assert lib.bts_sparse_fold_supported(64, 13, 21, 64) == 0            # H*W % 4 != 0
    assert lib.bts_sparse_fold_supported(64, 192, 640, 128) == 0         # the fold backward kernel stops at d_hidden 64
    assert lib.bts_mark_tiles(None, None, None, 1, 1, None, 0, None, None) == -1
    assert lib.bts_fold_features_sparse(None, 1, 64, 4, 4, None, 103, None, 64, None, None, None, None, None) == -1
    assert lib.bts_zero_tiles(None, 1, 4, 4, 64, None, None) == -1
    assert lib.bts_fold_features_bwd_sparse(None, None, 1, 64, 4, 4, None, 103, 64, None, None, None, None, None) == -1
