"""ctypes binding of the C ABI in ``include/bts_b200.h`` (libbts_b200.so).

There is no CPU implementation behind this module: if the CUDA library is missing or fails to load, importing the
render path raises -- by design (BASELINE.json north_star: "no CPU fallback").
"""
import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BTS_B200_LIB") or os.path.join(_PKG, "lib", "libbts_b200.so")   # override: kernel experiments

BTS_MAX_BLOCKS = 8
ABI_VERSION = 11
_fp = C.POINTER(C.c_float)


class BtsScene(C.Structure):
    _fields_ = [("n", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("Hd", C.c_int32),
                ("G", C.c_void_p), ("g_empty", C.c_void_p), ("K_f", C.c_void_p), ("w2c_f", C.c_void_p),
                ("Vc", C.c_int32), ("Cc", C.c_int32), ("Hc", C.c_int32), ("Wc", C.c_int32),
                ("img4", C.c_void_p), ("K_c", C.c_void_p), ("w2c_c", C.c_void_p), ("same_view_mask", C.c_int32),
                ("G16", C.c_void_p), ("g16_flags", C.c_void_p)]


class BtsMlp(C.Structure):
    _fields_ = [("C", C.c_int32), ("P", C.c_int32), ("Hd", C.c_int32), ("n_blocks", C.c_int32),
                ("w_in", C.c_void_p), ("b_in", C.c_void_p),
                ("fc0_w", C.c_void_p * BTS_MAX_BLOCKS), ("fc0_b", C.c_void_p * BTS_MAX_BLOCKS),
                ("fc1_w", C.c_void_p * BTS_MAX_BLOCKS), ("fc1_b", C.c_void_p * BTS_MAX_BLOCKS),
                ("w_out", C.c_void_p), ("b_out", C.c_void_p)]


class BtsRenderCfg(C.Structure):
    _fields_ = [("n_coarse", C.c_int32), ("lindisp", C.c_int32), ("hard_alpha_cap", C.c_int32),
                ("white_bkgd", C.c_int32), ("inv_z", C.c_int32), ("code_mode", C.c_int32),
                ("learn_empty", C.c_int32), ("num_freqs", C.c_int32), ("include_input", C.c_int32),
                ("freq_factor", C.c_float), ("z_near", C.c_float), ("z_far", C.c_float),
                ("inv_z_far", C.c_float), ("inv_z_range", C.c_float), ("z_range", C.c_float),
                ("z_steps", C.c_void_p), ("step", C.c_float), ("band_mode", C.c_int32), ("seed_dev", C.c_void_p),
                ("precision", C.c_int32)]


class BtsRenderOut(C.Structure):
    _fields_ = [("rgb", C.c_void_p), ("depth", C.c_void_p), ("invalid", C.c_void_p), ("weights", C.c_void_p),
                ("alphas", C.c_void_p), ("z_samps", C.c_void_p), ("rgb_samps", C.c_void_p),
                ("invalid_mass", C.c_void_p), ("sigma", C.c_void_p), ("depth_z", C.c_void_p),
                ("scratch", C.c_void_p), ("scratch_ints", C.c_int64)]


def band_scratch_ints(rays: int) -> int:
    """BTS_BAND_SCRATCH_INTS of include/bts_b200.h."""
    return 4 + (rays + 7) // 8


class BtsRayGen(C.Structure):
    _fields_ = [("poses_c2w", C.c_void_p), ("K", C.c_void_p), ("v", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
                ("norm_dir", C.c_int32), ("z_near", C.c_float), ("z_far", C.c_float)]


class BtsRenderGrad(C.Structure):
    _fields_ = [("rgb", C.c_void_p), ("depth", C.c_void_p), ("weights", C.c_void_p), ("alphas", C.c_void_p)]


class BtsParamGrad(C.Structure):
    _fields_ = [("d_G", C.c_void_p), ("d_g_empty", C.c_void_p), ("d_w_pe", C.c_void_p), ("d_w_out", C.c_void_p),
                ("d_b_out", C.c_void_p), ("d_fc0_w", C.c_void_p), ("d_fc0_b", C.c_void_p), ("d_fc1_w", C.c_void_p),
                ("d_fc1_b", C.c_void_p)]


EXPORTS = {
    "bts_abi_version": (C.c_int, []),
    "bts_last_error": (C.c_char_p, []),
    "bts_last_kernel": (C.c_char_p, []),
    "bts_last_bwd_kernel": (C.c_char_p, []),
    "bts_debug_tc_timeouts": (C.c_int, []),
    "bts_fold_features": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32,
                                    C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "bts_pack_g16": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "bts_g16_bytes": (C.c_int64, [C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "bts_pack_images": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "bts_render_fwd": (C.c_int, [C.POINTER(BtsScene), C.POINTER(BtsMlp), C.POINTER(BtsRenderCfg), C.c_void_p,
                                 C.c_int64, C.c_int64, C.c_void_p, C.c_uint64, C.POINTER(BtsRenderOut), C.c_void_p]),
    "bts_render_views_fwd": (C.c_int, [C.POINTER(BtsScene), C.POINTER(BtsMlp), C.POINTER(BtsRenderCfg), C.POINTER(BtsRayGen),
                                       C.c_int64, C.c_void_p, C.c_uint64, C.POINTER(BtsRenderOut), C.c_void_p]),
    "bts_render_bwd": (C.c_int, [C.POINTER(BtsScene), C.POINTER(BtsMlp), C.POINTER(BtsRenderCfg), C.c_void_p,
                                 C.c_int64, C.c_int64, C.c_void_p, C.c_uint64, C.c_void_p, C.POINTER(BtsRenderGrad),
                                 C.POINTER(BtsParamGrad), C.c_void_p]),
    "bts_query_field": (C.c_int, [C.POINTER(BtsScene), C.POINTER(BtsMlp), C.POINTER(BtsRenderCfg), C.c_void_p,
                                  C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "bts_fold_features_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                        C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "bts_sparse_fold_supported": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "bts_mark_tiles": (C.c_int, [C.POINTER(BtsScene), C.POINTER(BtsRenderCfg), C.c_void_p, C.c_int64, C.c_int64, C.c_void_p,
                                 C.c_uint64, C.c_void_p, C.c_void_p]),
    "bts_fold_features_sparse": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32,
                                           C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "bts_zero_tiles": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "bts_fold_features_bwd_sparse": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                               C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "bts_patch_smooth_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int64, C.c_int32,
                                       C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "bts_patch_smooth_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int64, C.c_int32,
                                       C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "bts_patch_loss_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int64, C.c_int32,
                                     C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "bts_patch_loss_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int64, C.c_int32,
                                     C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
}

_lib = None


class BtsCudaError(RuntimeError):
    pass


def load():
    """Load libbts_b200.so (building nothing: see ``behindthescenes_b200.build``).  Raises if unavailable."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise BtsCudaError(f"{LIB_PATH} not found: build it with `python -m behindthescenes_b200.build` "
                           "(there is no CPU fallback for the render path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in EXPORTS.items():
        fn = getattr(lib, name)      # AttributeError if a declared symbol is missing
        fn.restype = res
        fn.argtypes = args
    v = lib.bts_abi_version()
    if v != ABI_VERSION:
        raise BtsCudaError(f"libbts_b200.so ABI {v} != binding ABI {ABI_VERSION}")
    _lib = lib
    return lib


class _NoGuard:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


_NO_GUARD = _NoGuard()


def on_device(dev):
    """Context in which ``dev`` is the CURRENT CUDA device: the library launches on whatever device is current
    (cudaGetDevice / cudaFuncSetAttribute / kernel<<<>>>), while the stream it is given belongs to the tensors' device --
    a module living on cuda:1 with cuda:0 current (``model.to('cuda:1')`` without ``set_device``, or the autograd thread)
    would otherwise launch on the wrong device."""
    import torch
    if dev.type != "cuda" or dev.index is None or torch.cuda.current_device() == dev.index:
        return _NO_GUARD
    return torch.cuda.device(dev)


def check(rc: int, what: str):
    if rc == 0:
        return
    msg = load().bts_last_error().decode(errors="replace")
    if rc == -2:
        raise NotImplementedError(f"{what}: {msg}")
    if rc == -1:
        raise ValueError(f"{what}: {msg}")
    raise BtsCudaError(f"{what}: {msg}")
