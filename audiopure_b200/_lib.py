"""ctypes binding of libaudiopure_b200.so (include/audiopure_b200.h).

There is no fallback: if the library is missing or an entry point is absent, importing a symbol from
here raises.  Nothing in this package computes the hot path any other way.
"""

import ctypes
import os

from . import build as _build

c_float_p = ctypes.POINTER(ctypes.c_float)
c_i32_p = ctypes.POINTER(ctypes.c_int32)
c_i64_p = ctypes.POINTER(ctypes.c_int64)

AP_ABI_VERSION = 6
AP_FLAG_TF32 = 1
AP_COMM_ID_BYTES = 128


class ApConfig(ctypes.Structure):
    _fields_ = [
        ("num_res_layers", ctypes.c_int32),
        ("dilation_cycle", ctypes.c_int32),
        ("T", ctypes.c_int32),
        ("max_chunk", ctypes.c_int32),
        ("flags", ctypes.c_uint32),
        ("alpha", c_float_p),
        ("alpha_bar", c_float_p),
        ("sigma", c_float_p),
        ("sde_beta", c_float_p),
        ("sde_alphas_cumprod", c_float_p),
    ]


class ApWeights(ctypes.Structure):
    _fields_ = [
        ("w1", ctypes.c_void_p),
        ("b1", ctypes.c_void_p),
        ("w2", ctypes.c_void_p),
        ("c2", ctypes.c_void_p),
        ("part0", ctypes.c_void_p),
        ("w0", ctypes.c_void_p),
        ("b0", ctypes.c_void_p),
        ("ws", ctypes.c_void_p),
        ("bs", ctypes.c_void_p),
        ("wf", ctypes.c_void_p),
        ("bf", ctypes.c_void_p),
        ("wo", ctypes.c_void_p),
        ("bo", ctypes.c_float),
    ]


class ApMelTables(ctypes.Structure):
    _fields_ = [
        ("twiddles", ctypes.c_void_p),
        ("fb_start", ctypes.c_void_p),
        ("fb_len", ctypes.c_void_p),
        ("fb_off", ctypes.c_void_p),
        ("fb_w", ctypes.c_void_p),
        ("n_mels", ctypes.c_int32),
        ("fb_nnz", ctypes.c_int32),
    ]


# name -> (restype, argtypes); must list every function include/audiopure_b200.h declares
_VP, _I, _F, _U64, _U32, _I64, _SZ = (ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_uint64,
                                       ctypes.c_uint32, ctypes.c_int64, ctypes.c_size_t)
SIGNATURES = {
    "ap_last_error": (ctypes.c_char_p, []),
    "ap_abi_version": (_I, []),
    "ap_create": (_I, [ctypes.POINTER(ApConfig), ctypes.POINTER(ApWeights), ctypes.POINTER(_VP)]),
    "ap_destroy": (None, [_VP]),
    "ap_workspace_bytes": (_SZ, [_VP, _I, _I]),
    "ap_eps": (_I, [_VP, _VP, _I, _I, _I, _VP, _VP, _SZ, _VP]),
    "ap_step": (_I, [_VP, _VP, _VP, _I, _I, _I, _F, _F, _F, _VP, _U64, _U32, _I64, _VP, _SZ, _VP]),
    "ap_ddpm_purify": (_I, [_VP, _VP, _VP, _I, _I, _I, _VP, _U64, _I64, _VP, _SZ, _VP]),
    "ap_sde_purify": (_I, [_VP, _VP, _VP, _I, _I, _I, _VP, _U64, _I64, _VP, _SZ, _VP]),
    "ap_one_shot": (_I, [_VP, _VP, _VP, _I, _I, _I, _VP, _SZ, _VP]),
    "ap_logmel": (_I, [_VP, _I, _I, _VP, ctypes.POINTER(ApMelTables), _VP]),
    "ap_logmel_backward": (_I, [_VP, _I, _I, _VP, _VP, ctypes.POINTER(ApMelTables), _VP]),
    "ap_smooth_inputs": (_I, [_VP, _I, _I, _F, _F, _VP, _U64, _U32, _I64, _VP, _VP]),
    "ap_vote_counts": (_I, [_VP, _I, _I, _VP, _VP]),
    "ap_smooth_inputs_batch": (_I, [_VP, _I, _I, _I64, _I64, _I64, _F, _F, _VP, _U64, _U32, _VP, _VP]),
    "ap_vote_counts_batch": (_I, [_VP, _I, _I, _I64, _I64, _I64, _I, _VP, _VP]),
    "ap_nes_inputs": (_I, [_VP, _I, _I, _I, _I, _F, _VP, _U64, _U32, _I64, _VP, _VP]),
    "ap_nes_grad": (_I, [_VP, _I, _I, _I, _I, _I, _F, _VP, _U64, _U32, _I64, _VP, _VP]),
    "ap_bias_act_nhwc_bf16": (_I, [_VP, _VP, _VP, _I64, _I, _I, _VP]),
    "ap_comm_unique_id": (_I, [ctypes.c_char_p]),
    "ap_comm_init": (_I, [_I, _I, ctypes.c_char_p, ctypes.POINTER(_VP)]),
    "ap_allreduce_counts": (_I, [_VP, _VP, _SZ, _VP]),
    "ap_comm_destroy": (None, [_VP]),
    "ap_profile_enable": (_I, [_VP, _I]),
    "ap_profile_read": (_I, [_VP, ctypes.POINTER(ctypes.c_double), c_i64_p]),
    "ap_debug_gemm": (_I, [_VP, _VP, _VP, _I, _VP]),
    "ap_debug_gemm_tf32": (_I, [_VP, _VP, _VP, _I, _VP]),
    "ap_precision": (_I, [_VP, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_uint32)]),
}

_lib = None


class AudioPureError(RuntimeError):
    pass


def lib_path():
    """In-tree library; AP_LIB=/path/to/variant.so selects an alternative build (A/B experiments only)."""
    return os.environ.get("AP_LIB") or _build.LIB


def load():
    """Loads the shared library (building it with nvcc if it is missing and nvcc is present)."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        try:
            _build.build()
        except Exception as exc:  # no nvcc on this machine
            raise AudioPureError(
                "libaudiopure_b200.so is not built (%s); run `python -m audiopure_b200.build`. "
                "There is no CPU or PyTorch fallback for this path." % exc)
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the export is missing
        fn.restype = res
        fn.argtypes = args
    if lib.ap_abi_version() != AP_ABI_VERSION:
        raise AudioPureError("ABI mismatch: library %d, binding %d" % (lib.ap_abi_version(), AP_ABI_VERSION))
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise AudioPureError(load().ap_last_error().decode("utf-8", "replace"))


def stream_ptr():
    import torch

    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
