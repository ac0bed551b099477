"""ctypes binding of libcomfystereo_b200.so (include/comfystereo_b200.h).

The shared library IS the product: there is no Python, PyTorch or CPU fallback for any kernel.
If the library is missing or the device is not a B200-class (sm_100) GPU the calls raise.
PyTorch is used only for device memory, streams and torch.distributed.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcomfystereo_b200.so")

CS_OK = 0
ERRORS = {-1: "CS_ERR_ARG", -2: "CS_ERR_CUDA", -3: "CS_ERR_DEVICE", -4: "CS_ERR_WORKSPACE",
          -5: "CS_ERR_UNSUPPORTED", -6: "CS_ERR_MODE"}

FILL_KEYS = ["none", "naive", "naive_interpolating", "polylines_soft", "polylines_sharp", "inverse",
             "hybrid_edge", "gpu_warp", "none_post", "inverse_post", "hybrid_edge_plus", "gpu_warp_mesh"]   # index = cs_fill
MODES = ["left-right", "right-left", "top-bottom", "bottom-top", "red-cyan-anaglyph", "left-only",
         "only-right", "cyan-red-reverseanaglyph"]           # index = cs_mode


class CsParams(ctypes.Structure):
    _fields_ = [
        ("fill", ctypes.c_int32), ("mode", ctypes.c_int32),
        ("divergence", ctypes.c_double), ("separation", ctypes.c_double),
        ("stereo_balance", ctypes.c_double), ("convergence_point", ctypes.c_double),
        ("stereo_offset_exponent", ctypes.c_double),
        ("blur_enabled", ctypes.c_int32), ("blur_box", ctypes.c_int32),
        ("blur_radius", ctypes.c_int32), ("blur_vert_smooth", ctypes.c_int32),
        ("blur_edge_threshold", ctypes.c_double), ("blur_falloff", ctypes.c_double),
        ("group_size", ctypes.c_int32), ("depth_h", ctypes.c_int32), ("depth_w", ctypes.c_int32),
        ("blur_flavor", ctypes.c_int32),
    ]


class CsError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"{ERRORS.get(code, code)}: {message}")
        self.code = code


_P = ctypes.c_void_p
_I = ctypes.c_int
_D = ctypes.c_double
_SZ = ctypes.c_size_t
_PP = ctypes.POINTER(CsParams)
PROGRESS_FN = ctypes.CFUNCTYPE(None, ctypes.c_int, ctypes.c_void_p)   # cs_progress_fn

# name -> (restype, argtypes); every symbol include/comfystereo_b200.h declares
SIGNATURES = {
    "cs_abi_version": (_I, []),
    "cs_last_error": (ctypes.c_char_p, []),
    "cs_device_check": (_I, []),
    "cs_output_dims": (_I, [_PP, _I, _I] + [ctypes.POINTER(_I)] * 4),
    "cs_workspace_bytes": (_SZ, [_PP, _I, _I, _I]),
    "cs_depth_prepare": (_I, [_P, _I, _I, _I, _I, _P, _P, _P]),
    "cs_depth_resize": (_I, [_P, _I, _I, _I, _I, _I, _I, _P, _P]),
    "cs_blur": (_I, [_P, _I, _I, _I, _PP, _P, _P, _P, _P, _P]),
    "cs_shift_indices": (_I, [_P, _I, _I, _I, _D, _D, _D, _I, _P, _P]),
    "cs_warp_fill": (_I, [_P, _P, _I, _I, _I, _I, _D, _D, _D, _D, _P, _P, _SZ, _P]),
    "cs_warp_fill_scratch_bytes": (_SZ, [_I, _I, _I]),
    "cs_forward_warp": (_I, [_P, _P, _I, _I, _I, _D, _D, _D, _D, _P, _P, _P, _SZ, _P]),
    "cs_forward_warp_scratch_bytes": (_SZ, [_I, _I, _I]),
    "cs_forward_warp_mesh": (_I, [_P, _P, _I, _I, _I, _D, _D, _D, _D, _P, _P, _P, _SZ, _P]),
    "cs_forward_warp_mesh_scratch_bytes": (_SZ, [_I, _I, _I]),
    "cs_quantize_image": (_I, [_P, _I, _I, _I, _P, _P]),
    "cs_compose": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _P]),
    "cs_stereo_batch": (_I, [_PP, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _SZ, _P]),
    "cs_stereo_batch_host": (_I, [_PP, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _I]),
    "cs_stereo_batch_host_progress": (_I, [_PP, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _I, _P, _P]),
    "cs_host_release": (None, []),
    "cs_host_compact_enabled": (_I, []),
    "cs_host_stream_bandwidth": (_D, [_SZ, _I]),
    "cs_launch_count": (ctypes.c_longlong, [_I]),
    "cs_polylines_status": (_I, [_PP, _I, _I, _I, _P, ctypes.POINTER(_I), ctypes.POINTER(_I)]),
    "cs_set_test_flags": (None, [_I]),
    "cs_profile_kernel_count": (_I, []),
    "cs_profile_kernel_name": (ctypes.c_char_p, [_I]),
    "cs_profile_enable": (None, [_I]),
    "cs_profile_collect": (_I, [ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_longlong)]),
}

_lib = None


def lib():
    """Loads the library once.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C comfystereo_b200/csrc`. comfystereo_b200 has no CPU/PyTorch fallback.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the .so is stale
            fn.restype = res
            fn.argtypes = args
        if handle.cs_abi_version() != 4:
            raise ImportError("libcomfystereo_b200.so: ABI version mismatch, rebuild it")
        _lib = handle
    return _lib


def check(rc):
    if rc != CS_OK:
        raise CsError(rc, lib().cs_last_error().decode("utf-8", "replace"))


def loaded():
    return _lib is not None
