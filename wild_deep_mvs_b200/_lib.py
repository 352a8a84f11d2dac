"""ctypes binding of libmvsb200.so (include/mvsb200.h).  There is NO fallback: if the library is
missing or a call fails, an exception is raised."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(HERE, "libmvsb200.so")

MAX_SRC = 16

# enums of mvsb200.h
GEOM_MVS, GEOM_VIS = 0, 1
AGG_VARIANCE, AGG_VARIANCE_MEAN, AGG_SOFTMIN, AGG_GROUPCORR = 0, 1, 2, 3
DEPTH_VALUES, DEPTH_VOLUME, DEPTH_START, DEPTH_START_MAP = 0, 1, 2, 3
SKIP_NONE, SKIP_BEFORE_RELU, SKIP_AFTER_RELU = 0, 1, 2
CONF_NONE, CONF_SUM4, CONF_WINDOW = 0, 1, 2
PRECISION_3XTF32, PRECISION_TF32 = 0, 1
ABI_VERSION = 6


class Mvsb200Error(RuntimeError):
    pass


class CostVolumeDesc(ctypes.Structure):
    _fields_ = [("geom", ctypes.c_int), ("agg", ctypes.c_int), ("depth_mode", ctypes.c_int),
                ("B", ctypes.c_int), ("S", ctypes.c_int), ("C", ctypes.c_int),
                ("D", ctypes.c_int), ("H", ctypes.c_int), ("W", ctypes.c_int),
                ("groups", ctypes.c_int),
                ("src_h", ctypes.c_int * MAX_SRC), ("src_w", ctypes.c_int * MAX_SRC),
                ("out_view_stride", ctypes.c_longlong)]


class Conv3dDesc(ctypes.Structure):
    _fields_ = [("B", ctypes.c_int), ("D", ctypes.c_int), ("H", ctypes.c_int), ("W", ctypes.c_int),
                ("Cin", ctypes.c_int), ("Cin2", ctypes.c_int), ("Cout", ctypes.c_int),
                ("kd", ctypes.c_int), ("kh", ctypes.c_int), ("kw", ctypes.c_int),
                ("stride", ctypes.c_int), ("transposed", ctypes.c_int),
                ("relu", ctypes.c_int), ("skip_mode", ctypes.c_int), ("static_params", ctypes.c_int)]


_vp = ctypes.c_void_p
_i = ctypes.c_int

# name -> (restype, argtypes); every symbol include/mvsb200.h declares
SIGNATURES = {
    "mvsb200_abi_version": (_i, []),
    "mvsb200_last_error": (ctypes.c_char_p, []),
    "mvsb200_device_info": (_i, [ctypes.POINTER(_i)] * 3),
    "mvsb200_mvs_relative_proj": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "mvsb200_vis_homography_params": (_i, [_vp, _vp, ctypes.c_float, _vp, _i, _i, _vp]),
    "mvsb200_build_cost_volume": (_i, [ctypes.POINTER(CostVolumeDesc), _vp, ctypes.POINTER(_vp), _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mvsb200_build_cost_volume_backward": (_i, [ctypes.POINTER(CostVolumeDesc), _vp, ctypes.POINTER(_vp), _vp, _vp, _vp, _vp, _vp, _vp,
                                                ctypes.POINTER(_vp), _vp, _vp]),
    "mvsb200_conv3d_wgrad": (_i, [_vp, _vp] + [_i] * 10 + [_vp, _vp]),
    "mvsb200_depth_regress_backward": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mvsb200_conv3d_out_shape": (_i, [ctypes.POINTER(Conv3dDesc)] + [ctypes.POINTER(_i)] * 3),
    "mvsb200_conv3d": (_i, [ctypes.POINTER(Conv3dDesc), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mvsb200_conv3d_tc_supported": (_i, [ctypes.POINTER(Conv3dDesc)]),
    "mvsb200_conv3d_tc_packed_floats": (ctypes.c_longlong, [ctypes.POINTER(Conv3dDesc)]),
    "mvsb200_conv3d_tc_pack": (_i, [ctypes.POINTER(Conv3dDesc), _vp, _vp, _vp]),
    "mvsb200_conv3d_tc": (_i, [ctypes.POINTER(Conv3dDesc), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "mvsb200_absmax": (_i, [_vp, ctypes.c_longlong, _vp, _vp]),
    "mvsb200_conv3d_zm_supported": (_i, [ctypes.POINTER(Conv3dDesc)]),
    "mvsb200_conv3d_zm_packed_bytes": (ctypes.c_longlong, [ctypes.POINTER(Conv3dDesc)]),
    "mvsb200_conv3d_zm_pack": (_i, [ctypes.POINTER(Conv3dDesc), _vp, _vp, _vp]),
    "mvsb200_conv3d_zm": (_i, [ctypes.POINTER(Conv3dDesc), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mvsb200_conv3d_c1_supported": (_i, [ctypes.POINTER(Conv3dDesc)]),
    "mvsb200_conv3d_c1": (_i, [ctypes.POINTER(Conv3dDesc), _vp, ctypes.POINTER(ctypes.c_float), ctypes.c_float, ctypes.c_float, _vp, _vp]),
    "mvsb200_conv3d_zm_slice": (_i, [ctypes.POINTER(Conv3dDesc), _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mvsb200_depth_regress": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp]),
    "mvsb200_cvp_depth_delta": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "mvsb200_vis_uncert_net": (_i, [_vp, _i, _i, _i, ctypes.POINTER(ctypes.c_float), _vp, _vp]),
    "mvsb200_conv2d": (_i, [_i] * 8 + [_vp] * 6),
    "mvsb200_geometric_filter": (_i, [_vp, _i, _i, ctypes.POINTER(_vp), ctypes.POINTER(_i), ctypes.POINTER(_i), _i, _vp, _vp, _vp,
                                      ctypes.c_float, ctypes.c_float, ctypes.c_float, _i, _vp, _vp, _vp, _vp, _vp]),
    "mvsb200_bias_act": (_i, [_vp, ctypes.c_longlong, _i, _vp, _vp, _vp, ctypes.c_float, _vp, _vp]),
    "mvsb200_gathered_masks": (_i, [_i] * 6 + [_vp] * 5 + [ctypes.c_float] + [_vp] * 7),
    "mvsb200_vis_fuse": (_i, [ctypes.POINTER(_vp), ctypes.POINTER(_vp), _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "mvsb200_gather_unique_id": (_i, [_vp]),
    "mvsb200_gather_init": (_i, [_vp, _i, _i, ctypes.POINTER(_vp)]),
    "mvsb200_allgather_depth": (_i, [_vp, _vp, ctypes.c_longlong, _i, _vp]),
    "mvsb200_gather_destroy": (_i, [_vp]),
}

_lib = None


def load():
    """Load libmvsb200.so; raises Mvsb200Error if it has not been built (python -m wild_deep_mvs_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise Mvsb200Error("libmvsb200.so not found at %s -- build it with `python -m wild_deep_mvs_b200.build`; "
                           "there is no CPU or PyTorch fallback for the hot path" % SO_PATH)
    lib = ctypes.CDLL(SO_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    if lib.mvsb200_abi_version() != ABI_VERSION:
        raise Mvsb200Error("libmvsb200.so ABI version %d, expected %d" % (lib.mvsb200_abi_version(), ABI_VERSION))
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().mvsb200_last_error().decode("utf-8", "replace")
        raise Mvsb200Error("%s failed (code %d): %s" % (what, rc, msg))
