"""ctypes binding of libsykepic_b200.so (the C ABI declared in include/sykepic_b200.h).

The product path has no CPU fallback: if the library is missing or no B200 is
visible, every entry point raises.
"""

import ctypes as C
from pathlib import Path

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "libsykepic_b200.so"

SPK_OK = 0
SPK_ERR_INVALID = -1
SPK_ERR_CUDA = -2
SPK_ERR_FAULTY_BIN = -3
SPK_ERR_EMPTY_RESIZE = -4
SPK_ERR_PARSE = -5
SPK_ERR_CAPACITY = -6
SPK_ERR_UNSUPPORTED = -7
SPK_ERR_STATE = -8
SPK_ERR_IO = -9

BORDER = {"mode": 0, "black": 1, "white": 2}
DTYPE_F32, DTYPE_BF16, DTYPE_U8, DTYPE_SPLIT = 0, 1, 2, 3
LAYOUT_NCHW, LAYOUT_NHWC = 0, 1
PRECISION_FP32, PRECISION_BF16, PRECISION_FP32_TC = 0, 1, 2
CONV_AUTO, CONV_SIMT, CONV_TCGEN05, CONV_TCGEN05_TAPS = 0, 1, 2, 3
INT32_MAX = 2**31 - 1
PROF_CATEGORIES = ["preprocess", "conv_tc", "conv_simt", "stem", "pool", "bn_relu", "head", "other"]


class SpkError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"[spk {code}] {message}")
        self.code = code


class FaultyBin(ValueError):
    """A ROI slice runs past the .roi bytes: the reference's `ValueError`
    ("Faulty raw data", sykepic/compute/probability.py:111-112)."""


class EmptyResize(SpkError):
    """Aspect ratio > T:1: cv2.error in the reference (probability.py:113-114)."""


class AdcParseError(ValueError):
    """Malformed .adc line (the reference raises ValueError / IndexError)."""


_p = C.c_void_p
_i = C.c_int
_i64 = C.c_int64
_f = C.c_float
_pp = C.POINTER(C.c_void_p)

# name -> (restype, argtypes); must list every symbol include/sykepic_b200.h declares
PROTOTYPES = {
    "spk_abi_version": (_i, []),
    "spk_create": (_i, [_i, _p, _pp]),
    "spk_destroy": (_i, [_p]),
    "spk_set_stream": (_i, [_p, _p]),
    "spk_synchronize": (_i, [_p]),
    "spk_last_error": (C.c_char_p, [_p]),
    "spk_launch_count": (_i64, [_p]),
    "spk_profile_mode": (_i, [_p, _i]),
    "spk_profile_begin": (_i, [_p]),
    "spk_profile_read": (_i, [_p, _p, _p, _p, _p, _p, _i64]),
    "spk_profile_end": (_i, [_p]),
    "spk_adc_parse": (_i, [C.c_char_p, _i64, _i64, _p, _p, _p, _p, C.POINTER(_i64), C.POINTER(_i64)]),
    "spk_bin_load": (_i, [C.c_char_p, C.c_char_p, _p, _i64, _i64, _p, _p, _p, _p, C.POINTER(_i64), C.POINTER(_i64), _i, _i]),
    "spk_prob_csv_write": (_i, [C.c_char_p, C.c_char_p, _p, _p, _i64, _i, C.POINTER(_i64)]),
    "spk_rois_validate": (_i, [_p, _p, _p, _i64, _i64, _i, _i, C.POINTER(_i64)]),
    "spk_new_dims": (None, [_i, _i, _i, _i, C.POINTER(_i), C.POINTER(_i)]),
    "spk_preprocess": (_i, [_p, _p, _i64, _p, _p, _p, _i64, _i, _i, _i, _i, _i, _i, _p, _p]),
    "spk_fault_count": (_i, [_p, C.POINTER(_i64)]),
    "spk_net_begin": (_i, [_p, _i, _i, _i, _i, _i]),
    "spk_net_buffer": (_i, [_p, _i, _i, _i, _i]),
    "spk_net_conv": (_i, [_p, _i, _i, _i, _i, _i, _p, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _f, _p, _i, _i]),
    "spk_net_maxpool": (_i, [_p, _i, _i, _i, _i, _i]),
    "spk_net_avgpool": (_i, [_p, _i, _i, _i, _i]),
    "spk_net_bn_relu": (_i, [_p, _i, _i, _i, _p, _p, _p, _p, _f, _i]),
    "spk_net_head": (_i, [_p, _i, _i, _pp, _pp, C.POINTER(_i)]),
    "spk_net_end": (_i, [_p]),
    "spk_net_bytes": (_i64, [_p]),
    "spk_net_simt_layers": (_i, [_p]),
    "spk_forward": (_i, [_p, _p, _i64, _f, _p, _p, _p, _p]),
    "spk_last_logits": (_p, [_p]),
    "spk_net_read_buffer": (_i, [_p, _i, _i64, _p, _i64, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "spk_threshold_quantize": (C.c_int32, [C.c_double, _i]),
    "spk_format_prob_csv": (_i, [C.c_char_p, _p, _p, _i64, _i, _p, _i64, C.POINTER(_i64)]),
    "spk_prob_csv_shape": (_i, [C.c_char_p, _i64, C.POINTER(_i64), C.POINTER(_i)]),
    "spk_prob_csv_parse": (_i, [C.c_char_p, _i64, _i64, _i, _p, _p]),
    "spk_png_unfilter": (_i, [_p, _i64, _i64, _i, _p]),
    "spk_png_probe": (_i, [_p, _i64, _p, _p, _i, C.POINTER(_i64)]),
    "spk_png_decode_batch": (_i, [_p, _i64, _p, _p, _p, _p, _i64, _i, C.POINTER(_i64)]),
}

_lib = None


def load():
    """The loaded library (cached).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise SpkError(SPK_ERR_STATE, f"{LIB_PATH} is missing: run `python -m sykepic_b200._build` "
                                          "(there is no CPU fallback for this path)")
        lib = C.CDLL(str(LIB_PATH))
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def last_error(ctx=None):
    msg = load().spk_last_error(ctx)
    return msg.decode(errors="replace") if msg else ""


def check(rc, ctx=None):
    if rc == SPK_OK:
        return
    msg = last_error(ctx)
    if rc == SPK_ERR_FAULTY_BIN:
        raise FaultyBin(msg)
    if rc == SPK_ERR_EMPTY_RESIZE:
        raise EmptyResize(rc, msg)
    if rc == SPK_ERR_PARSE:
        raise AdcParseError(msg)
    if rc == SPK_ERR_IO:
        raise OSError(msg)
    raise SpkError(rc, msg)


def ptr(t):
    """Device / host address of a torch tensor or numpy array (None -> NULL)."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return t.data_ptr()
    return t.ctypes.data
