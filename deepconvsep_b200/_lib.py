"""ctypes binding of libdcs.so (include/dcs.h).  No torch types cross this boundary: plain
pointers and sizes only.  There is NO fallback: a missing library or a missing CUDA device
raises."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libdcs.so")

ARCH_IDS = {"dsd": 0, "ikala": 1, "ikala_nopool": 2, "bach10": 3, "bach10_score": 4, "dsd_ild": 5}
PATCHER_IDS = {"standalone": 0, "util": 1}


class DcsError(RuntimeError):
    pass


_p = C.c_void_p
_i64 = C.c_int64
_SIGS = {
    "dcs_version": (C.c_int, []),
    "dcs_last_error": (C.c_char_p, []),
    "dcs_create": (C.c_int, [C.c_int, C.POINTER(_p)]),
    "dcs_destroy": (C.c_int, [_p]),
    "dcs_workspace_bytes": (_i64, [_p]),
    "dcs_launch_count": (_i64, [_p]),
    "dcs_set_spectrum_tap": (C.c_int, [_p, _p, _i64]),
    "dcs_set_pool_tap": (C.c_int, [_p, _p, _i64]),
    "dcs_profile": (C.c_int, [_p, C.c_int]),
    "dcs_profile_read": (C.c_int, [_p, C.c_char_p, C.c_int, _p, C.c_int]),
    "dcs_stft_plan": (C.c_int, [_p, C.c_int, C.c_int, _p, _p, C.POINTER(_p)]),
    "dcs_stft_plan_destroy": (C.c_int, [_p]),
    "dcs_num_frames": (_i64, [_i64, C.c_int]),
    "dcs_padded_bins": (_i64, [C.c_int]),
    "dcs_stft_forward": (C.c_int, [_p, _p, _i64, _p, _p, C.c_float, _i64, _p]),
    "dcs_stft_forward_polar": (C.c_int, [_p, _p, _i64, _p, _p, C.c_float, _i64, _p]),
    "dcs_istft": (C.c_int, [_p, _p, C.c_int, _i64, _i64, _i64, _p, _i64, _i64, _p]),
    "dcs_istft_polar": (C.c_int, [_p, _p, _p, _p, C.c_float, _i64, _i64, _p, _i64, _p]),
    "dcs_model_create": (C.c_int, [_p, C.c_int, C.c_int, C.c_int, C.c_int, _p, _p, _p, C.POINTER(_p)]),
    "dcs_model_destroy": (C.c_int, [_p]),
    "dcs_model_nsources": (C.c_int, [_p]),
    "dcs_num_patches": (_i64, [_i64, C.c_int, C.c_int, C.c_int]),
    "dcs_separate_spec": (C.c_int, [_p, _p, _p, _p, _i64, _i64, C.c_int, C.c_int, _p, _i64, _p]),
    "dcs_separate_spec_channels": (C.c_int, [_p, _p, _p, _i64, _p, _i64, _i64, C.c_int, C.c_int, _p, _i64, _p]),
    "dcs_separate_audio_score": (C.c_int, [_p, _p, _p, _p, _i64, _p, C.c_float, C.c_int, C.c_int, _p, _i64, _p]),
    "dcs_separate_audio_stereo": (C.c_int, [_p, _p, _p, _p, _i64, _i64, C.c_float, C.c_int, C.c_int, _p, _i64, _p]),
    "dcs_xcorr_lags": (C.c_int, [_p, _p, _p, C.c_int, _i64, C.c_int, _p, _p]),
    "dcs_gemm_f32": (C.c_int, [_p, C.c_int, _p, _i64, _p, _i64, _p, _p, _i64, C.c_int, C.c_int, C.c_int, C.c_int, _p]),
    "dcs_separate_audio": (C.c_int, [_p, _p, _p, _p, _i64, C.c_float, C.c_int, C.c_int, _p, _i64, _p]),
    "dcs_separate_host": (C.c_int, [_p, _p, _p, _p, _i64, C.c_float, C.c_int, C.c_int, _p, _i64, _p]),
    "dcs_separate_batch_pcm16_host": (C.c_int, [_p, _p, _p, C.c_int, _p, _p, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, _p, _p, _p]),
    "dcs_separate_pcm16_host": (C.c_int, [_p, _p, _p, _p, _i64, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int,
                                           _p, _i64, _p]),
}

_lib = None


def exported_symbols():
    """Names include/dcs.h declares (kept in sync by tests/test_abi.py)."""
    return sorted(_SIGS)


def load():
    """dlopen libdcs.so and attach the signatures.  Raises DcsError if it was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DcsError("libdcs.so is missing (%s): build it with `python -m deepconvsep_b200.build` "
                       "-- there is no CPU fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise DcsError("libdcs error %d: %s" % (rc, load().dcs_last_error().decode("utf-8", "replace")))
