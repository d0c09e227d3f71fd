"""ctypes binding of libnnpops_b200.so (the C ABI declared in include/nnpops_b200.h).

The library is the product: if it is missing the import fails loudly -- there is no Python/CPU fallback.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NNPOPS_LIB_PATH") or os.path.join(_HERE, "libnnpops_b200.so")   # the override is a development aid (A/B builds)

if not os.path.exists(LIB_PATH):
    raise ImportError(
        "nnpops_b200: %s is missing. Build it with `python -m nnpops_b200.build` (needs nvcc); there is no CPU fallback." % LIB_PATH)

lib = C.CDLL(LIB_PATH)
lib.nnpops_last_error.restype = C.c_char_p

_vp, _i, _f = C.c_void_p, C.c_int, C.c_float

_SIGS = {
    "nnpops_abi_version": [],
    "nnpops_ani_create": [C.POINTER(_vp), _i, _i, _f, _f, _vp, _i, _vp, _i, _vp, _i, _i, _i],
    "nnpops_ani_forward": [_vp, _vp, _vp, _vp, _vp, _vp],
    "nnpops_ani_backward": [_vp, _vp, _vp, _vp, _vp],
    "nnpops_ani_overflowed": [_vp, C.POINTER(_i)],
    "nnpops_ani_overflow_poll": [_vp, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)],
    "nnpops_ani_work": [_vp, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), _vp],
    "nnpops_ani_model_create": [C.POINTER(_vp), _i, _i, _f, _f, _vp, _i, _vp, _i, _vp, _i, _i, _vp, _vp, _i, _i, _i],
    "nnpops_ani_model_create_sharded": [C.POINTER(_vp), _i, _i, _f, _f, _vp, _i, _vp, _i, _vp, _i, _i, _vp, _vp, _i, _i, _i, _i, _i],
    "nnpops_ani_model_create_owned": [C.POINTER(_vp), _i, _i, _f, _f, _vp, _i, _vp, _i, _vp, _i, _i, _vp, _vp, _i, _i, _i, _vp],
    "nnpops_ani_model_energy_grad": [_vp, _vp, _vp, _vp, _vp, _vp],
    "nnpops_ani_model_energy_grad_host": [_vp, _vp, _vp, _vp, _vp, _vp],
    "nnpops_ani_model_buffers": [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_i), _vp],
    "nnpops_ani_model_read_features": [_vp, _i, _vp, _vp],
    "nnpops_ani_model_work": [_vp, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), C.POINTER(C.c_double), _vp],
    "nnpops_ani_model_info": [_vp, C.POINTER(_i), C.POINTER(_i), C.POINTER(C.c_double)],
    "nnpops_ani_model_mlp_fused": [_vp, C.POINTER(_i)],
    "nnpops_ani_set_skin": [_vp, _f],
    "nnpops_ani_model_set_skin": [_vp, _f],
    "nnpops_ani_model_skin_stats": [_vp, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)],
    "nnpops_ani_model_overflowed": [_vp, C.POINTER(_i)],
    "nnpops_ani_model_overflow_poll": [_vp, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)],
    "nnpops_ani_model_timing_begin": [_vp, _i],
    "nnpops_ani_model_timing_end": [_vp, _vp, C.POINTER(_i)],
    "nnpops_launch_count": [C.POINTER(C.c_ulonglong)],
    "nnpops_batched_linear_forward": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    "nnpops_batched_linear_backward": [_vp, _vp, _vp, _i, _i, _i, _i, _vp],
}
for _name, _args in _SIGS.items():
    _fn = getattr(lib, _name)
    _fn.argtypes = _args
    _fn.restype = _i
for _name in ("nnpops_ani_destroy", "nnpops_ani_model_destroy"):
    getattr(lib, _name).argtypes = [_vp]
    getattr(lib, _name).restype = None


def register(sigs, destroyers=()):
    """Add further entry points (used by the neighbors / cfconv / pme modules)."""
    for name, args in sigs.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = _i
    for name in destroyers:
        getattr(lib, name).argtypes = [_vp]
        getattr(lib, name).restype = None


def check(code):
    if code != 0:
        raise RuntimeError(lib.nnpops_last_error().decode())


def ptr(t):
    """Device (or host) pointer of a torch tensor / numpy array, or None."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return C.c_void_p(t.data_ptr())
    return t.ctypes.data_as(C.c_void_p)


def current_stream(device=None):
    import torch
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
