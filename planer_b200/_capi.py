"""ctypes binding of libplaner_b200.so (the C ABI declared in include/planer_b200.h).

There is NO fallback: if the shared library is missing or a call fails, an exception is raised.
The library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a) and lives next to this
file so that it travels with the source snapshot to the GPU box.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# PLNR_LIB: another build of the same library (A/B timing of two kernel versions on one box); never a different backend
LIB_PATH = os.environ.get('PLNR_LIB') or os.path.join(HERE, 'libplaner_b200.so')

F32, F16, U8 = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_LEAKY, ACT_SIGMOID = 0, 1, 2, 3
ALGO_AUTO, ALGO_TCGEN05, ALGO_DIRECT = 0, 1, 2
EW_RELU, EW_LEAKY, EW_SIGMOID, EW_ADD, EW_SCALE_SHIFT, EW_CLIP, EW_HARDSIGMOID = 0, 1, 2, 3, 4, 5, 6


class Tensor(C.Structure):
    """plnr_tensor: a pixel-major (NHWC) activation view."""
    _fields_ = [('ptr', C.c_void_p), ('n', C.c_int32), ('h', C.c_int32), ('w', C.c_int32), ('c', C.c_int32),
                ('ld', C.c_int32), ('coff', C.c_int32)]


class Epilogue(C.Structure):
    _fields_ = [('scale', C.c_void_p), ('shift', C.c_void_p), ('residual', C.POINTER(Tensor)),
                ('act', C.c_int32), ('alpha', C.c_float), ('res_after_act', C.c_int32), ('out_nchw', C.c_int32),
                ('out_f32', C.c_int32), ('acc_scale', C.c_float), ('acc_scale_dev', C.c_void_p), ('pool_sum', C.c_void_p)]


class ConvDesc(C.Structure):
    _fields_ = [('dtype', C.c_int32), ('kh', C.c_int32), ('kw', C.c_int32),
                ('pad_t', C.c_int32), ('pad_l', C.c_int32), ('pad_b', C.c_int32), ('pad_r', C.c_int32),
                ('stride_h', C.c_int32), ('stride_w', C.c_int32), ('dil_h', C.c_int32), ('dil_w', C.c_int32),
                ('groups', C.c_int32), ('algo', C.c_int32)]


_P = C.c_void_p
_TP = C.POINTER(Tensor)
# name -> (argtypes);  every function returns int except plnr_last_error
PROTOTYPES = {
    'plnr_abi_version': [],
    'plnr_create': [C.c_int, _P, C.POINTER(_P)],
    'plnr_destroy': [_P],
    'plnr_set_stream': [_P, _P],
    'plnr_stream_sync': [_P],
    'plnr_launch_count': [_P, C.POINTER(C.c_int64)],
    'plnr_device_info': [_P, C.POINTER(C.c_int64)],
    'plnr_malloc': [_P, C.c_size_t, C.POINTER(_P)],
    'plnr_free': [_P, _P],
    'plnr_memcpy_h2d': [_P, _P, _P, C.c_size_t],
    'plnr_memcpy_d2h': [_P, _P, _P, C.c_size_t],
    'plnr_memset': [_P, _P, C.c_int, C.c_size_t],
    'plnr_nchw_to_nhwc': [_P, _P, C.c_int, C.c_int, _TP, C.c_int],
    'plnr_stem_pack': [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _TP, C.c_int, C.c_int, C.c_int],
    'plnr_stem_pool_supported': [C.c_int] * 16,
    'plnr_stem_pool_geometry': [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)],
    'plnr_stem_pool_fwd': [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P] + [C.c_int] * 11 + [_TP],
    'plnr_stem_pool_fwd_u8': [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P] + [C.c_int] * 11 + [_TP],
    'plnr_stem3x3_supported': [C.c_int] * 11,
    'plnr_stem3x3_fwd': [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P, C.c_int, C.c_float, _TP],
    'plnr_nhwc_to_nchw': [_P, _TP, C.c_int, _P, C.c_int],
    'plnr_cast': [_P, _P, C.c_int, _P, C.c_int, C.c_int64],
    'plnr_pack_conv_weight': [_P, _P, C.c_int, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int],
    'plnr_split_f32': [_P, _TP, _TP, C.c_float, _P],
    'plnr_pack_conv_weight_split': [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float],
    'plnr_absmax_f32': [_P, _P, C.c_int64, _P],
    'plnr_fold_affine': [_P, _P, _P, _P, C.c_int, _P, _P, C.c_int],
    'plnr_conv2d_fwd': [_P, C.POINTER(ConvDesc), _TP, _P, _TP, C.POINTER(Epilogue)],
    'plnr_conv2d_out_nchw_supported': [C.POINTER(ConvDesc), _TP, _TP],
    'plnr_conv2d_pool_parts': [C.POINTER(ConvDesc), _TP, _TP],
    'plnr_conv2d_shortcut_supported': [C.POINTER(ConvDesc), _TP, _TP, C.c_int, _TP],
    'plnr_conv2d_shortcut_fwd': [_P, C.POINTER(ConvDesc), _TP, _P, _TP, C.c_int, _TP, C.POINTER(Epilogue)],
    'plnr_dense_fwd': [_P, C.c_int, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.POINTER(Epilogue), C.c_int],
    'plnr_maxpool2d': [_P, C.c_int, _TP, _TP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int],
    'plnr_upsample_linear': [_P, C.c_int, _TP, _TP, C.c_int, C.c_int, _P],
    'plnr_resize_linear': [_P, C.c_int, _TP, _TP, _P, _P, _P, _P, _P, _P],
    'plnr_avgpool2d': [_P, C.c_int, _TP, _TP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int],
    'plnr_zero_stuff': [_P, C.c_int, _TP, _TP, C.c_int, C.c_int, C.c_int, C.c_int],
    'plnr_flip_weight': [_P, C.c_int, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int],
    'plnr_upsample_nearest': [_P, C.c_int, _TP, _TP, C.c_int, C.c_int],
    'plnr_copy_channels': [_P, C.c_int, _TP, _TP],
    'plnr_eltwise': [_P, C.c_int, C.c_int, _P, _P, _P, _P, C.c_int64, C.c_int, C.c_float],
    'plnr_unary2': [_P, C.c_int, C.c_int, _P, _P, C.c_int64, C.c_float, C.c_float],
    'plnr_softmax': [_P, C.c_int, _P, _P, C.c_int64, C.c_int],
    'plnr_global_avgpool': [_P, C.c_int, _TP, _P],
    'plnr_gap_dense_fwd': [_P, C.c_int, _TP, _P, _P, _P, _P, C.c_int, C.c_int, C.c_float],
    'plnr_pooled_dense_fwd': [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P, C.c_int, C.c_int, C.c_float],
    'plnr_graph_begin': [_P],
    'plnr_graph_end': [_P, C.POINTER(_P)],
    'plnr_graph_launch': [_P, _P],
    'plnr_graph_destroy': [_P],
    'plnr_event_create': [C.POINTER(_P)],
    'plnr_event_record': [_P, _P],
    'plnr_event_elapsed_ms': [_P, _P, C.POINTER(C.c_float)],
    'plnr_event_destroy': [_P],
    'plnr_conv2d_algo': [C.POINTER(ConvDesc), _TP, _TP],
    'plnr_last_kernel': [_P, C.c_char_p, C.c_int],
    'plnr_debug_shift_geometry': [C.POINTER(ConvDesc), _TP, _TP, C.POINTER(C.c_int), C.c_int],
    'plnr_debug_conv_profile': [_P, C.c_int, C.POINTER(C.c_int64), C.c_int],
}

_lib = None


class PlanerB200Error(RuntimeError):
    pass


def load():
    """dlopen the library and attach prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PlanerB200Error(
            'libplaner_b200.so not found at %s -- build it with `python -c "import __graft_entry__ as g; g.build()"`. '
            'planer_b200 has no CPU fallback.' % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError here = header and library out of sync
        fn.argtypes = argtypes
        fn.restype = C.c_int
    lib.plnr_last_error.argtypes = []
    lib.plnr_last_error.restype = C.c_char_p
    if lib.plnr_abi_version() != 3:
        raise PlanerB200Error('ABI version mismatch: library %d, binding 3' % lib.plnr_abi_version())
    _lib = lib
    return lib


def check(rc, what=''):
    if rc != 0:
        msg = _lib.plnr_last_error().decode(errors='replace') if _lib is not None else ''
        raise PlanerB200Error('%s failed (rc=%d): %s' % (what or 'libplaner_b200 call', rc, msg))
    return rc


def src_dtype_code(dt):
    """dtype code of a graph INPUT as the caller hands it over: float32 / float16, or uint8 images (converted by the
    graph-entry kernels: plnr_nchw_to_nhwc, plnr_stem_pack, plnr_stem_pool_fwd_u8, plnr_cast)."""
    import numpy as np
    return U8 if np.dtype(dt) == np.uint8 else dtype_code(dt)


def dtype_code(dt):
    import numpy as np
    dt = np.dtype(dt)
    if dt == np.float32:
        return F32
    if dt == np.float16:
        return F16
    raise PlanerB200Error('unsupported dtype %s (the B200 path computes in float32 or float16)' % dt)
