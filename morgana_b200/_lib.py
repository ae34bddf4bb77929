"""ctypes binding of the C ABI declared in ``include/morgana_b200.h``.

There is no CPU fallback: if the shared library has not been built this module raises at import, and every op in
this package refuses non-CUDA tensors.
"""
import ctypes
import os

from morgana_b200 import build as _build

c_i32, c_i64, c_f32, c_void_p, c_int = ctypes.c_int32, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p, ctypes.c_int

MG_OK, MG_ERR_INVALID_ARG, MG_ERR_CUDA, MG_ERR_UNSUPPORTED = 0, -1, -2, -3

NORM_NONE, NORM_MVN, NORM_MINMAX = 0, 1, 2
PATH_AUTO, PATH_BULK, PATH_DIRECT = 0, 1, 2
MAX_TERMS = 12
(RED_SQDIFF, RED_ABSDIFF, RED_BCE, RED_SUM, RED_ROOT_SQDIFF, RED_SQDIFF_EXP, RED_XOR, RED_AND, RED_EQ, RED_SQ, RED_CE) = range(11)
DT_F32, DT_U8 = 0, 1
FLAG_M_GT_HALF, FLAG_A_GT_HALF, FLAG_IN_TOTAL = 1, 2, 4
ACT_NONE, ACT_SIGMOID = 0, 1


class Term(ctypes.Structure):
    """``mg_term`` (include/morgana_b200.h)."""
    _fields_ = [
        ('a', c_void_p), ('b', c_void_p), ('m', c_void_p), ('grad', c_void_p), ('grad_scale_dev', c_void_p),
        ('result', c_void_p),
        ('a_sb', c_i64), ('a_st', c_i64), ('b_sb', c_i64), ('b_st', c_i64),
        ('m_sb', c_i64), ('m_st', c_i64), ('g_sb', c_i64), ('g_st', c_i64),
        ('D', c_i32), ('kind', c_i32), ('ab_dtype', c_i32), ('m_dtype', c_i32), ('b_is_u8', c_i32),
        ('accumulate', c_i32), ('grad_scale', c_f32), ('flags', c_i32),
    ]


class TermResult(ctypes.Structure):
    """``mg_term_result``: 48 bytes = 6 doubles = 12 floats per term."""
    _fields_ = [('sum', ctypes.c_double), ('count', ctypes.c_double), ('loss', ctypes.c_double), ('isum', c_i64),
                ('sum_f32', c_f32), ('count_f32', c_f32), ('loss_f32', c_f32), ('weighted_loss_f32', c_f32)]


class Column(ctypes.Structure):
    """``mg_column``: the per-column program of the whole-row objective kernel."""
    _fields_ = [('loss_kind', ctypes.c_int8), ('loss_slot', ctypes.c_int8), ('metric_kind', ctypes.c_int8),
                ('metric_slot', ctypes.c_int8), ('mask_col', ctypes.c_int16), ('width', ctypes.c_int16),
                ('loss_weight', c_f32)]


class Slot(ctypes.Structure):
    """``mg_slot``."""
    _fields_ = [('result', c_void_p), ('D', c_i32), ('per_frame', c_i32), ('weighted', c_i32), ('accumulate', c_i32),
                ('in_total', c_i32), ('weight', c_f32)]


COL_NONE = -1
assert ctypes.sizeof(Column) == 12 and ctypes.sizeof(Slot) == 32
TERM_RESULT_BYTES = ctypes.sizeof(TermResult)
assert TERM_RESULT_BYTES == 48 and ctypes.sizeof(Term) == 144

# name -> (restype, argtypes); exactly the exports of include/morgana_b200.h
PROTOTYPES = {
    'mg_abi_version': (c_int, []),
    'mg_last_error': (ctypes.c_char_p, []),
    'mg_sm_count': (c_int, []),
    'mg_dur_scan': (c_int, [c_void_p, c_int, c_i64, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    'mg_upsample_norm_f32': (c_int, [c_void_p, c_i64, c_i64, c_void_p, c_void_p, c_void_p, c_i64, c_int, c_void_p,
                                     c_int, c_int, c_int, c_i64, c_int, c_void_p]),
    'mg_dur_scan_packed': (c_int, [c_void_p, c_int, c_void_p, c_int, c_i64, c_void_p, c_void_p, c_void_p, c_void_p]),
    'mg_upsample_packed_norm_f32': (c_int, [c_void_p, c_i64, c_void_p, c_i64, c_void_p, c_void_p, c_void_p, c_i64, c_int, c_void_p,
                                            c_int, c_int, c_int, c_i64, c_void_p]),
    'mg_upsample_norm_f32_bf16out': (c_int, [c_void_p, c_i64, c_i64, c_void_p, c_void_p, c_void_p, c_i64, c_int, c_void_p,
                                             c_int, c_int, c_int, c_i64, c_void_p]),
    'mg_upsample_bytes': (c_int, [c_void_p, c_i64, c_i64, c_void_p, c_void_p, c_int, c_int, c_i64, c_i64, c_int,
                                  c_void_p]),
    'mg_upsample_norm_bwd_f32': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_int, c_void_p, c_int, c_int,
                                         c_int, c_i64, c_void_p]),
    'mg_pad_collate': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_i64, c_i64, c_i64, c_void_p]),
    'mg_pack_rows': (c_int, [c_void_p, c_i64, c_i64, c_void_p, c_void_p, c_int, c_i64, c_i64, c_void_p]),
    'mg_segment_ends': (c_int, [c_void_p, c_i64, c_i64, c_void_p, c_void_p, c_int, c_int, c_i64, c_i64, c_void_p]),
    'mg_split_to_segments': (c_int, [c_void_p, c_i64, c_i64, c_void_p, c_void_p, c_int, c_int, c_i64, c_i64, c_i64, c_void_p]),
    'mg_segments_bwd': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_i64, c_i64, c_i64, c_void_p]),
    'mg_normalise_f32': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_i64, c_int, c_i64, c_void_p]),
    'mg_masked_reduce_workspace_bytes': (c_i64, [c_int, c_int, c_i64]),
    'mg_masked_reduce': (c_int, [ctypes.POINTER(Term), c_int, c_void_p, c_int, c_i64, c_void_p, c_i64, c_void_p]),
    'mg_masked_objective_f32': (c_int, [c_void_p, c_i64, c_i64, c_void_p, c_i64, c_i64, c_void_p, c_i64, c_i64, c_void_p,
                                        c_void_p, c_int, ctypes.POINTER(Slot), c_int, c_void_p, c_int, c_i64, c_void_p,
                                        c_i64, c_void_p]),
    'mg_objective_stream_plan': (c_int, [ctypes.POINTER(c_i64), c_int, c_i64, c_int, c_int, c_int, c_int, ctypes.POINTER(c_i64), c_i64]),
    'mg_ema_update_f32': (c_int, [ctypes.POINTER(c_void_p), ctypes.POINTER(c_void_p), ctypes.POINTER(c_i64), c_int,
                                  c_f32, c_void_p]),
    'mg_mlpg_workspace_bytes': (c_i64, [c_int, c_i64, c_int, c_int]),
    'mg_mlpg_f32': (c_int, [c_void_p, c_i64, c_i64, c_void_p, c_i64, c_i64, c_void_p, c_void_p, c_i64, c_i64, c_int, c_i64, c_int,
                            c_int, ctypes.POINTER(ctypes.c_double), c_int, c_void_p, c_i64, c_void_p]),
    'mg_linear_bf16': (c_int, [c_void_p, c_i64, c_void_p, c_i64, c_void_p, c_void_p, c_i64, c_int, c_int, c_int, c_int,
                               c_int, c_void_p]),
    'mg_act_grad_workspace_bytes': (c_i64, [c_i64, c_int]),
    'mg_act_grad_bf16': (c_int, [c_void_p, c_int, c_i64, c_void_p, c_int, c_i64, c_void_p, c_i64, c_void_p, c_i64, c_int,
                                 c_void_p, c_i64, c_void_p]),
    'mg_linear_wgrad_workspace_bytes': (c_i64, [c_i64, c_int, c_int]),
    'mg_linear_wgrad_plan': (c_int, [c_i64, c_int, c_int, ctypes.POINTER(c_i64)]),
    'mg_linear_wgrad_bf16': (c_int, [c_void_p, c_i64, c_void_p, c_i64, c_void_p, c_i64, c_i64, c_int, c_int, c_void_p, c_i64,
                                     c_void_p]),
    'mg_kld_workspace_bytes': (c_i64, [c_i64]),
    'mg_kld_standard_normal_f32': (c_int, [c_void_p, c_void_p, c_i64, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_i64,
                                           c_void_p]),
    'mg_both_nonzero_u8': (c_int, [ctypes.POINTER(c_void_p), c_int, c_i64, c_void_p, c_void_p]),
    'mg_cast_pad_bf16': (c_int, [c_void_p, c_i64, c_void_p, c_i64, c_i64, c_int, c_void_p]),
    'mg_cast_transpose_bf16': (c_int, [c_void_p, c_i64, c_void_p, c_i64, c_int, c_int, c_void_p]),
}

LIB_PATH = _build.LIB_PATH


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            'morgana_b200: the CUDA library {} has not been built and there is no CPU fallback. '
            'Run `python -m morgana_b200.build` (needs nvcc; cross-compiles for sm_100a without a GPU).'.format(LIB_PATH))
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in PROTOTYPES.items():
        fn = getattr(lib, name)   # AttributeError if the library does not export what the header declares
        fn.restype, fn.argtypes = restype, argtypes
    if lib.mg_abi_version() != 1:
        raise ImportError('morgana_b200: ABI version mismatch, rebuild with `python -m morgana_b200.build --force`')
    return lib


lib = _load()


class MorganaB200Error(RuntimeError):
    pass


def check(status, what):
    """Map a C status code to a Python exception."""
    if status == MG_OK:
        return
    message = lib.mg_last_error().decode('utf-8', 'replace')
    if status == MG_ERR_INVALID_ARG:
        raise ValueError('{}: {}'.format(what, message))
    if status == MG_ERR_UNSUPPORTED:
        raise NotImplementedError('{}: {}'.format(what, message))
    raise MorganaB200Error('{}: {}'.format(what, message))
