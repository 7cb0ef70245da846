"""ctypes binding of the C-ABI library ``contrack_b200/lib/libcontrack_b200.so`` (``include/contrack_b200.h``).

The library is the product path: there is no Python/CPU fallback.  If the shared object is missing, or a compute entry
point is called on a machine without a B200, the call raises -- it never silently computes somewhere else.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'lib', 'libcontrack_b200.so')

CT_OK, CT_ERR_ARG, CT_ERR_CUDA, CT_ERR_CAPACITY, CT_ERR_NEARTIE, CT_ERR_INTERNAL, CT_ERR_COMM = 0, -1, -2, -3, -4, -5, -6
CT_F32, CT_F64 = 0, 1
CT_GE, CT_LE, CT_GT, CT_LT = 0, 1, 2, 3
STAGE_FINAL, STAGE_LABEL2D, STAGE_SEAM2D, STAGE_FILTERED, STAGE_LABEL3D = 0, 1, 2, 3, 4

# gorl spellings accepted by run_contrack (contrack/contrack.py:649-656, 664-671)
GORL_TO_OP = {'>=': CT_GE, 'ge': CT_GE, '<=': CT_LE, 'le': CT_LE, '>': CT_GT, 'gt': CT_GT, '<': CT_LT, 'lt': CT_LT}


class ContrackLibError(RuntimeError):
    def __init__(self, code, message):
        super().__init__('contrack_b200 [%d]: %s' % (code, message))
        self.code = code


_p = C.c_void_p
_i32p, _u32p, _f64p, _i64p, _u8p = (C.POINTER(C.c_int32), C.POINTER(C.c_uint32), C.POINTER(C.c_double),
                                    C.POINTER(C.c_int64), C.POINTER(C.c_uint8))
_longp = C.POINTER(C.c_long)

_PROTOS = {
    'ct_version': (C.c_int, []),
    'ct_last_error': (C.c_char_p, []),
    'ct_create': (C.c_int, [C.c_int, C.POINTER(_p)]),
    'ct_destroy': (None, [_p]),
    'ct_set_option': (C.c_int, [_p, C.c_char_p, C.c_long]),
    'ct_run_contrack': (C.c_int, [_p, _p, C.c_int, C.c_long, C.c_int, C.c_int, _f64p, _f64p, C.c_long, C.c_int, C.c_int,
                                  C.c_double, C.c_int, C.c_int, _p, _longp, C.c_int, _p]),
    'ct_run_contrack_host': (C.c_int, [_p, _p, C.c_int, C.c_long, C.c_int, C.c_int, _f64p, _f64p, C.c_long, C.c_int,
                                       C.c_int, C.c_double, C.c_int, C.c_int, _p, _longp, C.c_long]),
    'ct_get_stat': (C.c_double, [_p, C.c_char_p]),
    'ct_track_tables': (C.c_int, [C.c_long, C.c_int, C.c_int, C.c_int, C.c_long] + [_i32p] * 6 + [C.c_long] + [_i32p] * 5
                        + [_i64p, _i32p, _i32p, _i32p, _i32p, C.c_long] + [_i32p] * 5 + [_longp] * 4),
    'ct_host_tables': (C.c_int, [C.c_long, C.c_int, C.c_int, _f64p, C.c_double, C.c_int, C.c_int, C.c_int,
                                 C.c_long] + [_i32p] * 5 + [_u32p, _f64p, _f64p, _u32p]
                       + [C.c_long, _u32p, _u32p, _u32p, _u32p, _f64p, _f64p]
                       + [C.c_long, _u32p, _u32p, _u32p]
                       + [_i64p, _i32p, _i32p, _i32p, _u32p]
                       + [_i32p, C.c_long] + [_i32p] * 5 + [_longp, _longp]),
    'ct_host_tables_fast': (C.c_int, [C.c_long, C.c_int, C.c_int, _f64p, C.c_double, C.c_int, C.c_int, C.c_int, C.c_long]
                            + [_i32p] * 5 + [_u32p] + [_f64p] * 4 + [_u32p] * 5 + [_f64p] * 2
                            + [C.c_long] + [_i32p] * 3 + [_u32p] * 2 + [_p, _p]
                            + [_i32p, C.c_long] + [_i32p] * 5 + [_longp, _longp]),
    'ct_nccl_unique_id': (C.c_int, [C.POINTER(C.c_ubyte)]),
    'ct_comm_init_nccl': (C.c_int, [C.POINTER(C.c_ubyte), C.c_int, C.c_int, C.c_int, C.POINTER(_p)]),
    'ct_comm_from_nccl': (C.c_int, [_p, C.c_int, C.c_int, C.POINTER(_p)]),
    'ct_comm_init_local': (C.c_int, [C.c_int, C.POINTER(_p)]),
    'ct_comm_destroy': (None, [_p]),
    'ct_comm_rank': (C.c_int, [_p]),
    'ct_comm_size': (C.c_int, [_p]),
    'ct_run_contrack_sharded': (C.c_int, [_p, _p, _p, C.c_int, C.c_long, C.c_long, C.c_long, C.c_int, C.c_int, _f64p, _f64p,
                                          C.c_long, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, _p, _longp, _p]),
    'ct_quantile_time': (C.c_int, [_p, _p, C.c_long, C.c_int, C.c_int, C.c_int, C.c_int, _f64p, C.c_int, _p, _p]),
    'ct_run_contrack_sharded_host': (C.c_int, [_p, _p, _p, C.c_int, C.c_long, C.c_long, C.c_long, C.c_int, C.c_int, _f64p, _f64p,
                                               C.c_long, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, _p, _longp, C.c_long]),
    'ct_quantile_time_t': (C.c_int, [_p, _p, _p, C.c_int, C.c_long, C.c_int, C.c_int, C.c_int, C.c_int, _f64p, C.c_int, _p, _p]),
    'ct_flag_count': (C.c_int, [_p, _p, C.c_long, C.c_int, C.c_int, C.c_int, _p, _p]),
    'ct_divide_f32': (C.c_int, [_p, _p, C.c_size_t, C.c_float, _p, _p]),
    'ct_gather_planes': (C.c_int, [_p, _p, C.c_int, C.c_int, C.c_int, _i32p, _i32p, C.c_int, C.c_int, _p, _p]),
    'ct_classify_rows': (None, [_f64p, C.c_int, C.c_int, _u8p]),
    'ct_numpy_pairwise_sum_rle': (C.c_double, [_f64p, _i64p, C.c_long]),
    'ct_calc_clim': (C.c_int, [_p, _p, C.c_long, C.c_int, C.c_int, _i32p, C.c_int, C.c_int, _p, _p]),
    'ct_calc_anom': (C.c_int, [_p, _p, C.c_long, C.c_int, C.c_int, _i32p, C.c_int, _p, C.c_int, _p, _p]),
    'ct_calc_clim_t': (C.c_int, [_p, _p, C.c_int, C.c_long, C.c_int, C.c_int, _i32p, C.c_int, C.c_int, _p, _p]),
    'ct_calc_anom_t': (C.c_int, [_p, _p, C.c_int, C.c_long, C.c_int, C.c_int, _i32p, C.c_int, _p, C.c_int, _p, _p]),
    'ct_gather_planes_t': (C.c_int, [_p, _p, C.c_int, C.c_int, C.c_int, C.c_int, _i32p, _i32p, C.c_int, C.c_int, _p, _p]),
    'ct_run_lifecycle': (C.c_int, [_p, _p, _p, C.c_int, C.c_long, C.c_int, C.c_int, _f64p, _longp, _p]),
    'ct_lifecycle_fetch': (C.c_int, [_p, C.c_long] + [_i32p] * 4 + [_f64p] * 5),
}

_lib = None


def load():
    """Load the shared library (once) and attach prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ContrackLibError(CT_ERR_INTERNAL, 'shared library not built: %s (run `make` or __graft_entry__.build())'
                               % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _PROTOS.items():
        fn = getattr(lib, name)          # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def exported_symbols():
    return sorted(_PROTOS)


def check(rc):
    if rc != CT_OK:
        raise ContrackLibError(rc, load().ct_last_error().decode('utf-8', 'replace'))


def ptr(a, typ):
    """ctypes pointer of the given type to a C-contiguous numpy array (None -> NULL)."""
    if a is None:
        return None
    return a.ctypes.data_as(typ)
