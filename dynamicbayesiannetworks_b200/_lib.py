"""ctypes binding of the C ABI declared in include/dbn_b200.h.

There is no CPU fallback: if the shared library is missing, or no CUDA device is visible when an engine is
created, the call raises.  The library is built in-tree by `dynamicbayesiannetworks_b200.build`.
"""
import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
# DBN_B200_LIB points at an alternative build of the same library (kernel A/B experiments); never a CPU fallback
LIB_PATH = os.environ.get('DBN_B200_LIB') or os.path.join(PKG, 'libdbn_b200.so')

KMAX, PMAX, SMAX = 5, 4, 10
KTRI = KMAX * (KMAX + 1) // 2

# replay / trace record layouts (keep in sync with include/dbn_b200.h)
RD_SIGMA2, RD_LAMBDA2, RD_DELTA2, RD_U_PI, RD_U_TAU = 0, 1, 2, 3, 4
RD_MU_PI = 5
RD_MU_TAU = 5 + KMAX
RD_BETA = 5 + 2 * KMAX
RD_SIGMA2V = 5 + 2 * KMAX + SMAX * KMAX
RD_STRIDE = RD_SIGMA2V + SMAX
RI_PI_MOVE, RI_PI_PICK, RI_TAU_MOVE, RI_TAU_PICK, RI_STRIDE = 0, 1, 3, 4, 8
TD_SIG_SHAPE, TD_SIG_RATE, TD_LAM_SHAPE, TD_LAM_RATE, TD_DEL_SHAPE, TD_DEL_RATE = 0, 1, 2, 3, 4, 5
TD_LOGML, TD_SIGMA2, TD_LAMBDA2, TD_DELTA2, TD_MU, TD_MUS_LOGDENS, TD_MUS_MEAN = 6, 10, 11, 12, 13, 18, 22
TD_MUS_COV = 22 + 4 * KMAX
TD_BETA_MEAN = TD_MUS_COV + 4 * KTRI
TD_BETA_COV = TD_BETA_MEAN + SMAX * KMAX
TD_BETA = TD_BETA_COV + SMAX * KTRI
TD_SIGV_SHAPE = TD_BETA + SMAX * KMAX
TD_SIGV_RATE = TD_SIGV_SHAPE + SMAX
TD_SIGMA2V = TD_SIGV_RATE + SMAX
TD_STRIDE = TD_SIGMA2V + SMAX
TI_PI_VALID, TI_PI_ACCEPT, TI_TAU_VALID, TI_TAU_ACCEPT, TI_NPI, TI_PI, TI_S, TI_TAU = 0, 1, 2, 3, 4, 5, 9, 10
TI_PI_MOVE, TI_TAU_MOVE, TI_S_BEFORE, TI_STRIDE = 20, 21, 22, 24


def fp_rd_stride(k):
    """doubles per replay record of dbn_run_fp (include/dbn_b200.h)"""
    return 3 + k + SMAX * k + SMAX


def fp_td_stride(k):
    """doubles per trace record of dbn_run_fp"""
    return 10 + 3 * k + 3 * SMAX * k + 3 * SMAX


FP_RI_STRIDE, FP_TI_STRIDE = 4, 16


class DbnConfig(C.Structure):
    _fields_ = [('device', C.c_int32), ('n_chains', C.c_int32), ('chain_offset', C.c_int64), ('seed', C.c_uint64)]


class DbnRunParams(C.Structure):
    _fields_ = [('method', C.c_int32), ('fan_in', C.c_int32), ('with_tau', C.c_int32), ('num_samples', C.c_int32),
                ('T_ml', C.c_double), ('T_sigma', C.c_double),
                ('a_sigma', C.c_double), ('b_sigma', C.c_double), ('a_lambda', C.c_double), ('b_lambda', C.c_double),
                ('a_delta', C.c_double), ('b_delta', C.c_double), ('cp_p', C.c_double),
                ('burn_in', C.c_int32), ('thin', C.c_int32)]


# every symbol include/dbn_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
_dp, _ip, _lp, _up = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_uint64)
SYMBOLS = {
    'dbn_abi_version': (C.c_int, []),
    'dbn_create': (C.c_int, [C.POINTER(DbnConfig), C.POINTER(_P)]),
    'dbn_destroy': (None, [_P]),
    'dbn_last_error': (C.c_char_p, [_P]),
    'dbn_set_stream': (C.c_int, [_P, _P]),
    'dbn_sync': (C.c_int, [_P]),
    'dbn_build_stats': (C.c_int, [_P, _dp, _ip, C.c_int32, C.c_int32, C.c_int32]),
    'dbn_rebuild_stats': (C.c_int, [_P]),
    'dbn_stats_info': (C.c_int, [_P, _ip, _lp, _lp]),
    'dbn_read_stats_row': (C.c_int, [_P, C.c_int32, _dp]),
    'dbn_score_batch': (C.c_int, [_P, C.c_int32, C.c_double, C.c_double, C.c_double, C.c_int32,
                                  _ip, _ip, _ip, _ip, _ip, _dp, _dp, _dp, _dp]),
    'dbn_chain_init': (C.c_int, [_P, C.POINTER(DbnRunParams)]),
    'dbn_set_unit_range': (C.c_int, [_P, C.c_int64, C.c_int64]),
    'dbn_set_changepoints': (C.c_int, [_P, _ip, C.c_int32]),
    'dbn_run': (C.c_int, [_P, C.c_int32, _dp, _ip, _dp, _ip]),
    'dbn_run_fp': (C.c_int, [_P, C.c_int32, _dp, _ip, _dp, _ip]),
    'dbn_counts_size': (C.c_int, [_P, _lp]),
    'dbn_read_counts': (C.c_int, [_P, _lp]),
    'dbn_device_counts': (C.c_int, [_P, C.POINTER(_P)]),
    'dbn_read_state': (C.c_int, [_P, _ip, _ip, _ip, _ip, _dp, _dp]),
    'dbn_read_counters': (C.c_int, [_P, _up]),
    'dbn_reset_counters': (C.c_int, [_P]),
    'dbn_verify_cache': (C.c_int, [_P, C.POINTER(C.c_uint64)]),
    'dbn_measure_fp64_peak': (C.c_int, [_P, _dp]),
    'dbn_score_batch_time': (C.c_int, [_P, _dp]),
}

_lib = None


def load():
    """Load libdbn_b200.so (building nothing: run `python -m dynamicbayesiannetworks_b200.build` first)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            'dynamicbayesiannetworks_b200: %s is missing. Build it with `python -m dynamicbayesiannetworks_b200.build` '
            '(nvcc, sm_100a). There is no CPU fallback.' % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError here means the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    if lib.dbn_abi_version() != 1:
        raise RuntimeError('libdbn_b200.so ABI version mismatch')
    _lib = lib
    return lib
