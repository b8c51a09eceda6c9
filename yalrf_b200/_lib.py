"""ctypes binding of libyhb.so (include/yhb.h).  There is no fallback: a missing library is an error."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('YHB_LIB', os.path.join(_HERE, 'libyhb.so'))  # YHB_LIB: A/B builds of the same ABI

# enums of include/yhb.h
LIN_RESISTOR, LIN_CAPACITOR, LIN_INDUCTOR, LIN_GYRATOR, LIN_IHF = range(5)
SRC_DC, SRC_AC1, SRC_AC2, SRC_DC_ONLY = range(4)
DIODE_PARAMS = ['Temp', 'Is', 'N', 'Isr', 'Nr', 'Ikf', 'Bv', 'Ibv', 'Rs', 'Area', 'Cj0', 'Vj', 'M', 'Fc', 'Tt', 'Cp']
BJT_PARAMS = ['Temp', 'Is', 'Nf', 'Nr', 'Ikf', 'Ikr', 'Vaf', 'Var', 'Ise', 'Ne', 'Isc', 'Nc', 'Bf', 'Br',
              'Cje', 'Vje', 'Mje', 'Cjc', 'Vjc', 'Mjc', 'Xcjc', 'Cjs', 'Vjs', 'Mjs', 'Fc', 'Tf', 'Xtf', 'Vtf',
              'Itf', 'Tr', 'Area', 'type']
DIODE_NPAR = len(DIODE_PARAMS)
BJT_NPAR = len(BJT_PARAMS)
FLAG_PNP_EXTENSION = 1
FLAG_EXACT_JACOBIAN = 2
FLAG_NO_PEEL = 4
FLAG_STAGED = 8
FLAG_NO_BAND = 16
FLAG_BLOCK_THOMAS = 32
FLAG_NO_BLOCK_THOMAS = 64
FLAG_NO_BLOCK_SPARSE = 128
FLAG_BLOCK_SPARSE = 256
FLAG_DIODE_CHARGE = 512
ENGINES = ['fused', 'staged-dense', 'staged-banded', 'staged-block-thomas']
MAX_LEVELS = 64

_i32p = C.POINTER(C.c_int32)
_f64p = C.POINTER(C.c_double)


class PlanDesc(C.Structure):
    _fields_ = [('num_nodes', C.c_int32), ('num_tones', C.c_int32), ('K1', C.c_int32), ('K2', C.c_int32),
                ('num_lin', C.c_int32), ('lin_kind', _i32p), ('lin_nodes', _i32p),
                ('num_src', C.c_int32), ('src_kind', _i32p), ('src_nodes', _i32p),
                ('num_diode', C.c_int32), ('diode_nodes', _i32p),
                ('num_bjt', C.c_int32), ('bjt_nodes', _i32p),
                ('num_cubic', C.c_int32), ('cubic_nodes', _i32p),
                ('nl_kind', _i32p), ('nl_index', _i32p),
                ('idft', _f64p), ('dft', _f64p), ('flags', C.c_int32)]


class Params(C.Structure):
    _fields_ = [('batch', C.c_int32), ('tones', C.c_void_p), ('lin_val', C.c_void_p), ('src_val', C.c_void_p),
                ('diode_par', C.c_void_p), ('bjt_par', C.c_void_p), ('cubic_par', C.c_void_p)]


class Options(C.Structure):
    _fields_ = [('reltol', C.c_double), ('abstol', C.c_double), ('maxiter', C.c_int32), ('reserved', C.c_int32)]


class Result(C.Structure):
    """yhb_result; struct_size is filled in by the constructor (the library refuses a struct without it)"""
    _fields_ = [('struct_size', C.c_size_t), ('V', C.c_void_p), ('Vf', C.c_void_p), ('converged', C.c_void_p), ('newton_iters', C.c_void_p),
                ('lu_count', C.c_void_p), ('level_iters', C.c_void_p), ('level_alpha', C.c_void_p),
                ('err', C.c_void_p), ('Inl', C.c_void_p), ('Ic', C.c_void_p)]

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.struct_size = C.sizeof(Result)


class OscOptions(C.Structure):
    """yhb_osc_options"""
    _fields_ = [('struct_size', C.c_size_t), ('probe_src', C.c_int32), ('probe_filter', C.c_int32), ('node', C.c_int32),
                ('max_outer', C.c_int32), ('tol', C.c_double), ('max_df_rel', C.c_double), ('max_dv_rel', C.c_double)]

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.struct_size = C.sizeof(OscOptions)


class OscResult(C.Structure):
    """yhb_osc_result"""
    _fields_ = [('struct_size', C.c_size_t), ('fosc', C.c_void_p), ('Vosc', C.c_void_p), ('yosc', C.c_void_p),
                ('converged', C.c_void_p), ('outer_iters', C.c_void_p), ('hb_newton_iters', C.c_void_p), ('diag', C.c_void_p)]

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.struct_size = C.sizeof(OscResult)


_lib = None

PROF_KINDS = ['prepare', 'dc', 'init', 'eval', 'assemble', 'lu', 'finish', 'fused']

EXPORTS = ['yhb_last_error', 'yhb_version', 'yhb_set_device', 'yhb_profile_enable', 'yhb_profile_reset',
           'yhb_profile_read', 'yhb_fp64_peak', 'yhb_plan_create', 'yhb_plan_destroy',
           'yhb_plan_dims', 'yhb_plan_core', 'yhb_plan_band', 'yhb_plan_engine', 'yhb_plan_block_sparse', 'yhb_plan_flops', 'yhb_plan_freqs', 'yhb_workspace_bytes', 'yhb_dc_solve', 'yhb_solve', 'yhb_solve_cold', 'yhb_newton',
           'yhb_run_host', 'yhb_solve_oscillator', 'yhb_oscillator_host', 'yhb_to_time', 'yhb_stage_ifft', 'yhb_stage_device_eval', 'yhb_stage_assemble', 'yhb_stage_lu', 'yhb_stage_gj', 'yhb_stage_gjp', 'yhb_debug_timing']


def lib():
    """Load libyhb.so once.  Raises if it was not built (run __graft_entry__.build() or csrc/build.sh)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError('yalrf_b200: %s is missing -- build it with yalrf_b200/csrc/build.sh; there is no '
                           'CPU fallback' % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    L.yhb_last_error.restype = C.c_char_p
    L.yhb_version.restype = C.c_int
    L.yhb_set_device.argtypes = [C.c_int32]
    L.yhb_profile_enable.argtypes = [C.c_int32]
    L.yhb_profile_read.argtypes = [C.POINTER(C.c_int64), C.POINTER(C.c_double)]
    L.yhb_fp64_peak.argtypes = [C.POINTER(C.c_double), C.c_void_p]
    L.yhb_plan_create.argtypes = [C.POINTER(PlanDesc), C.POINTER(C.c_void_p)]
    L.yhb_plan_destroy.argtypes = [C.c_void_p]
    L.yhb_plan_destroy.restype = None
    L.yhb_plan_dims.argtypes = [C.c_void_p, _i32p, _i32p, _i32p, _i32p]
    L.yhb_plan_core.argtypes = [C.c_void_p, _i32p, _i32p, _i32p]
    L.yhb_plan_band.argtypes = [C.c_void_p]
    L.yhb_plan_engine.argtypes = [C.c_void_p]
    L.yhb_plan_block_sparse.argtypes = [C.c_void_p, _i32p]
    L.yhb_plan_flops.argtypes = [C.c_void_p, _f64p, _f64p, _f64p]
    L.yhb_plan_freqs.argtypes = [C.c_void_p, C.c_double, C.c_double, _f64p]
    L.yhb_workspace_bytes.argtypes = [C.c_void_p, C.c_int32]
    L.yhb_workspace_bytes.restype = C.c_size_t
    L.yhb_dc_solve.argtypes = [C.c_void_p, C.POINTER(Params), C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p,
                               C.c_void_p]
    L.yhb_solve.argtypes = [C.c_void_p, C.POINTER(Params), C.POINTER(Options), C.c_void_p, C.c_size_t, C.c_int32,
                            C.c_void_p, C.POINTER(Result), C.c_void_p]
    L.yhb_solve_cold.argtypes = [C.c_void_p, C.POINTER(Params), C.POINTER(Options), C.c_void_p, C.c_size_t, C.c_void_p,
                                 C.c_void_p, C.POINTER(Result), C.c_void_p]
    L.yhb_newton.argtypes = [C.c_void_p, C.POINTER(Params), C.POINTER(Options), C.c_void_p, C.c_size_t,
                             C.c_void_p, C.POINTER(Result), C.c_void_p]
    L.yhb_run_host.argtypes = [C.c_void_p, C.POINTER(Params), C.POINTER(Options), C.c_void_p, C.POINTER(Result)]
    L.yhb_solve_oscillator.argtypes = [C.c_void_p, C.POINTER(Params), C.POINTER(Options), C.POINTER(OscOptions), C.c_void_p,
                                       C.c_size_t, C.POINTER(Result), C.POINTER(OscResult), C.c_void_p]
    L.yhb_oscillator_host.argtypes = [C.c_void_p, C.POINTER(Params), C.POINTER(Options), C.POINTER(OscOptions),
                                      C.POINTER(Result), C.POINTER(OscResult)]
    L.yhb_to_time.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                              C.c_void_p]
    L.yhb_stage_ifft.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    L.yhb_stage_device_eval.argtypes = [C.c_void_p, C.POINTER(Params), C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p]
    L.yhb_stage_assemble.argtypes = [C.c_void_p, C.POINTER(Params), C.c_void_p, C.c_size_t, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.yhb_stage_gj.argtypes = [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.yhb_stage_gjp.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.yhb_stage_lu.argtypes = [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.yhb_debug_timing.argtypes = [C.POINTER(C.c_ulonglong), C.c_int32]
    _lib = L
    return L


def profile_read():
    """{kernel kind: (launches, total ms)} since the last yhb_profile_reset"""
    n = (C.c_int64 * len(PROF_KINDS))()
    ms = (C.c_double * len(PROF_KINDS))()
    lib().yhb_profile_read(n, ms)
    return {k: (int(n[i]), float(ms[i])) for i, k in enumerate(PROF_KINDS)}


def check(rc):
    if rc != 0:
        raise RuntimeError('libyhb error %d: %s' % (rc, lib().yhb_last_error().decode()))


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def ptr(a):
    """raw address of a numpy array / torch tensor (or None)"""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a.data_ptr()
