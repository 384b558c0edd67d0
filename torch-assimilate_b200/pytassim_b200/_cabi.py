"""ctypes binding of libb200da.so — the C ABI declared in include/b200da.h.

There is no CPU fallback: if the shared library is missing or no sm_100 device is present every entry
point raises.  Build the library with ``python -c "import __graft_entry__ as g; g.build()"``.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "libb200da.so")

OK = 0
ERR_INVALID, ERR_SIZE, ERR_UNSUPPORTED, ERR_NO_DEVICE, ERR_CUDA, ERR_STATE, ERR_NOMEM, ERR_OVERFLOW = -1, -2, -3, -4, -5, -6, -7, -8

METRIC_ABS1D, METRIC_PERIODIC1D, METRIC_EUCLID, METRIC_HAVERSINE = 0, 1, 2, 3
TAPER_GC, TAPER_GCINF = 0, 1
F64, F32 = 0, 1
SOLVER_NEWTON_SCHULZ, SOLVER_JACOBI = 0, 1
# b200da_kernel_op
(KOP_LINEAR, KOP_GAUSS, KOP_POLY, KOP_TANH, KOP_RATIONAL, KOP_SCALE, KOP_DIAG, KOP_ADD, KOP_MUL, KOP_POW) = range(10)
MAX_KERNEL_OPS = 16

_c = ctypes
_vp, _i, _i64, _dbl = _c.c_void_p, _c.c_int, _c.c_int64, _c.c_double
_dp = _c.POINTER(_c.c_double)

# name -> (restype, argtypes); must list every symbol declared in include/b200da.h
SIGNATURES = {
    "b200da_plan_create": (_i, [_c.POINTER(_vp), _i, _i, _i, _i, _dp, _i, _dp, _i, _dbl, _dbl, _i, _i]),
    "b200da_plan_destroy": (None, [_vp]),
    "b200da_plan_set_extra": (_i, [_vp, _i, _dp]),
    "b200da_plan_set_kernel": (_i, [_vp, _i, _c.POINTER(_c.c_int), _dp, _dp]),
    "b200da_set_grid": (_i, [_vp, _vp, _i64, _vp]),
    "b200da_bin_obs": (_i, [_vp, _vp, _vp, _vp, _i64, _vp]),
    "b200da_obs_prep": (_i, [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp]),
    "b200da_obs_gather_prep": (_i, [_vp, _vp, _vp, _i64, _vp, _vp, _i64, _vp, _vp, _vp]),
    "b200da_num_blocks": (_i64, [_vp]),
    "b200da_num_grid": (_i64, [_vp]),
    "b200da_num_obs": (_i64, [_vp]),
    "b200da_block_offset": (_i64, [_vp, _i64]),
    "b200da_grid_order": (_i, [_vp, _vp, _vp]),
    "b200da_letkf": (_i, [_vp, _vp, _vp, _vp, _i64, _i64, _vp, _vp]),
    "b200da_letkf_gram": (_i, [_vp, _vp, _i64, _i64, _vp]),
    "b200da_letkf_ienks": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _dbl, _dbl, _i64, _i64, _vp]),
    "b200da_etkf_ienks_weights": (_i, [_vp, _vp, _vp, _i64, _vp, _dbl, _dbl, _vp, _vp]),
    "b200da_letkf_host": (_i, [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp]),
    "b200da_letkf_host_blocks": (_i, [_vp, _vp, _i64, _i64, _vp]),
    "b200da_neighbour_count": (_i, [_vp, _vp, _vp, _vp]),
    "b200da_neighbour_fill": (_i, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "b200da_neighbour_ambiguous": (_i, [_vp, _i64, _vp, _vp, _vp, _vp, _vp]),
    "b200da_pending_status": (_i, [_vp, _i64, _vp, _vp, _vp, _vp, _vp]),
    "b200da_plan_set_overrides": (_i, [_vp, _i64, _vp, _vp, _vp, _vp]),
    "b200da_blocks_of_grid": (_i, [_vp, _i64, _vp, _vp, _vp]),
    "b200da_etkf_weights": (_i, [_vp, _vp, _vp, _i64, _vp, _vp]),
    "b200da_apply_weights": (_i, [_vp, _vp, _vp, _i, _i64, _vp, _vp]),
    "b200da_etkf_gram": (_i, [_vp, _vp, _vp, _i64, _i64, _vp, _vp]),
    "b200da_etkf_weights_from_gram": (_i, [_vp, _vp, _i64, _vp, _vp]),
    "b200da_apply_weights_cols": (_i, [_vp, _vp, _vp, _i, _i64, _i64, _i64, _vp, _vp]),
    "b200da_peer_copy_cols": (_i, [_vp, _vp, _i64, _i64, _i64, _i64, _i, _vp]),
    "b200da_apply_weights_cols_peers": (_i, [_vp, _vp, _vp, _i64, _i64, _i64, _vp, _i, _vp, _vp]),
    "b200da_pack_columns": (_i, [_vp, _vp, _i64, _i64, _vp, _i64, _vp]),
    "b200da_unpack_columns": (_i, [_vp, _vp, _i64, _i64, _vp, _i64, _vp]),
    "b200da_block_offsets": (_i, [_vp, _vp]),
    "b200da_gram_extra_rows": (_i, [_vp]),
    "b200da_strerror": (_c.c_char_p, [_i]),
    "b200da_last_cuda_error": (_c.c_char_p, []),
    "b200da_version": (_i, []),
    "b200da_launch_count": (_i64, []),
    "b200da_kernel_name": (_c.c_char_p, [_vp]),
    "b200da_enable_timing": (_i, [_vp, _i]),
    "b200da_last_kernel_ms": (_c.c_float, [_vp]),
    "b200da_last_phase_ms": (_c.c_float, [_vp, _i]),
    "b200da_collect_stats": (_i, [_vp, _i]),
    "b200da_get_stats": (_i, [_vp, _c.POINTER(_c.c_int64)]),
    "b200da_set_solver": (_i, [_vp, _i]),
}

_lib = None


def load():
    """Load libb200da.so (once) and attach the prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libb200da.so not found at {0}: the B200 LETKF engine has no CPU fallback; build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'`".format(LIB_PATH))
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class B200DAError(RuntimeError):
    pass


def check(status):
    """Map a b200da_status to the exception convention of the reference."""
    if status == OK:
        return
    lib = load()
    msg = lib.b200da_strerror(status).decode()
    if status == ERR_SIZE:
        raise ValueError(msg)                      # pytassim/core/base.py:33-38
    if status == ERR_INVALID:
        raise ValueError(msg)
    if status == ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    if status == ERR_CUDA:
        raise B200DAError(msg + ": " + lib.b200da_last_cuda_error().decode())
    raise B200DAError(msg)
