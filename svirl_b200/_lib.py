"""ctypes binding of libsvirl_b200.so (the C ABI in include/svirl_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C svirl_b200/csrc``.
There is NO CPU fallback: if the library is missing or no CUDA device is present, the first
call raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsvirl_b200.so")

NODE_C, NODE_R, EDGE, CELL_R, CELL_B, FLAT = range(6)

_p, _d, _i, _u32, _sz = C.c_void_p, C.c_double, C.c_int, C.c_uint32, C.c_size_t
_pd = C.POINTER(C.c_double)

SIGNATURES = {
    "svl_version": ([], _i),
    "svl_create": ([C.POINTER(_p), _i, _i, _i, _d, _d, _i, _i, _i], _i),
    "svl_destroy": ([_p], _i),
    "svl_synchronize": ([_p], _i),
    "svl_set_option": ([_p, C.c_char_p, _i], _i),
    "svl_get_stat": ([_p, C.c_char_p, _pd], _i),
    "svl_slab_split_plan": ([_i, _i, _i], _i),
    "svl_debug_sincos": ([_p, _sz, _pd, _pd, _pd], _i),
    "svl_debug_trace": ([_p, C.POINTER(C.c_ulonglong), _i, C.POINTER(_i)], _i),
    "svl_event_record": ([_p, _i], _i),
    "svl_event_elapsed_ms": ([_p, _i, _i, _pd], _i),
    "svl_alloc": ([_p, _i, _sz, _i, C.POINTER(_p)], _i),
    "svl_free": ([_p, _p], _i),
    "svl_h2d": ([_p, _p, _p], _i),
    "svl_d2h": ([_p, _p, _p], _i),
    "svl_d2d": ([_p, _p, _p], _i),
    "svl_fill_zero": ([_p, _p], _i),
    "svl_swap": ([_p, _p, _p], _i),
    "svl_buf_size": ([_p], _sz),
    "svl_h2d_rows": ([_p, _p, _i, _i, _i, _p], _i),
    "svl_d2h_rows": ([_p, _p, _p, _i, _i, _i], _i),
    "svl_set_material": ([_p, _p], _i),
    "svl_td_psi_sweep": ([_p, _d, _d, _p, _p, _p, _p, _p, _d, _u32, _u32, _pd], _i),
    "svl_td_a_sweep": ([_p, _d, _d, _d, _d, _p, _p, _p, _p, _p, _d, _u32, _u32, _pd], _i),
    "svl_td_psi_solve": ([_p, _d, _d, _p, _p, _p, _d, _u32, _d, C.POINTER(_i)], _i),
    "svl_td_a_solve": ([_p, _d, _d, _d, _d, _p, _p, _d, _u32, _d, C.POINTER(_i)], _i),
    "svl_td_a_solve_ph": ([_p, _d, _d, _d, _d, _p, _p, _p, _d, _u32, _d, C.POINTER(_i)], _i),
    "svl_edge_axpy_flat": ([_p, _p, _p, _d, C.c_longlong], _i),
    "svl_phase_lock": ([_p, _p, _p, _i], _i),
    "svl_td_run": ([_p, _i, _d, _i, _d, _p, _d, _d, _d, _p, _p, _d, _d, C.POINTER(_u32), _d, _d,
                    C.POINTER(C.c_longlong)], _i),
    "svl_free_energy": ([_p, _d, _d, _p, _d, _p, _p, _p, _pd], _i),
    "svl_jacobian_psi": ([_p, _d, _d, _p, _d, _p, _p, _p, _p], _i),
    "svl_jacobian_A": ([_p, _d, _d, _p, _p, _p, _p], _i),
    "svl_cg_coef_psi": ([_p, _d, _d, _d, _p, _p, _p, _p, _pd], _i),
    "svl_cg_coef": ([_p, _d, _d, _d, _p, _p, _p, _p, _p, _pd], _i),
    "svl_cg_beta": ([_p, _p, _p, _pd], _i),
    "svl_axmy": ([_p, _p, _p, _p, _d], _i),
    "svl_axpy": ([_p, _p, _p, _p, _d], _i),
    "svl_cg_begin": ([_p, _i, _i, _d, _d, _p, _d, _p, _p, _p, _p, _p, _p, _p, _p, _p, _pd, _pd], _i),
    "svl_cg_end": ([_p, _i, _d, _d, _p, _d, _p, _p, _p, _p, _p, _d, _d, _pd], _i),
    "svl_cg_pass_a": ([_p, _i, _i, _i, _i, _d, _d, _p, _d, _p, _p, _p, _p, _p, _d, _d, _p, _p, _pd, _pd], _i),
    "svl_cg_pass_b": ([_p, _i, _d, _d, _d, _p, _p, _p, _p, _p, _p, _p, _pd], _i),
    "svl_cg_line_search": ([_pd, _i, _pd, C.POINTER(_i)], _i),
    "svl_magnetic_field": ([_p, _p, _p, _p], _i),
    "svl_current_density": ([_p, _d, _d, _p, _p, _p], _i),
    "svl_supercurrent_density": ([_p, _p, _p, _p, _p], _i),
    "svl_vortex_candidates": ([_p, _d, _p, _p, C.POINTER(C.c_int64), _pd, _sz, C.POINTER(_sz)], _i),
    "svl_slab_export": ([_p, _p, _p, _p], _i),
    "svl_slab_connect": ([_p, _p, _i, _p, _i], _i),
    "svl_slab_exchange": ([_p, _p], _i),
    "svl_slab_board_export": ([_p, _p], _i),
    "svl_slab_board_connect": ([_p, _i, _i, _p], _i),
    "svl_set_reduce_callback": ([_p, _p], _i),
    "svl_set_reduce_callback_device": ([_p, _p], _i),
    "svl_get_stream": ([_p], _p),
    "svl_mt19937_doubles": ([C.POINTER(_u32), C.POINTER(_i), C.c_ulonglong, _pd, C.c_ulonglong], _i),
    "svl_seeded_psi": ([_pd, _pd, C.c_ulonglong, _d, _p, _i], _i),
    "svl_sum": ([_p, _p, _sz, _pd], _i),
    "svl_sum_v": ([_p, _p, _sz, _i, _pd], _i),
}

_lib = None


class SvirlB200Error(RuntimeError):
    pass


def load():
    """Load the shared library (once) and declare every entry point."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SvirlB200Error(
            "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C svirl_b200/csrc` (there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    lib.svl_last_error.restype = C.c_char_p
    lib.svl_last_error.argtypes = []
    for name, (argtypes, restype) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = restype
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise SvirlB200Error(load().svl_last_error().decode("utf-8", "replace"))


def call(name, *args):
    check(getattr(load(), name)(*args))
