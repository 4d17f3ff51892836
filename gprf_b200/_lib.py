"""ctypes binding of libgprf_b200.so (C-ABI: include/gprf_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` /
``gprf_b200/csrc/build.sh``.  A missing library is a hard error - there is no
CPU path in the product.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgprf_b200.so")
MAX_NCOV = 5
N_FAMILIES = 13

OK, ERR_NOT_PD, ERR_NONPOS_DIAG, ERR_ARG, ERR_CUDA, ERR_NO_STRUCTURE = range(6)

_lib = None

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_llp = C.POINTER(C.c_longlong)

_SIGNATURES = {
    "gprf_abi_version": (C.c_int, []),
    "gprf_strerror": (C.c_char_p, [C.c_int]),
    "gprf_last_error": (C.c_char_p, [C.c_void_p]),
    "gprf_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_longlong, C.c_int, C.c_int,
                              C.c_void_p, C.c_int, C.c_int]),
    "gprf_destroy": (C.c_int, [C.c_void_p]),
    "gprf_set_structure": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                     C.c_void_p, C.c_void_p]),
    "gprf_set_edges": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int]),
    "gprf_set_unit_mask": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "gprf_unit_lmul": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "gprf_llgrad_device_nosync": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gprf_set_x_prior": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gprf_neg_objective": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p]),
    "gprf_set_blocks": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "gprf_set_grid_partitioner": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]),
    "gprf_set_tree_partitioner": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_int, C.c_int, C.c_double, C.c_double, C.c_int]),
    "gprf_reblock": (C.c_int, [C.c_void_p, C.c_void_p]),
    "gprf_reblock_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "gprf_block_count": (C.c_int, [C.c_void_p, _ip, _llp]),
    "gprf_get_blocks": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "gprf_llgrad_reblock": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                      _dp, C.c_void_p, C.c_void_p, _ip]),
    "gprf_llgrad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                              _dp, C.c_void_p, C.c_void_p, _ip]),
    "gprf_llgrad_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                     C.c_void_p, C.c_void_p, _ip]),
    "gprf_unit_results": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "gprf_block_max_kernel": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "gprf_kernel_matrix": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong,
                                     C.c_void_p, C.c_int, C.c_void_p]),
    "gprf_kernel_deriv": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                    C.c_int, C.c_void_p]),
    "gprf_debug_unit": (C.c_int, [C.c_void_p, C.c_int, _ip, _ip, _ip, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gprf_last_timing": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), _ip]),
    "gprf_debug_trace": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "gprf_set_keep_kinv": (C.c_int, [C.c_void_p, C.c_int]),
    "gprf_set_fused_nt": (C.c_int, [C.c_void_p, C.c_int]),
    "gprf_set_factor_reuse": (C.c_int, [C.c_void_p, C.c_int]),
    "gprf_factor_reuse_stats": (C.c_int, [C.c_void_p, _ip, _llp]),
    "gprf_set_profiling": (C.c_int, [C.c_void_p, C.c_int]),
    "gprf_family_timing": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), _ip]),
    "gprf_family_name": (C.c_char_p, [C.c_int]),
    "gprf_set_resident": (C.c_int, [C.c_void_p, C.c_int]),
    "gprf_resident_stats": (C.c_int, [C.c_void_p, _llp, _llp, _ip]),
    "gprf_resident_layout": (C.c_int, [_llp, C.c_int]),
    "gprf_set_resident_debug": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "gprf_get_resident_debug": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
}

EXPORTS = sorted(_SIGNATURES)


def load():
    """Load the shared library (once) and declare every entry point."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s is missing - build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "or gprf_b200/csrc/build.sh (no CPU fallback exists)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def ptr(a):
    """Host pointer of a C-contiguous numpy array (or None)."""
    return None if a is None else a.ctypes.data          # plain integer: c_void_p arguments accept it
