"""ctypes loader for paml_b200/libpaml_b200.so (the C ABI declared in include/paml_b200.h).

The library is built in-tree by ``make`` / ``__graft_entry__.build()`` with nvcc for sm_100a.  There is no
Python or CPU fallback: if the shared object is missing, loading fails loudly.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PB200_LIB") or os.path.join(_HERE, "libpaml_b200.so")   # PB200_LIB: A/B builds while tuning

# every symbol include/paml_b200.h declares (tests check the library exports all of them)
SYMBOLS = [
    "pb200_create", "pb200_destroy", "pb200_last_error", "pb200_version", "pb200_set_data", "pb200_set_tree",
    "pb200_default_scale_flags", "pb200_eval", "pb200_eval_begin", "pb200_eval_end", "pb200_get_lnf", "pb200_get_fhK", "pb200_pmat",
    "pb200_nccl_unique_id", "pb200_comm_init", "pb200_p2p_local_handle", "pb200_p2p_init", "pb200_p2p_enable", "pb200_set_timing", "pb200_last_timing", "pb200_launch_count", "pb200_last_kernel",
    "pb200_plan_tree", "pb200_plan_tree_ps", "pb200_plan_tree4", "pb200_measure_fp64_peak", "pb200_flush_l2", "pb200_bench_resident", "pb200_bench_e2e", "pb200_debug_set_program",
    "pb200_compress_patterns", "pb200_eigen_qrev", "pb200_cond_update", "pb200_branch_lnl", "pb200_branch_newton", "pb200_cond_get",
]

_lib = None


class PB200Error(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PB200Error("%s not found: build it with `make` (nvcc, sm_100a); there is no CPU fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.pb200_last_error.restype = C.c_char_p
        L.pb200_version.restype = C.c_char_p
        L.pb200_launch_count.restype = C.c_longlong
        L.pb200_launch_count.argtypes = [C.c_void_p]
        L.pb200_last_kernel.restype = C.c_char_p
        L.pb200_last_kernel.argtypes = [C.c_void_p]
        L.pb200_destroy.argtypes = [C.c_void_p]
        L.pb200_destroy.restype = None
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise PB200Error(lib().pb200_last_error().decode())
