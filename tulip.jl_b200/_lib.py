"""ctypes binding of libtlpb200.so (the C ABI declared in include/tlpb200.h).

This is the Python twin of the Julia ``ccall`` glue in ``julia/TlpB200.jl``: it exists because
the container has no Julia runtime, so the parity tests and the benchmark drive the very same
shared library from Python.  There is NO fallback: a missing library raises ImportError, a
missing device raises ``TlpB200Error`` from the first numeric call.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtlpb200.so")

OK, NOT_POSDEF, OOM, BAD_ARG, CUDA, INTERNAL, NCCL = 0, 1, 2, 3, 4, 5, 6
K1, K2 = 1, 2


class Options(C.Structure):
    _fields_ = [("ordering", C.c_int32), ("device", C.c_int32), ("piece_width", C.c_int32),
                ("small_elems", C.c_int32), ("relax_always", C.c_int32), ("use_graph", C.c_int32),
                ("analyze_only", C.c_int32), ("rank", C.c_int32), ("nranks", C.c_int32),
                ("dense_col_threshold", C.c_int32), ("dense_solve_ncol", C.c_int32),
                ("ozaki_ncol", C.c_int32), ("refine_steps", C.c_int32), ("reserved", C.c_int32 * 3)]


class HSDOptions(C.Structure):
    _fields_ = [("iterations_limit", C.c_int32), ("correction_limit", C.c_int32), ("time_limit", C.c_double),
                ("tol_pfeas", C.c_double), ("tol_dfeas", C.c_double), ("tol_rgap", C.c_double), ("tol_ifeas", C.c_double),
                ("step_damp", C.c_double), ("gamma_min", C.c_double), ("centrality_outlier", C.c_double),
                ("preg_min", C.c_double), ("dreg_min", C.c_double)]


class HSDInfo(C.Structure):
    _fields_ = [("status", C.c_int32), ("niter", C.c_int32), ("pobj", C.c_double), ("dobj", C.c_double),
                ("rp_nrm", C.c_double), ("rl_nrm", C.c_double), ("ru_nrm", C.c_double), ("rd_nrm", C.c_double),
                ("rg_nrm", C.c_double), ("mu", C.c_double), ("tau", C.c_double), ("kappa", C.c_double),
                ("n_update", C.c_int64), ("n_solve", C.c_int64), ("ms_update", C.c_double), ("ms_solve", C.c_double),
                ("seconds_total", C.c_double)]


TRM_STATUS = ["Trm_Unknown", "Trm_Optimal", "Trm_PrimalInfeasible", "Trm_DualInfeasible", "Trm_IterationLimit",
              "Trm_TimeLimit", "Trm_NumericalProblem", "Trm_MemoryLimit"]


class Stats(C.Structure):
    _fields_ = [("m", C.c_int64), ("n", C.c_int64), ("nnzA", C.c_int64), ("order", C.c_int64),
                ("nnzL", C.c_int64), ("nnzL_stored", C.c_int64), ("flops", C.c_double),
                ("nsuper", C.c_int64), ("npieces", C.c_int64), ("nlevels", C.c_int64),
                ("max_ncol", C.c_int64), ("max_nrow", C.c_int64),
                ("nproducts", C.c_int64), ("nentries", C.c_int64),
                ("launches_update", C.c_int64), ("launches_solve", C.c_int64),
                ("ms_assemble", C.c_double), ("ms_factor", C.c_double), ("ms_solve", C.c_double),
                ("bad_pivot", C.c_int64), ("n_update", C.c_int64), ("n_solve", C.c_int64),
                ("bytes_device", C.c_int64),
                ("flops_update_inner", C.c_double), ("flops_update_ext", C.c_double),
                ("ms_class", C.c_double * 24), ("n_class", C.c_int64 * 24),
                ("flops_update_oz", C.c_double), ("oz_tasks", C.c_int64), ("oz_bytes", C.c_int64)]

    def asdict(self):
        d = {}
        for k, _ in self._fields_:
            v = getattr(self, k)
            d[k] = list(v) if hasattr(v, "__len__") else v
        return d


KERNEL_CLASSES = ["assemble", "small_factor", "diag_factor", "trsm", "update", "rhs_recover",
                  "fwd_small", "fwd_large", "update128", "bwd_large", "invert_diag", "bwd_small", "dense_cols",
                  "pack_big", "fwd_big", "bwd_big", "oz_slice", "oz_update", "comm", "refine"]


# every symbol include/tlpb200.h declares (tests check that the library exports all of them)
SYMBOLS = [
    "tlpb200_default_options", "tlpb200_create", "tlpb200_update", "tlpb200_update_dev",
    "tlpb200_update_status", "tlpb200_solve", "tlpb200_solve_dev", "tlpb200_solve_status", "tlpb200_debug_raise_timeout", "tlpb200_comm_unique_id", "tlpb200_comm_init", "tlpb200_comm_profile",
    "tlpb200_hsd_default_options", "tlpb200_hsd_create", "tlpb200_hsd_reset", "tlpb200_hsd_iterate", "tlpb200_hsd_optimize",
    "tlpb200_hsd_get_point", "tlpb200_hsd_get_log", "tlpb200_set_stream",
    "tlpb200_synchronize", "tlpb200_set_profiling", "tlpb200_stats_get", "tlpb200_get_symbolic",
    "tlpb200_get_structure", "tlpb200_debug_assemble", "tlpb200_debug_get_lx", "tlpb200_last_error",
    "tlpb200_backend_name", "tlpb200_linear_system", "tlpb200_destroy",
    "tlpb200_dist_info", "tlpb200_update_begin", "tlpb200_top_panels", "tlpb200_update_end",
    "tlpb200_solve_begin", "tlpb200_work_vector", "tlpb200_solve_mid", "tlpb200_solve_end",
    "tlpb200_get_dense_cols", "tlpb200_debug_big_plan", "tlpb200_debug_chain_times", "tlpb200_debug_factor_trace", "tlpb200_debug_ozaki", "tlpb200_debug_update_plan", "tlpb200_abi_sizes", "tlpb200_debug_solve_ops", "tlpb200_debug_phase_deps",
]

_lib = None


def load():
    """Load the shared library (once).  Fails loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `make -C tulip.jl_b200/csrc` or "
            "`python -c 'import __graft_entry__ as g; g.build()'`.  There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    p = C.c_void_p
    dp = C.POINTER(C.c_double)
    lib.tlpb200_default_options.argtypes = [C.POINTER(Options)]
    lib.tlpb200_default_options.restype = None
    lib.tlpb200_create.argtypes = [C.POINTER(p), C.c_int64, C.c_int64, C.POINTER(C.c_int64),
                                   C.POINTER(C.c_int64), dp, C.c_int, C.c_int, C.POINTER(Options)]
    lib.tlpb200_update.argtypes = [p, dp, dp, dp, C.POINTER(C.c_int64)]
    lib.tlpb200_update_dev.argtypes = [p, p, p, p]
    lib.tlpb200_update_status.argtypes = [p, C.POINTER(C.c_int64)]
    lib.tlpb200_solve.argtypes = [p, dp, dp, dp, dp, C.c_int32, C.c_int64, C.c_int64]
    lib.tlpb200_solve_dev.argtypes = [p, p, p, p, p, C.c_int32, C.c_int64, C.c_int64]
    lib.tlpb200_solve_status.argtypes = [p]
    lib.tlpb200_solve_status.restype = C.c_int
    lib.tlpb200_debug_raise_timeout.argtypes = [p]
    lib.tlpb200_debug_raise_timeout.restype = C.c_int
    lib.tlpb200_comm_unique_id.argtypes = [p]
    lib.tlpb200_comm_unique_id.restype = C.c_int
    lib.tlpb200_comm_init.argtypes = [p, p]
    lib.tlpb200_comm_init.restype = C.c_int
    lib.tlpb200_comm_profile.argtypes = [p, C.c_int32, C.POINTER(C.c_float), C.POINTER(C.c_int64)]
    lib.tlpb200_comm_profile.restype = C.c_int
    lib.tlpb200_hsd_default_options.argtypes = [C.POINTER(HSDOptions)]
    lib.tlpb200_hsd_default_options.restype = None
    lib.tlpb200_hsd_create.argtypes = [p, dp, dp, dp, dp, C.c_double]
    lib.tlpb200_hsd_reset.argtypes = [p]
    lib.tlpb200_hsd_iterate.argtypes = [p, C.POINTER(HSDOptions), C.POINTER(HSDInfo)]
    lib.tlpb200_hsd_optimize.argtypes = [p, C.POINTER(HSDOptions), C.POINTER(HSDInfo)]
    lib.tlpb200_hsd_get_point.argtypes = [p, dp, dp, dp, dp, dp, dp, dp]
    lib.tlpb200_hsd_get_log.argtypes = [p, dp, C.POINTER(C.c_int64)]
    for name in ("tlpb200_hsd_create", "tlpb200_hsd_reset", "tlpb200_hsd_iterate", "tlpb200_hsd_optimize",
                 "tlpb200_hsd_get_point", "tlpb200_hsd_get_log"):
        getattr(lib, name).restype = C.c_int
    lib.tlpb200_set_stream.argtypes = [p, p]
    lib.tlpb200_synchronize.argtypes = [p]
    lib.tlpb200_set_profiling.argtypes = [p, C.c_int]
    lib.tlpb200_stats_get.argtypes = [p, C.POINTER(Stats)]
    lib.tlpb200_get_symbolic.argtypes = [p, p, p, p, p]
    lib.tlpb200_get_structure.argtypes = [p, p, p]
    lib.tlpb200_debug_assemble.argtypes = [p, dp, dp, dp]
    lib.tlpb200_debug_get_lx.argtypes = [p, dp, C.POINTER(C.c_int64)]
    lib.tlpb200_dist_info.argtypes = [p, p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    lib.tlpb200_update_begin.argtypes = [p, dp, dp, dp]
    lib.tlpb200_top_panels.argtypes = [p, C.POINTER(p), C.POINTER(C.c_int64)]
    lib.tlpb200_update_end.argtypes = [p, C.POINTER(C.c_int64)]
    lib.tlpb200_solve_begin.argtypes = [p, dp, dp]
    lib.tlpb200_work_vector.argtypes = [p, C.POINTER(p), C.POINTER(C.c_int64)]
    lib.tlpb200_solve_mid.argtypes = [p]
    lib.tlpb200_solve_end.argtypes = [p, dp, dp]
    for name in ("tlpb200_dist_info", "tlpb200_update_begin", "tlpb200_top_panels", "tlpb200_update_end",
                 "tlpb200_solve_begin", "tlpb200_work_vector", "tlpb200_solve_mid", "tlpb200_solve_end"):
        getattr(lib, name).restype = C.c_int
    lib.tlpb200_get_dense_cols.argtypes = [p, C.POINTER(C.c_int32), C.POINTER(C.c_int64)]
    lib.tlpb200_get_dense_cols.restype = C.c_int
    lib.tlpb200_debug_big_plan.argtypes = [p, C.POINTER(C.c_int64), p, p, p]
    lib.tlpb200_debug_big_plan.restype = C.c_int
    lib.tlpb200_debug_chain_times.argtypes = [p, p, C.POINTER(C.c_int64)]
    lib.tlpb200_debug_chain_times.restype = C.c_int
    lib.tlpb200_debug_factor_trace.argtypes = [p, p, C.POINTER(C.c_int64)]
    lib.tlpb200_debug_factor_trace.restype = C.c_int
    lib.tlpb200_debug_solve_ops.argtypes = [p, C.POINTER(C.c_int64), p, p, p, p, p, p, p, p, p, p]
    lib.tlpb200_debug_solve_ops.restype = C.c_int
    lib.tlpb200_debug_phase_deps.argtypes = [p, C.c_int32, p, p, p]
    lib.tlpb200_debug_phase_deps.restype = C.c_int
    lib.tlpb200_abi_sizes.argtypes = [C.POINTER(C.c_int32)]
    lib.tlpb200_abi_sizes.restype = None
    lib.tlpb200_debug_update_plan.argtypes = [p, C.POINTER(C.c_int64), p, p, p, p, p, p, p, p, p]
    lib.tlpb200_debug_update_plan.restype = C.c_int
    lib.tlpb200_debug_ozaki.argtypes = [dp, C.c_int64, C.c_int64, dp, C.c_int32, C.c_int32, C.POINTER(C.c_float), C.POINTER(C.c_int32)]
    lib.tlpb200_debug_ozaki.restype = C.c_int
    lib.tlpb200_last_error.argtypes = [p]
    lib.tlpb200_last_error.restype = C.c_char_p
    lib.tlpb200_backend_name.argtypes = []
    lib.tlpb200_backend_name.restype = C.c_char_p
    lib.tlpb200_linear_system.argtypes = [p]
    lib.tlpb200_linear_system.restype = C.c_char_p
    lib.tlpb200_destroy.argtypes = [p]
    lib.tlpb200_destroy.restype = None
    for name in ("tlpb200_create", "tlpb200_update", "tlpb200_update_dev", "tlpb200_update_status",
                 "tlpb200_solve", "tlpb200_solve_dev", "tlpb200_set_stream", "tlpb200_synchronize",
                 "tlpb200_set_profiling", "tlpb200_stats_get", "tlpb200_get_symbolic",
                 "tlpb200_get_structure", "tlpb200_debug_assemble", "tlpb200_debug_get_lx"):
        getattr(lib, name).restype = C.c_int
    _lib = lib
    return lib
