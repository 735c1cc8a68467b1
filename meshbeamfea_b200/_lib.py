"""ctypes binding of libmbfea.so (include/mbfea.h).

The library is built in-tree by ``__graft_entry__.build()`` /
``make -C meshbeamfea_b200/csrc``.  There is no fallback of any kind: if the
shared object is missing, or no sm_100-class CUDA device is visible, the error
surfaces to the caller.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# MBFEA_LIB selects an alternative build of the same ABI (kernel-tuning A/B runs)
LIB_PATH = os.environ.get("MBFEA_LIB") or os.path.join(_HERE, "libmbfea.so")

# mbfea_status
OK, EPRECOND, ERANGE_FIXED, ERANGE_FORCE, ECUDA, ESTATE, ENOCONV, ESINGULAR, EIO, EARG = range(10)


class Options(C.Structure):
    _fields_ = [
        ("device", C.c_int32), ("verbose", C.c_int32), ("tol", C.c_double), ("max_iter", C.c_int64),
        ("check_every", C.c_int32), ("spmv_kernel", C.c_int32), ("rank", C.c_int32), ("world", C.c_int32),
        ("use_graph", C.c_int32), ("comm_mode", C.c_int32), ("storage", C.c_int32),
        ("cg_variant", C.c_int32), ("batch_mode", C.c_int32), ("reorder", C.c_int32),
        ("precond", C.c_int32), ("ml_smooth_from", C.c_int32), ("ml_cell", C.c_double), ("ml_ratio", C.c_double),
        ("ml_coarsest_rows", C.c_int64),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("n_vertices", C.c_int64), ("n_beams", C.c_int64), ("n_fixed", C.c_int64),
        ("n_block_rows", C.c_int64), ("n_block_rows_global", C.c_int64), ("nnz_blocks", C.c_int64),
        ("n_ghost_blocks", C.c_int64), ("iterations", C.c_int64),
        ("rel_residual", C.c_double), ("true_rel_residual", C.c_double),
        ("ms_host_prep", C.c_double), ("ms_h2d", C.c_double),
        ("ms_pack", C.c_double), ("ms_assemble", C.c_double), ("ms_precond", C.c_double),
        ("ms_pcg", C.c_double), ("ms_recover", C.c_double), ("ms_d2h", C.c_double),
        ("ms_execute_total", C.c_double),
        ("bytes_spmv", C.c_double), ("bytes_pcg_iter", C.c_double), ("bytes_assemble", C.c_double),
        ("bytes_recover", C.c_double),
        ("kernel_launches", C.c_int64), ("converged", C.c_int32), ("restarts", C.c_int32),
        ("n_components", C.c_int32), ("precond", C.c_int32), ("ml_levels", C.c_int32), ("reserved", C.c_int32),
        ("ms_ml_symbolic", C.c_double), ("ms_ml_setup", C.c_double), ("bytes_ml_apply", C.c_double),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}


class PlanSizes(C.Structure):
    _fields_ = [(k, C.c_int64) for k in
                ("nb_global", "row_begin", "row_end", "nnz_blocks", "n_ghost", "n_send", "n_neighbors",
                 "n_components", "max_component_rows")]


_lib = None
_vp = C.c_void_p
_pd = C.POINTER(C.c_double)
_pf = C.POINTER(C.c_float)
_pi32 = C.POINTER(C.c_int32)
_pi64 = C.POINTER(C.c_int64)
_i64 = C.c_int64
_i32 = C.c_int32

# name -> (restype, argtypes); also the list the export test checks against the header
SIGNATURES = {
    "mbfea_version": (C.c_int, []),
    "mbfea_abi_sizes": (None, [_pi64]),
    "mbfea_options_default": (None, [C.POINTER(Options)]),
    "mbfea_create": (C.c_int, [C.POINTER(_vp), C.POINTER(Options)]),
    "mbfea_destroy": (None, [_vp]),
    "mbfea_last_error": (C.c_char_p, [_vp]),
    "mbfea_status_string": (C.c_char_p, [C.c_int]),
    "mbfea_set_job": (C.c_int, [_vp, _i64, _i64, _pd, _pi32, _pd, _pf, _pf, _i64, _pi32, _i64, _pi64, _pi64, _pd]),
    "mbfea_set_forces": (C.c_int, [_vp, _i64, _pi64, _pi64, _pd]),
    "mbfea_execute": (C.c_int, [_vp]),
    "mbfea_assemble": (C.c_int, [_vp]),
    "mbfea_solve": (C.c_int, [_vp]),
    "mbfea_recover": (C.c_int, [_vp]),
    "mbfea_get_displacements": (C.c_int, [_vp, _pd]),
    "mbfea_get_edge_forces": (C.c_int, [_vp, _pd]),
    "mbfea_get_stats": (C.c_int, [_vp, C.POINTER(Stats)]),
    "mbfea_get_force_ranges": (C.c_int, [_vp, _pd, _pd]),
    "mbfea_get_edge_forces_normalized": (C.c_int, [_vp, _pd]),
    "mbfea_write_displacements_csv": (C.c_int, [_vp, C.c_char_p]),
    "mbfea_write_element_forces_csv": (C.c_int, [_vp, C.c_char_p]),
    "mbfea_get_props": (C.c_int, [_vp, _pd]),
    "mbfea_get_dofmap": (C.c_int, [_vp, _pi32]),
    "mbfea_get_kelem": (C.c_int, [_vp, _pd]),
    "mbfea_get_bsr_sizes": (C.c_int, [_vp, _pi64, _pi64, _pi64, _pi64]),
    "mbfea_get_bsr": (C.c_int, [_vp, _pi64, _pi32, _pd, _pi64]),
    "mbfea_get_load": (C.c_int, [_vp, _pd]),
    "mbfea_spmv": (C.c_int, [_vp, _pd, _pd]),
    "mbfea_spmv_with": (C.c_int, [_vp, _i32, _pd, _pd]),
    "mbfea_recover_from": (C.c_int, [_vp, _pd]),
    "mbfea_time_spmv": (C.c_int, [_vp, _i32, _i32, _i32, _pd]),
    "mbfea_time_assemble": (C.c_int, [_vp, _i32, _i32, _pd]),
    "mbfea_debug_trace": (C.c_int, [_vp, C.POINTER(C.c_uint64), _i64, _pi64]),
    "mbfea_debug_compare_plan": (C.c_int, [_vp, _pi64, _pi32]),
    "mbfea_debug_compare_hierarchy": (C.c_int, [_vp, _pi64, _pi32, _pi32]),
    "mbfea_comm_unique_id": (C.c_int, [C.POINTER(C.c_uint8)]),
    "mbfea_comm_init": (C.c_int, [_vp, C.POINTER(C.c_uint8), _i32, _i32]),
    "mbfea_plan_partition": (C.c_int, [_i64, _i64, _pi32, _i64, _pi32, _i32, _i32, C.POINTER(PlanSizes),
                                       _pi64, _pi32, _pi64, _pi32, _pi64, _pi32]),
    "mbfea_load_scenario": (C.c_int, [_vp, C.c_char_p, C.c_char_p]),
    "mbfea_scenario_open": (C.c_int, [C.c_char_p, C.c_char_p, C.POINTER(_vp), C.c_char_p, _i64]),
    "mbfea_scenario_sizes": (None, [_vp, _pi64, _pi64, _pi64, _pi64]),
    "mbfea_scenario_copy": (None, [_vp, _pd, _pi32, _pd, _pf, _pf, _pi32, _pi64, _pi64, _pd]),
    "mbfea_scenario_close": (None, [_vp]),
    "mbfea_host_validate": (C.c_int, [_i64, _i64, _pi32, _pf, _pf, _i64, _pi32, _i64, _pi64, _pi64, C.c_char_p, _i64]),
    "mbfea_host_props": (C.c_int, [_i64, _pf, _pf, _pd]),
    "mbfea_host_free_map": (C.c_int, [_i64, _i64, _pi32, _pi32, _pi64]),
    "mbfea_host_renumber": (C.c_int, [_i64, _i64, _pi32, _i64, _pi32, _i32, _pi32, _pi32]),
    "mbfea_host_contributors": (C.c_int, [_i64, _i64, _pi32, _i64, _pi32, _i32, _i32, _pi64, _pi32, _pi32]),
    "mbfea_host_ml_hierarchy": (C.c_int, [_i64, _i64, _pd, _pi32, _i64, _pi32, C.c_double, C.c_double, _i64, _i32, _i32,
                                          _pi32, _pi64]),
    "mbfea_host_ml_image": (C.c_int, [_i64, _i64, _pd, _pi32, _i64, _pi32, _i32, _i32, _pi64, C.c_void_p]),
    "mbfea_host_spd_inverse": (C.c_int, [_i64, _pd, _pd]),
    "mbfea_hierarchy_open": (C.c_int, [_i64, _i64, _pd, _pi32, _i64, _pi32, C.c_double, C.c_double, _i32, _i64,
                                       C.POINTER(C.c_void_p)]),
    "mbfea_hierarchy_levels": (C.c_int, [C.c_void_p]),
    "mbfea_hierarchy_sizes": (C.c_int, [C.c_void_p, _i32, _pi64]),
    "mbfea_hierarchy_copy": (C.c_int, [C.c_void_p, _i32, _pi32, _pd, _pi32, _pi32, _pi32, _pi32]),
    "mbfea_hierarchy_close": (None, [C.c_void_p]),
    "mbfea_host_solver_format": (C.c_int, [_i64, _i64, _pd, _pi32, _i64, _pi32, _i32, _i32, _i32, _pi64, _pi32, _pi32,
                                           _pi32, _pi32, _pi32, _pi32, _pi32]),
}


def lib():
    """Load libmbfea.so (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C meshbeamfea_b200/csrc`. meshbeamfea_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        sizes = (C.c_int64 * 3)()
        L.mbfea_abi_sizes(sizes)
        mine = (C.sizeof(Options), C.sizeof(Stats), C.sizeof(PlanSizes))
        if tuple(sizes) != mine:
            raise ImportError(f"{LIB_PATH}: struct layouts {tuple(sizes)} do not match this binding {mine}; rebuild the library")
        _lib = L
    return _lib


def ptr(a, t):
    """ctypes pointer of type ``t`` to a numpy array's buffer (None -> NULL)."""
    if a is None:
        return None
    return a.ctypes.data_as(t)
