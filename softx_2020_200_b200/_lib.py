"""ctypes binding of libglsns.so (include/glsns.h). There is no CPU fallback: if the library is not
built, importing this module's `lib()` raises."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GLSNS_LIB", os.path.join(HERE, "libglsns.so"))   # (GLSNS_LIB: a differently tuned build, for experiments)

c_double_p = C.POINTER(C.c_double)
c_i32_p = C.POINTER(C.c_int32)
c_i64_p = C.POINTER(C.c_int64)
c_u8_p = C.POINTER(C.c_uint8)

OK, ERR_BAD_ARGUMENT, ERR_CUDA, ERR_NO_CONVERGENCE, ERR_ZERO_PIVOT, ERR_STATE, ERR_UNSUPPORTED, \
    ERR_COMM = range(8)
STATUS_NAMES = ["OK", "BAD_ARGUMENT", "CUDA", "NO_CONVERGENCE", "ZERO_PIVOT", "STATE", "UNSUPPORTED",
                "COMM"]

# glsns_scheme (same order as Parameters::SimulationControl::TimeSteppingMethod)
SCHEMES = {"steady": 0, "bdf1": 1, "bdf2": 2, "bdf3": 3, "sdirk2": 4, "sdirk2_1": 5, "sdirk2_2": 6,
           "sdirk3": 7, "sdirk3_1": 8, "sdirk3_2": 9, "sdirk3_3": 10}
# glsns_vector
VEC = {"evaluation_point": 0, "solution_m1": 1, "solution_m2": 2, "solution_m3": 3,
       "system_rhs": 4, "newton_update": 5, "present_solution": 6}


class FeDesc(C.Structure):
    _fields_ = [("dim", C.c_int32), ("velocity_degree", C.c_int32), ("n_su", C.c_int32),
                ("n_sp", C.c_int32), ("n_q", C.c_int32),
                ("shape_u", c_double_p), ("grad_u", c_double_p), ("hess_u", c_double_p),
                ("shape_p", c_double_p), ("grad_p", c_double_p), ("weights", c_double_p)]


class MeshDesc(C.Structure):
    _fields_ = [("n_dofs", C.c_int64), ("n_owned", C.c_int64), ("n_cells", C.c_int64),
                ("cell_dofs", c_i32_p), ("geometry_per_q", C.c_int32),
                ("inv_jacobian", c_double_p), ("det_jacobian", c_double_p),
                ("cell_measure", c_double_p), ("q_points", c_double_p),
                ("constrained", c_u8_p), ("constraint_values", c_double_p),
                ("row_ptr", c_i64_p), ("col_idx", c_i32_p),
                ("n_colors", C.c_int32), ("color_ptr", c_i32_p), ("color_cells", c_i32_p),
                ("n_neighbors", C.c_int32), ("neighbor_rank", c_i32_p), ("send_ptr", c_i64_p),
                ("send_idx", c_i32_p), ("recv_ptr", c_i64_p), ("mapping_laplacian", c_double_p),
                ("constraint_ptr", c_i64_p), ("constraint_idx", c_i32_p),
                ("constraint_weight", c_double_p), ("constraint_inhomogeneity", c_double_p)]


class LinearSolverParams(C.Structure):
    _fields_ = [("relative_residual", C.c_double), ("minimum_residual", C.c_double),
                ("max_iterations", C.c_int32), ("restart", C.c_int32), ("ilu_fill", C.c_int32),
                ("ilu_atol", C.c_double), ("ilu_rtol", C.c_double), ("method", C.c_int32)]


class SolveInfo(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("tolerance", C.c_double),
                ("true_residual", C.c_double), ("estimated_residual", C.c_double)]


class Timers(C.Structure):
    _fields_ = [(n, C.c_double) for n in
                ("assemble_system_ms", "assemble_rhs_ms", "setup_ilu_ms", "solve_linear_system_ms",
                 "spmv_ms", "trsv_ms", "orthog_ms")] + \
               [(n, C.c_int64) for n in
                ("assemble_system_calls", "assemble_rhs_calls", "setup_ilu_calls", "solve_calls",
                 "spmv_calls", "trsv_calls", "orthog_calls", "kernel_launches")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


# every symbol include/glsns.h declares: name -> (restype, argtypes)
ctx_p = C.c_void_p
SYMBOLS = {
    "glsns_create": (C.c_int, [C.c_int32, C.POINTER(ctx_p)]),
    "glsns_destroy": (None, [ctx_p]),
    "glsns_last_error": (C.c_char_p, [ctx_p]),
    "glsns_version": (C.c_char_p, []),
    "glsns_comm_unique_id": (C.c_int, [c_u8_p]),
    "glsns_comm_init": (C.c_int, [ctx_p, C.c_int32, C.c_int32, c_u8_p]),
    "glsns_set_fe": (C.c_int, [ctx_p, C.POINTER(FeDesc)]),
    "glsns_set_mesh": (C.c_int, [ctx_p, C.POINTER(MeshDesc)]),
    "glsns_set_physics": (C.c_int, [ctx_p, C.c_double, C.c_int, c_double_p]),
    "glsns_set_forcing": (C.c_int, [ctx_p, c_double_p]),
    "glsns_set_vector": (C.c_int, [ctx_p, C.c_int, c_double_p, C.c_int64]),
    "glsns_get_vector": (C.c_int, [ctx_p, C.c_int, c_double_p, C.c_int64]),
    "glsns_assemble": (C.c_int, [ctx_p, C.c_int32, C.c_int, c_double_p]),
    "glsns_rhs_norm": (C.c_int, [ctx_p, c_double_p]),
    "glsns_setup_ilu": (C.c_int, [ctx_p, C.c_int32, C.c_double, C.c_double]),
    "glsns_solve_linear_system": (C.c_int, [ctx_p, C.POINTER(LinearSolverParams), C.c_int32,
                                            c_double_p, C.POINTER(SolveInfo)]),
    "glsns_assemble_l2_projection": (C.c_int, [ctx_p, c_double_p]),
    "glsns_distribute_constraints": (C.c_int, [ctx_p, C.c_int]),
    "glsns_calculate_cfl": (C.c_int, [ctx_p, C.c_int, c_double_p, C.c_int32, C.c_double, c_double_p]),
    "glsns_line_search_point": (C.c_int, [ctx_p, C.c_double]),
    "glsns_update_ghosts": (C.c_int, [ctx_p, C.c_int]),
    "glsns_accept_evaluation_point": (C.c_int, [ctx_p]),
    "glsns_get_matrix_values": (C.c_int, [ctx_p, c_double_p, C.c_int64]),
    "glsns_set_matrix_values": (C.c_int, [ctx_p, c_double_p, C.c_int64]),
    "glsns_get_ilu_values": (C.c_int, [ctx_p, c_double_p, C.c_int64]),
    "glsns_get_ilu_pattern": (C.c_int, [ctx_p, c_i64_p, c_i64_p, c_i32_p]),
    "glsns_spmv": (C.c_int, [ctx_p, c_double_p, c_double_p]),
    "glsns_ilu_apply": (C.c_int, [ctx_p, c_double_p, c_double_p]),
    "glsns_ilu_levels": (C.c_int, [ctx_p, c_i32_p, c_i32_p]),
    "glsns_ilu_apply_trace": (C.c_int, [ctx_p, c_double_p, c_double_p, C.POINTER(C.c_uint64), c_i32_p]),
    "glsns_get_timers": (C.c_int, [ctx_p, C.POINTER(Timers)]),
    "glsns_reset_timers": (C.c_int, [ctx_p]),
    "glsns_time_kernel": (C.c_int, [ctx_p, C.c_int32, C.c_int32, C.c_int32, c_double_p]),
}

_LIB = None


def lib():
    """The loaded library. Raises if libglsns.so has not been built (python -m
    softx_2020_200_b200.build); nothing in this package computes without it."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libglsns.so is not built: run `python -m softx_2020_200_b200.build` "
                              "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        _LIB = L
    return _LIB
