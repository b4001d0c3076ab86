"""TEST HARNESS. Python proxy of the product's host-side solver class (include/glsns_solver.hpp)
driven by the transcribed reference drivers (tests/mirror/reference_drivers.hpp: Newton /
SkipNewton, NavierStokesBase's time-stepping glue) through host buffers, exactly as the
reference's Newton drivers move their vectors (include/core/newton_non_linear_solver.h:76-139)."""
import ctypes as C

import numpy as np

from softx_2020_200_b200 import _lib
from softx_2020_200_b200.hotpath import GlsnsError, NoConvergence

_MIRROR = None


def mirror_lib():
    """tests/mirror/libglsns_mirror.so (built on demand; it links against libglsns.so)."""
    global _MIRROR
    if _MIRROR is None:
        from . import build
        _lib.lib()                      # the product library first (no CPU fallback: raises if missing)
        _MIRROR = C.CDLL(build.build())
    return _MIRROR

METHODS = _lib.SCHEMES  # Parameters::SimulationControl::TimeSteppingMethod, same order


def _bind(L):
    if getattr(L, "_glsnsh_solver_bound", False):
        return
    L.glsnsh_solver_create.restype = C.c_void_p
    L.glsnsh_solver_create.argtypes = [C.c_void_p, C.c_char_p, _lib.c_double_p, C.c_int]
    L.glsnsh_solver_error.restype = C.c_char_p
    L.glsnsh_solver_error.argtypes = [C.c_void_p]
    L.glsnsh_solver_destroy.restype = None
    L.glsnsh_solver_destroy.argtypes = [C.c_void_p]
    L.glsnsh_solver_set_vector.restype = None
    L.glsnsh_solver_set_vector.argtypes = [C.c_void_p, C.c_int, _lib.c_double_p, C.c_int64]
    L.glsnsh_solver_get_present.restype = None
    L.glsnsh_solver_get_present.argtypes = [C.c_void_p, _lib.c_double_p, C.c_int64]
    L.glsnsh_solver_set_time_steps.restype = None
    L.glsnsh_solver_set_time_steps.argtypes = [C.c_void_p, _lib.c_double_p, C.c_int]
    L.glsnsh_solver_solve_non_linear_system.restype = C.c_int
    L.glsnsh_solver_solve_non_linear_system.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    L.glsnsh_solver_set_initial_condition_l2.restype = C.c_int
    L.glsnsh_solver_set_initial_condition_l2.argtypes = [C.c_void_p, _lib.c_double_p]
    L.glsnsh_solver_set_initial_condition.restype = C.c_int
    L.glsnsh_solver_set_initial_condition.argtypes = [C.c_void_p, C.c_int, _lib.c_double_p,
                                                      _lib.c_double_p]
    L.glsnsh_solver_time_step.restype = C.c_int
    L.glsnsh_solver_time_step.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double]
    L.glsnsh_solver_calculate_cfl.restype = C.c_int
    L.glsnsh_solver_calculate_cfl.argtypes = [C.c_void_p, _lib.c_double_p, C.c_double, _lib.c_double_p]
    L.glsnsh_solver_log.restype = C.c_char_p
    L.glsnsh_solver_log.argtypes = [C.c_void_p]
    L._glsnsh_solver_bound = True


class GLSNavierStokesSolver:
    """mesh: softx_2020_200_b200.mesh.BoxMesh; prm: text of a reference .prm file (only the
    subsections of the hot path are read); forcing_at_q: [n_cells, n_q, dim] or None."""

    def __init__(self, mesh, prm="", forcing_at_q=None, device=0):
        self._L = mirror_lib()
        _bind(self._L)
        self.mesh = mesh
        f = None if forcing_at_q is None else np.ascontiguousarray(forcing_at_q, dtype=np.float64)
        self._h = self._L.glsnsh_solver_create(
            mesh._h, prm.encode(), None if f is None else f.ctypes.data_as(_lib.c_double_p), device)
        err = self._L.glsnsh_solver_error(self._h).decode()
        if err:
            self._L.glsnsh_solver_destroy(self._h)
            self._h = None
            raise GlsnsError(_lib.ERR_CUDA, err)

    def set_vector(self, which, values):
        idx = {"present_solution": 0, "solution_m1": 1, "solution_m2": 2, "solution_m3": 3}[which]
        v = np.ascontiguousarray(values, dtype=np.float64)
        assert v.size == self.mesh.n_dofs
        self._L.glsnsh_solver_set_vector(self._h, idx, v.ctypes.data_as(_lib.c_double_p), v.size)

    def set_time_steps(self, dts):
        d = np.ascontiguousarray(dts, dtype=np.float64)
        self._L.glsnsh_solver_set_time_steps(self._h, d.ctypes.data_as(_lib.c_double_p), d.size)

    @property
    def present_solution(self):
        out = np.empty(self.mesh.n_dofs)
        self._L.glsnsh_solver_get_present(self._h, out.ctypes.data_as(_lib.c_double_p), out.size)
        return out

    def solve_non_linear_system(self, time_stepping_method="steady", first_iteration=False,
                                force_matrix_renewal=True):
        rc = self._L.glsnsh_solver_solve_non_linear_system(
            self._h, METHODS[time_stepping_method], int(first_iteration), int(force_matrix_renewal))
        if rc == 3:
            raise NoConvergence(self._L.glsnsh_solver_error(self._h).decode(), {})
        if rc:
            raise RuntimeError(self._L.glsnsh_solver_error(self._h).decode())

    def set_initial_condition_l2_projection(self, initial_at_q):
        """set_initial_condition(L2projection) (gls_navier_stokes.cc:795-803); initial_at_q
        [n_cells, n_q, dim+1]: the initial-condition function at the quadrature points."""
        a = np.ascontiguousarray(initial_at_q, dtype=np.float64)
        assert a.size == self.mesh.n_cells * self.mesh.n_q * (self.mesh.dim + 1)
        rc = self._L.glsnsh_solver_set_initial_condition_l2(self._h, a.ctypes.data_as(_lib.c_double_p))
        if rc == 3:
            raise NoConvergence(self._L.glsnsh_solver_error(self._h).decode(), {})
        if rc:
            raise RuntimeError(self._L.glsnsh_solver_error(self._h).decode())

    def set_initial_condition(self, initial_nodal=None, initial_at_q=None, type=None):
        """set_initial_condition (gls_navier_stokes.cc:784-828) with the `initial conditions` type of
        the .prm (or `type`: "L2projection" | "viscous" | "nodal"); initial_nodal [n_dofs]: the
        interpolated initial-condition function, initial_at_q [n_cells, n_q, dim+1]: its values at
        the quadrature points (L2projection only)."""
        t = -1 if type is None else {"none": 0, "L2projection": 1, "viscous": 2, "nodal": 3}[type]
        a = None if initial_nodal is None else np.ascontiguousarray(initial_nodal, dtype=np.float64)
        q = None if initial_at_q is None else np.ascontiguousarray(initial_at_q, dtype=np.float64)
        rc = self._L.glsnsh_solver_set_initial_condition(
            self._h, t, None if a is None else a.ctypes.data_as(_lib.c_double_p),
            None if q is None else q.ctypes.data_as(_lib.c_double_p))
        if rc == 3:
            raise NoConvergence(self._L.glsnsh_solver_error(self._h).decode(), {})
        if rc:
            raise RuntimeError(self._L.glsnsh_solver_error(self._h).decode())

    def _time_step(self, what, method, dt=0.0, startup_scaling=0.4):
        rc = self._L.glsnsh_solver_time_step(self._h, what, METHODS[method], dt, startup_scaling)
        if rc == 3:
            raise NoConvergence(self._L.glsnsh_solver_error(self._h).decode(), {})
        if rc:
            raise RuntimeError(self._L.glsnsh_solver_error(self._h).decode())

    def advance(self, method, dt, first=False, startup_scaling=0.4):
        """One transient iteration as NavierStokesBase::solve runs it (navier_stokes_base.cc:428-590):
        integrate() adds dt to the time-step vector, first_iteration() / iterate() solve the stage(s)
        (method: "bdf1" | "bdf2" | "bdf3" | "sdirk2" | "sdirk3" | "steady"), finish_time_step() shifts
        solution_m1..m3."""
        self._time_step(0, method, dt)
        self._time_step(1 if first else 2, method, dt, startup_scaling)
        self._time_step(3, method)

    def finish_time_step(self, method):
        """After set_initial_condition: solution_m1 = present_solution (navier_stokes_base.cc:428-435)."""
        self._time_step(3, method)

    def calculate_cfl(self, shape_u_at_centre, time_step):
        """calculate_CFL (postprocessing_cfl.cc:34-87) of present_solution."""
        t = np.ascontiguousarray(shape_u_at_centre, dtype=np.float64)
        v = C.c_double()
        if self._L.glsnsh_solver_calculate_cfl(self._h, t.ctypes.data_as(_lib.c_double_p), time_step,
                                               C.byref(v)):
            raise RuntimeError(self._L.glsnsh_solver_error(self._h).decode())
        return v.value

    @property
    def log(self):
        return self._L.glsnsh_solver_log(self._h).decode()

    def close(self):
        if getattr(self, "_h", None):
            self._L.glsnsh_solver_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
