"""GLSHotPath: one device context of the C ABI (include/glsns.h) with numpy arrays at the boundary.

Every method forwards to exactly one glsns_* entry point; host buffers are numpy arrays borrowed for
the call. No computation happens in Python and nothing here can run without libglsns.so + a GPU.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import (FeDesc, LinearSolverParams, MeshDesc, SolveInfo, Timers, c_double_p, c_i32_p,
                   c_i64_p, c_u8_p)


class GlsnsError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("glsns status %s: %s" % (_lib.STATUS_NAMES[status], message))
        self.status = status


class NoConvergence(GlsnsError):
    """SolverControl::NoConvergence (thrown by deal.II's Trilinos wrapper)."""

    def __init__(self, message, info):
        super().__init__(_lib.ERR_NO_CONVERGENCE, message)
        self.info = info


def _c(a, dtype):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=dtype)
    return a


def _ptr(a, t):
    return None if a is None else a.ctypes.data_as(t)


class GLSHotPath:
    def __init__(self, device=0):
        self._L = _lib.lib()
        self._ctx = C.c_void_p()
        st = self._L.glsns_create(device, C.byref(self._ctx))
        if st != _lib.OK:
            raise GlsnsError(st, "glsns_create failed (no CUDA device? there is no CPU fallback)")
        self.n_dofs = self.n_owned = self.nnz = 0

    def close(self):
        if self._ctx:
            self._L.glsns_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st):
        if st != _lib.OK:
            raise GlsnsError(st, self._L.glsns_last_error(self._ctx).decode())

    # ---- multi-rank ----
    @staticmethod
    def comm_unique_id():
        uid = np.zeros(128, dtype=np.uint8)
        st = _lib.lib().glsns_comm_unique_id(_ptr(uid, c_u8_p))
        if st != _lib.OK:
            raise GlsnsError(st, "glsns_comm_unique_id")
        return uid

    def comm_init(self, n_ranks, rank, unique_id):
        uid = _c(unique_id, np.uint8)
        self._check(self._L.glsns_comm_init(self._ctx, n_ranks, rank, _ptr(uid, c_u8_p)))

    # ---- setup ----
    def set_fe(self, dim, velocity_degree, shape_u, grad_u, hess_u, shape_p, grad_p, weights):
        arrs = [_c(a, np.float64) for a in (shape_u, grad_u, hess_u, shape_p, grad_p, weights)]
        nq, n_su = arrs[0].shape
        n_sp = arrs[3].shape[1]
        assert arrs[1].shape == (nq, n_su, dim) and arrs[2].shape == (nq, n_su, dim, dim)
        assert arrs[4].shape == (nq, n_sp, dim) and arrs[5].shape == (nq,)
        d = FeDesc(dim, velocity_degree, n_su, n_sp, nq, *[_ptr(a, c_double_p) for a in arrs])
        self._check(self._L.glsns_set_fe(self._ctx, C.byref(d)))
        self.dim, self.n_su, self.n_sp, self.n_q = dim, n_su, n_sp, nq
        self.n_loc = dim * n_su + n_sp

    def set_mesh(self, n_dofs, cell_dofs, inv_jacobian, det_jacobian, cell_measure, constrained,
                 row_ptr, col_idx, color_ptr, color_cells, q_points=None, constraint_values=None,
                 geometry_per_q=False, n_owned=None, neighbor_rank=None, send_ptr=None,
                 send_idx=None, recv_ptr=None, mapping_laplacian=None, constraint_ptr=None,
                 constraint_idx=None, constraint_weight=None, constraint_inhomogeneity=None):
        n_owned = n_dofs if n_owned is None else n_owned
        cd = _c(cell_dofs, np.int32)
        keep = dict(
            cd=cd, ij=_c(inv_jacobian, np.float64), dj=_c(det_jacobian, np.float64),
            cm=_c(cell_measure, np.float64), qp=_c(q_points, np.float64),
            con=_c(constrained, np.uint8), cv=_c(constraint_values, np.float64),
            rp=_c(row_ptr, np.int64), ci=_c(col_idx, np.int32), cp=_c(color_ptr, np.int32),
            cc=_c(color_cells, np.int32), nr=_c(neighbor_rank, np.int32),
            sp=_c(send_ptr, np.int64), si=_c(send_idx, np.int32), rv=_c(recv_ptr, np.int64),
            ml=_c(mapping_laplacian, np.float64), hp=_c(constraint_ptr, np.int64),
            hi=_c(constraint_idx, np.int32), hw=_c(constraint_weight, np.float64),
            hg=_c(constraint_inhomogeneity, np.float64))
        n_cells = cd.shape[0] if cd.ndim == 2 else cd.size // self.n_loc
        assert keep["rp"].size == n_owned + 1 and keep["con"].size == n_dofs
        m = MeshDesc(n_dofs, n_owned, n_cells, _ptr(cd, c_i32_p), 1 if geometry_per_q else 0,
                     _ptr(keep["ij"], c_double_p), _ptr(keep["dj"], c_double_p),
                     _ptr(keep["cm"], c_double_p), _ptr(keep["qp"], c_double_p),
                     _ptr(keep["con"], c_u8_p), _ptr(keep["cv"], c_double_p),
                     _ptr(keep["rp"], c_i64_p), _ptr(keep["ci"], c_i32_p),
                     max(len(keep["cp"]) - 1, 0), _ptr(keep["cp"], c_i32_p),
                     _ptr(keep["cc"], c_i32_p),
                     0 if neighbor_rank is None else len(keep["nr"]), _ptr(keep["nr"], c_i32_p),
                     _ptr(keep["sp"], c_i64_p), _ptr(keep["si"], c_i32_p),
                     _ptr(keep["rv"], c_i64_p), _ptr(keep["ml"], c_double_p),
                     _ptr(keep["hp"], c_i64_p), _ptr(keep["hi"], c_i32_p), _ptr(keep["hw"], c_double_p),
                     _ptr(keep["hg"], c_double_p))
        self._check(self._L.glsns_set_mesh(self._ctx, C.byref(m)))
        self.n_dofs, self.n_owned, self.n_cells = n_dofs, n_owned, n_cells
        self.nnz = int(keep["rp"][-1])

    def set_physics(self, viscosity, srf=False, omega=(0.0, 0.0, 0.0)):
        om = np.asarray(omega, dtype=np.float64)
        self._check(self._L.glsns_set_physics(self._ctx, viscosity, 1 if srf else 0,
                                              _ptr(om, c_double_p)))

    def set_forcing(self, force_at_q):
        f = _c(force_at_q, np.float64)
        if f is not None:
            assert f.size == self.n_cells * self.n_q * self.dim
        self._check(self._L.glsns_set_forcing(self._ctx, _ptr(f, c_double_p)))

    # ---- vectors ----
    def set_vector(self, which, host):
        h = _c(host, np.float64)
        self._check(self._L.glsns_set_vector(self._ctx, _lib.VEC[which], _ptr(h, c_double_p),
                                             h.size))

    def get_vector(self, which):
        n = self.n_owned if which in ("system_rhs", "newton_update") else self.n_dofs
        out = np.empty(n)
        self._check(self._L.glsns_get_vector(self._ctx, _lib.VEC[which], _ptr(out, c_double_p), n))
        return out

    # ---- hot path ----
    def assemble(self, assemble_matrix, scheme="steady", time_steps=None):
        ts = None if time_steps is None else _c(
            list(time_steps) + [0.0] * (4 - len(time_steps)), np.float64)
        self._check(self._L.glsns_assemble(self._ctx, 1 if assemble_matrix else 0,
                                           _lib.SCHEMES[scheme], _ptr(ts, c_double_p)))

    def assemble_l2_projection(self, initial_at_q):
        """assemble_L2_projection (gls_navier_stokes.cc:829-914); initial_at_q [n_cells][n_q][dim+1]."""
        a = _c(initial_at_q, np.float64)
        assert a.size == self.n_cells * self.n_q * (self.dim + 1)
        self._check(self._L.glsns_assemble_l2_projection(self._ctx, _ptr(a, c_double_p)))

    def distribute_constraints(self, which):
        """nonzero_constraints.distribute on a device vector."""
        self._check(self._L.glsns_distribute_constraints(self._ctx, _lib.VEC[which]))

    def calculate_cfl(self, which, shape_u_at_centre, fe_degree, time_step):
        """calculate_CFL (postprocessing_cfl.cc:34-87) of a ghosted device vector."""
        t = _c(shape_u_at_centre, np.float64)
        assert t.size == self.n_su
        v = C.c_double()
        self._check(self._L.glsns_calculate_cfl(self._ctx, _lib.VEC[which], _ptr(t, c_double_p),
                                                fe_degree, time_step, C.byref(v)))
        return v.value

    def rhs_norm(self):
        v = C.c_double()
        self._check(self._L.glsns_rhs_norm(self._ctx, C.byref(v)))
        return v.value

    def setup_ilu(self, fill=0, atol=1e-8, rtol=1.0):
        self._check(self._L.glsns_setup_ilu(self._ctx, fill, atol, rtol))

    def solve_linear_system(self, relative_residual=1e-3, minimum_residual=1e-8,
                            max_iterations=1000, restart=30, ilu_fill=0, ilu_atol=1e-8,
                            ilu_rtol=1.0, renewed_matrix=True, download=True, method="gmres"):
        """Returns (newton_update or None, info dict). Raises NoConvergence like deal.II.
        method: the .prm `linear solver/method`, "gmres" or "bicgstab"."""
        p = LinearSolverParams(relative_residual, minimum_residual, max_iterations, restart,
                               ilu_fill, ilu_atol, ilu_rtol,
                               {"gmres": 0, "bicgstab": 1, "amg": 2}[method])
        info = SolveInfo()
        out = np.empty(self.n_owned) if download else None
        st = self._L.glsns_solve_linear_system(self._ctx, C.byref(p), 1 if renewed_matrix else 0,
                                               _ptr(out, c_double_p), C.byref(info))
        d = dict(iterations=info.iterations, tolerance=info.tolerance,
                 true_residual=info.true_residual, estimated_residual=info.estimated_residual)
        if st == _lib.ERR_NO_CONVERGENCE:
            raise NoConvergence(self._L.glsns_last_error(self._ctx).decode(), d)
        self._check(st)
        return out, d

    def line_search_point(self, alpha):
        self._check(self._L.glsns_line_search_point(self._ctx, alpha))

    def update_ghosts(self, which):
        self._check(self._L.glsns_update_ghosts(self._ctx, _lib.VEC[which]))

    def accept_evaluation_point(self):
        self._check(self._L.glsns_accept_evaluation_point(self._ctx))

    # ---- inspection ----
    def get_matrix_values(self):
        out = np.empty(self.nnz)
        self._check(self._L.glsns_get_matrix_values(self._ctx, _ptr(out, c_double_p), self.nnz))
        return out

    def set_matrix_values(self, values):
        v = _c(values, np.float64)
        self._check(self._L.glsns_set_matrix_values(self._ctx, _ptr(v, c_double_p), v.size))

    def get_ilu_pattern(self):
        """(row_ptr, col_idx) of the factors: the mesh's CSR for fill 0, the level-of-fill pattern
        after setup_ilu(fill > 0)."""
        nnz = C.c_int64()
        self._check(self._L.glsns_get_ilu_pattern(self._ctx, C.byref(nnz), None, None))
        rp, col = np.empty(self.n_owned + 1, dtype=np.int64), np.empty(nnz.value, dtype=np.int32)
        self._check(self._L.glsns_get_ilu_pattern(self._ctx, None, _ptr(rp, _lib.c_i64_p),
                                                  _ptr(col, c_i32_p)))
        return rp, col

    def get_ilu_values(self):
        nnz = C.c_int64()
        self._check(self._L.glsns_get_ilu_pattern(self._ctx, C.byref(nnz), None, None))
        v = np.empty(nnz.value)
        self._check(self._L.glsns_get_ilu_values(self._ctx, _ptr(v, c_double_p), nnz.value))
        return v

    def spmv(self, x):
        x = _c(x, np.float64)
        assert x.size == self.n_dofs
        y = np.empty(self.n_owned)
        self._check(self._L.glsns_spmv(self._ctx, _ptr(x, c_double_p), _ptr(y, c_double_p)))
        return y

    def ilu_apply(self, r):
        r = _c(r, np.float64)
        assert r.size == self.n_owned
        z = np.empty(self.n_owned)
        self._check(self._L.glsns_ilu_apply(self._ctx, _ptr(r, c_double_p), _ptr(z, c_double_p)))
        return z

    def ilu_apply_trace(self, r):
        """(z, t_lower, t_upper, warp_lower, warp_upper): publish times in ns per row and sweep."""
        r = np.ascontiguousarray(r, dtype=np.float64)
        n = self.n_owned
        z, t, w = np.empty(n), np.zeros(10 * n, dtype=np.uint64), np.zeros(2 * n, dtype=np.int32)
        self._check(self._L.glsns_ilu_apply_trace(
            self._ctx, _ptr(r, c_double_p), _ptr(z, c_double_p),
            t.ctypes.data_as(C.POINTER(C.c_uint64)), _ptr(w, c_i32_p)))
        self.last_trace_polls = (t[2 * n:3 * n], t[3 * n:4 * n])
        self.last_trace_posts = (t[4 * n:5 * n], t[5 * n:6 * n])  # when a row's totals reached the mailbox
        self.last_trace_begin = (t[6 * n:7 * n], t[7 * n:8 * n])  # when its helper began the group's last item
        self.last_trace_inputs = (t[8 * n:9 * n], t[9 * n:])      # ... and had all its inputs
        return z, t[:n], t[n:2 * n], w[:n], w[n:]

    def ilu_levels(self):
        a, b = C.c_int32(), C.c_int32()
        self._check(self._L.glsns_ilu_levels(self._ctx, C.byref(a), C.byref(b)))
        return a.value, b.value

    def timers(self):
        t = Timers()
        self._check(self._L.glsns_get_timers(self._ctx, C.byref(t)))
        return t.as_dict()

    def reset_timers(self):
        self._check(self._L.glsns_reset_timers(self._ctx))

    KERNELS = {"spmv": 0, "ilu_apply": 1, "orthog": 2, "assemble_system": 3, "assemble_rhs": 4,
               "ilu_factor": 5}

    def time_kernel(self, kernel, reps=10, nvec=15):
        v = C.c_double()
        self._check(self._L.glsns_time_kernel(self._ctx, self.KERNELS[kernel], reps, nvec,
                                              C.byref(v)))
        return v.value
