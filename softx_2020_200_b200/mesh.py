"""BoxMesh: binding of the C++ host-side stand-in for deal.II's mesh / DoFHandler / FEValues objects
(csrc/host_mesh.cpp). It builds the arrays of glsns_fe_desc / glsns_mesh_desc for uniform box meshes
(GridGenerator::hyper_cube / subdivided_hyper_cube) and hands them to a GLSHotPath context without
copying through numpy. Host code only."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import FeDesc, MeshDesc

_DTYPES = {"cell_dofs": np.int32, "col_idx": np.int32, "row_ptr": np.int64, "constrained": np.uint8,
           "constraint_values": np.float64, "inv_jacobian": np.float64, "det_jacobian": np.float64,
           "cell_measure": np.float64, "q_points": np.float64, "color_ptr": np.int32,
           "color_cells": np.int32, "dof_component": np.int32, "dof_coords": np.float64,
           "local_to_global": np.int64, "cell_ids": np.int64, "periodic_slave": np.int32,
           "periodic_master": np.int32, "neighbor_rank": np.int32,
           "send_ptr": np.int64, "send_idx": np.int32, "recv_ptr": np.int64, "shape_u": np.float64,
           "grad_u": np.float64, "hess_u": np.float64, "shape_p": np.float64, "grad_p": np.float64,
           "weights": np.float64, "unit_q_points": np.float64}

NOSLIP, FUNCTION = 1, 2


def _bind(L):
    if getattr(L, "_glsnsh_bound", False):
        return
    L.glsnsh_mesh_create.restype = C.c_void_p
    L.glsnsh_mesh_create.argtypes = [C.c_int, C.POINTER(C.c_int), C.c_int, C.c_int,
                                     _lib.c_double_p, _lib.c_double_p, C.c_int, C.POINTER(C.c_int),
                                     _lib.c_double_p, C.POINTER(C.c_int), C.c_int, C.c_int]
    L.glsnsh_mesh_create_local.restype = C.c_void_p
    L.glsnsh_mesh_create_local.argtypes = L.glsnsh_mesh_create.argtypes + [C.c_int, C.c_int]
    L.glsnsh_mesh_make_periodic.restype = C.c_int
    L.glsnsh_mesh_make_periodic.argtypes = [C.c_void_p, C.c_int]
    L.glsnsh_mesh_partition.restype = C.c_void_p
    L.glsnsh_mesh_partition.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.glsnsh_mesh_destroy.restype = None
    L.glsnsh_mesh_destroy.argtypes = [C.c_void_p]
    L.glsnsh_mesh_error.restype = C.c_char_p
    L.glsnsh_mesh_error.argtypes = [C.c_void_p]
    L.glsnsh_mesh_info.restype = None
    L.glsnsh_mesh_info.argtypes = [C.c_void_p, _lib.c_i64_p]
    L.glsnsh_mesh_array.restype = C.c_void_p
    L.glsnsh_mesh_array.argtypes = [C.c_void_p, C.c_char_p, _lib.c_i64_p]
    L.glsnsh_mesh_fill_desc.restype = None
    L.glsnsh_mesh_fill_desc.argtypes = [C.c_void_p, C.POINTER(FeDesc), C.POINTER(MeshDesc)]
    L._glsnsh_bound = True


class BoxMesh:
    """Uniform box mesh with Q_pu^dim x Q_pp elements.

    bcs: list of (face_id, "noslip") or (face_id, "function", (ux, uy[, uz])) in the order the
    reference creates the constraints (first listed wins on shared edges); face ids as deal.II's
    colorize=true (0 x=lo, 1 x=hi, 2 y=lo, 3 y=hi, 4 z=lo, 5 z=hi). bcs=None: no-slip everywhere
    (boundary id 0 of an uncolorized hyper_cube); bcs=[] : no Dirichlet boundary at all.
    periodic: directions d whose faces 2d / 2d+1 are a `type = periodic` pair (the dofs of the hi
    face are identified with those of the lo face, see glsnsh_mesh_make_periodic).
    """

    def __init__(self, dim, n, pu, pp, lo=-1.0, hi=1.0, bcs=None, renumber=True, nq1=0,
                 with_q_points=False, periodic=(), _handle=None, local=None):
        """local=(n_ranks, rank): build only that rank's part (owned rows + ghost layer) without
        the global dof-level arrays -- the same arrays as BoxMesh(...).partition(n_ranks, rank)."""
        self._L = _lib.lib()
        _bind(self._L)
        if _handle is not None:
            self._h = _handle
        else:
            nd = (C.c_int * 3)(*([n] * 3 if np.isscalar(n) else list(n) + [1] * (3 - len(n))))
            lo3 = (C.c_double * 3)(*([lo] * 3 if np.isscalar(lo) else list(lo) + [0] * (3 - len(lo))))
            hi3 = (C.c_double * 3)(*([hi] * 3 if np.isscalar(hi) else list(hi) + [0] * (3 - len(hi))))
            types = (C.c_int * 6)(*([0] * 6))
            vals = (C.c_double * 18)(*([0.0] * 18))
            order = list(range(6))
            if bcs is None:
                for f in range(2 * dim):
                    types[f] = NOSLIP
            else:
                listed = []
                for bc in bcs:
                    f = bc[0]
                    listed.append(f)
                    if bc[1] == "noslip":
                        types[f] = NOSLIP
                    else:
                        types[f] = FUNCTION
                        for c, v in enumerate(bc[2]):
                            vals[f * 3 + c] = v
                order = listed + [f for f in range(6) if f not in listed]
            args = (dim, nd, pu, pp, lo3, hi3, nq1, types, vals, (C.c_int * 6)(*order),
                    1 if renumber else 0, 1 if with_q_points else 0)
            if local is not None:
                assert not len(periodic), "periodic meshes are serial"
                self._h = self._L.glsnsh_mesh_create_local(*args, int(local[0]), int(local[1]))
            else:
                self._h = self._L.glsnsh_mesh_create(*args)
            if len(periodic):
                self._L.glsnsh_mesh_make_periodic(self._h, sum(1 << d for d in periodic))
        err = self._L.glsnsh_mesh_error(self._h).decode()
        if err:
            raise ValueError(err)
        info = np.zeros(16, dtype=np.int64)
        self._L.glsnsh_mesh_info(self._h, info.ctypes.data_as(_lib.c_i64_p))
        (self.n_dofs, self.n_owned, self.n_cells, self.nnz, self.n_loc, self.n_q, self.n_su,
         self.n_sp, self.n_colors, self.n_neighbors, self.n_global, self.owned_begin) = \
            [int(v) for v in info[:12]]
        self.dim, self.pu, self.pp = dim, pu, pp

    def partition(self, n_ranks, rank):
        """The rank-local view (owned rows + ghost layer), see glsnsh_mesh_partition."""
        h = self._L.glsnsh_mesh_partition(self._h, n_ranks, rank)
        return BoxMesh(self.dim, 0, self.pu, self.pp, _handle=h)

    def array(self, name):
        """Read-only numpy view of a named host array (no copy)."""
        cnt = C.c_int64()
        p = self._L.glsnsh_mesh_array(self._h, name.encode(), C.byref(cnt))
        if cnt.value < 0:
            raise KeyError(name)
        dt = np.dtype(_DTYPES[name])
        if cnt.value == 0 or not p:
            return np.zeros(0, dtype=dt)
        buf = (C.c_char * (cnt.value * dt.itemsize)).from_address(p)
        a = np.frombuffer(buf, dtype=dt)
        a.flags.writeable = False
        return a

    def attach(self, hotpath):
        """glsns_set_fe + glsns_set_mesh on a GLSHotPath context, straight from the C++ arrays."""
        fe, md = FeDesc(), MeshDesc()
        self._L.glsnsh_mesh_fill_desc(self._h, C.byref(fe), C.byref(md))
        hotpath._check(self._L.glsns_set_fe(hotpath._ctx, C.byref(fe)))
        hotpath.dim, hotpath.n_su, hotpath.n_sp, hotpath.n_q = self.dim, self.n_su, self.n_sp, \
            self.n_q
        hotpath.n_loc = self.n_loc
        hotpath._check(self._L.glsns_set_mesh(hotpath._ctx, C.byref(md)))
        hotpath.n_dofs, hotpath.n_owned, hotpath.n_cells, hotpath.nnz = \
            self.n_dofs, self.n_owned, self.n_cells, self.nnz

    def initial_state(self):
        """Nodal initial condition: zero, with the boundary values applied
        (set_initial_condition nodal + apply_constraints)."""
        return self.array("constraint_values").copy()

    def close(self):
        if getattr(self, "_h", None):
            self._L.glsnsh_mesh_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
