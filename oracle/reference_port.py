"""oracle/reference_port.py — CPU oracle for Lethe's GLS Navier–Stokes hot path.

TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / ``--impl reference`` legs may import this module.  The product package
(softx_2020_200_b200/) never imports anything under oracle/.

Parity status: PINNED against the reference's golden files (tests/test_oracle_golden.py):
``tests/solvers/restart_01.output`` (GMRES iterations 8/6/10, true residuals, L2 error) and
``applications_tests/gls_navier_stokes_3d/mms3d_gls.output`` (3D Q1-Q1 MMS errors),
``applications_tests/gls_navier_stokes_2d/taylor-green-vortex_gls_{sdirk3,sdirk2,bdf1}.mpirun=2.output``
(periodic Taylor-Green vortex: CFL, enstrophy, kinetic energy and L2 error of every step, which pin the
transient terms, the L2 projection, calculate_CFL and the SDIRK stage logic).

What is restated here (numpy) and in gls_oracle.c (the heavy loops), with the reference file:line:

* the host-side objects deal.II provides to the hot path on a uniform hyper_cube mesh: FE_Q shape
  tables on QGauss points, DoF numbering + Cuthill–McKee (gls_navier_stokes.cc:69-70), homogeneous /
  inhomogeneous Dirichlet constraints (:79-183), the sparsity pattern with
  keep_constrained_dofs=false (:204-213);
* assembleGLS (:231-777)                       -> gls_oracle.c: glso_assemble
* setup_ILU (:1161-1176)                       -> gls_oracle.c: glso_ilu0; `ilu preconditioner fill`
  = k > 0 -> iluk_pattern (Ifpack's level-of-fill graph, restated; ILU(k) = ILU(0) on that pattern.
  PARITY of k > 0 pinned only through solver-independent results: the reference prints no
  iteration counts for fill > 0)
* solve_system_GMRES (:1242-1289)              -> gls_oracle.c: glso_gmres
* solve_system_BiCGStab (:1291-1340)           -> bicgstab (PARITY UNPINNED: no reference test or
  example uses `method = bicgstab`; AztecOO's AZ_bicgstab is restated from the published
  right-preconditioned algorithm, van der Vorst 1992 as in the Aztec user's guide)
* assemble_L2_projection (:829-914) + set_initial_condition(L2projection) (:795-803)
                                                -> assemble_l2_projection, l2_projection
* calculate_CFL (source/solvers/postprocessing_cfl.cc:34-87)                     -> calculate_cfl
* NavierStokesBase::iterate, SDIRK stages (source/solvers/navier_stokes_base.cc:461-505) -> sdirk_step
* calculate_kinetic_energy / calculate_enstrophy (source/solvers/postprocessing_*.cc)  -> kinetic_energy,
  enstrophy
* periodic boundary conditions (type = periodic): BoxMesh(periodic=...)
* NewtonNonLinearSolver::solve (include/core/newton_non_linear_solver.h:76-139) -> newton_solve
* calculate_L2_error (source/solvers/navier_stokes_base.cc:255-380)              -> l2_error
* bdf_coefficients (source/core/bdf.cc:46-75), sdirk_coefficients (source/core/sdirk.cc:11-44)
"""
import ctypes as C
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_double_p = C.POINTER(C.c_double)
c_i32_p = C.POINTER(C.c_int32)
c_i64_p = C.POINTER(C.c_int64)
c_u8_p = C.POINTER(C.c_uint8)


class _FE(C.Structure):
    _fields_ = [("dim", C.c_int), ("n_su", C.c_int), ("n_sp", C.c_int), ("nq", C.c_int),
                ("vel_degree", C.c_int),
                ("Nu", c_double_p), ("dNu", c_double_p), ("d2Nu", c_double_p),
                ("Np", c_double_p), ("dNp", c_double_p), ("wq", c_double_p)]


class _Cells(C.Structure):
    _fields_ = [("ncell", C.c_int64), ("cell_dofs", c_i32_p), ("cell_invJ", c_double_p),
                ("cell_detJ", c_double_p), ("cell_measure", c_double_p),
                ("qpoints", c_double_p), ("force", c_double_p),
                ("geometry_per_q", C.c_int), ("map_lap", c_double_p)]


class _Params(C.Structure):
    _fields_ = [("viscosity", C.c_double), ("transient", C.c_int), ("sdt", C.c_double),
                ("coefs", C.c_double * 4), ("srf", C.c_int), ("omega", C.c_double * 3)]


def build_lib(force=False):
    """Compile gls_oracle.c (see oracle/Makefile) and return the path of the .so."""
    so = os.path.join(_HERE, "_build", "libgls_oracle.so")
    src = os.path.join(_HERE, "gls_oracle.c")
    stamp = os.path.join(_HERE, "_build", "cpu.stamp")
    cpu = _cpu_id()
    old = open(stamp).read() if os.path.exists(stamp) else ""
    # -march=native: a library built on another host (it travels with gpurun) must be rebuilt
    if force or not os.path.exists(so) or old != cpu or (
            os.path.exists(src) and os.path.getmtime(so) < os.path.getmtime(src)):
        subprocess.check_call(["make", "-s", "-B", "-C", _HERE])
        with open(stamp, "w") as f:
            f.write(cpu)
    return so


def _cpu_id():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    import hashlib
                    return hashlib.sha1(line.encode()).hexdigest()
    except OSError:
        pass
    return "unknown"


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build_lib())
        _LIB.glso_num_threads.restype = C.c_int
    return _LIB


def _p(a, t):
    return None if a is None else a.ctypes.data_as(t)


# ----------------------------------------------------------------------------------------------
# finite element tables (what FEValues precomputes on the reference cell [0,1]^dim)
# ----------------------------------------------------------------------------------------------
def gauss01(nq1):
    """QGauss<1>(nq1) on [0,1]."""
    x, w = np.polynomial.legendre.leggauss(nq1)
    return 0.5 * (x + 1.0), 0.5 * w


def lagrange1d(p, x):
    """Lagrange basis of degree p on equispaced nodes of [0,1] (FE_Q support points for p<=2):
    values, first and second derivatives at points x. Shapes [p+1, len(x)]."""
    nodes = np.linspace(0.0, 1.0, p + 1)
    x = np.asarray(x, dtype=np.float64)
    V = np.zeros((p + 1, x.size))
    D = np.zeros((p + 1, x.size))
    D2 = np.zeros((p + 1, x.size))
    for a in range(p + 1):
        coef = np.poly1d([1.0])
        for b in range(p + 1):
            if b != a:
                coef = coef * np.poly1d([1.0, -nodes[b]]) / (nodes[a] - nodes[b])
        V[a] = coef(x)
        D[a] = coef.deriv(1)(x) if p >= 1 else 0.0
        D2[a] = coef.deriv(2)(x) if p >= 2 else 0.0
    return V, D, D2


def _tensor_tables(dim, p, xq1):
    """Scalar FE_Q(p) tables at the tensor-product points built from xq1 (x fastest):
    N[q,a], dN[q,a,d], d2N[q,a,d,e]; local shape index a lexicographic, x fastest."""
    V, D, D2 = lagrange1d(p, xq1)
    n1, q1 = p + 1, len(xq1)
    ns, nq = n1 ** dim, q1 ** dim
    N = np.zeros((nq, ns))
    dN = np.zeros((nq, ns, dim))
    d2N = np.zeros((nq, ns, dim, dim))
    for q in range(nq):
        qi = [(q // q1 ** d) % q1 for d in range(dim)]
        for a in range(ns):
            ai = [(a // n1 ** d) % n1 for d in range(dim)]
            N[q, a] = np.prod([V[ai[d], qi[d]] for d in range(dim)])
            for d in range(dim):
                dN[q, a, d] = np.prod([(D if e == d else V)[ai[e], qi[e]] for e in range(dim)])
                for e in range(dim):
                    if d == e:
                        f = [(D2 if g == d else V)[ai[g], qi[g]] for g in range(dim)]
                    else:
                        f = [(D if g in (d, e) else V)[ai[g], qi[g]] for g in range(dim)]
                    d2N[q, a, d, e] = np.prod(f)
    return N, dN, d2N


class FETables:
    """FESystem(FE_Q(pu)^dim, FE_Q(pp)) on QGauss(nq1) (navier_stokes_base.cc:62,70,93-94)."""

    def __init__(self, dim, pu, pp, nq1=None):
        self.dim, self.pu, self.pp = dim, pu, pp
        self.nq1 = nq1 if nq1 else pu + 1
        x, w = gauss01(self.nq1)
        self.xq1, self.wq1 = x, w
        self.Nu, self.dNu, self.d2Nu = _tensor_tables(dim, pu, x)
        self.Np, self.dNp, _ = _tensor_tables(dim, pp, x)
        self.nq = self.nq1 ** dim
        self.n_su, self.n_sp = (pu + 1) ** dim, (pp + 1) ** dim
        self.n = dim * self.n_su + self.n_sp
        wq = np.ones(self.nq)
        xq = np.zeros((self.nq, dim))
        for q in range(self.nq):
            for d in range(dim):
                i = (q // self.nq1 ** d) % self.nq1
                wq[q] *= w[i]
                xq[q, d] = x[i]
        self.wq, self.xq = wq, xq

    def cstruct(self):
        keep = [np.ascontiguousarray(a) for a in
                (self.Nu, self.dNu, self.d2Nu, self.Np, self.dNp, self.wq)]
        s = _FE(self.dim, self.n_su, self.n_sp, self.nq, self.pu,
                *[_p(a, c_double_p) for a in keep])
        s._keep = keep
        return s


# ----------------------------------------------------------------------------------------------
# mesh / DoFs / constraints / sparsity: the deal.II side of the boundary
# ----------------------------------------------------------------------------------------------
class BoxMesh:
    """GridGenerator::hyper_cube / subdivided_hyper_cube on [lo,hi]^dim with n cells per direction
    and the DoFHandler objects the hot path consumes.

    bcs: {face_id: ("noslip",) | ("function", callable(x[:, dim]) -> [:, dim])}; face ids as
    deal.II colorize=true: 0:x=lo 1:x=hi 2:y=lo 3:y=hi 4:z=lo 5:z=hi; ``None`` key = every boundary
    face (colorize=false, boundary id 0). First listed bc wins on shared edges/corners, as
    VectorTools::interpolate_boundary_values does not overwrite existing constraint lines.
    renumber: "cm" (Cuthill–McKee, gls_navier_stokes.cc:70), "none", or an explicit new-index array.
    periodic: directions d whose faces 2d / 2d+1 carry `type = periodic` (boundary_conditions.h,
    DoFTools::make_periodicity_constraints in navier_stokes_base.cc): the dofs of the hi face are
    identified with those of the lo face.  Here: the cells refer to the lo-face dofs directly (the
    same global system as resolving x_hi = x_lo in distribute_local_to_global); the hi-face dofs
    stay in the numbering (deal.II counts them) as constrained rows with a unit diagonal
    (``periodic_slave`` / ``periodic_master``).
    n, lo, hi: scalars (hyper_cube) or one value per direction (a structured rectangle / box).
    """

    def __init__(self, dim, n, pu, pp, lo=-1.0, hi=1.0, bcs=None, renumber="cm", nq1=None,
                 periodic=(), mapping=None):
        """mapping: callable(param[:, dim]) -> x[:, dim].  The box is then the PARAMETER domain of a
        curved mesh whose cells carry MappingQ(pu) geometry on all cells (`qmapping all = true`,
        gls_navier_stokes.cc:244-252): the geometry nodes are the images of the velocity nodes
        (where a manifold puts MappingQ's support points), Jacobians, determinants and the
        mapping's second derivatives are per quadrature point."""
        assert pu % pp == 0
        self.dim, self.ncd, self.pu, self.pp = dim, n, pu, pp
        # per-direction cell counts and extents (scalars = the same in every direction:
        # hyper_cube; sequences: subdivided_hyper_rectangle-like boxes)
        nd = np.array([n] * dim if np.isscalar(n) else list(n), dtype=np.int64)
        lov = np.array([lo] * dim if np.isscalar(lo) else list(lo), dtype=np.float64)
        hiv = np.array([hi] * dim if np.isscalar(hi) else list(hi), dtype=np.float64)
        self.nd = nd
        self.lo, self.hi = (float(lo), float(hi)) if np.isscalar(lo) and np.isscalar(hi) else (lov, hiv)
        self.fe = FETables(dim, pu, pp, nq1)
        fe = self.fe
        hxv = (hiv - lov) / nd
        self.hx = float(hxv[0]) if np.allclose(hxv, hxv[0]) else hxv
        guv = pu * nd + 1                    # velocity node grid size per direction
        ratio = pu // pp
        self.gu = int(guv[0]) if np.all(guv == guv[0]) else guv
        nnode = int(np.prod(guv))
        idx = np.indices(tuple(int(g) for g in guv[::-1])).reshape(dim, -1)[::-1]  # idx[d] along d, x fastest
        # numpy's indices with C order makes the LAST axis fastest; reverse so axis 0 = x fastest
        self.node_idx = idx.T.copy()                              # [nnode, dim]
        has_p = np.all(self.node_idx % ratio == 0, axis=1)
        ndof_node = dim + has_p.astype(np.int64)
        first = np.concatenate([[0], np.cumsum(ndof_node)])
        self.ndof = int(first[-1])
        self.vel_dof = first[:-1, None] + np.arange(dim)[None, :]     # [nnode, dim] provisional ids
        self.p_dof = np.where(has_p, first[:-1] + dim, -1)
        # cells
        cidx = np.indices(tuple(int(v) for v in nd[::-1])).reshape(dim, -1)[::-1].T   # [ncell, dim], x fastest
        self.ncell = cidx.shape[0]
        self.cell_idx = cidx
        stride = np.concatenate([[1], np.cumprod(guv[:-1])])

        def local_nodes(p):
            n1 = p + 1
            a = np.arange(n1 ** dim)
            ai = np.stack([(a // n1 ** d) % n1 for d in range(dim)], axis=1)   # [ns, dim]
            step = pu // p
            nid = ((cidx[:, None, :] * pu + ai[None, :, :] * step) * stride).sum(axis=2)
            return nid                                                         # [ncell, ns]

        un, pn = local_nodes(pu), local_nodes(pp)
        cd = [self.vel_dof[un, c] for c in range(dim)] + [self.p_dof[pn]]
        cell_dofs = np.concatenate(cd, axis=1)
        assert cell_dofs.min() >= 0
        # periodic identification (provisional numbering)
        master_node = np.arange(nnode)
        for d in periodic:
            mi = self.node_idx[master_node].copy()
            mi[mi[:, d] == guv[d] - 1, d] = 0
            master_node = (mi * stride).sum(axis=1)
        master_of = np.arange(self.ndof)
        if len(periodic):
            sl = np.nonzero(master_node != np.arange(nnode))[0]
            for c in range(dim):
                master_of[self.vel_dof[sl, c]] = self.vel_dof[master_node[sl], c]
            slp = sl[has_p[sl]]
            master_of[self.p_dof[slp]] = self.p_dof[master_node[slp]]
            cell_dofs = master_of[cell_dofs]
        # dof meta
        comp = np.zeros(self.ndof, dtype=np.int32)
        dof_node = np.zeros(self.ndof, dtype=np.int64)
        for c in range(dim):
            comp[self.vel_dof[:, c]] = c
            dof_node[self.vel_dof[:, c]] = np.arange(nnode)
        comp[self.p_dof[has_p]] = dim
        dof_node[self.p_dof[has_p]] = np.arange(nnode)[has_p]
        coords = lov + self.node_idx * (hxv / pu)
        self.mapping = mapping
        if mapping is not None:
            coords = np.asarray(mapping(coords), dtype=np.float64)
        # constraints
        constrained = np.zeros(self.ndof, dtype=np.uint8)
        cvalue = np.zeros(self.ndof)
        if bcs is None:
            bcs = {None: ("noslip",)}
        for face, bc in bcs.items():
            if face is None:
                on = np.any((self.node_idx == 0) | (self.node_idx == guv - 1), axis=1)
            else:
                d, side = face // 2, face % 2
                on = self.node_idx[:, d] == (guv[d] - 1 if side else 0)
            nodes = np.nonzero(on)[0]
            vals = np.zeros((nodes.size, dim)) if bc[0] == "noslip" else np.asarray(
                bc[1](coords[nodes]), dtype=np.float64)
            for c in range(dim):
                d_ = self.vel_dof[nodes, c]
                new = constrained[d_] == 0
                constrained[d_[new]] = 1
                cvalue[d_[new]] = vals[new, c]
        slave = master_of != np.arange(self.ndof)
        constrained[slave] = 1
        cvalue[slave] = 0.0
        # renumbering
        if isinstance(renumber, str) and renumber == "cm":
            new_of_old = self._cuthill_mckee(cell_dofs)
        elif isinstance(renumber, str):
            new_of_old = np.arange(self.ndof)
        else:
            new_of_old = np.asarray(renumber)
        self.new_of_old = new_of_old
        self.cell_dofs = np.ascontiguousarray(new_of_old[cell_dofs].astype(np.int32))
        inv = np.empty_like(new_of_old)
        inv[new_of_old] = np.arange(self.ndof)
        self.dof_comp = comp[inv]
        self.dof_coords = coords[dof_node[inv]]
        self.constrained = np.ascontiguousarray(constrained[inv])
        self.constraint_value = cvalue[inv]
        self.periodic_slave = np.nonzero(slave[inv])[0]                  # new ids
        self.periodic_master = new_of_old[master_of[inv[self.periodic_slave]]]
        # geometry (affine Cartesian cells)
        self.cell_invJ = np.ascontiguousarray(
            np.broadcast_to(np.eye(dim) / self.hx if np.isscalar(self.hx) else np.diag(1.0 / hxv),
                            (self.ncell, dim, dim)))
        # (hx ** dim for cubes: the committed fixtures were made with exactly this expression)
        vol = self.hx ** dim if np.isscalar(self.hx) else float(np.prod(hxv))
        self.cell_detJ = np.full(self.ncell, vol)
        self.cell_measure = np.full(self.ncell, vol)
        self.qpoints = np.ascontiguousarray(
            lov + (cidx[:, None, :] + fe.xq[None, :, :]) * hxv)             # [ncell, nq, dim]
        self.geometry_per_q, self.map_lap = False, None
        if mapping is not None:
            assert dim == 2, "curved meshes: 2D only (cell->measure() of a 3D cell is not restated)"
            self.cell_X = coords[un]                                        # [ncell, ns, dim] geometry nodes
            g = self.mapped_geometry(fe)
            self.cell_invJ, self.cell_detJ, self.qpoints, self.map_lap = g
            self.geometry_per_q = True
            # cell->measure(): the straight-sided quadrilateral through the four vertices
            n1 = pu + 1
            corner = [0, pu, n1 * n1 - 1, n1 * pu]                           # counter-clockwise
            P = self.cell_X[:, corner, :]
            x, y = P[:, :, 0], P[:, :, 1]
            self.cell_measure = 0.5 * np.abs((x * np.roll(y, -1, axis=1) - np.roll(x, -1, axis=1) * y).sum(axis=1))
        # colours: no two cells of a colour share a dof.  Parity per direction; across a periodic
        # wrap with an odd cell count the last cell of the direction takes a third colour.
        per_dir = cidx % 2
        radix = 2
        for d in periodic:
            if nd[d] % 2:
                per_dir[cidx[:, d] == nd[d] - 1, d] = 2
                radix = 3
        self.cell_color = (per_dir * (radix ** np.arange(dim))).sum(axis=1).astype(np.int32)
        self.rowptr, self.col = self._sparsity()

    def mapped_geometry(self, fe):
        """MappingQ(pu) at the quadrature points of `fe` (tables of the same FE_Q(pu) basis):
        inverse Jacobians [ncell, nq, dim, dim] (invJ[r][d] = d xi_r / d x_d), determinants,
        real-space points, and c_k = sum_rs d2x_k/dxi_r dxi_s (J^-1 J^-T)_rs."""
        X = self.cell_X
        J = np.einsum("cak,qar->cqkr", X, fe.dNu)               # J[k][r] = d x_k / d xi_r
        H = np.einsum("cak,qars->cqkrs", X, fe.d2Nu)
        K = np.linalg.inv(J)                                    # K[r][d]
        det = np.linalg.det(J)
        assert np.all(det > 0), "negatively oriented cells"
        KKt = np.einsum("cqrd,cqsd->cqrs", K, K)
        c = np.einsum("cqkrs,cqrs->cqk", H, KKt)
        xq = np.einsum("cak,qa->cqk", X, fe.Nu)
        return (np.ascontiguousarray(K), np.ascontiguousarray(det), np.ascontiguousarray(xq),
                np.ascontiguousarray(c))

    def _cuthill_mckee(self, cell_dofs):
        """DoFRenumbering::Cuthill_McKee stand-in: plain Cuthill–McKee =
        reversed scipy RCM on the dof graph where all dofs of a cell couple."""
        import scipy.sparse as sp
        from scipy.sparse.csgraph import reverse_cuthill_mckee
        n = cell_dofs.shape[1]
        rows = np.repeat(cell_dofs, n, axis=1).ravel()
        cols = np.tile(cell_dofs, (1, n)).ravel()
        G = sp.csr_matrix((np.ones(rows.size, dtype=np.int8), (rows, cols)),
                          shape=(self.ndof, self.ndof))
        G.sum_duplicates()
        perm = reverse_cuthill_mckee(G, symmetric_mode=True)[::-1]     # perm[new] = old
        new_of_old = np.empty(self.ndof, dtype=np.int64)
        new_of_old[perm] = np.arange(self.ndof)
        return new_of_old

    def _sparsity(self):
        """DoFTools::make_sparsity_pattern(dof_handler, dsp, constraints, keep_constrained=false)
        (gls_navier_stokes.cc:204-208): couplings between unconstrained dofs of a cell, plus the
        diagonal of every row. CSR with sorted columns, int64 row pointers, int32 columns."""
        import scipy.sparse as sp
        n = self.cell_dofs.shape[1]
        cd = self.cell_dofs.astype(np.int64)
        free = self.constrained[cd] == 0
        rows = np.repeat(cd, n, axis=1)
        cols = np.tile(cd, (1, n))
        keep = np.repeat(free, n, axis=1) & np.tile(free, (1, n))
        rows, cols = rows[keep], cols[keep]
        rows = np.concatenate([rows, np.arange(self.ndof)])
        cols = np.concatenate([cols, np.arange(self.ndof)])
        A = sp.csr_matrix((np.ones(rows.size, dtype=np.int8), (rows, cols)),
                          shape=(self.ndof, self.ndof))
        A.sum_duplicates()
        A.sort_indices()
        return A.indptr.astype(np.int64), A.indices.astype(np.int32)

    # ---- data handed to C -----------------------------------------------------------------
    def cells_struct(self, force=None, with_qpoints=True):
        keep = [self.cell_dofs, self.cell_invJ, self.cell_detJ, self.cell_measure,
                self.qpoints if with_qpoints else None,
                None if force is None else np.ascontiguousarray(force, dtype=np.float64)]
        s = _Cells(self.ncell, _p(keep[0], c_i32_p), *[_p(a, c_double_p) for a in keep[1:]],
                   1 if self.geometry_per_q else 0, _p(self.map_lap, c_double_p))
        s._keep = keep
        return s

    def color_lists(self):
        order = np.argsort(self.cell_color, kind="stable").astype(np.int32)
        ncolor = int(self.cell_color.max()) + 1
        ptr = np.concatenate([[0], np.cumsum(np.bincount(self.cell_color, minlength=ncolor))])
        return ptr.astype(np.int32), order, ncolor

    def evaluate_force(self, fn):
        """Forcing values at the quadrature points (gls_navier_stokes.cc:364-369)."""
        f = fn(self.qpoints.reshape(-1, self.dim))
        return np.ascontiguousarray(f[:, :self.dim].reshape(self.ncell, self.fe.nq, self.dim))

    def distribute_periodic(self, U):
        """x_hi = x_lo on the periodic faces (AffineConstraints::distribute for those lines)."""
        U = U.copy()
        U[self.periodic_slave] = U[self.periodic_master]
        return U

    def unit_diagonal_on_periodic_slaves(self, val):
        """Rows of the identified hi-face dofs: no cell refers to them, give them a unit diagonal
        (deal.II's constrained rows carry a positive diagonal as well; nothing couples to them)."""
        if self.periodic_slave.size:
            rs = self.rowptr[self.periodic_slave]
            assert np.all(self.rowptr[self.periodic_slave + 1] - rs == 1)
            val[rs] = 1.0
        return val

    def apply_nonzero_constraints(self, U):
        """PhysicsSolver::apply_constraints -> nonzero_constraints.distribute
        (include/core/physics_solver.h:98-102)."""
        U = U.copy()
        m = self.constrained != 0
        U[m] = self.constraint_value[m]
        return U


class RefinedBoxMesh(BoxMesh):
    """A box mesh of n^dim cells of which the cells selected by `refine(cell_centre[:, dim]) -> bool`
    are refined once (one level of Kelly / uniform-in-a-region refinement,
    navier_stokes_base.cc:610-729): a non-conforming mesh with HANGING NODES on the faces and edges
    between refined and unrefined cells, constrained as DoFTools::make_hanging_node_constraints does
    (setup_dofs, gls_navier_stokes.cc:57-228): the value at a fine node on a coarse cell's boundary is
    the coarse cell's trace there, x_i = sum_a N_a(xi_i) x_a.  In the closed zero_constraints the
    Dirichlet masters drop out; nonzero_constraints keep them as an inhomogeneity.  Everything else
    (dof numbering by Cuthill-McKee, sparsity with keep_constrained_dofs = false, boundary values,
    colours) follows BoxMesh.  parity: unpinned -- no reference test with hanging nodes can be
    reproduced here (cylinder_gls needs gmsh + Kelly); the class is checked by a patch test (a
    solution in the FE space is reproduced to rounding on the non-conforming mesh)."""

    def __init__(self, dim, n, pu, pp, refine, lo=-1.0, hi=1.0, bcs=None, renumber="cm", nq1=None):
        assert pu % pp == 0
        self.dim, self.ncd, self.pu, self.pp = dim, n, pu, pp
        self.lo, self.hi = float(lo), float(hi)
        self.fe = fe = FETables(dim, pu, pp, nq1)
        self.geometry_per_q, self.map_lap, self.mapping = False, None, None
        H = (hi - lo) / n
        L = 2 * pu                                   # lattice units per coarse cell (fine spacing = 1)
        base = np.indices((n,) * dim).reshape(dim, -1)[::-1].T            # x fastest
        centre = lo + (base + 0.5) * H
        ref = np.asarray(refine(centre), dtype=bool)
        orig, size = [], []                                                # per active cell (lattice units)
        child = np.indices((2,) * dim).reshape(dim, -1)[::-1].T
        for c in range(base.shape[0]):
            if ref[c]:
                for ch in child:
                    orig.append(base[c] * L + ch * pu)
                    size.append(pu)
            else:
                orig.append(base[c] * L)
                size.append(L)
        orig, size = np.array(orig, dtype=np.int64), np.array(size, dtype=np.int64)
        self.ncell = ncell = orig.shape[0]
        self.cell_is_fine = size == pu

        def lattice(p):                                                    # local node offsets in units of size/p
            n1 = p + 1
            a = np.arange(n1 ** dim)
            return np.stack([(a // n1 ** d) % n1 for d in range(dim)], axis=1)
        au, ap = lattice(pu), lattice(pp)
        un_xyz = orig[:, None, :] + au[None, :, :] * (size[:, None, None] // pu)     # [ncell, n_su, dim]
        pn_xyz = orig[:, None, :] + ap[None, :, :] * (size[:, None, None] // pp)
        G = n * L + 1
        key = lambda xyz: (xyz * (G ** np.arange(dim))).sum(axis=-1)
        ukeys, pkeys = key(un_xyz), key(pn_xyz)
        node_keys = np.unique(ukeys)
        nnode = node_keys.size
        node_of = lambda k: np.searchsorted(node_keys, k)
        has_p = np.zeros(nnode, dtype=bool)
        has_p[node_of(np.unique(pkeys))] = True
        node_xyz = np.stack([(node_keys // G ** d) % G for d in range(dim)], axis=1)
        ndof_node = dim + has_p.astype(np.int64)
        first = np.concatenate([[0], np.cumsum(ndof_node)])
        self.ndof = int(first[-1])
        vel_dof = first[:-1, None] + np.arange(dim)[None, :]
        p_dof = np.where(has_p, first[:-1] + dim, -1)
        un, pn = node_of(ukeys), node_of(pkeys)
        cell_dofs = np.concatenate([vel_dof[un, c] for c in range(dim)] + [p_dof[pn]], axis=1)
        assert cell_dofs.min() >= 0
        comp = np.zeros(self.ndof, dtype=np.int32)
        dof_node = np.zeros(self.ndof, dtype=np.int64)
        for c in range(dim):
            comp[vel_dof[:, c]] = c
            dof_node[vel_dof[:, c]] = np.arange(nnode)
        comp[p_dof[has_p]] = dim
        dof_node[p_dof[has_p]] = np.arange(nnode)[has_p]
        coords = lo + node_xyz * (H / L)
        # ---- hanging-node lines: fine nodes on the boundary of a coarse cell that are not its own ----
        lines = {}                                                         # dof -> {master dof: weight}
        keyset = {int(k): i for i, k in enumerate(node_keys)}

        def trace_lines(p, own_dof, has):                                  # element of degree p
            step = L // p
            half = np.indices((2 * p + 1,) * dim).reshape(dim, -1)[::-1].T            # half-lattice of the coarse cell
            half = half[np.any((half == 0) | (half == 2 * p), axis=1) & np.any(half % 2 == 1, axis=1)]
            xi = half / (2.0 * p)
            V = [lagrange1d(p, xi[:, d])[0] for d in range(dim)]           # [p+1, npts] per direction
            lat = lattice(p)
            W = np.ones((half.shape[0], lat.shape[0]))
            for d in range(dim):
                W *= V[d][lat[:, d], :].T
            for c in np.nonzero(~self.cell_is_fine)[0]:
                pts = orig[c] + half * (step // 2)
                masters_nodes = node_of(key(orig[c] + lat * step))
                for t in range(pts.shape[0]):
                    nd = keyset.get(int(key(pts[t])))
                    if nd is None or not has(nd):
                        continue
                    w = W[t]
                    nz = np.nonzero(np.abs(w) > 1e-14)[0]
                    for dofs_of in own_dof:                                # one line per component
                        lines.setdefault(int(dofs_of(nd)), {int(dofs_of(masters_nodes[a])): float(w[a]) for a in nz})
        trace_lines(pu, [(lambda nd, c=c: vel_dof[nd, c]) for c in range(dim)], lambda nd: True)
        trace_lines(pp, [lambda nd: p_dof[nd]], lambda nd: has_p[nd])
        # ---- boundary values (not on dofs that already carry a hanging-node line) ----
        constrained = np.zeros(self.ndof, dtype=np.uint8)
        cvalue = np.zeros(self.ndof)
        for d_ in lines:
            constrained[d_] = 2
        if bcs is None:
            bcs = {None: ("noslip",)}
        dirichlet_value = {}
        for face, bc in bcs.items():
            if face is None:
                on = np.any((node_xyz == 0) | (node_xyz == G - 1), axis=1)
            else:
                d, side = face // 2, face % 2
                on = node_xyz[:, d] == (G - 1 if side else 0)
            nodes = np.nonzero(on)[0]
            vals = np.zeros((nodes.size, dim)) if bc[0] == "noslip" else np.asarray(
                bc[1](coords[nodes]), dtype=np.float64)
            for c in range(dim):
                d_ = vel_dof[nodes, c]
                for k, dd in enumerate(d_):
                    if dd not in dirichlet_value:                          # first listed boundary wins
                        dirichlet_value[int(dd)] = float(vals[k, c])
                new = constrained[d_] == 0
                constrained[d_[new]] = 1
                cvalue[d_[new]] = vals[new, c]
        # close(): Dirichlet masters leave the zero_constraints lines and become an inhomogeneity of
        # the nonzero_constraints lines
        hang_rows, hang_inhom = {}, {}
        for d_, ms in lines.items():
            hang_rows[d_] = {m: w for m, w in ms.items() if constrained[m] == 0}
            hang_inhom[d_] = sum(w * dirichlet_value.get(m, 0.0) for m, w in ms.items() if constrained[m] == 1)
            assert all(constrained[m] != 2 for m in ms), "chained hanging nodes (more than one level)"
        # ---- renumbering on the graph of the expanded cells ----
        def expand(cd):
            out = []
            for g in cd:
                out.extend(hang_rows[g].keys() if constrained[g] == 2 else [g])
            return sorted(set(out) | set(cd))
        exp_cells = [expand(cell_dofs[c].tolist()) for c in range(ncell)]
        if isinstance(renumber, str) and renumber == "cm":
            import scipy.sparse as sp
            from scipy.sparse.csgraph import reverse_cuthill_mckee
            rows = np.concatenate([np.repeat(e, len(e)) for e in exp_cells])
            cols = np.concatenate([np.tile(e, len(e)) for e in exp_cells])
            Gm = sp.csr_matrix((np.ones(rows.size, dtype=np.int8), (rows, cols)), shape=(self.ndof, self.ndof))
            Gm.sum_duplicates()
            perm = reverse_cuthill_mckee(Gm, symmetric_mode=True)[::-1]
            new_of_old = np.empty(self.ndof, dtype=np.int64)
            new_of_old[perm] = np.arange(self.ndof)
        elif isinstance(renumber, str):
            new_of_old = np.arange(self.ndof)
        else:
            new_of_old = np.asarray(renumber)
        self.new_of_old = new_of_old
        inv = np.empty_like(new_of_old)
        inv[new_of_old] = np.arange(self.ndof)
        self.cell_dofs = np.ascontiguousarray(new_of_old[cell_dofs].astype(np.int32))
        self.dof_comp = comp[inv]
        self.dof_coords = coords[dof_node[inv]]
        self.constrained = np.ascontiguousarray(constrained[inv])
        self.constraint_value = cvalue[inv]
        self.periodic_slave = np.zeros(0, dtype=np.int64)
        self.periodic_master = np.zeros(0, dtype=np.int64)
        # hanging-node lines in the new numbering, CSR over all dofs
        ptr, idx, wts = [0], [], []
        self.hang_inhomogeneity = np.zeros(self.ndof)
        for g in range(self.ndof):
            old = int(inv[g])
            if constrained[old] == 2:
                for m, w in sorted((int(new_of_old[m]), w) for m, w in hang_rows[old].items()):
                    idx.append(m)
                    wts.append(w)
                self.hang_inhomogeneity[g] = hang_inhom[old]
            ptr.append(len(idx))
        self.hang_ptr = np.array(ptr, dtype=np.int64)
        self.hang_idx = np.array(idx, dtype=np.int32)
        self.hang_w = np.array(wts, dtype=np.float64)
        # ---- geometry (affine, two cell sizes) ----
        hc = size * (H / L)
        self.cell_invJ = np.ascontiguousarray(np.eye(dim)[None] / hc[:, None, None])
        self.cell_detJ = hc ** dim
        self.cell_measure = hc ** dim
        self.qpoints = np.ascontiguousarray(lo + orig[:, None, :] * (H / L) + fe.xq[None, :, :] * hc[:, None, None])
        self.cell_idx = None
        # ---- colours on the expanded cells (greedy), sparsity on them ----
        new_exp = [np.array(sorted(int(new_of_old[g]) for g in e), dtype=np.int64) for e in exp_cells]
        used = [set() for _ in range(self.ndof)]
        color = np.zeros(ncell, dtype=np.int32)
        for c in range(ncell):
            k = 0
            while any(k in used[g] for g in new_exp[c]):
                k += 1
            color[c] = k
            for g in new_exp[c]:
                used[g].add(k)
        self.cell_color = color
        import scipy.sparse as sp
        rows, cols = [], []
        for e in new_exp:
            f = e[self.constrained[e] == 0]
            rows.append(np.repeat(f, f.size))
            cols.append(np.tile(f, f.size))
        rows = np.concatenate(rows + [np.arange(self.ndof)])
        cols = np.concatenate(cols + [np.arange(self.ndof)])
        A = sp.csr_matrix((np.ones(rows.size, dtype=np.int8), (rows, cols)), shape=(self.ndof, self.ndof))
        A.sum_duplicates()
        A.sort_indices()
        self.rowptr, self.col = A.indptr.astype(np.int64), A.indices.astype(np.int32)

    def distribute_hanging(self, U, inhomogeneous=True):
        """AffineConstraints::distribute for the hanging-node lines: x_i = sum w x_master (+ the part
        that Dirichlet masters contribute in nonzero_constraints)."""
        U = U.copy()
        h = np.nonzero(self.constrained == 2)[0]
        for g in h:
            s = self.hang_ptr[g], self.hang_ptr[g + 1]
            U[g] = float(self.hang_w[s[0]:s[1]] @ U[self.hang_idx[s[0]:s[1]]]) + \
                (self.hang_inhomogeneity[g] if inhomogeneous else 0.0)
        return U

    def apply_nonzero_constraints(self, U):
        U = U.copy()
        m = self.constrained == 1
        U[m] = self.constraint_value[m]
        return self.distribute_hanging(U, True)


# ----------------------------------------------------------------------------------------------
# time integration scalars
# ----------------------------------------------------------------------------------------------
SCHEMES = ("steady", "bdf1", "bdf2", "bdf3", "sdirk2_1", "sdirk2_2", "sdirk3_1", "sdirk3_2",
           "sdirk3_3")


def bdf_coefficients(order, dts):
    out = np.zeros(order + 1)
    d = np.ascontiguousarray(dts, dtype=np.float64)
    lib().glso_bdf_coefficients(C.c_int(order), _p(d, c_double_p), _p(out, c_double_p))
    return out


def sdirk_coefficients(order, dt):
    """source/core/sdirk.cc:11-44."""
    sdt = 1.0 / dt
    if order == 2:
        alpha = (2.0 - math.sqrt(2)) / 2.0
        m = np.zeros((2, 3))
        m[0, 0] = 1.0 / alpha * sdt
        m[0, 1] = -1.0 / alpha * sdt
        m[1, 0] = 1.0 / alpha * sdt
        m[1, 1] = -(2 * alpha - 1) / alpha / alpha * sdt
        m[1, 2] = -(1 - alpha) / alpha / alpha * sdt
        return m
    m = np.zeros((3, 4))
    m[0, 0] = 2.29428036027904 * sdt
    m[0, 1] = -2.29428036027904 * sdt
    m[1, 0] = 2.29428036027904 * sdt
    m[1, 1] = -0.809559354637498 * sdt
    m[1, 2] = -1.48472100564154 * sdt
    m[2, 0] = 2.29428036027904 * sdt
    m[2, 1] = 2.87009860433106 * sdt
    m[2, 2] = -8.55612780155264 * sdt
    m[2, 3] = 3.39174883694255 * sdt
    return m


def scheme_params(scheme, dts, viscosity, srf=False, omega=(0.0, 0.0, 0.0)):
    """The compile-time dispatch of assemble_matrix_and_rhs / assemble_rhs
    (gls_navier_stokes.cc:916-1128) reduced to (transient, 1/dt, coefs[4])."""
    p = _Params()
    p.viscosity = viscosity
    p.srf = 1 if srf else 0
    p.omega[:] = omega
    coefs = np.zeros(4)
    if scheme == "steady":
        p.transient, p.sdt = 0, 0.0
    else:
        p.transient, p.sdt = 1, 1.0 / dts[0]
        if scheme.startswith("bdf"):
            order = int(scheme[3])
            coefs[:order + 1] = bdf_coefficients(order, dts)
        else:
            order, step = int(scheme[5]), int(scheme[7])
            row = sdirk_coefficients(order, dts[0])[step - 1]
            coefs[:step + 1] = row[:step + 1]
    p.coefs[:] = coefs
    return p


# ----------------------------------------------------------------------------------------------
# hot path wrappers
# ----------------------------------------------------------------------------------------------
def assemble(mesh, U, params, assemble_matrix=True, force=None, U1=None, U2=None, U3=None,
             return_local=False, threads=1, structured=False):
    """assembleGLS. Returns (val or None, rhs[, localM, localb]).  structured=True: the same cell
    matrices from the structured block form (gls_oracle.c: cell_structured, bench.py's best-CPU
    figure), not the reference's literal loop."""
    L = lib()
    L.glso_set_cell_mode(C.c_int(1 if structured else 0))
    if getattr(mesh, "hang_ptr", None) is not None:
        L.glso_set_hanging(_p(mesh.hang_ptr, c_i64_p), _p(mesh.hang_idx, c_i32_p), _p(mesh.hang_w, c_double_p))
    try:
        return _assemble(L, mesh, U, params, assemble_matrix, force, U1, U2, U3, return_local, threads)
    finally:
        L.glso_set_cell_mode(C.c_int(0))
        L.glso_set_hanging(None, None, None)


def _assemble(L, mesh, U, params, assemble_matrix, force, U1, U2, U3, return_local, threads):
    fe_s, cs = mesh.fe.cstruct(), mesh.cells_struct(force)
    N, n = mesh.ndof, mesh.fe.n
    nnz = int(mesh.rowptr[-1])
    val = np.zeros(nnz) if assemble_matrix else None
    rhs = np.zeros(N)
    U = np.ascontiguousarray(U, dtype=np.float64)
    hist = [None if a is None else np.ascontiguousarray(a, dtype=np.float64) for a in (U1, U2, U3)]
    common = [C.byref(fe_s), C.byref(cs), C.byref(params), C.c_int64(N), _p(U, c_double_p),
              *[_p(a, c_double_p) for a in hist], _p(mesh.constrained, c_u8_p),
              _p(mesh.rowptr, c_i64_p), _p(mesh.col, c_i32_p), C.c_int(1 if assemble_matrix else 0),
              _p(val, c_double_p), _p(rhs, c_double_p)]
    if threads == 1:
        lM = np.zeros((mesh.ncell, n, n)) if (return_local and assemble_matrix) else None
        lb = np.zeros((mesh.ncell, n)) if return_local else None
        L.glso_assemble(*common, _p(lM, c_double_p), _p(lb, c_double_p))
        if return_local:
            if assemble_matrix and getattr(mesh, "periodic_slave", np.zeros(0)).size:
                mesh.unit_diagonal_on_periodic_slaves(val)
            return val, rhs, lM, lb
    else:
        ptr, order, ncolor = mesh.color_lists()
        L.glso_set_num_threads(C.c_int(threads))
        L.glso_assemble_mt(*common, _p(ptr, c_i32_p), _p(order, c_i32_p), C.c_int(ncolor))
    if assemble_matrix and getattr(mesh, "periodic_slave", np.zeros(0)).size:
        mesh.unit_diagonal_on_periodic_slaves(val)
    return val, rhs


def spmv(mesh, val, x):
    y = np.zeros(mesh.ndof)
    x = np.ascontiguousarray(x, dtype=np.float64)
    lib().glso_spmv(C.c_int64(mesh.ndof), _p(mesh.rowptr, c_i64_p), _p(mesh.col, c_i32_p),
                    _p(val, c_double_p), _p(x, c_double_p), _p(y, c_double_p))
    return y


def _blocks(ndof, block_ptr):
    if block_ptr is None:
        block_ptr = np.array([0, ndof], dtype=np.int64)
    block_ptr = np.ascontiguousarray(block_ptr, dtype=np.int64)
    return block_ptr, block_ptr.size - 1


def ilu0(mesh, val, atol=1e-8, rtol=1.0, block_ptr=None):
    """setup_ILU with fill 0. Returns (lu, diag_pos)."""
    bp, nb = _blocks(mesh.ndof, block_ptr)
    lu = np.zeros_like(val)
    dp = np.zeros(mesh.ndof, dtype=np.int64)
    st = lib().glso_ilu0(C.c_int64(mesh.ndof), _p(mesh.rowptr, c_i64_p), _p(mesh.col, c_i32_p),
                         _p(val, c_double_p), C.c_double(atol), C.c_double(rtol), C.c_int(nb),
                         _p(bp, c_i64_p), _p(lu, c_double_p), _p(dp, c_i64_p))
    if st:
        raise ZeroDivisionError("zero pivot in row %d" % (st - 1))
    return lu, dp


def iluk_pattern(mesh, fill, block_ptr=None):
    """Level-of-fill pattern ILU(fill) of every diagonal block (Ifpack_IlukGraph behind
    TrilinosWrappers::PreconditionILU::AdditionalData(ilu_fill, ...), call site
    gls_navier_stokes.cc:1166-1175): entries of A have level 0, eliminating row i with row k creates
    (i, j) for each (k, j), j > k, at level lev(i,k) + lev(k,j) + 1, kept when <= fill.
    Returns a mesh-like object (ndof, rowptr, col: A's entries plus the fill positions) and `a2p`,
    the position of every entry of A in it.  Pure Python: small cases only."""
    import heapq
    import types
    bp, nb = _blocks(mesh.ndof, block_ptr)
    rows, upper = [], {}
    for b in range(nb):
        b0, b1 = int(bp[b]), int(bp[b + 1])
        for i in range(b0, b1):
            cols = mesh.col[mesh.rowptr[i]:mesh.rowptr[i + 1]]
            lev = {int(c): 0 for c in cols if b0 <= c < b1}
            outside = [int(c) for c in cols if not b0 <= c < b1]
            todo = [c for c in lev if c < i]
            heapq.heapify(todo)
            while todo:
                k = heapq.heappop(todo)
                for j, lkj in upper[k]:
                    nl = lev[k] + lkj + 1
                    if nl > fill:
                        continue
                    if j in lev:
                        lev[j] = min(lev[j], nl)
                    else:
                        lev[j] = nl
                        if j < i:
                            heapq.heappush(todo, j)
            upper[i] = sorted((j, l) for j, l in lev.items() if j > i)
            rows.append(sorted(list(lev) + outside))
    rowptr = np.zeros(mesh.ndof + 1, dtype=np.int64)
    rowptr[1:] = np.cumsum([len(r) for r in rows])
    col = np.array([c for r in rows for c in r], dtype=np.int32)
    a2p = np.empty(mesh.rowptr[-1], dtype=np.int64)
    for i in range(mesh.ndof):
        a2p[mesh.rowptr[i]:mesh.rowptr[i + 1]] = rowptr[i] + np.searchsorted(
            col[rowptr[i]:rowptr[i + 1]], mesh.col[mesh.rowptr[i]:mesh.rowptr[i + 1]])
    return types.SimpleNamespace(ndof=mesh.ndof, rowptr=rowptr, col=col), a2p


def pad_values(padded, a2p, val):
    """A's values on the level-of-fill pattern (explicit zeros at the fill positions)."""
    out = np.zeros(padded.rowptr[-1])
    out[a2p] = val
    return out


def ilu_apply(mesh, lu, dp, r, block_ptr=None):
    bp, nb = _blocks(mesh.ndof, block_ptr)
    z = np.zeros(mesh.ndof)
    r = np.ascontiguousarray(r, dtype=np.float64)
    lib().glso_ilu_apply(C.c_int64(mesh.ndof), _p(mesh.rowptr, c_i64_p), _p(mesh.col, c_i32_p),
                         _p(lu, c_double_p), _p(dp, c_i64_p), C.c_int(nb), _p(bp, c_i64_p),
                         _p(r, c_double_p), _p(z, c_double_p))
    return z


def gmres(mesh, val, lu, dp, b, tol, max_iters=1000, restart=30, block_ptr=None):
    """Returns (x, iterations, true_residual, converged, residual_history)."""
    bp, nb = _blocks(mesh.ndof, block_ptr)
    x = np.zeros(mesh.ndof)
    it = C.c_int(0)
    tr = C.c_double(0)
    hist = np.zeros(max_iters + 1)
    b = np.ascontiguousarray(b, dtype=np.float64)
    st = lib().glso_gmres(C.c_int64(mesh.ndof), _p(mesh.rowptr, c_i64_p), _p(mesh.col, c_i32_p),
                          _p(val, c_double_p), _p(lu, c_double_p), _p(dp, c_i64_p), C.c_int(nb),
                          _p(bp, c_i64_p), _p(b, c_double_p), C.c_double(tol), C.c_int(max_iters),
                          C.c_int(restart), _p(x, c_double_p), C.byref(it), C.byref(tr),
                          _p(hist, c_double_p))
    return x, it.value, tr.value, st == 0, hist[:it.value + 1]


def bicgstab(mesh, val, lu, dp, b, tol, max_iters=1000, block_ptr=None):
    """solve_system_BiCGStab (gls_navier_stokes.cc:1291-1340): TrilinosWrappers::SolverBicgstab ->
    AztecOO AZ_bicgstab, right-preconditioned with the ILU, zero initial guess, stop on ||r||_2 < tol.
    PARITY UNPINNED (see the module header).  Returns (x, iterations, true_residual, converged)."""
    n = mesh.ndof
    b = np.ascontiguousarray(b, dtype=np.float64)
    x = np.zeros(n)
    r = b.copy()
    rt = b.copy()
    rho = float(rt @ r)
    res = math.sqrt(float(r @ r))
    rho_old = alpha = omega = 1.0
    p = v = None
    it = 0
    ok = res < tol
    while not ok and it < max_iters:
        if rho == 0.0 or omega == 0.0:
            break
        if it == 0:
            p = r.copy()
        else:
            beta = (rho / rho_old) * (alpha / omega)
            p = p - omega * v
            p = r + beta * p
        ph = ilu_apply(mesh, lu, dp, p, block_ptr)
        v = spmv(mesh, val, ph)
        d = float(rt @ v)
        if d == 0.0:
            break
        alpha = rho / d
        x = x + alpha * ph
        r = r - alpha * v                      # s
        sh = ilu_apply(mesh, lu, dp, r, block_ptr)
        t = spmv(mesh, val, sh)
        tt = float(t @ t)
        omega = float(r @ t) / tt if tt != 0.0 else 0.0
        x = x + omega * sh
        r = r - omega * t
        rho_old, rho = rho, float(rt @ r)
        res = math.sqrt(float(r @ r))
        it += 1
        ok = res < tol
    tr = float(np.linalg.norm(b - spmv(mesh, val, x)))
    return x, it, tr, ok


class NoConvergence(RuntimeError):
    """SolverControl::NoConvergence stand-in."""


def solve_linear_system(mesh, val, rhs, rel=1e-3, abs_=1e-8, max_iters=1000, ilu_atol=1e-8,
                        ilu_rtol=1.0, restart=30, block_ptr=None, method="gmres", ilu_fill=0):
    """solve_system_GMRES / solve_system_BiCGStab (gls_navier_stokes.cc:1242-1340).
    Returns (newton_update, iters, res)."""
    tol = max(rel * float(np.linalg.norm(rhs)), abs_)
    constrained = mesh.constrained
    if ilu_fill > 0:                     # ILU(k) = ILU(0) on the level-of-fill pattern
        mesh, a2p = iluk_pattern(mesh, ilu_fill, block_ptr)
        val = pad_values(mesh, a2p, val)
    lu, dp = ilu0(mesh, val, ilu_atol, ilu_rtol, block_ptr)
    if method == "bicgstab":
        x, it, res, ok = bicgstab(mesh, val, lu, dp, rhs, tol, max_iters, block_ptr)
    elif method == "gmres":
        x, it, res, ok, _ = gmres(mesh, val, lu, dp, rhs, tol, max_iters, restart, block_ptr)
    else:
        raise RuntimeError("This solver is not allowed")     # :1158
    if not ok:
        raise NoConvergence("%s did not converge in %d iterations" % (method, it))
    x[constrained != 0] = 0.0           # zero_constraints.distribute (:1287)
    if getattr(mesh, "hang_ptr", None) is not None:
        x = mesh.distribute_hanging(x, inhomogeneous=False)
    return x, it, res


def assemble_l2_projection(mesh, initial):
    """assemble_L2_projection (gls_navier_stokes.cc:829-914), literally: per cell, per quadrature
    point, local(i,j) += (phi_u_j . phi_u_i + phi_p_j phi_p_i) JxW, local_rhs(i) += (phi_u_i . u0 +
    phi_p_i p0) JxW, then nonzero_constraints.distribute_local_to_global (:902-909, default
    use_inhomogeneities_for_rhs = false: constrained row -> |local(i,i)| on the diagonal, zero rhs;
    constrained column j -> rhs_i -= local(i,j) g_j).  `initial(x)`: points [m][dim] -> [m][dim+1].
    Returns (val on mesh's CSR, rhs).  numpy, small cases."""
    fe, dim = mesh.fe, mesh.dim
    n_su, n_sp, nq = fe.Nu.shape[1], fe.Np.shape[1], fe.Nu.shape[0]
    n = dim * n_su + n_sp
    comp = np.concatenate([np.repeat(np.arange(dim), n_su), np.full(n_sp, dim)])
    phi = np.zeros((nq, n))                       # value of the dof's own component
    for c in range(dim):
        phi[:, c * n_su:(c + 1) * n_su] = fe.Nu
    phi[:, dim * n_su:] = fe.Np
    same = comp[:, None] == comp[None, :]
    val = np.zeros(mesh.rowptr[-1])
    rhs = np.zeros(mesh.ndof)
    g = np.where(mesh.constrained != 0, mesh.constraint_value, 0.0)
    for c in range(mesh.ncell):
        dofs = mesh.cell_dofs[c]
        JxW = np.broadcast_to(mesh.cell_detJ[c], (nq,)) * fe.wq
        u0 = initial(mesh.qpoints[c].reshape(nq, dim))           # [nq][dim+1]
        local = np.einsum("qi,qj,q->ij", phi, phi, JxW) * same
        lrhs = np.einsum("qi,qi,q->i", phi, u0[:, comp], JxW)
        con = mesh.constrained[dofs] != 0
        for i in range(n):
            gi = dofs[i]
            rs, re = mesh.rowptr[gi], mesh.rowptr[gi + 1]
            if con[i]:
                val[rs + np.searchsorted(mesh.col[rs:re], gi)] += abs(local[i, i])
                continue
            r = lrhs[i]
            for j in range(n):
                if local[i, j] == 0.0 and not same[i, j]:
                    continue
                if con[j]:
                    r -= local[i, j] * g[dofs[j]]
                else:
                    val[rs + np.searchsorted(mesh.col[rs:re], dofs[j])] += local[i, j]
            rhs[gi] += r
    if getattr(mesh, "periodic_slave", np.zeros(0)).size:
        mesh.unit_diagonal_on_periodic_slaves(val)
    return val, rhs


def l2_projection(mesh, initial, ilu_atol=1e-8, rel=1e-15, abs_=1e-15):
    """set_initial_condition(L2projection) (gls_navier_stokes.cc:795-803): assemble_L2_projection,
    solve_system_GMRES(true, 1e-15, 1e-15, true), present_solution = newton_update (constrained dofs
    take the nonzero constraint values, :1287)."""
    val, rhs = assemble_l2_projection(mesh, initial)
    tol = max(rel * float(np.linalg.norm(rhs)), abs_)
    lu, dp = ilu0(mesh, val, ilu_atol, 1.0)
    x, it, res, ok, _ = gmres(mesh, val, lu, dp, rhs, tol, 1000, 30)
    x = mesh.apply_nonzero_constraints(x)
    if getattr(mesh, "periodic_slave", np.zeros(0)).size:
        x = mesh.distribute_periodic(x)
    return x, it, ok


def calculate_cfl(mesh, U, time_step):
    """calculate_CFL (source/solvers/postprocessing_cfl.cc:34-87): QGauss(1), h from the cell measure
    and fe.degree (the FESystem's: the larger of the two orders), max over the cells."""
    dim = mesh.dim
    n_su = mesh.fe.Nu.shape[1]
    degree = float(max(mesh.pu, mesh.pp))
    Nc = shape_at_centre(mesh.dim, mesh.pu)
    cfl = 0.0
    for c in range(mesh.ncell):
        meas = mesh.cell_measure[c]
        h = (math.sqrt(4.0 * meas / math.pi) if dim == 2 else (6 * meas / math.pi) ** (1.0 / 3.0)) / degree
        u = np.array([Nc @ U[mesh.cell_dofs[c][d * n_su:(d + 1) * n_su]] for d in range(dim)])
        cfl = max(cfl, float(np.linalg.norm(u)) / h * time_step)
    return cfl


def shape_at_centre(dim, p):
    """Scalar FE_Q(p) shape functions at the point of QGauss(1), in the ordering of FETables."""
    Nu, _, _ = _tensor_tables(dim, p, np.array([0.5]))[:3]
    return np.asarray(Nu).reshape(-1)


def newton_solve(mesh, U0, params, force=None, tol=1e-6, max_it=10, lin=None, hist=(None,) * 3,
                 log=None):
    """NewtonNonLinearSolver::solve (include/core/newton_non_linear_solver.h:76-139).
    `log` collects (gmres_iters, true_res) per Newton step, like restart_01.output."""
    lin = lin or {}
    present = U0.copy()
    current_res = last_res = 1.0
    it = 0
    U1, U2, U3 = hist
    while current_res > tol and it < max_it:
        evaluation_point = present
        val, rhs = assemble(mesh, evaluation_point, params, True, force, U1, U2, U3)
        if it == 0:
            current_res = float(np.linalg.norm(rhs))
            last_res = current_res
        dx, k, res = solve_linear_system(mesh, val, rhs, **lin)
        if log is not None:
            log.append((k, res))
        alpha = 1.0
        while alpha > 1e-3:
            local = mesh.apply_nonzero_constraints(present + alpha * dx)
            evaluation_point = local
            _, rhs = assemble(mesh, evaluation_point, params, False, force, U1, U2, U3)
            current_res = float(np.linalg.norm(rhs))
            if current_res < 0.9 * last_res or last_res < tol:
                break
            alpha *= 0.5
        present = evaluation_point
        last_res = current_res
        it += 1
    return present, it, current_res


def sdirk_step(mesh, order, U_m1, dt, nu, force=None, tol=1e-6, max_it=10, lin=None):
    """One time step of `method = sdirk2 | sdirk3` as NavierStokesBase::iterate runs it
    (source/solvers/navier_stokes_base.cc:461-505): stage k solves with method sdirk<order>_k from
    the previous stage's result, the stage results become solution_m2 / solution_m3; solution_m1 is
    the solution of the previous time step.  Returns the solution at the end of the step."""
    dts = [dt] * 4
    hist = [U_m1, None, None]
    U = U_m1
    for k in range(1, order + 1):
        pr = scheme_params("sdirk%d_%d" % (order, k), dts, nu)
        U, _, _ = newton_solve(mesh, U, pr, force, tol=tol, max_it=max_it, lin=lin, hist=tuple(hist))
        if k < order:
            hist[k] = U
    return U


def kinetic_energy(mesh, U):
    """calculate_kinetic_energy (source/solvers/postprocessing_kinetic_energy.cc:34-84):
    sum_q 0.5 |u_h|^2 JxW / volume on QGauss(fe.degree + 1)."""
    dim, fe = mesh.dim, FETables(mesh.dim, mesh.pu, mesh.pp, max(mesh.pu, mesh.pp) + 1)
    JxW = mesh.cell_detJ[:, None] * fe.wq[None, :]
    Uc, n_su = U[mesh.cell_dofs], fe.n_su
    ke = 0.0
    for c in range(dim):
        uh = Uc[:, c * n_su:(c + 1) * n_su] @ fe.Nu.T
        ke += 0.5 * float((uh * uh * JxW).sum())
    return ke / float(JxW.sum())


def enstrophy(mesh, U):
    """calculate_enstrophy (source/solvers/postprocessing_enstrophy.cc:34-93): sum_q 0.5 |curl u_h|^2
    JxW / volume on QGauss(fe.degree + 1)."""
    dim, fe = mesh.dim, FETables(mesh.dim, mesh.pu, mesh.pp, max(mesh.pu, mesh.pp) + 1)
    JxW = mesh.cell_detJ[:, None] * fe.wq[None, :]
    Uc, n_su = U[mesh.cell_dofs], fe.n_su
    # real-space gradients: d/dx_d = sum_r dxi_r/dx_d d/dxi_r, cell_invJ[c][r][d] = dxi_r/dx_d
    gN = np.einsum("qar,crd->cqad", fe.dNu, mesh.cell_invJ)               # [cell][q][a][d]
    g = np.stack([np.einsum("ca,cqad->cqd", Uc[:, c * n_su:(c + 1) * n_su], gN) for c in range(dim)],
                 axis=2)                                                  # [cell][q][comp][d]
    pairs = [(1, 0)] if dim == 2 else [(2, 1), (0, 2), (1, 0)]
    en = 0.0
    for i, j in pairs:                                                    # (du_i/dx_j - du_j/dx_i)
        w = g[:, :, i, j] - g[:, :, j, i]
        en += 0.5 * float((w * w * JxW).sum())
    return en / float(JxW.sum())


def l2_error(mesh, U, exact):
    """calculate_L2_error (navier_stokes_base.cc:255-380): QGauss(nq1+1), mean-free pressure.
    exact(x[:, dim]) -> [:, dim+1]. Returns (err_u, err_p)."""
    dim = mesh.dim
    fe = FETables(dim, mesh.pu, mesh.pp, mesh.fe.nq1 + 1)
    if getattr(mesh, "geometry_per_q", False):     # MappingQ(velocity degree, qmapping all), :262-263
        _, det, xq, _ = mesh.mapped_geometry(fe)
        JxW = det * fe.wq[None, :]
    else:
        xq = mesh.lo + (mesh.cell_idx[:, None, :] + fe.xq[None, :, :]) * mesh.hx   # (scalars or [dim] arrays)
        JxW = mesh.cell_detJ[:, None] * fe.wq[None, :]
    ex = exact(xq.reshape(-1, dim)).reshape(mesh.ncell, fe.nq, dim + 1)
    n_su = fe.n_su
    Uc = U[mesh.cell_dofs]
    uh = np.stack([Uc[:, c * n_su:(c + 1) * n_su] @ fe.Nu.T for c in range(dim)], axis=2)
    ph = Uc[:, dim * n_su:] @ fe.Np.T
    vol = JxW.sum()
    pavg, pex_avg = (ph * JxW).sum() / vol, (ex[:, :, dim] * JxW).sum() / vol
    eu = (((uh - ex[:, :, :dim]) ** 2).sum(axis=2) * JxW).sum()
    ep = ((((ph - pavg) - (ex[:, :, dim] - pex_avg)) ** 2) * JxW).sum()
    return math.sqrt(eu), math.sqrt(ep)
