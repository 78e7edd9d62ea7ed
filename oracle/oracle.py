"""ctypes wrapper around oracle/_build/liboracle.so.

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(insilico_b200/) must never import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")

POINT, LINE, TRI, QUAD, TET, HEX = range(6)
VERTEX, EDGE, FACE, CELL = range(4)
ACTIVE, CONSTRAINED, INACTIVE = range(3)
K_MASS = 7
K_CONVECTION = 8
K_LAPLACE, K_HYPEL_STVENANT, K_HYPEL_NEOHOOKE, K_PRESSURE_GRADIENT, K_VELOCITY_DIVERGENCE, K_VECTOR_LAPLACE = (
    1, 2, 3, 4, 5, 6)
SHAPE_DIM = {LINE: 1, TRI: 2, QUAD: 2, TET: 3, HEX: 3}

_lib = None


def build():
    subprocess.check_call(["make", "-C", _HERE, "-s"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        vp, i32, i64, f64 = C.c_void_p, C.c_int, C.c_int64, C.c_double
        L.orc_problem_new.restype = vp
        L.orc_system_new.restype = vp
        L.orc_system_new.argtypes = [i64]
        L.orc_system_error.restype = C.c_char_p
        for name in ("orc_dof_generate", "orc_sparsity_pattern", "orc_mesh_boundary", "orc_boundary_dof_points",
                     "orc_number_dofs", "orc_nnz"):
            getattr(L, name).restype = i64
        for name in ("orc_rhs_norm", "orc_measure", "orc_l2_error"):
            getattr(L, name).restype = f64
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def shape_nfun(shape, deg):
    return lib().orc_shape_nfun(shape, deg)


def shape_eval(shape, deg, xi):
    n = shape_nfun(shape, deg)
    dim = SHAPE_DIM[shape]
    xi = np.ascontiguousarray(xi, dtype=np.float64)
    f = np.zeros(n)
    g = np.zeros((n, dim))
    lib().orc_shape_eval(shape, deg, _p(xi), _p(f), _p(g))
    return f, g


def support_points(shape, deg):
    n = shape_nfun(shape, deg)
    p = np.zeros((n, SHAPE_DIM[shape]))
    lib().orc_support_points(shape, deg, _p(p))
    return p


def hierarchic_order(shape, deg):
    out = np.zeros(512, dtype=np.int32)
    n = lib().orc_hierarchic_order(shape, deg, _p(out))
    return out[:n].copy()


def quadrature(shape, degree):
    n = lib().orc_quadrature(shape, degree, None, None)
    w = np.zeros(n)
    p = np.zeros((n, SHAPE_DIM[shape]))
    lib().orc_quadrature(shape, degree, _p(w), _p(p))
    return w, p


def face_dofs(shape, deg, nface, face_no):
    out = np.zeros(256, dtype=np.int32)
    n = lib().orc_face_dofs(shape, deg, nface, face_no, _p(out))
    return out[:n].copy()


def ndpe(shape, deg):
    return lib().orc_ndpe(shape, deg)


def unit_cube(dim, simplex, degree, e1, e2=1, e3=1):
    nn, ne, npe = C.c_int64(), C.c_int64(), C.c_int()
    lib().orc_unit_cube_sizes(dim, int(simplex), degree, e1, e2, e3, C.byref(nn), C.byref(ne), C.byref(npe))
    coords = np.zeros((nn.value, dim))
    conn = np.zeros((ne.value, npe.value), dtype=np.int64)
    lib().orc_unit_cube(dim, int(simplex), degree, e1, e2, e3, _p(coords), _p(conn))
    return coords, conn


def sparsity_pattern(elem_dof, n_obj):
    """IndexMap::generateSparsityPattern -> (nnz, 2) array of (dof, connected dof) pairs."""
    ed = np.ascontiguousarray(elem_dof, dtype=np.int64)
    ne, nd = ed.shape
    n = lib().orc_sparsity_pattern(C.c_int64(ne), nd, _p(ed), C.c_int64(n_obj), None)
    pairs = np.zeros((n, 2), dtype=np.int64)
    lib().orc_sparsity_pattern(C.c_int64(ne), nd, _p(ed), C.c_int64(n_obj), _p(pairs))
    return pairs


def number_dofs(status, init=0):
    status = np.ascontiguousarray(status, dtype=np.uint8)
    nobj, ds = status.shape
    eqn = np.zeros((nobj, ds), dtype=np.int64)
    n = lib().orc_number_dofs(C.c_int64(nobj), ds, _p(status), C.c_int64(init), _p(eqn))
    return eqn, n


def surface_points(surf_shape, geom_deg, surf_x, quad_deg):
    """(x, normal, detg) at the surface quadrature points of every surface element (NeumannForce.hpp:152-163)"""
    surf_x = np.ascontiguousarray(surf_x, dtype=np.float64)
    n, _, dim = surf_x.shape
    nq = lib().orc_surface_points(surf_shape, geom_deg, dim, C.c_int64(n), None, quad_deg, None, None, None)
    x, nr, dg = np.zeros((n, nq, dim)), np.zeros((n, nq, dim)), np.zeros((n, nq))
    lib().orc_surface_points(surf_shape, geom_deg, dim, C.c_int64(n), _p(surf_x), quad_deg, _p(x), _p(nr), _p(dg))
    return x, nr, dg


class Problem:
    """Mesh + up to five fields (FieldBinder<Mesh,F1..F5>)."""

    def __init__(self, shape, geom_deg, coords, conn):
        self.h = C.c_void_p(lib().orc_problem_new())
        self.shape, self.geom_deg = shape, geom_deg
        self.coords = np.ascontiguousarray(coords, dtype=np.float64)
        self.conn = np.ascontiguousarray(conn, dtype=np.int64)
        self.dim = self.coords.shape[1]
        self.n_elems = self.conn.shape[0]
        lib().orc_set_mesh(self.h, shape, geom_deg, self.dim, C.c_int64(self.coords.shape[0]), _p(self.coords),
                           C.c_int64(self.n_elems), _p(self.conn))
        self.fields = {}

    def __del__(self):
        try:
            lib().orc_problem_free(self.h)
        except Exception:
            pass

    def dof_generate(self, fe_deg):
        n = ndpe(self.shape, fe_deg)
        elem_dof = np.zeros((self.n_elems, n), dtype=np.int64)
        nobj = lib().orc_dof_generate(self.h, fe_deg, _p(elem_dof))
        return elem_dof, int(nobj)

    def mesh_boundary(self):
        n = lib().orc_mesh_boundary(self.h, None)
        out = np.zeros((n, 2), dtype=np.int64)
        lib().orc_mesh_boundary(self.h, _p(out))
        return out

    def boundary_dof_points(self, fe_deg, pairs):
        pairs = np.ascontiguousarray(pairs, dtype=np.int64)
        n = lib().orc_boundary_dof_points(self.h, fe_deg, C.c_int64(len(pairs)), _p(pairs), None, None, None)
        elem = np.zeros(n, dtype=np.int64)
        loc = np.zeros(n, dtype=np.int32)
        x = np.zeros((n, self.dim))
        lib().orc_boundary_dof_points(self.h, fe_deg, C.c_int64(len(pairs)), _p(pairs), _p(elem), _p(loc), _p(x))
        return elem, loc, x

    def boundary_surface(self, pairs):
        """generateBoundaryMesh for (element, face) pairs: (domain_elem [n], surf_x [n, P, dim], surf_param [n, P, dim])"""
        pairs = np.ascontiguousarray(pairs, dtype=np.int64)
        P = lib().orc_boundary_surface(self.h, C.c_int64(len(pairs)), _p(pairs), None, None, None)
        de = np.zeros(len(pairs), dtype=np.int64)
        sx = np.zeros((len(pairs), P, self.dim))
        sp = np.zeros((len(pairs), P, self.dim))
        lib().orc_boundary_surface(self.h, C.c_int64(len(pairs)), _p(pairs), _p(de), _p(sx), _p(sp))
        return de, sx, sp

    def set_field(self, fid, fe_deg, dof_size, n_obj, elem_dof, eqn, status, prescribed, values):
        a = dict(elem_dof=np.ascontiguousarray(elem_dof, dtype=np.int64),
                 eqn=np.ascontiguousarray(eqn, dtype=np.int64),
                 status=np.ascontiguousarray(status, dtype=np.uint8),
                 prescribed=np.ascontiguousarray(prescribed, dtype=np.float64),
                 values=np.ascontiguousarray(values, dtype=np.float64))
        self.fields[fid] = a
        lib().orc_set_field(self.h, fid, fe_deg, dof_size, C.c_int64(n_obj), _p(a["elem_dof"]), _p(a["eqn"]),
                            _p(a["status"]), _p(a["prescribed"]), _p(a["values"]))

    def quadrature_points(self, quad_deg):
        """physical coordinates of the quadrature points [n_elems, nq, dim]"""
        w, _ = quadrature(self.shape, quad_deg)
        x = np.zeros((self.n_elems, len(w), self.dim))
        lib().orc_quadrature_points(self.h, quad_deg, _p(x))
        return x

    def set_field_constraints(self, fid, con_dof, con_ptr, master_eqn, weight):
        """linear constraints with master DoFs (base/dof/Constraint.hpp): see orc_set_field_constraints"""
        con_dof = np.ascontiguousarray(con_dof, dtype=np.int64); con_ptr = np.ascontiguousarray(con_ptr, dtype=np.int64)
        master_eqn = np.ascontiguousarray(master_eqn, dtype=np.int64); weight = np.ascontiguousarray(weight, dtype=np.float64)
        lib().orc_set_field_constraints(self.h, fid, C.c_int64(len(con_dof)), _p(con_dof), _p(con_ptr), _p(master_eqn),
                                        _p(weight))

    def set_field_values(self, fid, values):
        v = np.ascontiguousarray(values, dtype=np.float64)
        lib().orc_set_field_values(self.h, fid, _p(v))

    def measure(self, quad_deg):
        return lib().orc_measure(self.h, quad_deg)

    def quad_points_x(self, fid, quad_deg):
        w, _ = quadrature(self.shape, quad_deg)
        x = np.zeros((self.n_elems, len(w), self.dim))
        lib().orc_l2_error(self.h, fid, quad_deg, None, _p(x))
        return x

    def l2_error(self, fid, quad_deg, uref):
        uref = np.ascontiguousarray(uref, dtype=np.float64)
        return lib().orc_l2_error(self.h, fid, quad_deg, _p(uref), None)


class System:
    """base::solver::Eigen3 restated (insert/register/finish)."""

    def __init__(self, n):
        self.n = int(n)
        self.h = C.c_void_p(lib().orc_system_new(C.c_int64(n)))

    def __del__(self):
        try:
            lib().orc_system_free(self.h)
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(lib().orc_system_error(self.h).decode())

    def register_fields(self, prob, test, trial):
        lib().orc_register_fields(self.h, prob.h, test, trial)

    def stiffness(self, prob, kid, params, quad_deg, test, trial, incremental=True, nthreads=1):
        params = np.ascontiguousarray(params, dtype=np.float64)
        self._check(lib().orc_stiffness(self.h, prob.h, kid, _p(params), quad_deg, test, trial, int(incremental),
                                        nthreads))

    def stiffness_sampled(self, prob, kid, values, quad_deg, test, trial, incremental=True):
        """heat::Laplace with a conductivity function: values [n_elems, nq] = kappa at the quadrature points"""
        values = np.ascontiguousarray(values, dtype=np.float64)
        self._check(lib().orc_stiffness_sampled(self.h, prob.h, kid, _p(values), quad_deg, test, trial, int(incremental)))

    def stiffness_aux(self, prob, kid, params, quad_deg, test, trial, aux, incremental=True):
        """kernels reading a third field of the tuple (fluid::Convection: the advection velocity)"""
        params = np.ascontiguousarray(params, dtype=np.float64)
        self._check(lib().orc_stiffness_aux(self.h, prob.h, kid, _p(params), quad_deg, test, trial, aux, int(incremental)))

    def residual_aux(self, prob, kid, params, quad_deg, test, trial, aux):
        params = np.ascontiguousarray(params, dtype=np.float64)
        self._check(lib().orc_residual_aux(self.h, prob.h, kid, _p(params), quad_deg, test, trial, aux))

    def bodyforce_sampled(self, prob, values, quad_deg, test):
        """asmb::bodyForceComputation with a general f(x): values [n_elems, nq, ds] = f at the quadrature points"""
        v = np.ascontiguousarray(values, dtype=np.float64)
        if lib().orc_bodyforce_sampled(self.h, prob.h, _p(v), quad_deg, test):
            raise RuntimeError(lib().orc_system_error(self.h).decode())

    def neumann(self, prob, domain_elem, surf_x, surf_param, quad_deg, test, mode, data):
        """asmb::neumannForceComputation: mode 0 constant f = data, 1 f = data[0] * normal, 2 data [n_surf, nq, ds] sampled"""
        de = np.ascontiguousarray(domain_elem, dtype=np.int64)
        sx = np.ascontiguousarray(surf_x, dtype=np.float64)
        sp = np.ascontiguousarray(surf_param, dtype=np.float64)
        data = np.ascontiguousarray(data, dtype=np.float64)
        self._check(lib().orc_neumann(self.h, prob.h, C.c_int64(len(de)), _p(de), _p(sx), _p(sp), quad_deg, test, mode, _p(data)))

    def residual(self, prob, kid, params, quad_deg, test, trial):
        params = np.ascontiguousarray(params, dtype=np.float64)
        self._check(lib().orc_residual(self.h, prob.h, kid, _p(params), quad_deg, test, trial))

    def bodyforce(self, prob, f, quad_deg, test):
        f = np.ascontiguousarray(f, dtype=np.float64)
        self._check(lib().orc_bodyforce(self.h, prob.h, _p(f), quad_deg, test))

    def finish(self):
        lib().orc_finish(self.h)
        nnz = lib().orc_nnz(self.h)
        rowptr = np.zeros(self.n + 1, dtype=np.int64)
        col = np.zeros(nnz, dtype=np.int32)
        val = np.zeros(nnz)
        rhs = np.zeros(self.n)
        lib().orc_get_csr(self.h, _p(rowptr), _p(col), _p(val), _p(rhs))
        return rowptr, col, val, rhs

    def rhs(self):
        rhs = np.zeros(self.n)
        lib().orc_get_csr(self.h, None, None, None, _p(rhs))
        return rhs

    def rhs_norm(self):
        return lib().orc_rhs_norm(self.h)


def num_procs():
    return lib().orc_num_procs()
