"""ctypes binding of the C ABI (include/insilico_b200.h) plus a thin host-side mirror of the reference's
assembly interface.

Names follow the reference (paths relative to the reference root):
  Engine.stiffness_matrix_computation  <- base::asmb::stiffnessMatrixComputation   (base/asmb/StiffnessMatrix.hpp:49-87)
  Engine.compute_residual_forces       <- base::asmb::computeResidualForces        (base/asmb/ForceIntegrator.hpp:37-71)
  Engine.body_force_computation        <- base::asmb::bodyForceComputation         (base/asmb/BodyForce.hpp:65-84)
  Engine.new_solver / register_fields / finish_assembly / get_value / norm
                                       <- base::solver::Eigen3                     (base/solver/Eigen3.hpp:71-336)
  dof_generate / mesh_boundary / number_dofs_consecutively / boundary_dofs
                                       <- base/dof/generate.hpp, base/mesh/MeshBoundary.hpp, base/dof/numbering.hpp,
                                          base/dof/constrainBoundary.hpp

There is no CPU fallback: creating an Engine without the CUDA library or without a GPU raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libinsilico_b200.so")

POINT, LINE, TRI, QUAD, TET, HEX = range(6)
VERTEX, EDGE, FACE, CELL = range(4)
ACTIVE, CONSTRAINED, INACTIVE = range(3)
K_LAPLACE, K_HYPEL_STVENANT, K_HYPEL_NEOHOOKE, K_PRESSURE_GRADIENT, K_VELOCITY_DIVERGENCE, K_VECTOR_LAPLACE = (
    1, 2, 3, 4, 5, 6)
K_MASS = 7
K_CONVECTION = 8
SHAPE_DIM = {LINE: 1, TRI: 2, QUAD: 2, TET: 3, HEX: 3}

# every symbol include/insilico_b200.h declares (checked by tests/test_cabi.py)
EXPORTED = [
    "isl_last_error", "isl_version", "isl_engine_create", "isl_engine_destroy", "isl_engine_set_option", "isl_synchronize", "isl_flush",
    "isl_engine_stream",
    "isl_kernel_launches", "isl_measure_fp64_peak", "isl_measure_red_peak", "isl_quadrature", "isl_shape_nfun", "isl_shape_eval", "isl_support_points",
    "isl_dof_generate", "isl_dof_generate_device", "isl_ndpe", "isl_mesh_boundary", "isl_boundary_dofs", "isl_number_dofs", "isl_mesh_set",
    "isl_mesh_set_owned", "isl_mesh_update_coords", "isl_field_set", "isl_field_set_constraints", "isl_field_update",
    "isl_system_create", "isl_pattern_register",
    "isl_assemble_matrix", "isl_assemble_matrix_aux", "isl_assemble_residual_aux", "isl_assemble_matrix_sampled", "isl_assemble_residual", "isl_assemble_bodyforce", "isl_assemble_bodyforce_sampled", "isl_insert_lhs",
    "isl_insert_rhs",
    "isl_finish", "isl_get_csr", "isl_get_csr_async", "isl_copy_wait", "isl_get_device_csr", "isl_rhs_value", "isl_rhs_norm", "isl_solve_cg", "isl_pack_entries",
    "isl_unpack_add_entries", "isl_boundary_surface", "isl_surface_points", "isl_assemble_neumann", "isl_assemble_neumann_rows",
    "isl_comm_unique_id", "isl_comm_init", "isl_comm_destroy", "isl_exchange_setup", "isl_exchange", "isl_distribute", "isl_field_get_values",
]


class EngineError(RuntimeError):
    pass


_lib = None


def lib():
    """Load the CUDA library; fails loudly when it has not been built (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EngineError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'`" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.isl_last_error.restype = C.c_char_p
        L.isl_engine_stream.restype = C.c_void_p
        L.isl_engine_stream.argtypes = [C.c_void_p]
        L.isl_kernel_launches.restype = C.c_int64
        L.isl_kernel_launches.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _chk(rc):
    if rc != 0:
        raise EngineError(lib().isl_last_error().decode())


def _ptr(a):
    """numpy array, raw integer address (host or device) or None -> c_void_p"""
    if a is None:
        return None
    if isinstance(a, (int, np.integer)):
        return C.c_void_p(int(a))
    return a.ctypes.data_as(C.c_void_p)


def _i64(x):
    return C.c_int64(int(x))


# ---- host tables / DoF handling (no GPU needed) ------------------------------------------------------------
def quadrature(shape, degree):
    n = lib().isl_quadrature(shape, degree, None, None)
    if n < 0:
        raise EngineError(lib().isl_last_error().decode())
    w = np.zeros(n)
    p = np.zeros((n, SHAPE_DIM[shape]))
    lib().isl_quadrature(shape, degree, _ptr(w), _ptr(p))
    return w, p


def shape_nfun(shape, degree):
    n = lib().isl_shape_nfun(shape, degree)
    if n < 0:
        raise EngineError(lib().isl_last_error().decode())
    return n


def shape_eval(shape, degree, xi):
    n = shape_nfun(shape, degree)
    xi = np.ascontiguousarray(xi, dtype=np.float64)
    f = np.zeros(n)
    g = np.zeros((n, SHAPE_DIM[shape]))
    _chk(lib().isl_shape_eval(shape, degree, _ptr(xi), _ptr(f), _ptr(g)))
    return f, g


def support_points(shape, degree):
    p = np.zeros((shape_nfun(shape, degree), SHAPE_DIM[shape]))
    _chk(lib().isl_support_points(shape, degree, _ptr(p)))
    return p


def ndpe(shape, fe_deg):
    return lib().isl_ndpe(shape, fe_deg)


def dof_generate(shape, geom_deg, conn, fe_deg):
    """base::dof::generate: returns (elem_dof[n_elems, ndpe] int32, n_obj)."""
    conn = np.ascontiguousarray(conn, dtype=np.int32)
    ed = np.zeros((conn.shape[0], ndpe(shape, fe_deg)), dtype=np.int32)
    n = C.c_int64()
    _chk(lib().isl_dof_generate(shape, geom_deg, _i64(conn.shape[0]), _ptr(conn), fe_deg, _ptr(ed), C.byref(n)))
    return ed, n.value


def mesh_boundary(shape, geom_deg, conn):
    """base::mesh::MeshBoundary::create: (element, face number) pairs."""
    conn = np.ascontiguousarray(conn, dtype=np.int32)
    n = C.c_int64()
    _chk(lib().isl_mesh_boundary(shape, geom_deg, _i64(conn.shape[0]), _ptr(conn), None, C.byref(n)))
    pairs = np.zeros((n.value, 2), dtype=np.int64)
    _chk(lib().isl_mesh_boundary(shape, geom_deg, _i64(conn.shape[0]), _ptr(conn), _ptr(pairs), C.byref(n)))
    return pairs


def boundary_dofs(shape, geom_deg, coords, conn, fe_deg, elem_dof, pairs):
    """DoF objects visited by base::dof::constrainBoundary and the positions of their support points."""
    coords = np.ascontiguousarray(coords, dtype=np.float64)
    conn = np.ascontiguousarray(conn, dtype=np.int32)
    elem_dof = np.ascontiguousarray(elem_dof, dtype=np.int32)
    pairs = np.ascontiguousarray(pairs, dtype=np.int64)
    dim = coords.shape[1]
    n = C.c_int64()
    args = (shape, geom_deg, dim, _ptr(coords), _i64(conn.shape[0]), _ptr(conn), fe_deg, _ptr(elem_dof),
            _i64(len(pairs)), _ptr(pairs))
    _chk(lib().isl_boundary_dofs(*args, None, None, C.byref(n)))
    obj = np.zeros(n.value, dtype=np.int32)
    x = np.zeros((n.value, dim))
    _chk(lib().isl_boundary_dofs(*args, _ptr(obj), _ptr(x), C.byref(n)))
    return obj, x


def constrain_boundary(shape, geom_deg, coords, conn, fe_deg, dof_size, elem_dof, n_obj, fun, pairs=None):
    """base::dof::constrainBoundary with a Dirichlet function of the position:
    fun(x[n,dim]) -> values[n,dof_size].  Returns (status u8 [n_obj,ds], prescribed [n_obj,ds])."""
    if pairs is None:
        pairs = mesh_boundary(shape, geom_deg, conn)
    obj, x = boundary_dofs(shape, geom_deg, coords, conn, fe_deg, elem_dof, pairs)
    vals = np.asarray(fun(x), dtype=np.float64).reshape(len(obj), dof_size)
    status = np.zeros((n_obj, dof_size), dtype=np.uint8)
    prescribed = np.zeros((n_obj, dof_size))
    status[obj] = CONSTRAINED
    # later visits overwrite earlier ones (DegreeOfFreedom::constrainValue); numpy keeps the last assignment
    prescribed[obj] = vals
    return status, prescribed


NEUMANN_CONSTANT, NEUMANN_NORMAL, NEUMANN_SAMPLED = 0, 1, 2


def boundary_surface(shape, geom_deg, coords, conn, pairs):
    """base::mesh::generateBoundaryMesh for (element, face number) pairs: (surf_shape, domain_elem [n], surf_x [n, P, dim],
    surf_param [n, P, dim]) -- the nodes of every surface element and their coordinates in the domain element."""
    coords = np.ascontiguousarray(coords, dtype=np.float64)
    conn = np.ascontiguousarray(conn, dtype=np.int32)
    pairs = np.ascontiguousarray(pairs, dtype=np.int64)
    dim = coords.shape[1]
    ss, P = C.c_int(), C.c_int()
    args = (shape, geom_deg, dim, _ptr(coords), _ptr(conn), _i64(len(pairs)), _ptr(pairs))
    _chk(lib().isl_boundary_surface(*args, None, None, None, C.byref(ss), C.byref(P)))
    de = np.zeros(len(pairs), dtype=np.int32)
    sx = np.zeros((len(pairs), P.value, dim))
    sp = np.zeros((len(pairs), P.value, dim))
    _chk(lib().isl_boundary_surface(*args, _ptr(de), _ptr(sx), _ptr(sp), C.byref(ss), C.byref(P)))
    return ss.value, de, sx, sp


def surface_points(surf_shape, geom_deg, surf_x, quad_deg):
    """position, unit normal and surface metric at the points of SurfaceQuadrature<quad_deg> of every surface element:
    (x [n, nq, dim], normal [n, nq, dim], detg [n, nq]) -- the arguments of a Neumann force function f(x, normal)"""
    surf_x = np.ascontiguousarray(surf_x, dtype=np.float64)
    n, _, dim = surf_x.shape
    nq = C.c_int()
    _chk(lib().isl_surface_points(surf_shape, geom_deg, dim, _i64(n), None, quad_deg, None, None, None, C.byref(nq)))
    x, nr, dg = np.zeros((n, nq.value, dim)), np.zeros((n, nq.value, dim)), np.zeros((n, nq.value))
    _chk(lib().isl_surface_points(surf_shape, geom_deg, dim, _i64(n), _ptr(surf_x), quad_deg, _ptr(x), _ptr(nr), _ptr(dg), C.byref(nq)))
    return x, nr, dg


def number_dofs_consecutively(status, init=0):
    """base::dof::numberDoFsConsecutively: eqn[n_obj, ds] (-1 where not ACTIVE), count."""
    status = np.ascontiguousarray(status, dtype=np.uint8)
    n_obj, ds = status.shape
    eqn = np.zeros((n_obj, ds), dtype=np.int64)
    n = C.c_int64()
    _chk(lib().isl_number_dofs(_i64(n_obj), ds, _ptr(status), _i64(init), _ptr(eqn), C.byref(n)))
    return eqn, n.value


# ---- engine ---------------------------------------------------------------------------------------------
class Engine:
    """One engine per GPU: mesh + fields (asmb::FieldBinder) + one linear system (solver::Eigen3)."""

    def __init__(self, device=0):
        self.h = C.c_void_p()
        _chk(lib().isl_engine_create(int(device), C.byref(self.h)))
        self.n_eqn = 0

    def set_option(self, name, value):
        """kernel-selection knob (see isl_engine_set_option)"""
        _chk(lib().isl_engine_set_option(self.h, name.encode(), C.c_double(float(value))))

    def close(self):
        if self.h:
            lib().isl_engine_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def stream(self):
        return lib().isl_engine_stream(self.h)

    @property
    def kernel_launches(self):
        return lib().isl_kernel_launches(self.h)

    def synchronize(self):
        _chk(lib().isl_synchronize(self.h))

    def flush(self):
        """launch what the engine has deferred (the Q1 stiffness launch waits one call for a body force to fuse), without
        waiting for the device"""
        _chk(lib().isl_flush(self.h))

    def measure_fp64_peak(self):
        """measured DFMA peak of the device in TFLOP/s (denominator of the FP64-pipe roofline fraction)"""
        v = C.c_double()
        _chk(lib().isl_measure_fp64_peak(self.h, C.byref(v)))
        return v.value

    def measure_red_peak(self, pattern):
        """1e9 FP64 atomic adds per second into a 2 GiB array (0 coalesced, 1 triples at random places, 2 singles at random places)"""
        v = C.c_double()
        _chk(lib().isl_measure_red_peak(self.h, pattern, C.byref(v)))
        return v.value

    # base::Unstructured<SHAPE,GEOMDEG>

    def set_mesh(self, shape, geom_deg, coords, conn, n_nodes=None, n_elems=None, dim=None):
        if not isinstance(coords, (int, np.integer)):
            coords = np.ascontiguousarray(coords, dtype=np.float64)
            n_nodes, dim = coords.shape
        if not isinstance(conn, (int, np.integer)):
            conn = np.ascontiguousarray(conn, dtype=np.int32)
            n_elems = conn.shape[0]
        self.shape, self.geom_deg, self.dim = shape, geom_deg, dim
        self.n_nodes, self.n_elems = n_nodes, n_elems
        _chk(lib().isl_mesh_set(self.h, shape, geom_deg, dim, _i64(n_nodes), _ptr(coords), _i64(n_elems), _ptr(conn)))

    def set_owned_elements(self, n_owned):
        _chk(lib().isl_mesh_set_owned(self.h, _i64(n_owned)))

    def dof_generate(self, fe_deg):
        """base::dof::generate on the device for the engine's mesh: (elem_dof [n_elems, ndpe] int32, n_obj), same ids as
        the host function dof_generate()"""
        ed = np.zeros((self.n_elems, ndpe(self.shape, fe_deg)), dtype=np.int32)
        n = C.c_int64()
        _chk(lib().isl_dof_generate_device(self.h, fe_deg, _ptr(ed), C.byref(n)))
        return ed, n.value

    def update_coords(self, coords):
        if not isinstance(coords, (int, np.integer)):
            coords = np.ascontiguousarray(coords, dtype=np.float64)
        _chk(lib().isl_mesh_update_coords(self.h, _ptr(coords)))

    # base::Field<FEBasis,DOFSIZE>
    def set_field(self, fid, fe_deg, dof_size, n_obj, elem_dof, eqn, status, prescribed, values):
        conv = lambda a, dt: a if isinstance(a, (int, np.integer)) else np.ascontiguousarray(a, dtype=dt)
        elem_dof, eqn = conv(elem_dof, np.int32), conv(eqn, np.int64)
        status, prescribed, values = conv(status, np.uint8), conv(prescribed, np.float64), conv(values, np.float64)
        _chk(lib().isl_field_set(self.h, fid, fe_deg, dof_size, _i64(n_obj), _ptr(elem_dof), _ptr(eqn), _ptr(status),
                                 _ptr(prescribed), _ptr(values)))

    def set_field_constraints(self, fid, con_dof, con_ptr, master_eqn, weight):
        """linear constraints with master DoFs (base/dof/Constraint.hpp:57-140): component con_dof[k] = obj*ds+comp is
        u = prescribed + sum_j weight[j] u(master_eqn[j]), j in [con_ptr[k], con_ptr[k+1]).  Host arrays."""
        con_dof = np.ascontiguousarray(con_dof, dtype=np.int64); con_ptr = np.ascontiguousarray(con_ptr, dtype=np.int64)
        master_eqn = np.ascontiguousarray(master_eqn, dtype=np.int64); weight = np.ascontiguousarray(weight, dtype=np.float64)
        _chk(lib().isl_field_set_constraints(self.h, fid, _i64(len(con_dof)), _ptr(con_dof), _ptr(con_ptr),
                                             _ptr(master_eqn), _ptr(weight)))

    def update_field(self, fid, prescribed=None, values=None):
        conv = lambda a: a if a is None or isinstance(a, (int, np.integer)) else np.ascontiguousarray(a, dtype=np.float64)
        prescribed, values = conv(prescribed), conv(values)
        _chk(lib().isl_field_update(self.h, fid, _ptr(prescribed), _ptr(values)))

    # base::solver::Eigen3
    def new_solver(self, n_eqn):
        self.n_eqn = int(n_eqn)
        _chk(lib().isl_system_create(self.h, _i64(n_eqn)))

    def register_fields(self, test, trial):
        _chk(lib().isl_pattern_register(self.h, test, trial))

    @staticmethod
    def _params(kernel_id, params):
        need = 2 if kernel_id in (K_HYPEL_STVENANT, K_HYPEL_NEOHOOKE) else 1
        params = np.ascontiguousarray(params if params is not None else [0.0], dtype=np.float64).reshape(-1)
        if len(params) < need:
            raise EngineError("kernel %d needs %d parameters (lambda, mu), got %d" % (kernel_id, need, len(params)))
        return params

    def stiffness_matrix_computation(self, kernel_id, params, quad_deg, test, trial, incremental=True, aux=-1):
        """aux: index of the tuple's third field for kernels that read one (fluid::Convection: the advection velocity)"""
        params = self._params(kernel_id, params)
        _chk(lib().isl_assemble_matrix_aux(self.h, kernel_id, _ptr(params), quad_deg, test, trial, aux, int(incremental)))

    def stiffness_matrix_computation_sampled(self, kernel_id, values, quad_deg, test, trial, incremental=True):
        """asmb::stiffnessMatrixComputation with heat::Laplace + conductivity function: values [n_elems, nq] at the points"""
        values = np.ascontiguousarray(values, dtype=np.float64)
        _chk(lib().isl_assemble_matrix_sampled(self.h, kernel_id, _ptr(values), quad_deg, test, trial, int(incremental)))

    def compute_residual_forces(self, kernel_id, params, quad_deg, test, trial, factor=-1.0, aux=-1):
        params = self._params(kernel_id, params)
        _chk(lib().isl_assemble_residual_aux(self.h, kernel_id, _ptr(params), quad_deg, test, trial, aux, C.c_double(factor)))

    def body_force_computation(self, f, quad_deg, test):
        f = np.ascontiguousarray(f, dtype=np.float64)
        _chk(lib().isl_assemble_bodyforce(self.h, _ptr(f), quad_deg, test))

    def body_force_computation_sampled(self, values, quad_deg, test):
        """asmb::bodyForceComputation<FTB> with a general f(x): values [n_elems, nq, ds] = f at the quadrature points"""
        values = np.ascontiguousarray(values, dtype=np.float64)
        _chk(lib().isl_assemble_bodyforce_sampled(self.h, _ptr(values), quad_deg, test))

    def neumann_force_computation(self, domain_elem, surf_x, surf_param, quad_deg, test, mode, data):
        """asmb::neumannForceComputation<SFTB>: rhs += int f phi ds over the surface elements; mode NEUMANN_CONSTANT
        (data = f), NEUMANN_NORMAL (data = [p]: f = p * normal) or NEUMANN_SAMPLED (data [n_surf, nq, ds])"""
        de = np.ascontiguousarray(domain_elem, dtype=np.int32)
        sx = np.ascontiguousarray(surf_x, dtype=np.float64)
        sp = np.ascontiguousarray(surf_param, dtype=np.float64)
        data = np.ascontiguousarray(data, dtype=np.float64)
        _chk(lib().isl_assemble_neumann(self.h, _i64(len(de)), _ptr(de), _ptr(sx), _ptr(sp), quad_deg, test, mode, _ptr(data)))

    def neumann_force_computation_rows(self, shape, geom_deg, surf_x, surf_param, quad_deg, fe_deg, ds, rows, mode, data):
        """the same with the equation numbers per surface element (rows [n_surf, ndpe * ds], < 0 skipped)"""
        sx = np.ascontiguousarray(surf_x, dtype=np.float64)
        sp = np.ascontiguousarray(surf_param, dtype=np.float64)
        rows = np.ascontiguousarray(rows, dtype=np.int32)
        data = np.ascontiguousarray(data, dtype=np.float64)
        _chk(lib().isl_assemble_neumann_rows(self.h, shape, geom_deg, _i64(len(sx)), _ptr(sx), _ptr(sp), quad_deg, fe_deg, ds, _ptr(rows), mode, _ptr(data)))

    def insert_to_lhs(self, mat, rows, cols):
        mat = np.ascontiguousarray(mat, dtype=np.float64)
        rows = np.ascontiguousarray(rows, dtype=np.int64)
        cols = np.ascontiguousarray(cols, dtype=np.int64)
        _chk(lib().isl_insert_lhs(self.h, _ptr(mat), _ptr(rows), len(rows), _ptr(cols), len(cols)))

    def insert_to_rhs(self, vec, rows):
        vec = np.ascontiguousarray(vec, dtype=np.float64)
        rows = np.ascontiguousarray(rows, dtype=np.int64)
        _chk(lib().isl_insert_rhs(self.h, _ptr(vec), _ptr(rows), len(rows)))

    def finish_assembly(self):
        n, nnz = C.c_int64(), C.c_int64()
        _chk(lib().isl_finish(self.h, C.byref(n), C.byref(nnz)))
        self.nnz = nnz.value
        return n.value, nnz.value

    def get_csr(self, rowptr=None, col=None, val=None, rhs=None, pattern=True):
        """Copies the system out.  With no arguments allocates numpy arrays and returns (rowptr, col, val, rhs)."""
        if rowptr is None and col is None and val is None and rhs is None:
            n, nnz = self.finish_assembly()
            rowptr = np.zeros(n + 1, dtype=np.int64) if pattern else None
            col = np.zeros(nnz, dtype=np.int32) if pattern else None
            val = np.zeros(nnz)
            rhs = np.zeros(n)
        _chk(lib().isl_get_csr(self.h, _ptr(rowptr), _ptr(col), _ptr(val), _ptr(rhs)))
        return rowptr, col, val, rhs

    def get_csr_async(self, val, rhs):
        """values / rhs of the finished system -> pinned host arrays on the copy stream; the next call must be new_solver"""
        _chk(lib().isl_get_csr_async(self.h, _ptr(val), _ptr(rhs)))

    def copy_wait(self):
        _chk(lib().isl_copy_wait(self.h))

    def cg_solve(self, tol=0.0, max_iter=0):
        """solver.cgSolve() on the device (base/solver/Eigen3.hpp:263-275): rhs <- A^-1 rhs; returns (iterations, error)"""
        it, err = C.c_int64(0), C.c_double(0.0)
        _chk(lib().isl_solve_cg(self.h, C.c_double(tol), _i64(max_iter), C.byref(it), C.byref(err)))
        return it.value, err.value

    def distribute(self, fid, add=False):
        """base::dof::setDoFsFromSolver / addToDoFsFromSolver on the device (the rhs holds the solution after cg_solve)"""
        _chk(lib().isl_distribute(self.h, int(fid), int(bool(add))))

    def get_field_values(self, fid, n_obj, ds):
        out = np.zeros((n_obj, ds))
        _chk(lib().isl_field_get_values(self.h, int(fid), _ptr(out)))
        return out

    def device_csr(self):
        p = [C.c_void_p() for _ in range(4)]
        _chk(lib().isl_get_device_csr(self.h, *[C.byref(x) for x in p]))
        return tuple(x.value for x in p)

    def get_value(self, index):
        v = C.c_double()
        _chk(lib().isl_rhs_value(self.h, _i64(index), C.byref(v)))
        return v.value

    def norm(self):
        v = C.c_double()
        _chk(lib().isl_rhs_norm(self.h, C.byref(v)))
        return v.value

    # multi-GPU exchange inside the engine (NCCL, isl_comm.cuh)
    @staticmethod
    def comm_unique_id():
        buf = C.create_string_buffer(128)
        _chk(lib().isl_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, unique_id, rank, world):
        _chk(lib().isl_comm_init(self.h, C.c_char_p(unique_id), int(rank), int(world)))

    def comm_destroy(self):
        _chk(lib().isl_comm_destroy(self.h))

    def exchange_setup(self, l2g, own_lo, own_hi, segments):
        """segments: [(owner rank, row_lo, row_hi), ...] ghost rows in local numbering"""
        l2g = np.ascontiguousarray(l2g, dtype=np.int64)
        own = np.array([s[0] for s in segments], dtype=np.int32)
        lo = np.array([s[1] for s in segments], dtype=np.int64)
        hi = np.array([s[2] for s in segments], dtype=np.int64)
        _chk(lib().isl_exchange_setup(self.h, _i64(len(l2g)), _ptr(l2g), _i64(own_lo), _i64(own_hi), len(segments),
                                      _ptr(own), _ptr(lo), _ptr(hi)))

    def exchange(self):
        _chk(lib().isl_exchange(self.h))

    def pack_entries(self, which, idx_dev, n, out_dev):
        _chk(lib().isl_pack_entries(self.h, which, _ptr(idx_dev), _i64(n), _ptr(out_dev)))

    def unpack_add_entries(self, which, idx_dev, n, in_dev):
        _chk(lib().isl_unpack_add_entries(self.h, which, _ptr(idx_dev), _i64(n), _ptr(in_dev)))
