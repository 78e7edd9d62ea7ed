"""Application flows run on BOTH the CPU oracle and the CUDA engine (through the C ABI) for parity tests.

Each case follows one of the reference's applications:
  laplace_*     reference/04-heat/dirichlet.cpp:79-167            (Dirichlet from the Laplace fundamental solution)
  stvenant_*    reference/06-elastic/linearElastic.cpp:83-196     (HyperElastic<StVenant>, Lame from E, nu)
  neohooke_*    reference/06-elastic/compressible.cpp:245-320     (tangent + residual per Newton step)
  stokes_*      reference/07-drivenCavity/drivenCavity.cpp:176-275 (Taylor-Hood blocks UU, UP, PU)
"""
import numpy as np

from insilico_b200 import engine as E
from insilico_b200 import meshgen
from oracle import oracle as orc
from tests import helpers as H


def lame(Emod, nu):
    return Emod * nu / (1. + nu) / (1. - 2. * nu), Emod / 2. / (1. + nu)


def make_mesh(shape, n, perturb, permute=False):
    if shape == E.HEX:
        coords, conn, _ = meshgen.unit_cube_hex(n, n, n)
    elif shape == E.TET:
        coords, conn = meshgen.unit_cube_tet(n, n, n)
    elif shape == E.QUAD:
        coords, conn = meshgen.unit_square_quad(n, n)
    elif shape == E.TRI:
        coords, conn = meshgen.unit_square_tri(n, n)
    else:
        raise ValueError(shape)
    if perturb:
        coords = meshgen.perturb_interior(coords, 1.0 / n, max_dist=0.15)
    if permute:
        conn = meshgen.permute_elements(conn)
    return coords, conn


class Case:
    """mesh + fields + list of assembly operations"""

    def __init__(self, shape, geom_deg, coords, conn):
        self.shape, self.geom_deg = shape, geom_deg
        self.coords = np.ascontiguousarray(coords, dtype=np.float64)
        self.conn = np.ascontiguousarray(conn, dtype=np.int32)
        self.dim = coords.shape[1]
        self.fields = []
        self.ops = []
        self.n_eqn = 0

    def add_field(self, fe_deg, ds, dirichlet=None, values=None, pin_first=False, linear=None, where=None):
        """dirichlet: fun(x)->[n,ds] applied on the whole boundary (dof::constrainBoundary);
        values: fun(x_support)->[n_obj,ds] current field state; pin_first: constrain component 0 of DoF 0 to 0
        (drivenCavity.cpp:199-202)."""
        ed, nobj = E.dof_generate(self.shape, self.geom_deg, self.conn, fe_deg)
        status = np.zeros((nobj, ds), dtype=np.uint8)
        presc = np.zeros((nobj, ds))
        if dirichlet is not None:
            status, presc = E.constrain_boundary(self.shape, self.geom_deg, self.coords, self.conn, fe_deg, ds, ed,
                                                 nobj, dirichlet)
            if where is not None:   # Dirichlet part of the boundary only: where(x of the DoF) -> bool
                keep = np.asarray(where(self.dof_positions(fe_deg, ed, nobj)), dtype=bool)
                status[~keep] = E.ACTIVE
                presc[~keep] = 0.0
        if pin_first:
            status[0, 0] = E.CONSTRAINED
            presc[0, 0] = 0.0
        # linear constraints with master DoFs (base/dof/Constraint.hpp): linear(status) -> [(obj, comp, rhs,
        # [(master obj, master comp, weight), ...]), ...]; applied after the Dirichlet boundary, before the numbering
        lin = linear(status) if linear is not None else []
        for obj, comp, rhs, _ in lin:
            status[obj, comp] = E.CONSTRAINED
            presc[obj, comp] = rhs
        eqn, n = E.number_dofs_consecutively(status, init=self.n_eqn)
        self.n_eqn += n
        vals = np.zeros((nobj, ds))
        if values is not None:
            vals = np.asarray(values(self.dof_positions(fe_deg, ed, nobj)), dtype=np.float64).reshape(nobj, ds)
        self.fields.append(dict(fe_deg=fe_deg, ds=ds, n_obj=nobj, elem_dof=ed, status=status, presc=presc, eqn=eqn,
                                values=vals, boundary=(2 if where is not None else int(dirichlet is not None)), pin=0 if pin_first else -1,
                                linear=lin))
        return len(self.fields) - 1

    def sampled_factor(self, op):
        """kappa(x) of a ("matrixfun", kid, fun, quad_deg, t, c, incremental) operation at the quadrature points [n_elems, nq]"""
        w, xi = E.quadrature(self.shape, op[3])
        Ng = np.array([E.shape_eval(self.shape, self.geom_deg, p)[0] for p in xi])
        x = np.einsum("qa,ead->eqd", Ng, self.coords[self.conn])
        return np.ascontiguousarray(np.asarray(op[2](x.reshape(-1, self.dim)), dtype=np.float64).reshape(len(self.conn), len(w)))

    def sampled_force(self, op):
        """f(x) of a ("bodyfun", fun, quad_deg, field) operation at the quadrature points of every element
        [n_elems, nq, ds]; x(xi_q) = sum_a N_a(xi_q) x_a like base::Geometry (base/geometry.hpp:105-135)"""
        w, xi = E.quadrature(self.shape, op[2])
        Ng = np.array([E.shape_eval(self.shape, self.geom_deg, p)[0] for p in xi])     # [nq, npe]
        x = np.einsum("qa,ead->eqd", Ng, self.coords[self.conn])
        ds = self.fields[op[3]]["ds"]
        return np.ascontiguousarray(np.asarray(op[1](x.reshape(-1, self.dim)), dtype=np.float64).reshape(len(self.conn), len(w), ds))

    def surface_elements(self, op, lib):
        """surface elements of a ("neumann", name, quad_deg, field, filter, params) operation: all boundary faces
        (filter 0) or those in the plane x0 = 1 (filter 1, all nodes of the face), from `lib` = the engine's host side
        (E) or the oracle's (a Problem)"""
        if lib is E:
            pairs = E.mesh_boundary(self.shape, self.geom_deg, self.conn)
            _, de, sx, sp = E.boundary_surface(self.shape, self.geom_deg, self.coords, self.conn, pairs)
        else:
            de, sx, sp = lib.boundary_surface(lib.mesh_boundary())
        if op[4] == 1:
            keep = np.all(sx[:, :, 0] > 1.0 - 1e-9, axis=1)
            de, sx, sp = de[keep], sx[keep], sp[keep]
        return de, np.ascontiguousarray(sx), np.ascontiguousarray(sp)

    def surface_force(self, op, sx, points):
        """(mode, data) of a neumann operation; points(surf_shape, geom_deg, sx, quad_deg) -> (x, normal, detg)"""
        name, ds = op[1], self.fields[op[3]]["ds"]
        if name == "constant":
            return E.NEUMANN_CONSTANT, np.asarray(op[5], dtype=np.float64)
        if name == "pressure" and ds == self.dim:
            return E.NEUMANN_NORMAL, np.asarray(op[5][:1], dtype=np.float64)
        surf_shape = {E.HEX: E.QUAD, E.TET: E.TRI}.get(self.shape, E.LINE)
        x, nr, _ = points(surf_shape, self.geom_deg, sx, op[2])
        dim = self.dim
        if name == "pressure":
            f = np.stack([op[5][0] * nr[..., d % dim] for d in range(ds)], axis=-1)
        else:
            s = 1.0 + x[..., 0] * x[..., 1] + 0.5 * np.sin(2.0 * x[..., dim - 1])
            f = np.stack([s * nr[..., d % dim] + 0.25 * x[..., (d + 1) % dim] for d in range(ds)], axis=-1)
        return E.NEUMANN_SAMPLED, np.ascontiguousarray(f)

    @staticmethod
    def constraint_arrays(f):
        """flat form of f["linear"] for set_field_constraints (masters as equation numbers)"""
        con_dof, con_ptr, meq, w = [], [0], [], []
        for obj, comp, _, masters in f["linear"]:
            con_dof.append(obj * f["ds"] + comp)
            for mo, mc, wt in masters:
                assert f["status"][mo, mc] == E.ACTIVE, "master DoFs must be ACTIVE"
                meq.append(int(f["eqn"][mo, mc])); w.append(wt)
            con_ptr.append(len(meq))
        return (np.array(con_dof, dtype=np.int64), np.array(con_ptr, dtype=np.int64), np.array(meq, dtype=np.int64),
                np.array(w, dtype=np.float64))

    def dof_positions(self, fe_deg, ed, nobj):
        """physical position of every DoF object (support point mapped through the geometry)."""
        sp = E.support_points(self.shape, fe_deg)
        Ng = np.array([E.shape_eval(self.shape, self.geom_deg, s)[0] for s in sp])  # [ndpe, npe]
        xe = self.coords[self.conn]                                                  # [ne, npe, dim]
        xd = np.einsum("la,ead->eld", Ng, xe)                                        # [ne, ndpe, dim]
        pos = np.zeros((nobj, self.dim))
        pos[ed.reshape(-1)] = xd.reshape(-1, self.dim)
        return pos

    # ---- run on the oracle -------------------------------------------------------------------------
    def run_oracle(self, register=False, nthreads=1):
        prob = orc.Problem(self.shape, self.geom_deg, self.coords, self.conn.astype(np.int64))
        for i, f in enumerate(self.fields):
            prob.set_field(i, f["fe_deg"], f["ds"], f["n_obj"], f["elem_dof"].astype(np.int64), f["eqn"], f["status"],
                           f["presc"], f["values"])
            if f["linear"]:
                prob.set_field_constraints(i, *self.constraint_arrays(f))
        s = orc.System(self.n_eqn)
        if register:
            for op in self.ops:
                if op[0] in ("matrix", "matrixfun"):
                    s.register_fields(prob, op[4], op[5])
        for op in self.ops:
            if op[0] == "matrix" and len(op) > 7:     # a third field of the tuple (fluid::Convection: advection velocity)
                s.stiffness_aux(prob, op[1], op[2], op[3], op[4], op[5], op[7], incremental=op[6])
            elif op[0] == "residual" and len(op) > 6:
                s.residual_aux(prob, op[1], op[2], op[3], op[4], op[5], op[6])
            elif op[0] == "matrix":
                s.stiffness(prob, op[1], op[2], op[3], op[4], op[5], incremental=op[6], nthreads=nthreads)
            elif op[0] == "matrixfun":
                s.stiffness_sampled(prob, op[1], self.sampled_factor(op), op[3], op[4], op[5], incremental=op[6])
            elif op[0] == "residual":
                s.residual(prob, op[1], op[2], op[3], op[4], op[5])
            elif op[0] == "body":
                s.bodyforce(prob, op[1], op[2], op[3])
            elif op[0] == "bodyfun":
                s.bodyforce_sampled(prob, self.sampled_force(op), op[2], op[3])
            elif op[0] == "neumann":
                de, sx, sp = self.surface_elements(op, prob)
                mode, data = self.surface_force(op, sx, orc.surface_points)
                s.neumann(prob, de, sx, sp, op[2], op[3], mode, data)
        return s.finish()

    # ---- run on the CUDA engine ----------------------------------------------------------------------
    def run_engine(self, eng=None, register=False):
        own = eng is None
        if own:
            eng = E.Engine(0)
        eng.set_mesh(self.shape, self.geom_deg, self.coords, self.conn)
        for i, f in enumerate(self.fields):
            eng.set_field(i, f["fe_deg"], f["ds"], f["n_obj"], f["elem_dof"], f["eqn"], f["status"], f["presc"],
                          f["values"])
            if f["linear"]:
                eng.set_field_constraints(i, *self.constraint_arrays(f))
        eng.new_solver(self.n_eqn)
        if register:
            for op in self.ops:
                if op[0] in ("matrix", "matrixfun"):
                    eng.register_fields(op[4], op[5])
        for op in self.ops:
            if op[0] == "matrix" and len(op) > 7:
                eng.stiffness_matrix_computation(op[1], op[2], op[3], op[4], op[5], incremental=op[6], aux=op[7])
            elif op[0] == "residual" and len(op) > 6:
                eng.compute_residual_forces(op[1], op[2], op[3], op[4], op[5], aux=op[6])
            elif op[0] == "matrix":
                eng.stiffness_matrix_computation(op[1], op[2], op[3], op[4], op[5], incremental=op[6])
            elif op[0] == "matrixfun":
                eng.stiffness_matrix_computation_sampled(op[1], self.sampled_factor(op), op[3], op[4], op[5], incremental=op[6])
            elif op[0] == "residual":
                eng.compute_residual_forces(op[1], op[2], op[3], op[4], op[5])
            elif op[0] == "body":
                eng.body_force_computation(op[1], op[2], op[3])
            elif op[0] == "bodyfun":
                eng.body_force_computation_sampled(self.sampled_force(op), op[2], op[3])
            elif op[0] == "neumann":
                de, sx, sp = self.surface_elements(op, E)
                mode, data = self.surface_force(op, sx, E.surface_points)
                eng.neumann_force_computation(de, sx, sp, op[2], op[3], mode, data)
        out = eng.get_csr()
        if own:
            eng.close()
        return out


def linear_constraints(status, ds, count=4):
    """deterministic set of linear constraints on ACTIVE DoF components: slave k gets 2 or 3 ACTIVE masters (never a
    slave, possibly another component of the slave's own DoF object or a DoF of the same element) and an rhs term"""
    free = [(int(o), int(cmp)) for o, cmp in np.argwhere(status == E.ACTIVE)]
    rng = np.random.default_rng(2024 + ds)
    picks = rng.permutation(len(free))
    nslave = min(count, len(free) // 5)
    slaves = [free[k] for k in picks[:nslave]]
    pool = [free[k] for k in picks[nslave:]]
    out = []
    for k, (o, cmp) in enumerate(slaves):
        nm = 2 + (k % 2)
        masters = [pool[(3 * k + 7 * j) % len(pool)] for j in range(nm)]
        if ds > 1 and k == 0 and (o, (cmp + 1) % ds) in pool:
            masters[0] = (o, (cmp + 1) % ds)
        w = [0.5, 0.25, -0.75][:nm] if k % 2 else [0.6, 0.4]
        out.append((o, cmp, 0.05 * (k + 1), [(m[0], m[1], w[j]) for j, m in enumerate(masters)]))
    return out


def smooth_u(dim, amp=0.02):
    return lambda x: amp * np.sin(np.pi * x[:, :dim]) * (1.0 + x[:, ::-1][:, :dim])


def kappa_fun(x):
    """the conductivity function of the *_kappafun cases (also in oracle/ref_driver.cpp, NamedConductivity "kappa1")"""
    return 1.0 + x[:, 0] ** 2 + 0.5 * np.sin(3.0 * x[:, 1]) * x[:, x.shape[1] - 1]


def build_case(name, n=4, perturb=True, permute=False):
    src3 = np.full(3, -0.5)
    if name == "laplace_q1_hex":
        c = Case(E.HEX, 1, *make_mesh(E.HEX, n, perturb, permute))
        c.add_field(1, 1, dirichlet=lambda x: H.fund_sol_laplace(x, src3))
        c.ops = [("matrix", E.K_LAPLACE, [1.0], 3, 0, 0, True), ("body", [1.0], 3, 0)]
    elif name == "laplace_q1_hex_values":  # non-zero current values: incremental lift uses prescribed - current
        c = Case(E.HEX, 1, *make_mesh(E.HEX, n, perturb, permute))
        c.add_field(1, 1, dirichlet=lambda x: H.fund_sol_laplace(x, src3), values=lambda x: 0.3 * x[:, :1] + 0.1)
        c.ops = [("matrix", E.K_LAPLACE, [2.5], 3, 0, 0, True), ("residual", E.K_LAPLACE, [2.5], 3, 0, 0)]
    elif name == "laplace_q2_hex":
        c = Case(E.HEX, 1, *make_mesh(E.HEX, n, perturb, permute))
        c.add_field(2, 1, dirichlet=lambda x: H.fund_sol_laplace(x, src3))
        c.ops = [("matrix", E.K_LAPLACE, [1.0], 4, 0, 0, False), ("body", [1.0], 4, 0)]
    elif name == "laplace_p1_tet":
        c = Case(E.TET, 1, *make_mesh(E.TET, n, perturb, permute))
        c.add_field(1, 1, dirichlet=lambda x: H.fund_sol_laplace(x, src3))
        c.ops = [("matrix", E.K_LAPLACE, [1.0], 3, 0, 0, True), ("body", [1.0], 2, 0)]
    elif name == "laplace_q1_quad":
        c = Case(E.QUAD, 1, *make_mesh(E.QUAD, n, perturb, permute))
        c.add_field(1, 1, dirichlet=lambda x: H.fund_sol_laplace(x, np.full(2, -0.5)))
        c.ops = [("matrix", E.K_LAPLACE, [1.0], 3, 0, 0, True), ("body", [1.0], 3, 0)]
    elif name == "laplace_p2_tri":
        c = Case(E.TRI, 1, *make_mesh(E.TRI, n, perturb, permute))
        c.add_field(2, 1, dirichlet=lambda x: H.fund_sol_laplace(x, np.full(2, -0.5)))
        c.ops = [("matrix", E.K_LAPLACE, [1.0], 4, 0, 0, True), ("body", [1.0], 4, 0)]
    elif name == "vector_laplace_q1_hex":
        c = Case(E.HEX, 1, *make_mesh(E.HEX, n, perturb, permute))
        c.add_field(1, 3, dirichlet=lambda x: np.stack([x[:, 0], 0 * x[:, 1], -x[:, 2]], axis=1), values=smooth_u(3))
        c.ops = [("matrix", E.K_VECTOR_LAPLACE, [0.7], 3, 0, 0, True), ("residual", E.K_VECTOR_LAPLACE, [0.7], 3, 0, 0),
                 ("body", [0.0, 0.0, -1.0], 3, 0)]
    elif name in ("stvenant_q1_hex", "stvenant_q2_hex", "neohooke_q1_hex"):
        lam, mu = lame(1000.0, 0.25)
        deg = 2 if "q2" in name else 1
        kid = E.K_HYPEL_NEOHOOKE if "neohooke" in name else E.K_HYPEL_STVENANT
        y0 = np.full(3, -0.1); d = np.array([0., 1., 0.])
        c = Case(E.HEX, 1, *make_mesh(E.HEX, n, perturb, permute))
        c.add_field(deg, 3, dirichlet=lambda x: H.fund_sol_elastostatic(x, y0, d, lam, mu), values=smooth_u(3))
        q = 4 if deg == 2 else 3
        c.ops = [("matrix", kid, [lam, mu], q, 0, 0, True), ("residual", kid, [lam, mu], q, 0, 0)]
    elif name == "stvenant_q1_quad":
        lam, mu = lame(1000.0, 0.25)
        y0 = np.full(2, -0.1); d = np.array([0., 1.])
        c = Case(E.QUAD, 1, *make_mesh(E.QUAD, n, perturb, permute))
        c.add_field(1, 2, dirichlet=lambda x: H.fund_sol_elastostatic(x, y0, d, lam, mu), values=smooth_u(2))
        c.ops = [("matrix", E.K_HYPEL_STVENANT, [lam, mu], 3, 0, 0, True),
                 ("residual", E.K_HYPEL_STVENANT, [lam, mu], 3, 0, 0)]
    elif name == "stvenant_q2_quad":   # the element of reference/06-elastic/compressible.cpp (Q2 quads on Q1 geometry)
        lam, mu = lame(1000.0, 0.25)
        y0 = np.full(2, -0.1); d = np.array([0., 1.])
        c = Case(E.QUAD, 1, *make_mesh(E.QUAD, n, perturb, permute))
        c.add_field(2, 2, dirichlet=lambda x: H.fund_sol_elastostatic(x, y0, d, lam, mu), values=smooth_u(2))
        c.ops = [("matrix", E.K_HYPEL_STVENANT, [lam, mu], 3, 0, 0, True),
                 ("residual", E.K_HYPEL_STVENANT, [lam, mu], 3, 0, 0)]
    elif name == "laplace_p1_tri":
        c = Case(E.TRI, 1, *make_mesh(E.TRI, n, perturb, permute))
        c.add_field(1, 1, dirichlet=lambda x: H.fund_sol_laplace(x, np.full(2, -0.5)), values=lambda x: 0.2 * x[:, 1:2])
        c.ops = [("matrix", E.K_LAPLACE, [0.5], 2, 0, 0, False), ("residual", E.K_LAPLACE, [0.5], 2, 0, 0), ("body", [2.0], 2, 0)]
    elif name == "stvenant_p2_tet":
        lam, mu = lame(1000.0, 0.3)
        c = Case(E.TET, 1, *make_mesh(E.TET, n, perturb, permute))
        c.add_field(2, 3, dirichlet=lambda x: 0.0 * x, values=lambda x: 0.03 * np.sin(np.pi * x))
        c.ops = [("matrix", E.K_HYPEL_STVENANT, [lam, mu], 4, 0, 0, True), ("residual", E.K_HYPEL_STVENANT, [lam, mu], 4, 0, 0),
                 ("body", [0.0, 0.0, -9.81], 4, 0)]
    elif name == "neohooke_p2_tet":
        lam, mu = lame(1000.0, 0.3)
        c = Case(E.TET, 1, *make_mesh(E.TET, n, perturb, permute))
        c.add_field(2, 3, dirichlet=lambda x: 0.0 * x, values=lambda x: 0.05 * np.sin(np.pi * x))
        c.ops = [("matrix", E.K_HYPEL_NEOHOOKE, [lam, mu], 4, 0, 0, True),
                 ("residual", E.K_HYPEL_NEOHOOKE, [lam, mu], 4, 0, 0)]
    elif name in ("stokes_p2p1_tet", "stokes_q2q1_hex", "stokes_q2q1_quad"):
        shape = E.TET if "tet" in name else (E.HEX if "hex" in name else E.QUAD)
        dim = E.SHAPE_DIM[shape]
        c = Case(shape, 1, *make_mesh(shape, n, perturb, permute))
        lid = lambda x: np.stack([(x[:, dim - 1] > 1 - 1e-9) * 1.0] + [0 * x[:, 0]] * (dim - 1), axis=1)
        u = c.add_field(2, dim, dirichlet=lid, values=smooth_u(dim))
        p = c.add_field(1, 1, pin_first=True, values=lambda x: x[:, :1] - 0.5)
        c.ops = [("matrix", E.K_VECTOR_LAPLACE, [1.0], 4, u, u, True), ("matrix", E.K_PRESSURE_GRADIENT, None, 4, u, p, True),
                 ("matrix", E.K_VELOCITY_DIVERGENCE, [0.0], 4, p, u, True),
                 ("residual", E.K_VECTOR_LAPLACE, [1.0], 4, u, u), ("residual", E.K_PRESSURE_GRADIENT, None, 4, u, p),
                 ("residual", E.K_VELOCITY_DIVERGENCE, [0.0], 4, p, u)]
    elif name in ("laplace_q1_hex_kappafun", "laplace_p2_tet_kappafun"):
        # heat::Laplace with a conductivity FUNCTION (heat/Laplace.hpp:85-126): kappa(x) evaluated per quadrature point
        shape, deg, q = (E.HEX, 1, 3) if "hex" in name else (E.TET, 2, 4)
        c = Case(shape, 1, *make_mesh(shape, n, perturb, permute))
        c.add_field(deg, 1, dirichlet=lambda x: H.fund_sol_laplace(x, src3), values=lambda x: 0.3 * x[:, :1] + 0.1)
        c.ops = [("matrixfun", E.K_LAPLACE, kappa_fun, q, 0, 0, True), ("body", [1.0], q, 0)]
    elif name == "mass_q1_hex":       # M/dt + K of an implicit heat step: base::kernel::Mass + heat::Laplace into one matrix
        c = Case(E.HEX, 1, *make_mesh(E.HEX, n, perturb, permute))
        c.add_field(1, 1, dirichlet=lambda x: H.fund_sol_laplace(x, src3), values=lambda x: 0.3 * x[:, :1] + 0.1)
        c.ops = [("matrix", E.K_MASS, [12.5], 3, 0, 0, True), ("matrix", E.K_LAPLACE, [1.0], 3, 0, 0, True), ("body", [1.0], 3, 0)]
    elif name == "mass_p2_tet_vector":  # consistent mass matrix of a P2 displacement field + St. Venant tangent
        lam, mu = lame(1000.0, 0.3)
        c = Case(E.TET, 1, *make_mesh(E.TET, n, perturb, permute))
        c.add_field(2, 3, dirichlet=lambda x: 0.0 * x, values=lambda x: 0.03 * np.sin(np.pi * x))
        c.ops = [("matrix", E.K_MASS, [7.8], 4, 0, 0, True), ("matrix", E.K_HYPEL_STVENANT, [lam, mu], 4, 0, 0, True)]
    elif name in ("convection_q1_hex", "convection_q2_quad", "convection_q1_hex_at_rest"):
        # one Picard step of the momentum equation: viscous term + fluid::Convection (fluid/Convection.hpp) on the tuple
        # (test, trial, advection velocity) = (u, u, u) with a non-zero velocity state; residual of the convective term
        shape, deg = (E.HEX, 1) if "hex" in name else (E.QUAD, 2)
        dim = E.SHAPE_DIM[shape]
        c = Case(shape, 1, *make_mesh(shape, n, perturb, permute))
        lid = lambda x: np.stack([(x[:, dim - 1] > 1 - 1e-9) * 1.0] + [0 * x[:, 0]] * (dim - 1), axis=1)
        q = 3 if "hex" in name else 4
        # at rest: the first Picard step from u = 0 (the convective terms vanish; the binding cannot probe the density there)
        c.add_field(deg, dim, dirichlet=lid, values=None if name.endswith("at_rest") else smooth_u(dim, amp=0.3))
        c.ops = [("matrix", E.K_VECTOR_LAPLACE, [0.1], q, 0, 0, True), ("matrix", E.K_CONVECTION, [1.2], q, 0, 0, True, 0),
                 ("residual", E.K_CONVECTION, [1.2], q, 0, 0, 0)]
    elif name == "neumann_q1_hex":     # 05-mixedPoisson: Dirichlet on x0 = 0, a surface force f(x, n) on the whole boundary
        c = Case(E.HEX, 1, *make_mesh(E.HEX, n, perturb, permute))
        c.add_field(1, 1, dirichlet=lambda x: H.fund_sol_laplace(x, src3), where=lambda x: x[:, 0] < 1e-9)
        c.ops = [("matrix", E.K_LAPLACE, [1.0], 3, 0, 0, True), ("body", [1.0], 3, 0), ("neumann", "fun", 3, 0, 0, [])]
    elif name == "neumann_p2_tet_solid":   # clamped at x0 = 0, a pressure on the opposite face and a constant traction everywhere
        lam, mu = lame(1000.0, 0.3)
        c = Case(E.TET, 1, *make_mesh(E.TET, n, perturb, permute))
        c.add_field(2, 3, dirichlet=lambda x: 0.0 * x, where=lambda x: x[:, 0] < 1e-9, values=lambda x: 0.02 * np.sin(np.pi * x))
        c.ops = [("matrix", E.K_HYPEL_STVENANT, [lam, mu], 4, 0, 0, True), ("neumann", "pressure", 4, 0, 1, [-0.7]),
                 ("neumann", "constant", 4, 0, 0, [0.1, -0.2, 0.3])]
    elif name == "neumann_p2_tri":         # line elements on the boundary of a triangle mesh, quadratic test functions
        c = Case(E.TRI, 1, *make_mesh(E.TRI, n, perturb, permute))
        c.add_field(2, 1, pin_first=True)
        c.ops = [("matrix", E.K_LAPLACE, [1.0], 4, 0, 0, True), ("neumann", "fun", 4, 0, 0, []), ("neumann", "pressure", 4, 0, 1, [2.0])]
    elif name == "neumann_q1_quad_solid":  # pressure on the whole boundary of a 2-D solid, two linear constraints on board
        lam, mu = lame(1000.0, 0.3)
        c = Case(E.QUAD, 1, *make_mesh(E.QUAD, n, perturb, permute))
        c.add_field(1, 2, dirichlet=lambda x: 0.0 * x, where=lambda x: x[:, 1] < 1e-9, linear=lambda st: linear_constraints(st, 2, 2))
        c.ops = [("matrix", E.K_HYPEL_STVENANT, [lam, mu], 3, 0, 0, True), ("neumann", "pressure", 3, 0, 0, [1.5]),
                 ("neumann", "fun", 3, 0, 1, [])]
    elif name in ("laplace_q1_hex_bodyfun", "laplace_p2_tri_bodyfun", "vector_laplace_q1_hex_bodyfun"):
        # general (non-constant) body force f(x): BodyForce.hpp:172-205 evaluates the caller's function per point
        if "tri" in name:
            c = Case(E.TRI, 1, *make_mesh(E.TRI, n, perturb, permute))
            c.add_field(2, 1, dirichlet=lambda x: H.fund_sol_laplace(x, np.full(2, -0.5)))
            c.ops = [("matrix", E.K_LAPLACE, [1.0], 4, 0, 0, True),
                     ("bodyfun", lambda x: (np.sin(3.0 * x[:, 0]) * (1.0 + x[:, 1] ** 2))[:, None], 4, 0)]
        elif "vector" in name:
            c = Case(E.HEX, 1, *make_mesh(E.HEX, n, perturb, permute))
            c.add_field(1, 3, dirichlet=lambda x: np.stack([x[:, 0], 0 * x[:, 1], -x[:, 2]], axis=1))
            c.ops = [("matrix", E.K_VECTOR_LAPLACE, [0.7], 3, 0, 0, True),
                     ("bodyfun", lambda x: np.stack([x[:, 1] * x[:, 2], np.cos(x[:, 0]), 1.0 + x[:, 0] * x[:, 1] * x[:, 2]], axis=1), 3, 0)]
        else:
            c = Case(E.HEX, 1, *make_mesh(E.HEX, n, perturb, permute))
            c.add_field(1, 1, dirichlet=lambda x: H.fund_sol_laplace(x, src3))
            c.ops = [("matrix", E.K_LAPLACE, [1.0], 3, 0, 0, True),
                     ("bodyfun", lambda x: (np.exp(x[:, 0]) * np.sin(2.0 * x[:, 1]) + x[:, 2] ** 2)[:, None], 3, 0)]
    elif name in ("laplace_q1_hex_linear", "laplace_q2_hex_linear", "laplace_p1_tet_linear"):
        # general linear constraints: some interior DoFs are slaves of two / three ACTIVE masters with an rhs term
        shape = E.TET if "tet" in name else E.HEX
        deg = 2 if "q2" in name else 1
        c = Case(shape, 1, *make_mesh(shape, n, perturb, permute))
        c.add_field(deg, 1, dirichlet=lambda x: H.fund_sol_laplace(x, src3), values=lambda x: 0.3 * x[:, :1] + 0.1,
                    linear=lambda st: linear_constraints(st, 1))
        q = 4 if deg == 2 else 3
        c.ops = [("matrix", E.K_LAPLACE, [1.5], q, 0, 0, True), ("residual", E.K_LAPLACE, [1.5], q, 0, 0),
                 ("body", [1.0], q, 0)]
    elif name == "stvenant_q1_hex_linear":
        lam, mu = lame(1000.0, 0.25)
        y0 = np.full(3, -0.1); d = np.array([0., 1., 0.])
        c = Case(E.HEX, 1, *make_mesh(E.HEX, n, perturb, permute))
        c.add_field(1, 3, dirichlet=lambda x: H.fund_sol_elastostatic(x, y0, d, lam, mu), values=smooth_u(3),
                    linear=lambda st: linear_constraints(st, 3))
        c.ops = [("matrix", E.K_HYPEL_STVENANT, [lam, mu], 3, 0, 0, True), ("residual", E.K_HYPEL_STVENANT, [lam, mu], 3, 0, 0)]
    elif name == "stokes_p2p1_tet_linear":
        c = Case(E.TET, 1, *make_mesh(E.TET, n, perturb, permute))
        lid = lambda x: np.stack([(x[:, 2] > 1 - 1e-9) * 1.0, 0 * x[:, 0], 0 * x[:, 0]], axis=1)
        u = c.add_field(2, 3, dirichlet=lid, values=smooth_u(3), linear=lambda st: linear_constraints(st, 3))
        p = c.add_field(1, 1, pin_first=True, values=lambda x: x[:, :1] - 0.5, linear=lambda st: linear_constraints(st, 1, 2))
        c.ops = [("matrix", E.K_VECTOR_LAPLACE, [1.0], 4, u, u, True), ("matrix", E.K_PRESSURE_GRADIENT, None, 4, u, p, True),
                 ("matrix", E.K_VELOCITY_DIVERGENCE, [0.0], 4, p, u, True),
                 ("residual", E.K_VECTOR_LAPLACE, [1.0], 4, u, u), ("residual", E.K_PRESSURE_GRADIENT, None, 4, u, p),
                 ("residual", E.K_VELOCITY_DIVERGENCE, [0.0], 4, p, u)]
    else:
        raise ValueError(name)
    c.name = name
    for i, op in enumerate(c.ops):  # normalise params
        if op[0] in ("matrix", "residual") and op[2] is None:
            c.ops[i] = (op[0], op[1], [0.0]) + tuple(op[3:])
    return c


def compare(a, b):
    """a, b = (rowptr, col, val, rhs).  Pattern must be identical; values by the SURVEY 8(d) metric."""
    pattern_equal = np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    res = dict(pattern_equal=bool(pattern_equal), nnz=int(len(a[1])), n=int(len(a[3])))
    if pattern_equal:
        res["val_diff"] = H.csr_rel_diff(a[0], a[2], b[2])
        res["rhs_diff"] = H.vec_rel_diff(a[3], b[3])
    return res


def run_case(name, n=4, perturb=True, permute=False, register=False, eng=None):
    c = build_case(name, n, perturb, permute)
    ref = c.run_oracle(register=register)
    out = c.run_engine(eng=eng, register=register)
    return compare(ref, out)
