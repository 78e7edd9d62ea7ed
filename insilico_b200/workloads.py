"""Synthetic workloads of the BASELINE.json configs as flat arrays for the C ABI.

  C1  reference/04-heat/dirichlet.cpp flow on unitCube N^3 Q1 hexahedra (N in 8, 16, 32), scalar Laplace, Dirichlet data
      from the Laplace fundamental solution on the whole boundary (dirichlet.cpp:127-142), quadrature degree 3
  C3  linear elasticity: solid::HyperElastic<mat::hypel::StVenant> (E = 1000, nu = 0.25, reference/06-elastic/
      linearElastic.cpp:83-84) at u = 0 on n^3 hexahedra, Q1 geometry, Q2 displacement field with 3 DoFs per node,
      quadrature degree 4 (27 points), displacement fixed on the face x = 0
  C4  compressible neo-Hooke (E = 1000, nu = 0.3, reference/06-elastic/input.dat): tangent + residual of one Newton
      step on 6 n^3 perturbed P1-geometry tetrahedra in random element order, P2 displacement, current state
      u = 0.05 sin(pi x), quadrature degree 4 (11 points)
  C5  Taylor-Hood Stokes blocks (reference/07-drivenCavity/drivenCavity.cpp:176-275): P2 velocity x 3 + P1 pressure on
      6 n^3 perturbed tetrahedra in random element order, blocks UU (fluid::VectorLaplace), UP (PressureGradient),
      PU (VelocityDivergence), quadrature degree 4, velocity prescribed on the boundary (lid), pressure DoF 0 pinned,
      block offsets as drivenCavity.cpp:205-211

Each builder returns a Workload: mesh, fields (numbering by the engine's host-side restatement of base/dof, the same
calls the parity tests pin against the reference run), the list of assembly operations of one step, and the
algorithmic bytes / flops per element that bench.py's roofline uses (formulas below, stated in DESIGN.md section 4).
Nothing here touches oracle/ or tests/.
"""
import numpy as np

from . import engine as E
from . import meshgen


def lame(emod, nu):
    """mat::Lame (mat/Lame.hpp:24-51)"""
    return emod * nu / (1. + nu) / (1. - 2. * nu), emod / 2. / (1. + nu)


def fund_sol_laplace(x, src=-0.5):
    d = np.sqrt(((x - src) ** 2).sum(axis=1))
    return (1.0 / (4.0 * np.pi)) / d


class Workload:
    def __init__(self, name, shape, geom_deg, coords, conn):
        self.name, self.shape, self.geom_deg = name, shape, geom_deg
        self.coords = np.ascontiguousarray(coords, dtype=np.float64)
        self.conn = np.ascontiguousarray(conn, dtype=np.int32)
        self.dim = self.coords.shape[1]
        self.fields, self.ops, self.n_eqn = [], [], 0
        self.description = ""

    def dof_positions(self, fe_deg, ed, nobj):
        sp = E.support_points(self.shape, fe_deg)
        ng = np.array([E.shape_eval(self.shape, self.geom_deg, s)[0] for s in sp])
        pos = np.zeros((nobj, self.dim))
        # chunked: the einsum over all elements at once would need n_elems * ndpe * dim doubles twice
        step = 1 << 20
        for a in range(0, len(self.conn), step):
            xe = self.coords[self.conn[a:a + step]]
            xd = np.einsum("la,ead->eld", ng, xe)
            pos[ed[a:a + step].reshape(-1)] = xd.reshape(-1, self.dim)
        return pos

    def add_field(self, fe_deg, ds, dirichlet=None, where=None, values=None, pin_first=False):
        """dirichlet(x) -> [n, ds] on the boundary DoFs selected by where(x) (default: all of the boundary);
        values(x) -> current state"""
        ed, nobj = E.dof_generate(self.shape, self.geom_deg, self.conn, fe_deg)
        status = np.zeros((nobj, ds), dtype=np.uint8)
        presc = np.zeros((nobj, ds))
        if dirichlet is not None:
            pairs = E.mesh_boundary(self.shape, self.geom_deg, self.conn)
            obj, x = E.boundary_dofs(self.shape, self.geom_deg, self.coords, self.conn, fe_deg, ed, pairs)
            if where is not None:
                keep = where(x)
                obj, x = obj[keep], x[keep]
            status[obj] = E.CONSTRAINED
            presc[obj] = np.asarray(dirichlet(x), dtype=np.float64).reshape(len(obj), ds)
        if pin_first:
            status[0, 0] = E.CONSTRAINED
            presc[0, 0] = 0.0
        eqn, n = E.number_dofs_consecutively(status, init=self.n_eqn)
        self.n_eqn += n
        vals = np.zeros((nobj, ds))
        if values is not None:
            vals = np.asarray(values(self.dof_positions(fe_deg, ed, nobj)), dtype=np.float64).reshape(nobj, ds)
        self.fields.append(dict(fe_deg=fe_deg, ds=ds, n_obj=nobj, elem_dof=ed, status=status, presc=presc, eqn=eqn,
                                values=vals, ndpe=ed.shape[1]))
        return len(self.fields) - 1

    # ---- engine ------------------------------------------------------------------------------------------
    def upload(self, eng):
        eng.set_mesh(self.shape, self.geom_deg, self.coords, self.conn)
        for i, f in enumerate(self.fields):
            eng.set_field(i, f["fe_deg"], f["ds"], f["n_obj"], f["elem_dof"], f["eqn"], f["status"], f["presc"], f["values"])

    def register(self, eng):
        for op in self.ops:
            if op[0] == "matrix":
                eng.register_fields(op[4], op[5])

    def step(self, eng):
        """one assembly pass: fresh solver, every operation of the step"""
        eng.new_solver(self.n_eqn)
        for op in self.ops:
            if op[0] == "matrix":
                eng.stiffness_matrix_computation(op[1], op[2], op[3], op[4], op[5], incremental=op[6])
            elif op[0] == "residual":
                eng.compute_residual_forces(op[1], op[2], op[3], op[4], op[5])
            elif op[0] == "body":
                eng.body_force_computation(op[1], op[2], op[3])

    # ---- roofline figures ----------------------------------------------------------------------------------
    def algorithmic_flops_per_element(self):
        """FP64 operations of one step per element with the per-point quantities hoisted (SURVEY 8(d) counts the
        integrals, not the reference's per-entry recomputation).  A fused multiply-add counts 2.
          geometry per point: J = sum_n x_n (x) dN_n: 2 dim^2 npe; inverse + determinant ~ 50 (dim 3);
                              physical gradients 2 dim^2 per basis function
          Laplace / VectorLaplace tangent: (2 dim + 1) per (M, N) pair and point
          HyperElastic tangent: C_eff 81 * (1 + 2 * 9 * 2) per point, B_N = C_eff g_N: 2 * 81 per trial node,
                                g_M . B_N: 2 dim per entry and point
          residuals: gradient of the trial field 2 dim ds per node, stress ~ 150, 2 dim per row and point
          PressureGradient / VelocityDivergence: 3 per entry and point"""
        dim = self.dim
        npe = self.conn.shape[1]
        total = 0.0
        for op in self.ops:
            if op[0] == "body":
                w, _ = E.quadrature(self.shape, op[2])
                f = self.fields[op[3]]
                total += len(w) * (2 * dim * dim * npe + 50 + 3 * f["ndpe"] * f["ds"])
                continue
            w, _ = E.quadrature(self.shape, op[3])
            nq = len(w)
            ft, fc = self.fields[op[4]], self.fields[op[5]]
            nt, nc = ft["ndpe"], fc["ndpe"]
            geo = 2 * dim * dim * npe + 50 + 2 * dim * dim * (nt + (nc if op[4] != op[5] else 0))
            kid = op[1]
            if op[0] == "matrix":
                if kid in (E.K_LAPLACE, E.K_VECTOR_LAPLACE):
                    per = (2 * dim + 1) * nt * nc
                elif kid in (E.K_HYPEL_STVENANT, E.K_HYPEL_NEOHOOKE):
                    per = 81 * 37 + 2 * 81 * nc + 2 * dim * (nt * dim) * (nc * dim) + 2 * dim * dim * nc + 150
                else:
                    per = 3 * (nt * ft["ds"]) * (nc * fc["ds"])
            else:
                per = 2 * dim * fc["ds"] * nc + 150 + 2 * dim * nt * ft["ds"]
            total += nq * (geo + per)
        return total

    def atomics_per_element(self):
        """FP64 atomic adds one step issues per element: one per local force entry (k_force) and, where the atomic scatter
        is used, one per local matrix entry that is integrated.  The matrix operations take the atomic-free paths by default
        (isl_gather.cuh: element matrices to memory, CSR rows gathered), unless ISL_GEN_GATHER=0 selects the atomic scatter of
        the generic kernels (the component-diagonal entries of the Laplace-type kernels, nt dim nc of a Stokes coupling block)"""
        import os
        gather_default = os.environ.get("ISL_GEN_GATHER", "1") != "0"
        total = 0
        for op in self.ops:
            if op[0] == "body":
                f = self.fields[op[3]]
                total += f["ndpe"] * f["ds"]
                continue
            ft, fc = self.fields[op[4]], self.fields[op[5]]
            if op[0] != "matrix":
                total += ft["ndpe"] * ft["ds"]
            elif op[1] in (E.K_HYPEL_STVENANT, E.K_HYPEL_NEOHOOKE):
                # 3-D vector fields take the atomic-free path (element matrices to memory, rows gathered): no atomics
                if not (self.dim == 3 and ft["ds"] == 3 and op[4] == op[5]):
                    total += (ft["ndpe"] * ft["ds"]) * (fc["ndpe"] * fc["ds"])
            elif not gather_default:
                if op[1] in (E.K_PRESSURE_GRADIENT, E.K_VELOCITY_DIVERGENCE):
                    total += ft["ndpe"] * fc["ndpe"] * self.dim
                else:
                    total += ft["ndpe"] * fc["ndpe"] * fc["ds"]
        return total

    def algorithmic_bytes_per_element(self, nnz):
        """compulsory HBM traffic of one step per element, every array touched once: connectivity,
        element -> DoF tables, element -> CSR slot maps of the matrix operations, and per element its share of the
        unique coordinates, CSR values, right-hand side and field values"""
        ne = getattr(self, "n_elems_global", len(self.conn))      # (set when the mesh is generated per rank: partition.structured_stokes_slab)
        n_nodes = getattr(self, "n_nodes_global", len(self.coords))
        b = self.conn.shape[1] * 4.0
        used = set()
        for op in self.ops:
            if op[0] == "matrix":
                ft, fc = self.fields[op[4]], self.fields[op[5]]
                b += 4.0 * ft["ndpe"] * ft["ds"] * fc["ndpe"] * fc["ds"]
                used.update((op[4], op[5]))
            elif op[0] == "residual":
                used.update((op[4], op[5]))
            else:
                used.add(op[3])
        for i in used:
            b += self.fields[i]["ndpe"] * 4.0
        shared = 8.0 * self.dim * n_nodes + 8.0 * nnz + 8.0 * self.n_eqn
        if any(op[0] == "residual" for op in self.ops):
            shared += sum(8.0 * self.fields[i]["n_obj"] * self.fields[i]["ds"] for i in used)
        return b + shared / ne


def build(config, n, perturb=None, permute=None):
    config = config.upper()
    if config in ("C1", "C2"):
        coords, conn, _ = meshgen.unit_cube_hex(n, n, n)
        w = Workload(config, E.HEX, 1, coords, conn)
        w.add_field(1, 1, dirichlet=lambda x: fund_sol_laplace(x))
        w.ops = [("matrix", E.K_LAPLACE, [1.0], 3, 0, 0, True)]
        if config == "C2":
            w.ops.append(("body", [1.0], 3, 0))
        w.description = ("%s: 3D scalar Laplace Q1 hex, structured %d^3 mesh, stiffness%s, quadrature degree 3"
                         % (config, n, " + RHS (Dirichlet lift + constant body force)" if config == "C2" else " + Dirichlet lift"))
        return w
    if config == "C3":
        coords, conn, _ = meshgen.unit_cube_hex(n, n, n)
        lam, mu = lame(1000.0, 0.25)
        w = Workload(config, E.HEX, 1, coords, conn)
        w.add_field(2, 3, dirichlet=lambda x: 0.0 * x, where=lambda x: x[:, 0] < 1e-12)
        w.ops = [("matrix", E.K_HYPEL_STVENANT, [lam, mu], 4, 0, 0, True)]
        w.description = ("C3: linear elasticity (mat::Lame E=1000 nu=0.25, HyperElastic<StVenant> at u=0) Q2 hex, 3 DoF/node, "
                         "structured %d^3 mesh, Q1 geometry, quadrature degree 4 (27 points), face x=0 fixed" % n)
        return w
    if config in ("C4", "C5"):
        coords, conn = meshgen.unit_cube_tet(n, n, n)
        if perturb is None or perturb:
            coords = meshgen.perturb_interior(coords, 1.0 / n, max_dist=0.1)
        if permute is None or permute:
            conn = meshgen.permute_elements(conn)
        w = Workload(config, E.TET, 1, coords, conn)
        if config == "C4":
            lam, mu = lame(1000.0, 0.3)
            w.add_field(2, 3, dirichlet=lambda x: 0.0 * x, values=lambda x: 0.05 * np.sin(np.pi * x))
            w.ops = [("matrix", E.K_HYPEL_NEOHOOKE, [lam, mu], 4, 0, 0, True),
                     ("residual", E.K_HYPEL_NEOHOOKE, [lam, mu], 4, 0, 0)]
            w.description = ("C4: compressible neo-Hooke (E=1000 nu=0.3) tangent + residual of one Newton step, P2 tets x 3 DoF, "
                             "6*%d^3 perturbed tetrahedra in random order, u = 0.05 sin(pi x), quadrature degree 4 (11 points)" % n)
        else:
            lid = lambda x: np.stack([(x[:, 2] > 1 - 1e-9) * 1.0, 0 * x[:, 0], 0 * x[:, 0]], axis=1)
            u = w.add_field(2, 3, dirichlet=lid)
            p = w.add_field(1, 1, pin_first=True)
            w.ops = [("matrix", E.K_VECTOR_LAPLACE, [1.0], 4, u, u, True), ("matrix", E.K_PRESSURE_GRADIENT, [0.0], 4, u, p, True),
                     ("matrix", E.K_VELOCITY_DIVERGENCE, [0.0], 4, p, u, True)]
            w.description = ("C5: Taylor-Hood P2/P1 Stokes blocks UU (VectorLaplace) + UP (PressureGradient) + PU "
                             "(VelocityDivergence), 6*%d^3 perturbed tetrahedra in random order, quadrature degree 4, lid-driven "
                             "velocity on the boundary, pressure DoF 0 pinned" % n)
        return w
    raise ValueError("unknown config %r" % (config,))
