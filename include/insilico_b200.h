/* =============================================================================
 * insilico_b200.h -- C ABI of the B200-native element-assembly engine
 * =============================================================================
 * Drop-in boundary for inSilico's assembly hot path.  The
 * reference is header-only C++ templates with no FFI of its own; each entry point
 * below names the reference interface it replaces (paths relative to the
 * reference root).  Plain pointers and sizes only; no C++/torch types.
 *
 * Conventions
 *   - all functions return 0 on success, non-zero on error; isl_last_error()
 *     returns the message (the C++ facade turns it into the reference's
 *     VERIFY_MSG behaviour: message on stderr + abort(), base/verify.hpp:139-149).
 *   - array arguments may be HOST or DEVICE pointers (copied with
 *     cudaMemcpyDefault under unified addressing).
 *   - shapes, n-faces and DoF status use the reference's enum values
 *     (base/shape.hpp:25-45, base/dof/DegreeOfFreedom.hpp:33-38).
 *   - fields are addressed by 0-based id (the reference's FieldBinder uses 1-based
 *     template indices, base/asmb/FieldBinder.hpp:132-138; the facade subtracts 1).
 *   - element-local ordering is the reference's hierarchic ordering (vertices,
 *     edges, faces, cell; base/mesh/HierarchicOrder.hpp), DoF object s major,
 *     component d minor: local index s*dof_size+d (base/asmb/collectFromDoFs.hpp).
 * ===========================================================================*/
#ifndef INSILICO_B200_H
#define INSILICO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* base/shape.hpp:25-32 */
enum { ISL_POINT = 0, ISL_LINE = 1, ISL_TRI = 2, ISL_QUAD = 3, ISL_TET = 4, ISL_HEX = 5 };
/* base/shape.hpp:40-45 */
enum { ISL_VERTEX = 0, ISL_EDGE = 1, ISL_FACE = 2, ISL_CELL = 3 };
/* base/dof/DegreeOfFreedom.hpp:33-38 */
enum { ISL_ACTIVE = 0, ISL_CONSTRAINED = 1, ISL_INACTIVE = 2 };

/* integrand kernels ("kernel objects" of the reference) */
enum {
    ISL_K_LAPLACE = 1,             /* heat::Laplace (heat/Laplace.hpp:113-181), base::kernel::Laplace
                                      (base/kernel/Laplace.hpp:100-151); params = {factor}            */
    ISL_K_HYPEL_STVENANT = 2,      /* solid::HyperElastic<mat::hypel::StVenant> (solid/HyperElastic.hpp:110-257,
                                      mat/hypel/StVenant.hpp); params = {lambda, mu} (mat/Lame.hpp)      */
    ISL_K_HYPEL_NEOHOOKE = 3,      /* solid::HyperElastic<mat::hypel::NeoHookeanCompressible>; {lambda, mu} */
    ISL_K_PRESSURE_GRADIENT = 4,   /* fluid::PressureGradient (fluid/PressureGradient.hpp:76-155); no params */
    ISL_K_VELOCITY_DIVERGENCE = 5, /* fluid::VelocityDivergence (fluid/VelocityDivergence.hpp:67-124);
                                      params = {changeSign != 0}                                       */
    ISL_K_VECTOR_LAPLACE = 6,      /* fluid::VectorLaplace (fluid/VectorLaplace.hpp:40-112); {viscosity}   */
    ISL_K_MASS = 7,                /* base::kernel::Mass (base/kernel/Mass.hpp:88-138), matrix only; {factor}: entry =
                                      factor detJ w phi_M psi_N on every DoF component (time stepping, L2 projections) */
    ISL_K_CONVECTION = 8           /* fluid::Convection (fluid/Convection.hpp:88-220, Picard form), {density}: tangent entry =
                                      phi_M (uAdv . grad phi_N + 0.5 div(u) phi_N) rho detJ w on every component, uAdv from the
                                      tuple's third field; residual rho (u . grad) u_aux phi_M.  Through the _aux entry points. */
};

typedef struct isl_engine* isl_handle;

const char* isl_last_error(void);
int isl_version(void);

/* ---- engine ------------------------------------------------------------- */
int isl_engine_create(int device, isl_handle* out);
int isl_engine_destroy(isl_handle h);
int isl_synchronize(isl_handle h);
/* launches deferred work (the Q1 stiffness kernel waits one call for a body force to fuse) without waiting */
int isl_flush(isl_handle h);
/* kernel-selection knobs that can change between launches without new preprocessing (tuning sweeps; the same knobs
 * are read from ISL_* environment variables when the engine is created): "q1_rows", "rows_threads", "rows_ss",
 * "affine_kernel", "aff_split", "aff_threads", "tangent_tiled", "defer", "gen_gather", "hypel_gather" (1: element
 * matrices to memory + CSR rows gathered without atomics, 0: atomic scatter through the slot maps)           */
int isl_engine_set_option(isl_handle h, const char* name, double value);
/* CUDA stream the engine launches on (cudaStream_t as void*), for event timing by the caller */
void* isl_engine_stream(isl_handle h);
/* number of the engine's own kernels launched since creation (bench.py "gpu_launches") */
int64_t isl_kernel_launches(isl_handle h);

/* measured FP64 peak of this device (independent DFMA chains in registers, CUDA events on the engine stream):
 * the denominator of the FP64-pipe roofline fraction SURVEY 8(d) asks for; nothing comparable in the reference */
int isl_measure_fp64_peak(isl_handle h, double* tflops);
/* measured throughput (1e9 per second) of FP64 atomic adds without return value into a 2 GiB array: pattern 0 coalesced
 * sweep, 1 three neighbouring entries at pseudo-random places, 2 single entries at pseudo-random places.  The scatter of
 * the generic kernels (6 561 adds per Q2 x 3 element) is bounded by it; reported next to configs 3-5.               */
int isl_measure_red_peak(isl_handle h, int pattern, double* gatomics_per_s);

/* ---- host-side tables (no GPU needed) ------------------------------------ */
/* base::Quadrature<DEG,SHAPE> (base/Quadrature.hpp:113-143): returns #points; weights[n], points[n*dim] */
int isl_quadrature(int shape, int degree, double* weights, double* points);
/* base::LagrangeShapeFun<DEG,SHAPE>::fun / gradient in hierarchic order (base/LagrangeShapeFun.hpp) */
int isl_shape_nfun(int shape, int degree);
int isl_shape_eval(int shape, int degree, const double* xi, double* fun, double* grad);
int isl_support_points(int shape, int degree, double* pts);

/* ---- DoF handling (host side, base/dof) ----------------------------------- */
/* base::dof::generate<FEBasis>(mesh, field) (base/dof/generate.hpp:46-115, IndexMap.hpp:221-280):
 * elem_dof[n_elems * ndpe] receives DoF-object ids per element, *n_obj their number.                     */
int isl_dof_generate(int shape, int geom_deg, int64_t n_elems, const int32_t* conn, int fe_deg, int32_t* elem_dof,
                     int64_t* n_obj);
int isl_ndpe(int shape, int fe_deg);
/* the same on the device, for the mesh of the engine (isl_mesh_set): generateDoFIndicesFromFaces
 * (base/dof/generateDoFIndicesFromFaces.hpp:169-298) as stable radix sorts of (sorted vertex tuple, visiting position) and
 * a scan of the "met for the first time" flags: same ids as isl_dof_generate.  elem_dof: host or device pointer,
 * [n_elems * ndpe].                                                                                               */
int isl_dof_generate_device(isl_handle h, int fe_deg, int32_t* elem_dof, int64_t* n_obj);
/* base::mesh::MeshBoundary::create (base/mesh/MeshBoundary.hpp, createBoundaryFromUnstructured.hpp:55-106):
 * pairs[2*k] = element, pairs[2*k+1] = face number; pass NULL to query the count (*n_pairs).              */
int isl_mesh_boundary(int shape, int geom_deg, int64_t n_elems, const int32_t* conn, int64_t* pairs,
                      int64_t* n_pairs);
/* DoFs visited by base::dof::constrainBoundary (base/dof/constrainBoundary.hpp:49-123) in visiting order:
 * for each boundary pair the DoF objects on that face and the physical location of their support points.
 * obj[n], x[n*dim]; pass obj = NULL to query *n.                                                          */
int isl_boundary_dofs(int shape, int geom_deg, int dim, const double* coords, int64_t n_elems, const int32_t* conn,
                      int fe_deg, const int32_t* elem_dof, int64_t n_pairs, const int64_t* pairs, int32_t* obj,
                      double* x, int64_t* n);
/* base::dof::numberDoFsConsecutively (base/dof/numbering.hpp:44-68): eqn[n_obj*dof_size], -1 where not ACTIVE */
int isl_number_dofs(int64_t n_obj, int dof_size, const uint8_t* status, int64_t init, int64_t* eqn,
                    int64_t* n_numbered);

/* ---- mesh + fields on the device (base::Unstructured, base::Field, asmb::FieldBinder) ------------------ */
/* base::Unstructured<SHAPE,GEOMDEG,DIM> (base/Unstructured.hpp:57-65): coords[n_nodes*dim], conn[n_elems*npe] */
int isl_mesh_set(isl_handle h, int shape, int geom_deg, int dim, int64_t n_nodes, const double* coords,
                 int64_t n_elems, const int32_t* conn);
/* multi-GPU element blocks: elements [0, n_owned) are assembled by this engine, the remaining (halo) elements
 * only contribute to the sparsity pattern of the rows this GPU owns                                         */
int isl_mesh_set_owned(isl_handle h, int64_t n_owned);
/* only the nodal coordinates change (moving mesh); pattern and maps stay valid */
int isl_mesh_update_coords(isl_handle h, const double* coords);
/* base::Field<FEBasis,DOFSIZE> (base/Field.hpp:50-55) flattened: per DoF component eqn / status / prescribed
 * value (dof::Constraint rhs, base/dof/Constraint.hpp) / current value                                      */
int isl_field_set(isl_handle h, int field, int fe_deg, int dof_size, int64_t n_obj, const int32_t* elem_dof,
                  const int64_t* eqn, const uint8_t* status, const double* prescribed, const double* values);
/* general linear constraints (base/dof/Constraint.hpp:57-140, collected per element by asmb::collectFromDoFs,
 * base/asmb/collectFromDoFs.hpp:112-131, applied by asmb::assembleMatrix / assembleForces,
 * base/asmb/assembleMatrix.hpp:212-338, assembleForces.hpp:58-139): DoF component con_dof[k] = obj*dof_size+comp
 * (status CONSTRAINED) is  u = prescribed + sum_j weight[j] * u_master[j],  j in [con_ptr[k], con_ptr[k+1]), with
 * master_eqn[j] the equation number of the (ACTIVE) master.  HOST arrays.  Call after isl_field_set (which drops
 * earlier constraints); n_con = 0 removes them.  Fields with such slaves use the generic kernels.              */
int isl_field_set_constraints(isl_handle h, int field, int64_t n_con, const int64_t* con_dof, const int64_t* con_ptr,
                              const int64_t* master_eqn, const double* weight);
/* new Newton state / new Dirichlet values, numbering unchanged (either pointer may be NULL) */
int isl_field_update(isl_handle h, int field, const double* prescribed, const double* values);

/* ---- solver hand-off (base::solver::Eigen3) ------------------------------------------------------------- */
/* Solver solver(n) (base/solver/Eigen3.hpp:71-77): fresh system, zero rhs, no matrix entries.  The sparsity
 * pattern and element->slot maps of previously registered field pairs stay cached on the device.          */
int isl_system_create(isl_handle h, int64_t n_eqn);
/* solver.registerFields<FTB>(fieldBinder) (Eigen3.hpp:332-336, TripletContainer.hpp:158-301): unions the
 * (test,trial) block pattern into the system CSR                                                            */
int isl_pattern_register(isl_handle h, int test_field, int trial_field);
/* asmb::stiffnessMatrixComputation<FTB>(quad, solver, binder, kernelObj, incremental)
 * (base/asmb/StiffnessMatrix.hpp:49-87): K into the CSR values, Dirichlet lift into rhs.  Every CSR row is summed in the
 * caller's element order by the threads that own it (no atomics), so the matrix is reproducible bit for bit; fields with
 * slaves of master DoFs take the atomic scatter (isl_field_set_constraints)                                   */
int isl_assemble_matrix(isl_handle h, int kernel_id, const double* params, int quad_deg, int test_field,
                        int trial_field, int incremental);
/* kernels that read a third field of the tuple (FieldTupleBinder<I,J,K>: AuxField1Element, base/asmb/FieldTupleBinder.hpp),
 * e.g. fluid::Convection with the advection velocity; aux_field = its index (same FE basis as the trial field)            */
int isl_assemble_matrix_aux(isl_handle h, int kernel_id, const double* params, int quad_deg, int test_field, int trial_field,
                            int aux_field, int incremental);
int isl_assemble_residual_aux(isl_handle h, int kernel_id, const double* params, int quad_deg, int test_field, int trial_field,
                              int aux_field, double factor);
/* the same for heat::Laplace with a conductivity FUNCTION (heat/Laplace.hpp:85-126, setConductivityFunction): values[n_elems *
 * nq] = conductivity at the quadrature points of Quadrature<quad_deg> of every (owned) element, host or device pointer; the
 * caller's function runs on the host (like isl_assemble_bodyforce_sampled), the integration on the device.  Laplace kernels. */
int isl_assemble_matrix_sampled(isl_handle h, int kernel_id, const double* values, int quad_deg, int test_field,
                                int trial_field, int incremental);
/* asmb::computeResidualForces<FTB>(quad, solver, binder, kernelObj) (base/asmb/ForceIntegrator.hpp:37-71):
 * rhs += factor * f_e with factor = -1 in the reference                                                     */
int isl_assemble_residual(isl_handle h, int kernel_id, const double* params, int quad_deg, int test_field,
                          int trial_field, double factor);
/* asmb::bodyForceComputation<FTB>(quad, solver, binder, f) with constant f[dof_size]
 * (base/asmb/BodyForce.hpp:65-84,172-205)                                                                   */
int isl_assemble_bodyforce(isl_handle h, const double* f, int quad_deg, int test_field);
/* the same with a general f(x) (BodyForce.hpp:172-205 evaluates the caller's function at every quadrature point):
 * values[n_elems * nq * dof_size] = f at the physical location x(xi_q) of every (owned) element and quadrature point
 * of Quadrature<quad_deg> (isl_quadrature gives xi_q; x = sum_a N_a(xi_q) x_a), host or device pointer.  The caller's
 * function runs on the host, the integration on the device.                                                       */
int isl_assemble_bodyforce_sampled(isl_handle h, const double* values, int quad_deg, int test_field);
/* ---- surface (Neumann) terms ------------------------------------------------------------------------------------
 * A surface element is what base::mesh::SurfaceElement (base/mesh/SurfaceElement.hpp:93-201) holds: P nodes (P = nodes of
 * the Lagrange element of the face shape and the mesh's geometry degree) with physical coordinates surf_x[P*dim], the
 * coordinates of the same nodes in the parameter space of the domain element surf_param[P*dim], and the domain element.
 *
 * isl_boundary_surface: base::mesh::generateBoundaryMesh (base/mesh/generateBoundaryMesh.hpp:279-432, no triangulation)
 * for (element, face number) pairs as isl_mesh_boundary returns them.  Host-side; pass domain_elem = NULL to query
 * *surf_shape / *nodes_per_surf only.                                                                              */
int isl_boundary_surface(int shape, int geom_deg, int dim, const double* coords, const int32_t* conn, int64_t n_pairs,
                         const int64_t* pairs, int32_t* domain_elem, double* surf_x, double* surf_param, int* surf_shape,
                         int* nodes_per_surf);
/* what base::asmb::NeumannForce hands to the caller's force function f(x, normal) (base/asmb/NeumannForce.hpp:152-163;
 * base::SurfaceNormal, base/geometry.hpp:256-346): position, unit normal and surface metric at the points of
 * SurfaceQuadrature<quad_deg> (base/Quadrature.hpp:148-151) of every surface element.  Host-side; x[n_surf*nq*dim],
 * normal[n_surf*nq*dim], detg[n_surf*nq], any of them NULL; surf_x = NULL queries *nq only.                          */
int isl_surface_points(int surf_shape, int geom_deg, int dim, int64_t n_surf, const double* surf_x, int quad_deg, double* x,
                       double* normal, double* detg, int* nq);
enum isl_neumann_mode {
    ISL_NEUMANN_CONSTANT = 0,      /* data[dof_size]: f constant                                                      */
    ISL_NEUMANN_NORMAL = 1,        /* data[1]: f = data[0] * normal (dof_size == dim), e.g. a pressure load            */
    ISL_NEUMANN_SAMPLED = 2        /* data[n_surf*nq*dof_size]: f at the surface quadrature points, host or device     */
};
/* base::asmb::neumannForceComputation<SFTB>(surfaceQuadrature, solver, surfaceFieldBinder, f)
 * (base/asmb/NeumannForce.hpp:33-66,140-184; SurfaceFieldBinder.hpp:81-184): rhs += int f(x, n) phi ds over the surface
 * elements, phi = the test field's shape functions of the domain element at localDomainCoordinate(eta).  One kernel
 * launch for all surface elements; the caller's function (mode SAMPLED) runs on the host.                              */
int isl_assemble_neumann(isl_handle h, int64_t n_surf, const int32_t* domain_elem, const double* surf_x, const double* surf_param,
                         int quad_deg, int test_field, int mode, const double* data);
/* the same with the equation numbers given per surface element instead of a field index: rows[n_surf * ndpe * dof_size]
 * (shape function major, < 0: not ACTIVE, skipped), for a caller that reads them off its own element objects (the
 * reference-tree binding does: SurfaceFieldBinder tuples need not belong to a FieldBinder the engine has seen, so the
 * shape and geometry degree of the domain elements are arguments and no mesh or field needs to be set)                  */
int isl_assemble_neumann_rows(isl_handle h, int shape, int geom_deg, int64_t n_surf, const double* surf_x, const double* surf_param,
                              int quad_deg, int fe_deg, int dof_size, const int32_t* rows, int mode, const double* data);
/* solver.insertToLHS / insertToRHS (Eigen3.hpp:81-124) for host-side odd contributions:
 * mat is row-major [n_rows*n_cols]; entries must exist in the registered pattern                            */
int isl_insert_lhs(isl_handle h, const double* mat, const int64_t* rows, int n_rows, const int64_t* cols,
                   int n_cols);
int isl_insert_rhs(isl_handle h, const double* vec, const int64_t* rows, int n_rows);
/* solver.finishAssembly() (Eigen3.hpp:142-153): waits for the device, reports sizes                          */
int isl_finish(isl_handle h, int64_t* n_eqn, int64_t* nnz);
/* canonical CSR (rows ascending, columns ascending, explicit zeros kept) + rhs; NULL pointers are skipped   */
int isl_get_csr(isl_handle h, int64_t* rowptr, int32_t* col, double* val, double* rhs);
/* the same hand-off without waiting: values and rhs of the finished system go to PINNED host buffers on a copy stream
 * while the engine already assembles the next system into a second set of buffers (the caller's next call must be
 * isl_system_create); isl_copy_wait blocks until the last such copy has arrived.  For callers that insist on a host CSR
 * per step the device -> host copy then overlaps the next step's host -> device copies and kernels.              */
int isl_get_csr_async(isl_handle h, double* val, double* rhs);
int isl_copy_wait(isl_handle h);
/* zero-copy hand-off to a device solver: device pointers valid until the next create/register call         */
int isl_get_device_csr(isl_handle h, int64_t** rowptr, int32_t** col, double** val, double** rhs);
/* solver.getValue(i), solver.norm() (Eigen3.hpp:293-296,128-138; norm = ||b||_2 / n, quirk kept)            */
int isl_rhs_value(isl_handle h, int64_t index, double* value);
int isl_rhs_norm(isl_handle h, double* norm);

/* solver.cgSolve() (base/solver/Eigen3.hpp:263-275: Eigen::ConjugateGradient, diagonal preconditioner, zero initial
 * guess): solves A x = rhs on the device for the finished s.p.d. system, rhs is replaced by x.  tol <= 0 = machine
 * epsilon, max_iter <= 0 = 2 n (Eigen's defaults); *error = |r| / |b| at exit.  First row of SURVEY 8(f); not on the
 * benchmarked path.                                                                                            */
int isl_solve_cg(isl_handle h, double tol, int64_t max_iter, int64_t* iterations, double* error);

/* base::dof::setDoFsFromSolver (add = 0) / addToDoFsFromSolver (add != 0) (base/dof/Distribute.hpp:35-56,139-215) on the
 * device: after a solve the rhs holds the solution; ACTIVE components of the field take / add their entry, CONSTRAINED
 * ones are set to their constraint value (prescribed + weighted masters).  With isl_solve_cg this closes a Newton
 * iteration without the matrix or the field leaving the GPU (solid/CompressibleDriver.hpp:179-210).            */
int isl_distribute(isl_handle h, int field, int add);
/* current values of a field [n_obj * dof_size] -> host or device buffer */
int isl_field_get_values(isl_handle h, int field, double* values);

/* ---- multi-GPU: one engine per process and GPU, NCCL inside the engine (the reference's only parallel
 * construct is the OpenMP loop of base/auxi/parallel.hpp:25-60) ----------------------------------------------- */
/* rank 0 makes the 128-byte NCCL id, the caller hands it to the other processes (file, socket, MPI, torch ...) */
int isl_comm_unique_id(void* id128);
int isl_comm_init(isl_handle h, const void* id128, int rank, int world);
int isl_comm_destroy(isl_handle h);
/* exchange plan, collective, after the pattern is registered.  l2g[n_local]: global equation of every local one;
 * rows [own_lo, own_hi) are owned here (l2g ascending on that range); the ghost rows [seg_lo[k], seg_hi[k]) are
 * assembled here but owned by rank seg_owner[k] (at most one segment per owner).  The (global row, global column)
 * keys of the ghost entries go to the owners once, which locate them in their CSR.                            */
int isl_exchange_setup(isl_handle h, int64_t n_local, const int64_t* l2g, int64_t own_lo, int64_t own_hi, int n_seg,
                       const int* seg_owner, const int64_t* seg_lo, const int64_t* seg_hi);
/* after the local assembly calls of a step: ghost values and rhs rows -> owners (ncclSend / ncclRecv on a second
 * stream), added there; overlaps the interior patches of the Q1 row kernel, which launches interface patches first */
int isl_exchange(isl_handle h);

/* ---- helpers of the round-1 exchange driven from Python (kept for the torch.distributed / gloo path) ------- */
/* gather val[idx[k]] (or rhs when which = 1) into a packed device buffer / scatter-add a packed buffer      */
int isl_pack_entries(isl_handle h, int which, const int64_t* idx_dev, int64_t n, double* out_dev);
int isl_unpack_add_entries(isl_handle h, int which, const int64_t* idx_dev, int64_t n, const double* in_dev);

#ifdef __cplusplus
}
#endif
#endif /* INSILICO_B200_H */
