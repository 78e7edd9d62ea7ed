// =============================================================================
// isl_neumann.cuh -- surface (Neumann) terms on the device (included by isl_engine.cu).
//
// Reference: base::asmb::neumannForceComputation<SFTB>(surfaceQuadrature, solver, surfaceFieldBinder, f)
// (base/asmb/NeumannForce.hpp:33-66) applies base::asmb::NeumannForce (NeumannForce.hpp:140-184) through
// ForceIntegrator / assembleForces to every surface element: per quadrature point eta
//     x = Geometry(surfEp, eta);  detG = SurfaceNormal(surfEp, eta, normal);  f = forceFun(x, normal)
//     xi = surfEp->localDomainCoordinate(eta);  phi = test shape functions of the DOMAIN element at xi
//     vector.segment(s * doFSize, doFSize) += f * phi[s] * weight * detG
// A surface element is (base/mesh/SurfaceElement.hpp:93-201) P nodes with physical coordinates, the coordinates of the
// same nodes in the parameter space of its domain element, and the domain element.
//
// Here: one thread per (surface element, test function, DoF component); the surface shape-function tables at the
// quadrature points are the same for all elements, the test functions phi(xi(eta_q)) depend on the parameter
// coordinates only and are tabulated once per DISTINCT block of parameter coordinates (six for the faces of a
// hexahedron): phi[pattern][q][s].  The force is a constant vector, a constant times the normal, or sampled by the caller
// at the quadrature points (the caller's function runs on the host: isl_surface_points gives it x and normal).
// =============================================================================
#pragma once

struct NeumannParams {
    int64_t n_surf;
    int dim, P, nq, nt, ds, mode;
    const int32_t* elem;      // domain element of every surface element
    const double* sx;         // [n_surf][P][dim] physical node coordinates
    const int32_t* pat;       // [n_surf] index of the parameter-coordinate block
    const double* phi;        // [n_patterns][nq][nt]
    const double* sdN;        // [nq][P][dim-1] surface shape-function derivatives at the quadrature points
    const double* w;          // [nq]
    const double* data;       // mode SAMPLED: [n_surf][nq][ds]
    double f[3];              // mode CONSTANT: the force; mode NORMAL: f[0] * normal
    const int32_t* rows;      // explicit equation numbers [n_surf][nt*ds] (< 0: skip), or nullptr: from the field arrays
    const int32_t* ed; const int32_t* eqn;
    const int32_t* cptr; const int32_t* cm; const double* cw;
    double* rhs;
};

template <int DIM>
__global__ void __launch_bounds__(128) k_neumann(const NeumannParams p) {
    const int nr = p.nt * p.ds;
    const int64_t total = p.n_surf * nr;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = t / nr;
        const int i = (int)(t % nr), M = i / p.ds, ci = i % p.ds;
        const double* xs = p.sx + (size_t)k * p.P * DIM;
        const double* phi = p.phi + (size_t)p.pat[k] * p.nq * p.nt;
        double acc = 0.;
        for (int q = 0; q < p.nq; q++) {
            // Jacobi matrix of the surface element, its cross product and the surface metric
            double J[DIM][DIM - 1];
            for (int d = 0; d < DIM; d++) for (int a = 0; a < DIM - 1; a++) J[d][a] = 0.;
            const double* dN = p.sdN + (size_t)q * p.P * (DIM - 1);
            for (int n = 0; n < p.P; n++)
                for (int d = 0; d < DIM; d++)
                    for (int a = 0; a < DIM - 1; a++) J[d][a] += xs[n * DIM + d] * dN[n * (DIM - 1) + a];
            double nrm[DIM], len;
            if (DIM == 3) {
                nrm[0] = J[1][0] * J[2][DIM - 2] - J[2][0] * J[1][DIM - 2];
                nrm[1] = J[2][0] * J[0][DIM - 2] - J[0][0] * J[2][DIM - 2];
                nrm[DIM - 1] = J[0][0] * J[1][DIM - 2] - J[1][0] * J[0][DIM - 2];
                len = sqrt(nrm[0] * nrm[0] + (nrm[1] * nrm[1] + nrm[DIM - 1] * nrm[DIM - 1]));
            } else {
                nrm[0] = J[1][0]; nrm[1] = -J[0][0];
                len = sqrt(nrm[0] * nrm[0] + nrm[1] * nrm[1]);
            }
            double fc;
            if (p.mode == ISL_NEUMANN_SAMPLED) fc = p.data[((size_t)k * p.nq + q) * p.ds + ci];
            else if (p.mode == ISL_NEUMANN_NORMAL) fc = p.f[0] * (nrm[ci] / len);
            else fc = p.f[ci];
            acc += fc * phi[q * p.nt + M] * p.w[q] * len;
        }
        if (p.rows) {
            const int32_t r = p.rows[(size_t)k * nr + i];
            if (r >= 0) atomicAdd(p.rhs + r, acc);
            continue;
        }
        const size_t kr = (size_t)p.ed[(size_t)p.elem[k] * p.nt + M] * p.ds + ci;
        const int32_t r = p.eqn[kr];
        if (r >= 0) atomicAdd(p.rhs + r, acc);
        else if (p.cptr != nullptr)  // slave of master DoFs (asmb/assembleForces.hpp:118-131)
            isl_scatter_force_to_masters<DeviceAdd>(IslMasters{p.cptr, p.cm, p.cw}, p.rhs, kr, acc);
    }
}
