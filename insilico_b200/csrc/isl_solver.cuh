// =============================================================================
// isl_solver.cuh -- conjugate gradients on the device for the finished CSR system (the linear-solve
// hand-off; reference: base::solver::Eigen3::cgSolve, base/solver/Eigen3.hpp:263-275 = Eigen::ConjugateGradient with
// its default diagonal preconditioner, tolerance = machine epsilon, at most 2 n iterations, zero initial guess).
// Included by isl_engine.cu inside its anonymous namespace.  Entry point isl_solve_cg; GPU-verified against a sparse direct
// solve and inside the device-resident Newton loop (tests/test_zz_linear_constraints.py).
//
// HBM-bound: per iteration one pass over the matrix (12 B per non-zero + the gathered vector) and ~9 vector passes.
// One warp per row for the product (rows hold 27..375 non-zeros), warp-shuffle reductions, one atomicAdd per warp.
// =============================================================================
#pragma once

// dinv[r] = 1 / A[r,r] (Eigen's DiagonalPreconditioner uses 1 where the diagonal is zero)
__global__ void k_cg_diag_inv(const int64_t* rowptr, const int32_t* col, const double* val, int64_t n, double* dinv) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
        const int64_t pos = find_in_row(rowptr, col, (int32_t)r, (int32_t)r);
        const double d = pos >= 0 ? val[pos] : 0.;
        dinv[r] = d != 0. ? 1.0 / d : 1.0;
    }
}

// y = A x and *dot += x . y ; LPR lanes per row (8 for short rows: four rows per warp are in flight, which is what hides
// the three dependent loads rowptr -> col / val -> x[col]; one warp per row was measured at 5.4 ms per product of the
// 256^3 Laplace matrix, 5x its HBM time)
template <int LPR>
__global__ void __launch_bounds__(256) k_cg_spmv_dot(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                                     const double* __restrict__ val, const double* __restrict__ x,
                                                     double* __restrict__ y, int64_t n, double* dot) {
    const int sub = threadIdx.x % LPR;
    const int64_t grp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LPR;
    const int64_t ngrp = ((int64_t)gridDim.x * blockDim.x) / LPR;
    double acc = 0.;
    const int64_t trips = (n + 2 * ngrp - 1) / (2 * ngrp);   // the same for every thread: the shuffles below need whole warps
    for (int64_t it = 0; it < trips; it++) {                 // two rows per group and trip: more loads in flight
        const int64_t ra = grp + it * 2 * ngrp, rb = ra + ngrp;
        double sa = 0., sb = 0.;
        const int64_t a0 = ra < n ? rowptr[ra] : 0, a1 = ra < n ? rowptr[ra + 1] : 0;
        const int64_t b0 = rb < n ? rowptr[rb] : 0, b1 = rb < n ? rowptr[rb + 1] : 0;
        for (int64_t k = a0 + sub, j = b0 + sub; k < a1 || j < b1; k += LPR, j += LPR) {   // (no shuffles inside: may diverge)
            if (k < a1) sa += val[k] * x[col[k]];
            if (j < b1) sb += val[j] * x[col[j]];
        }
        __syncwarp();
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) {
            sa += __shfl_down_sync(0xffffffffu, sa, o, LPR);
            sb += __shfl_down_sync(0xffffffffu, sb, o, LPR);
        }
        if (sub == 0) {
            if (ra < n) { y[ra] = sa; acc += x[ra] * sa; }
            if (rb < n) { y[rb] = sb; acc += x[rb] * sb; }
        }
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc != 0.) atomicAdd(dot, acc);
}

// x += alpha p ; r -= alpha Ap ; out[0] += r.r ; out[1] += r.(dinv r)
__global__ void k_cg_update_xr(double alpha, const double* __restrict__ p, const double* __restrict__ Ap,
                               const double* __restrict__ dinv, double* __restrict__ x, double* __restrict__ r, int64_t n,
                               double* out) {
    double rr = 0., rz = 0.;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        x[i] += alpha * p[i];
        const double ri = r[i] - alpha * Ap[i];
        r[i] = ri;
        rr += ri * ri;
        rz += ri * (dinv[i] * ri);
    }
    for (int o = 16; o > 0; o >>= 1) { rr += __shfl_down_sync(0xffffffffu, rr, o); rz += __shfl_down_sync(0xffffffffu, rz, o); }
    if ((threadIdx.x & 31) == 0) { atomicAdd(out, rr); atomicAdd(out + 1, rz); }
}

// p = dinv r + beta p
__global__ void k_cg_update_p(double beta, const double* __restrict__ dinv, const double* __restrict__ r, double* __restrict__ p,
                              int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        p[i] = dinv[i] * r[i] + beta * p[i];
}

// Eigen 3.2 ConjugateGradient.h conjugate_gradient(): returns iterations, *error = sqrt(|r|^2 / |b|^2); rhs <- x
int64_t solve_cg(isl_engine* h, double tol, int64_t max_iter, double* error) {
    const int64_t n = h->n_eqn;
    ISL_REQUIRE(n > 0 && h->nnz > 0 && h->val.p, "no assembled system");
    if (tol <= 0.) tol = 2.220446049250313e-16;
    if (max_iter <= 0) max_iter = 2 * n;
    DevBuf<double> x, r, p, Ap, dinv, sc;
    x.alloc(n); r.alloc(n); p.alloc(n); Ap.alloc(n); dinv.alloc(n); sc.alloc(2);
    const int g = h->grid_for(n, 256), gw = h->grid_for(n * 16, 256), gw8 = h->grid_for(n * 4, 256);
    auto scalars = [&](double* out, int cnt) {
        ISL_CUDA(cudaMemcpyAsync(out, sc.p, cnt * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        ISL_CUDA(cudaStreamSynchronize(h->stream));
    };
    ISL_CUDA(cudaMemsetAsync(x.p, 0, n * sizeof(double), h->stream));
    ISL_CUDA(cudaMemcpyAsync(r.p, h->rhs.p, n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));   // residual = b - A 0
    ISL_LAUNCH(h, k_cg_diag_inv, g, 256, 0, h->rowptr.p, h->col.p, h->val.p, n, dinv.p);
    // |b|^2 and r.z with alpha = 0
    ISL_CUDA(cudaMemsetAsync(sc.p, 0, 2 * sizeof(double), h->stream));
    ISL_CUDA(cudaMemsetAsync(Ap.p, 0, n * sizeof(double), h->stream));
    ISL_CUDA(cudaMemsetAsync(p.p, 0, n * sizeof(double), h->stream));
    ISL_LAUNCH(h, k_cg_update_xr, g, 256, 0, 0.0, p.p, Ap.p, dinv.p, x.p, r.p, n, sc.p);
    double s2[2];
    scalars(s2, 2);
    const double rhs_norm2 = s2[0];
    int64_t it = 0;
    double res_norm2 = s2[0], abs_new = s2[1];
    if (rhs_norm2 == 0.) {
        ISL_CUDA(cudaMemsetAsync(h->rhs.p, 0, n * sizeof(double), h->stream));
        ISL_CUDA(cudaStreamSynchronize(h->stream));
        if (error) *error = 0.;
        return 0;
    }
    const double threshold = tol * tol * rhs_norm2;
    if (res_norm2 >= threshold) {
        ISL_LAUNCH(h, k_cg_update_p, g, 256, 0, 0.0, dinv.p, r.p, p.p, n);   // p = z
        while (it < max_iter) {
            ISL_CUDA(cudaMemsetAsync(sc.p, 0, 2 * sizeof(double), h->stream));
            if (h->nnz / n <= 48) ISL_LAUNCH(h, k_cg_spmv_dot<8>, gw8, 256, 0, h->rowptr.p, h->col.p, h->val.p, p.p, Ap.p, n, sc.p);
            else ISL_LAUNCH(h, k_cg_spmv_dot<32>, gw, 256, 0, h->rowptr.p, h->col.p, h->val.p, p.p, Ap.p, n, sc.p);
            double pAp;
            scalars(&pAp, 1);
            ISL_REQUIRE(pAp != 0. && pAp == pAp, "conjugate gradients broke down (p.Ap = 0 or NaN): matrix not s.p.d.?");
            const double alpha = abs_new / pAp;
            ISL_CUDA(cudaMemsetAsync(sc.p, 0, 2 * sizeof(double), h->stream));
            ISL_LAUNCH(h, k_cg_update_xr, g, 256, 0, alpha, p.p, Ap.p, dinv.p, x.p, r.p, n, sc.p);
            scalars(s2, 2);
            res_norm2 = s2[0];
            if (res_norm2 < threshold) break;
            const double abs_old = abs_new;
            abs_new = s2[1];
            ISL_LAUNCH(h, k_cg_update_p, g, 256, 0, abs_new / abs_old, dinv.p, r.p, p.p, n);
            it++;
        }
    }
    ISL_CUDA(cudaMemcpyAsync(h->rhs.p, x.p, n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    ISL_CUDA(cudaStreamSynchronize(h->stream));
    if (error) *error = std::sqrt(res_norm2 / rhs_norm2);
    return it;
}
