// Test infrastructure — NOT product code.
//
// Put BEFORE the reference root on the include path, this file shadows the reference's base/solver/Eigen3.hpp so that
// the reference's applications compile WITHOUT ANY SOURCE CHANGE against the B200 binding
// (include/insilico_b200_reference.hpp): `base::solver::Eigen3` becomes `base::solver::B200`.  The linear solves
// (outside the assembly path) are served on the host by the Eigen stand-in of oracle/compat through the binding's
// solve hook.  A maintainer would instead switch the `typedef base::solver::Eigen3 Solver;` line of an application.
#ifndef base_solver_eigen3_hpp
#define base_solver_eigen3_hpp

#include <cstdlib>
#include <Eigen/Sparse>
#include <base/io/Format.hpp>
#include <insilico_b200_reference.hpp>

namespace base {
namespace solver {

typedef B200 Eigen3;

namespace shadow_detail {
inline int hostSolve(const char* method, std::size_t n, const std::vector<int64_t>& rowptr, const std::vector<int32_t>& col,
                     const std::vector<double>& val, std::vector<double>& rhs) {
    typedef Eigen::SparseMatrix<double> SM;
    std::vector<Eigen::Triplet<double> > trip;
    trip.reserve(val.size());
    for (std::size_t i = 0; i < n; i++)
        for (int64_t k = rowptr[i]; k < rowptr[i + 1]; k++) trip.push_back(Eigen::Triplet<double>(static_cast<int>(i), col[k], val[k]));
    SM A(static_cast<int>(n), static_cast<int>(n));
    A.setFromTriplets(trip.begin(), trip.end());
    Eigen::VectorXd b(static_cast<Eigen::DenseIndex>(n));
    for (std::size_t i = 0; i < n; i++) b[i] = rhs[i];
    Eigen::VectorXd x;
    int iterations = 1;
    const std::string m(method);
    if (m == "cholesky") {
        Eigen::SimplicialLDLT<SM> chol(A);
        x = chol.solve(b);
    } else if (m == "cg") {
        Eigen::ConjugateGradient<SM> cg;
        cg.compute(A);
        x = cg.solve(b);
        iterations = cg.iterations();
    } else {
        Eigen::SparseLU<SM> lu;
        lu.compute(A);
        x = lu.solve(b);
    }
    for (std::size_t i = 0; i < n; i++) rhs[i] = x[i];
    return iterations;
}
struct InstallHook {
    InstallHook() {
        B200::solveHook() = &hostSolve;
        B200::nativeCG() = (std::getenv("ISL_NATIVE_CG") != NULL);   // default: the host stand-in, like the CPU reference run
        b200_detail::rescanOncePerSolver() = (std::getenv("ISL_RESCAN_PER_SOLVER") != NULL);
        b200_detail::rescanEveryCall() = (std::getenv("ISL_RESCAN_EVERY_CALL") != NULL);
    }
};
static InstallHook installHook;
}  // namespace shadow_detail

}  // namespace solver
}  // namespace base
#endif
