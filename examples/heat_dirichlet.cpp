// Poisson problem with Dirichlet data from the Laplace fundamental solution on a Q1 hex unit cube: the flow of the
// reference application reference/04-heat/dirichlet.cpp:79-167, written against include/insilico_b200.hpp
// (same calls: dof::generate, MeshBoundary::create, dof::constrainBoundary, numberDoFsConsecutively, Solver(n),
// FieldBinder / TupleBinder<1,1>, heat::Laplace, asmb::stiffnessMatrixComputation<FTB>, finishAssembly).
//
// usage: heat_dirichlet N   -> prints "#dofs nnz norm(rhs) sum(val)" for the N^3 mesh
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <numeric>

#include "insilico_b200.hpp"

static const base::Shape shape = base::HEX;
typedef base::Unstructured<shape, 1> Mesh;
typedef base::fe::Basis<shape, 1> FEBasis;
typedef base::Field<FEBasis, 1> Field;

// unitCube recipe (tools/meshGeneration/unitCube/unitCube.hpp:85-265): nodes x-fastest, hierarchic connectivity
static void unitCube(int n, std::vector<double>& X, std::vector<int32_t>& conn) {
    const int n1 = n + 1;
    const double h = 1.0 / n;
    for (int k = 0; k < n1; k++) for (int j = 0; j < n1; j++) for (int i = 0; i < n1; i++) { X.push_back(h * i); X.push_back(h * j); X.push_back(h * k); }
    const int H[8] = {0, 1, 3, 2, 4, 5, 7, 6};
    for (int k = 0; k < n; k++) for (int j = 0; j < n; j++) for (int i = 0; i < n; i++) {
        const int b = i + j * n1 + k * n1 * n1;
        const int lex[8] = {b, b + 1, b + n1, b + n1 + 1, b + n1 * n1, b + n1 * n1 + 1, b + n1 * n1 + n1, b + n1 * n1 + n1 + 1};
        int32_t e[8];
        for (int v = 0; v < 8; v++) e[H[v]] = lex[v];
        conn.insert(conn.end(), e, e + 8);
    }
}

// base::auxi::FundSolLaplace<3>::fun with source point (-.5,-.5,-.5) (auxi/FundamentalSolution.hpp:110-130)
static void dirichlet(const Mesh::Node::VecDim& x, Field::DegreeOfFreedom* doF) {
    const double d = std::sqrt((x[0] + .5) * (x[0] + .5) + (x[1] + .5) * (x[1] + .5) + (x[2] + .5) * (x[2] + .5));
    doF->constrainValue(0, 1. / (4. * M_PI) / d);
}

int main(int argc, char** argv) {
    const int n = argc > 1 ? std::atoi(argv[1]) : 8;
    Mesh mesh;
    {
        std::vector<double> X; std::vector<int32_t> conn;
        unitCube(n, X, conn);
        mesh.set(X, conn);
    }
    base::Quadrature<3, shape> quadrature;
    Field field;
    base::dof::generate<FEBasis>(mesh, field);
    base::mesh::MeshBoundary meshBoundary;
    meshBoundary.create(mesh);
    base::dof::constrainBoundary<FEBasis>(meshBoundary.begin(), meshBoundary.end(), mesh, field, dirichlet);
    const std::size_t numDofs = base::dof::numberDoFsConsecutively(field);

    typedef base::solver::B200 Solver;
    Solver solver(numDofs);
    typedef base::asmb::FieldBinder<Mesh, Field> FieldBinder;
    FieldBinder fieldBinder(mesh, field);
    fieldBinder.upload();
    typedef FieldBinder::TupleBinder<1, 1>::Type FTB;

    typedef heat::Laplace<FTB::Tuple> Laplace;
    Laplace laplace(1.0);
    base::asmb::stiffnessMatrixComputation<FTB>(quadrature, solver, fieldBinder, laplace);
    solver.finishAssembly();

    std::vector<int64_t> rowptr; std::vector<int32_t> col; std::vector<double> val, rhs;
    solver.getCSR(rowptr, col, val, rhs);
    const double sum = std::accumulate(val.begin(), val.end(), 0.0);
    std::printf("%zu %lld %.15e %.15e\n", numDofs, (long long)solver.nonZeros(), solver.norm(), sum);
    return 0;
}
