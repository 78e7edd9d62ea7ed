// Test infrastructure — NOT product code.
//
// Drives the UNMODIFIED inSilico headers under /root/reference (compiled against the std-only Boost/Eigen stand-ins
// of oracle/compat) through the reference's own public API for the assembly hot path and dumps what the parity
// tests compare: DoF numbering (element -> DoF ids, status, equation numbers) and the assembled system of
// base::solver::Eigen3 (finishAssembly -> debugLHS / debugRHS, reference base/solver/Eigen3.hpp:140-153,307-322).
// The call sequence per case is the reference applications':
//   reference/04-heat/dirichlet.cpp:99-167, reference/06-elastic/compressible.cpp:163-288,
//   reference/07-drivenCavity/drivenCavity.cpp:176-275.
// Built only where /root/reference exists (oracle/Makefile target `ref`), output into oracle/_ref/.
//
// usage: ref_driver <job file>          (see tools/make_ref_goldens.py for the job format)
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include <base/shape.hpp>
#include <base/Unstructured.hpp>
#include <base/mesh/MeshBoundary.hpp>
#include <base/Quadrature.hpp>
#include <base/io/smf/Reader.hpp>
#include <base/fe/Basis.hpp>
#include <base/Field.hpp>
#include <base/dof/numbering.hpp>
#include <base/dof/generate.hpp>
#include <base/dof/constrainBoundary.hpp>
#include <base/asmb/FieldBinder.hpp>
#include <base/asmb/StiffnessMatrix.hpp>
#include <base/asmb/ForceIntegrator.hpp>
#include <base/asmb/BodyForce.hpp>
#include <base/mesh/generateBoundaryMesh.hpp>
#include <base/asmb/SurfaceFieldBinder.hpp>
#include <base/asmb/NeumannForce.hpp>
#include <base/solver/Eigen3.hpp>
#include <base/kernel/Mass.hpp>
#include <heat/Laplace.hpp>
#include <fluid/Stokes.hpp>
#include <fluid/Convection.hpp>
#include <mat/Lame.hpp>
#include <mat/hypel/StVenant.hpp>
#include <mat/hypel/NeoHookeanCompressible.hpp>
#include <solid/HyperElastic.hpp>

namespace drv {

struct Op {
    std::string what, kernel;  // matrix | residual | body
    int test = 0, trial = 0, incremental = 1;
    std::vector<double> p;
};
struct LinearConstraint {       // base/dof/Constraint.hpp: u(obj,comp) = rhs + sum weight * u(master obj, master comp)
    long obj = 0;
    unsigned comp = 0;
    double rhs = 0.;
    std::vector<long> mobj;
    std::vector<unsigned> mcomp;
    std::vector<double> weight;
};
struct FieldSpec {
    std::string presc, values;  // raw f64 files [n_obj][ds] ("-" = none)
    int boundary = 0;           // 1: constrain the whole boundary through dof::constrainBoundary; 2: constrain exactly
                                // the DoF components whose entry in the `status` table (raw f64 [n_obj][ds]) is 1
    std::string status;
    long pin = -1;              // constrainValue(0, 0.) on this DoF object
    std::vector<LinearConstraint> linear;
};
struct Job {
    std::string type, mesh, out;
    int registerFields = 0, repeat = 1, dump = 1;
    std::map<int, FieldSpec> fields;
    std::vector<Op> ops;
};

static std::vector<double> readF64(const std::string& f) {
    std::vector<double> v;
    if (f == "-" || f.empty()) return v;
    std::ifstream in(f.c_str(), std::ios::binary);
    VERIFY_MSG(in.is_open(), "cannot open " + f);
    in.seekg(0, std::ios::end);
    const std::size_t n = static_cast<std::size_t>(in.tellg()) / sizeof(double);
    in.seekg(0);
    v.resize(n);
    in.read(reinterpret_cast<char*>(v.data()), static_cast<std::streamsize>(n * sizeof(double)));
    return v;
}

static Job readJob(const char* file) {
    Job j;
    std::ifstream in(file);
    VERIFY_MSG(in.is_open(), "cannot open job file");
    std::string line;
    while (std::getline(in, line)) {
        std::istringstream s(line);
        std::string key;
        if (!(s >> key) || key[0] == '#') continue;
        if (key == "type") s >> j.type;
        else if (key == "mesh") s >> j.mesh;
        else if (key == "out") s >> j.out;
        else if (key == "register") s >> j.registerFields;
        else if (key == "repeat") s >> j.repeat;
        else if (key == "dump") s >> j.dump;
        else if (key == "field") {
            int i;
            FieldSpec f;
            s >> i >> f.boundary >> f.pin >> f.presc >> f.values;
            if (f.boundary == 2) s >> f.status;
            j.fields[i] = f;
        } else if (key == "constraint") {
            int i, nm;
            LinearConstraint c;
            s >> i >> c.obj >> c.comp >> c.rhs >> nm;
            c.mobj.resize(nm); c.mcomp.resize(nm); c.weight.resize(nm);
            for (int k = 0; k < nm; k++) s >> c.mobj[k] >> c.mcomp[k] >> c.weight[k];
            j.fields[i].linear.push_back(c);
        } else if (key == "op") {
            Op o;
            s >> o.what >> o.kernel >> o.test >> o.trial >> o.incremental;
            double v;
            while (s >> v) o.p.push_back(v);
            j.ops.push_back(o);
        }
    }
    return j;
}

// Dirichlet callback handed to dof::constrainBoundary: the value comes from a table indexed by the DoF id, so that
// WHICH DoFs get constrained is decided by the reference's boundary extraction alone.
template <typename DOF, typename VEC>
void prescribeFromTable(const VEC&, DOF* doF, const std::vector<double>* table) {
    const std::size_t id = doF->getID();
    for (unsigned d = 0; d < DOF::size; d++)
        if (doF->isActive(d)) doF->constrainValue(d, (*table)[id * DOF::size + d]);
}

template <unsigned DS>
struct ConstantForce {
    typedef typename base::Vector<DS>::Type result_type;
    result_type f;
    template <typename X>
    result_type operator()(const X&) const { return f; }
};

// conductivity function of the *_kappafun cases (tests/flows.py kappa_fun), evaluated by the reference at every
// quadrature point through heat::Laplace::setConductivityFunction (heat/Laplace.hpp:85-97)
template <typename GEOMELEMENT>
struct NamedConductivity {
    typedef typename base::GeomTraits<GEOMELEMENT>::LocalVecDim LocalVecDim;
    double operator()(const GEOMELEMENT* gep, const LocalVecDim& xi) const {
        const typename base::GeomTraits<GEOMELEMENT>::GlobalVecDim x = base::Geometry<GEOMELEMENT>()(gep, xi);
        const unsigned dim = base::GeomTraits<GEOMELEMENT>::globalDim;
        return 1.0 + x[0] * x[0] + 0.5 * std::sin(3.0 * x[1]) * x[dim - 1];
    }
};

// the non-constant body forces of tests/flows.py ("bodyfun" operations), by name
template <unsigned DS, unsigned DIM>
struct NamedForce {
    typedef typename base::Vector<DS>::Type result_type;
    std::string name;
    result_type operator()(const typename base::Vector<DIM>::Type& x) const {
        result_type f = base::constantVector<DS>(0.);
        const double X = x[0], Y = x[1], Z = (DIM > 2 ? x[DIM - 1] : 0.);
        if (name == "hex_scalar") f[0] = std::exp(X) * std::sin(2.0 * Y) + Z * Z;
        else if (name == "tri_scalar") f[0] = std::sin(3.0 * X) * (1.0 + Y * Y);
        else if (name == "hex_vector") { f[0] = Y * Z; f[DS > 1 ? 1 : 0] = std::cos(X); f[DS - 1] = 1.0 + X * Y * Z; }
        else VERIFY_MSG(false, "unknown force function " + name);
        return f;
    }
};

// the surface forces f(x, normal) of the "neumann" operations of tests/flows.py, by name
template <unsigned DS, unsigned DIM>
struct NamedSurfaceForce {
    typedef typename base::Vector<DS>::Type result_type;
    std::string name;
    std::vector<double> p;
    result_type operator()(const typename base::Vector<DIM>::Type& x, const typename base::Vector<DIM>::Type& n) const {
        result_type f = base::constantVector<DS>(0.);
        if (name == "constant") { for (unsigned d = 0; d < DS; d++) f[d] = p[d]; }
        else if (name == "pressure") { for (unsigned d = 0; d < DS; d++) f[d] = p[0] * n[d % DIM]; }
        else if (name == "fun") {
            const double s = 1.0 + x[0] * x[1] + 0.5 * std::sin(2.0 * x[DIM - 1]);
            for (unsigned d = 0; d < DS; d++) f[d] = s * n[d % DIM] + 0.25 * x[(d + 1) % DIM];
        } else VERIFY_MSG(false, "unknown surface force " + name);
        return f;
    }
};

// geometry filter of generateBoundaryMesh: 0 = every boundary face, 1 = faces in the plane x0 = 1
struct FaceFilter {
    int id;
    template <typename X>
    bool operator()(const X& x) const { return id == 0 || x[0] > 1.0 - 1e-9; }
};

template <typename MESH, typename FEBASIS, typename FIELD>
void setUpField(const MESH& mesh, FIELD& field, const FieldSpec& spec, const base::mesh::MeshBoundary& boundary) {
    typedef typename FIELD::DegreeOfFreedom DoF;
    base::dof::generate<FEBASIS>(mesh, field);
    const std::vector<double> presc = readF64(spec.presc), values = readF64(spec.values);
    if (spec.boundary == 2) {
        const std::vector<double> st = readF64(spec.status);
        for (typename FIELD::DoFPtrIter it = field.doFsBegin(); it != field.doFsEnd(); ++it)
            for (unsigned d = 0; d < DoF::size; d++)
                if (st[(*it)->getID() * DoF::size + d] == 1.0) (*it)->constrainValue(d, presc[(*it)->getID() * DoF::size + d]);
    } else if (spec.boundary) {
        typedef typename base::Vector<MESH::Node::dim>::Type VecDim;
        base::dof::constrainBoundary<FEBASIS>(boundary.begin(), boundary.end(), mesh, field,
                                              boost::bind(&prescribeFromTable<DoF, VecDim>, _1, _2, &presc));
    }
    if (spec.pin >= 0) {
        typename FIELD::DoFPtrIter it = field.doFsBegin();
        std::advance(it, spec.pin);
        (*it)->constrainValue(0, 0.0);
    }
    for (const LinearConstraint& c : spec.linear) {
        typename FIELD::DoFPtrIter slave = field.doFsBegin();
        std::advance(slave, c.obj);
        (*slave)->makeConstraint(c.comp);
        (*slave)->getConstraint(c.comp)->setValue(c.rhs);
        for (std::size_t k = 0; k < c.mobj.size(); k++) {
            typename FIELD::DoFPtrIter master = field.doFsBegin();
            std::advance(master, c.mobj[k]);
            (*slave)->getConstraint(c.comp)->addWeightedDoF(*master, c.mcomp[k], c.weight[k]);
        }
    }
    if (!values.empty()) {
        for (typename FIELD::DoFPtrIter it = field.doFsBegin(); it != field.doFsEnd(); ++it)
            for (unsigned d = 0; d < DoF::size; d++) (*it)->setValue(d, values[(*it)->getID() * DoF::size + d]);
    }
}

template <typename FIELD>
void dumpField(const FIELD& field, const std::string& prefix) {
    typedef typename FIELD::DegreeOfFreedom DoF;
    {
        std::ofstream o((prefix + ".elemdof.txt").c_str());
        for (typename FIELD::ElementPtrConstIter e = field.elementsBegin(); e != field.elementsEnd(); ++e) {
            for (typename FIELD::Element::DoFPtrConstIter d = (*e)->doFsBegin(); d != (*e)->doFsEnd(); ++d)
                o << (*d)->getID() << " ";
            o << "\n";
        }
    }
    {
        std::ofstream o((prefix + ".dofs.txt").c_str());
        o << std::setprecision(17);
        for (typename FIELD::DoFPtrConstIter d = field.doFsBegin(); d != field.doFsEnd(); ++d) {
            o << (*d)->getID();
            for (unsigned c = 0; c < DoF::size; c++) {
                const bool act = (*d)->isActive(c), con = (*d)->isConstrained(c);
                std::vector<double> pv(DoF::size);
                o << " " << (act ? 0 : (con ? 1 : 2)) << " " << (act ? static_cast<long>((*d)->getIndex(c)) : -1L);
            }
            o << "\n";
        }
    }
}

template <typename SOLVER>
void dumpSystem(const SOLVER& solver, const std::string& prefix) {
    std::ofstream a((prefix + ".lhs.txt").c_str());
    a << std::setprecision(17);
    solver.debugLHS(a);
    std::ofstream b((prefix + ".rhs.txt").c_str());
    b << std::setprecision(17);
    solver.debugRHS(b);
}

static double now() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

enum Kind { SCALAR, VECTOR, SOLID };

// ---- one field (Bubnov) cases: scalar Laplace, vector Laplace, hyperelastic solids --------------------------------
template <base::Shape SHAPE, unsigned FDEG, unsigned DS, unsigned QDEG, unsigned QDEGBODY, Kind KIND>
int runSingle(const Job& job) {
    typedef base::Unstructured<SHAPE, 1> Mesh;
    typedef base::fe::Basis<SHAPE, FDEG> FEBasis;
    typedef base::Field<FEBasis, DS> Field;
    typedef base::asmb::FieldBinder<Mesh, Field> FieldBinder;
    typedef typename FieldBinder::template TupleBinder<1, 1>::Type FTB;
    typedef base::Quadrature<QDEG, SHAPE> Quadrature;
    typedef base::Quadrature<QDEGBODY, SHAPE> QuadratureBody;
    typedef base::solver::Eigen3 Solver;

    Mesh mesh;
    {
        std::ifstream smf(job.mesh.c_str());
        VERIFY_MSG(smf.is_open(), "cannot open mesh");
        base::io::smf::readMesh(smf, mesh);
    }
    base::mesh::MeshBoundary boundary;
    boundary.create(mesh.elementsBegin(), mesh.elementsEnd());
    Field field;
    setUpField<Mesh, FEBasis>(mesh, field, job.fields.at(0), boundary);
    const std::size_t numDoFs = base::dof::numberDoFsConsecutively(field.doFsBegin(), field.doFsEnd());
    if (job.dump) dumpField(field, job.out + ".f0");
    if (job.dump == 2) return 0;   // numbering fixtures of large meshes: no assembly

    FieldBinder fieldBinder(mesh, field);
    Quadrature quadrature;
    QuadratureBody quadratureBody;

    double best = 1e300;
    for (int rep = 0; rep < job.repeat; rep++) {
        const double t0 = now();
        Solver solver(numDoFs);
        if (job.registerFields) solver.template registerFields<FTB>(fieldBinder);
        const double t1 = now();
        for (const Op& op : job.ops) {
            if (op.what == "bodyfun") {
                NamedForce<DS, Mesh::Node::dim> f;
                f.name = op.kernel;
                base::asmb::bodyForceComputation<FTB>(quadratureBody, solver, fieldBinder, f);
                continue;
            }
            if (op.what == "neumann") {   // base::asmb::neumannForceComputation over (a part of) the boundary
                typedef typename base::mesh::BoundaryMeshBinder<typename Mesh::Element>::Type BoundaryMesh;
                BoundaryMesh boundaryMesh;
                FaceFilter filter;
                filter.id = op.incremental;
                base::mesh::generateBoundaryMesh(boundary.begin(), boundary.end(), mesh, boundaryMesh, filter);
                typedef base::asmb::SurfaceFieldBinder<BoundaryMesh, Field> SurfaceFieldBinder;
                SurfaceFieldBinder surfaceFieldBinder(boundaryMesh, field);
                typedef typename SurfaceFieldBinder::template TupleBinder<1>::Type SFTB;
                base::SurfaceQuadrature<QDEGBODY, SHAPE> surfaceQuadrature;
                NamedSurfaceForce<DS, Mesh::Node::dim> f;
                f.name = op.kernel;
                f.p = op.p;
                base::asmb::neumannForceComputation<SFTB>(surfaceQuadrature, solver, surfaceFieldBinder, f);
                continue;
            }
            if (op.what == "body") {
                ConstantForce<DS> f;
                for (unsigned d = 0; d < DS; d++) f.f[d] = op.p[d];
                base::asmb::bodyForceComputation<FTB>(quadratureBody, solver, fieldBinder, f);
                continue;
            }
            if (op.kernel == "mass") {   // base::kernel::Mass (base/kernel/Mass.hpp), matrix only
                base::kernel::Mass<typename FTB::Tuple> kernel(op.p[0]);
                base::asmb::stiffnessMatrixComputation<FTB>(quadrature, solver, fieldBinder, kernel, op.incremental != 0);
                continue;
            }
            if (op.kernel == "convection") {   // fluid::Convection on the tuple (test, trial, advection velocity) = (u, u, u)
                if constexpr (KIND == VECTOR) {
                    typedef typename FieldBinder::template TupleBinder<1, 1, 1>::Type UUU;
                    fluid::Convection<typename UUU::Tuple> kernel(op.p[0]);
                    if (op.what == "matrix")
                        base::asmb::stiffnessMatrixComputation<UUU>(quadrature, solver, fieldBinder, kernel, op.incremental != 0);
                    else
                        base::asmb::computeResidualForces<UUU>(quadrature, solver, fieldBinder, kernel);
                }
                continue;
            }
            if (op.what == "matrixfun") {   // heat::Laplace with a conductivity function
                if constexpr (KIND == SCALAR) {
                    typedef heat::Laplace<typename FTB::Tuple> Kernel;
                    Kernel kernel(1.0);
                    typename Kernel::ConductivityFun cf = NamedConductivity<typename Mesh::Element>();
                    kernel.setConductivityFunction(cf);
                    base::asmb::stiffnessMatrixComputation<FTB>(quadrature, solver, fieldBinder, kernel, op.incremental != 0);
                }
                continue;
            }
            if constexpr (KIND == SCALAR) {
                heat::Laplace<typename FTB::Tuple> kernel(op.p[0]);
                if (op.what == "matrix")
                    base::asmb::stiffnessMatrixComputation<FTB>(quadrature, solver, fieldBinder, kernel, op.incremental != 0);
                else
                    base::asmb::computeResidualForces<FTB>(quadrature, solver, fieldBinder, kernel);
            } else if constexpr (KIND == VECTOR) {
                fluid::VectorLaplace<typename FTB::Tuple> kernel(op.p[0]);
                if (op.what == "matrix")
                    base::asmb::stiffnessMatrixComputation<FTB>(quadrature, solver, fieldBinder, kernel, op.incremental != 0);
                else
                    base::asmb::computeResidualForces<FTB>(quadrature, solver, fieldBinder, kernel);
            } else {
                if (op.kernel == "stvenant") {
                    typedef mat::hypel::StVenant Material;
                    Material material(op.p[0], op.p[1]);
                    solid::HyperElastic<Material, typename FTB::Tuple> kernel(material);
                    if (op.what == "matrix")
                        base::asmb::stiffnessMatrixComputation<FTB>(quadrature, solver, fieldBinder, kernel, op.incremental != 0);
                    else
                        base::asmb::computeResidualForces<FTB>(quadrature, solver, fieldBinder, kernel);
                } else {
                    typedef mat::hypel::NeoHookeanCompressible Material;
                    Material material(op.p[0], op.p[1]);
                    solid::HyperElastic<Material, typename FTB::Tuple> kernel(material);
                    if (op.what == "matrix")
                        base::asmb::stiffnessMatrixComputation<FTB>(quadrature, solver, fieldBinder, kernel, op.incremental != 0);
                    else
                        base::asmb::computeResidualForces<FTB>(quadrature, solver, fieldBinder, kernel);
                }
            }
        }
        const double t2 = now();
        solver.finishAssembly();
        const double t3 = now();
        best = std::min(best, t2 - t1);
        std::printf("rep %d  elements %zu  dofs %zu  register %.6f s  assemble %.6f s  finish %.6f s\n", rep,
                    static_cast<std::size_t>(std::distance(mesh.elementsBegin(), mesh.elementsEnd())), numDoFs, t1 - t0,
                    t2 - t1, t3 - t2);
        if (job.dump == 1 && rep == 0) dumpSystem(solver, job.out);   // dump 2: numbering only
    }
    std::printf("best_assemble_seconds %.9f\n", best);
    return 0;
}

// ---- Taylor-Hood Stokes blocks (velocity FDEG x dim, pressure FDEG-1 x 1) -----------------------------------------
template <base::Shape SHAPE, unsigned FDEG, unsigned QDEG>
int runStokes(const Job& job) {
    typedef base::Unstructured<SHAPE, 1> Mesh;
    static const unsigned dim = Mesh::Node::dim;
    typedef base::fe::Basis<SHAPE, FDEG> FEBasisU;
    typedef base::fe::Basis<SHAPE, FDEG - 1> FEBasisP;
    typedef base::Field<FEBasisU, dim> Velocity;
    typedef base::Field<FEBasisP, 1> Pressure;
    typedef base::asmb::FieldBinder<Mesh, Velocity, Pressure> Binder;
    typedef typename Binder::template TupleBinder<1, 1>::Type UU;
    typedef typename Binder::template TupleBinder<1, 2>::Type UP;
    typedef typename Binder::template TupleBinder<2, 1>::Type PU;
    typedef base::Quadrature<QDEG, SHAPE> Quadrature;
    typedef base::solver::Eigen3 Solver;

    Mesh mesh;
    {
        std::ifstream smf(job.mesh.c_str());
        VERIFY_MSG(smf.is_open(), "cannot open mesh");
        base::io::smf::readMesh(smf, mesh);
    }
    base::mesh::MeshBoundary boundary;
    boundary.create(mesh.elementsBegin(), mesh.elementsEnd());
    Velocity velocity;
    Pressure pressure;
    setUpField<Mesh, FEBasisU>(mesh, velocity, job.fields.at(0), boundary);
    setUpField<Mesh, FEBasisP>(mesh, pressure, job.fields.at(1), boundary);
    const std::size_t nU = base::dof::numberDoFsConsecutively(velocity.doFsBegin(), velocity.doFsEnd());
    const std::size_t nP = base::dof::numberDoFsConsecutively(pressure.doFsBegin(), pressure.doFsEnd(), nU);
    if (job.dump) {
        dumpField(velocity, job.out + ".f0");
        dumpField(pressure, job.out + ".f1");
    }
    if (job.dump == 2) return 0;
    Binder binder(mesh, velocity, pressure);
    Quadrature quadrature;
    double best = 1e300;
    auto registerAll = [&](Solver& solver) {
        if (!job.registerFields) return;
        solver.template registerFields<UU>(binder);
        solver.template registerFields<UP>(binder);
        solver.template registerFields<PU>(binder);
    };
    auto applyOp = [&](const Op& op, Solver& solver) {
        const bool matrix = (op.what == "matrix"), incr = (op.incremental != 0);
        if (op.kernel == "vector_laplace") {
            fluid::VectorLaplace<typename UU::Tuple> k(op.p[0]);
            if (matrix) base::asmb::stiffnessMatrixComputation<UU>(quadrature, solver, binder, k, incr);
            else base::asmb::computeResidualForces<UU>(quadrature, solver, binder, k);
        } else if (op.kernel == "pressure_gradient") {
            fluid::PressureGradient<typename UP::Tuple> k;
            if (matrix) base::asmb::stiffnessMatrixComputation<UP>(quadrature, solver, binder, k, incr);
            else base::asmb::computeResidualForces<UP>(quadrature, solver, binder, k);
        } else if (op.kernel == "velocity_divergence") {
            fluid::VelocityDivergence<typename PU::Tuple> k(!op.p.empty() && op.p[0] != 0.0);
            if (matrix) base::asmb::stiffnessMatrixComputation<PU>(quadrature, solver, binder, k, incr);
            else base::asmb::computeResidualForces<PU>(quadrature, solver, binder, k);
        } else {
            VERIFY_MSG(false, "unknown kernel " + op.kernel);
        }
    };
#ifdef INSILICO_B200_REFERENCE_HPP
    // B200 binding only (ISL_DRIVER_DEVICES=2): ONE host thread drives two devices, the calls of two solvers interleaved
    // call by call (base::solver::B200::selectDevice).  Both systems are dumped: <out> and <out>.dev1
    if (const char* nd = std::getenv("ISL_DRIVER_DEVICES")) {
        if (std::atoi(nd) == 2) {
            Solver::selectDevice(0);
            Solver first(nU + nP);
            registerAll(first);
            Solver::selectDevice(1);
            Solver second(nU + nP);
            registerAll(second);
            for (const Op& op : job.ops) {
                Solver::selectDevice(0);
                applyOp(op, first);
                Solver::selectDevice(1);
                applyOp(op, second);
            }
            Solver::selectDevice(0);
            first.finishAssembly();
            if (job.dump == 1) dumpSystem(first, job.out);
            Solver::selectDevice(1);
            second.finishAssembly();
            if (job.dump == 1) dumpSystem(second, job.out + ".dev1");
            std::printf("two devices, one host thread: solvers on devices %d and %d\n", first.device(), second.device());
            return 0;
        }
    }
#endif
    for (int rep = 0; rep < job.repeat; rep++) {
        Solver solver(nU + nP);
        registerAll(solver);
        const double t1 = now();
        for (const Op& op : job.ops) applyOp(op, solver);
        const double t2 = now();
        solver.finishAssembly();
        best = std::min(best, t2 - t1);
        std::printf("rep %d  elements %zu  dofs %zu  assemble %.6f s\n", rep,
                    static_cast<std::size_t>(std::distance(mesh.elementsBegin(), mesh.elementsEnd())), nU + nP, t2 - t1);
        if (job.dump == 1 && rep == 0) dumpSystem(solver, job.out);   // dump 2: numbering only
    }
    std::printf("best_assemble_seconds %.9f\n", best);
    return 0;
}

}  // namespace drv

int main(int argc, char* argv[]) {
    if (argc != 2) {
        std::cerr << "usage: " << argv[0] << " job.txt\n";
        return 2;
    }
    using namespace drv;
    const Job job = readJob(argv[1]);
    const std::string& t = job.type;
    //                                   shape      fdeg ds qdeg qbody kind
    if (t == "laplace_q1_hex") return runSingle<base::HEX, 1, 1, 3, 3, SCALAR>(job);
    if (t == "laplace_q2_hex") return runSingle<base::HEX, 2, 1, 4, 4, SCALAR>(job);
    if (t == "laplace_p1_tet") return runSingle<base::TET, 1, 1, 3, 2, SCALAR>(job);
    if (t == "laplace_p2_tet") return runSingle<base::TET, 2, 1, 4, 4, SCALAR>(job);
    if (t == "laplace_q1_quad") return runSingle<base::QUAD, 1, 1, 3, 3, SCALAR>(job);
    if (t == "laplace_p2_tri") return runSingle<base::TRI, 2, 1, 4, 4, SCALAR>(job);
    if (t == "vector_laplace_q1_hex") return runSingle<base::HEX, 1, 3, 3, 3, VECTOR>(job);
    if (t == "vector_laplace_q2_quad") return runSingle<base::QUAD, 2, 2, 4, 4, VECTOR>(job);
    if (t == "solid_q1_hex") return runSingle<base::HEX, 1, 3, 3, 3, SOLID>(job);
    if (t == "solid_q2_hex") return runSingle<base::HEX, 2, 3, 4, 4, SOLID>(job);
    if (t == "solid_q1_quad") return runSingle<base::QUAD, 1, 2, 3, 3, SOLID>(job);
    if (t == "solid_p2_tet") return runSingle<base::TET, 2, 3, 4, 4, SOLID>(job);
    if (t == "solid_q2_quad") return runSingle<base::QUAD, 2, 2, 3, 3, SOLID>(job);
    if (t == "laplace_p1_tri") return runSingle<base::TRI, 1, 1, 2, 2, SCALAR>(job);
    if (t == "stokes_p2p1_tet") return runStokes<base::TET, 2, 4>(job);
    if (t == "stokes_q2q1_hex") return runStokes<base::HEX, 2, 4>(job);
    if (t == "stokes_q2q1_quad") return runStokes<base::QUAD, 2, 4>(job);
    std::cerr << "unknown case type '" << t << "'\n";
    return 2;
}
