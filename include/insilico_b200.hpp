// =============================================================================
// insilico_b200.hpp -- header-only C++ facade over the C ABI (insilico_b200.h) that keeps the reference's
// template API surface for the assembly hot path, so that an application written against
// inSilico's base::asmb / base::solver / base::dof interfaces compiles against this engine by switching includes.
//
// Same names, argument meaning and error behaviour as the reference (paths relative to the reference root):
//   base::Unstructured<SHAPE,GDEG>            base/Unstructured.hpp:57-65      (flat SoA instead of heap nodes)
//   base::fe::Basis<SHAPE,DEG>                base/fe/Basis.hpp:147-168
//   base::Field<FEBASIS,DOFSIZE>              base/Field.hpp:50-55
//   base::Quadrature<DEG,SHAPE>               base/Quadrature.hpp:113-143
//   base::dof::generate / numberDoFsConsecutively / constrainBoundary
//                                             base/dof/generate.hpp:46-53, numbering.hpp:44-68, constrainBoundary.hpp:49-123
//   base::mesh::MeshBoundary                  base/mesh/MeshBoundary.hpp
//   base::asmb::FieldBinder<MESH,F1..F5>::TupleBinder<I,J>::Type     base/asmb/FieldBinder.hpp:104-188
//   base::asmb::stiffnessMatrixComputation<FTB> / computeResidualForces<FTB> / bodyForceComputation<FTB>
//                                             base/asmb/StiffnessMatrix.hpp:49-87, ForceIntegrator.hpp:37-71, BodyForce.hpp:65-84
//   base::solver::B200 (members of base::solver::Eigen3, base/solver/Eigen3.hpp:71-336; SOLVER is a template
//                                             parameter everywhere in the reference, which is the seam used here)
//   heat::Laplace, fluid::VectorLaplace, fluid::PressureGradient, fluid::VelocityDivergence,
//   solid::HyperElastic<MATERIAL,...>, mat::Lame, mat::hypel::StVenant, mat::hypel::NeoHookeanCompressible
// Kernel objects are recognised at compile time (KernelTraits); an unsupported kernel type is a static_assert --
// there is no CPU fallback.  Errors of the C ABI become the reference's VERIFY_MSG behaviour: message on stderr and
// abort() (base/verify.hpp:139-149).
// =============================================================================
#ifndef INSILICO_B200_HPP
#define INSILICO_B200_HPP

#include <array>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <utility>
#include <vector>

#include "insilico_b200.h"

#define ISL_VERIFY(call)                                                                        \
    do {                                                                                        \
        if ((call) != 0) {                                                                      \
            std::fprintf(stderr, "(EE) %s\n(EE) in %s:%d\n", isl_last_error(), __FILE__, __LINE__); \
            std::abort();                                                                       \
        }                                                                                       \
    } while (0)

namespace base {

typedef double number;
enum Shape { POINT = ISL_POINT, LINE = ISL_LINE, TRI = ISL_TRI, QUAD = ISL_QUAD, TET = ISL_TET, HEX = ISL_HEX };
enum NFace { VERTEX = 0, EDGE, FACE, CELL };

template <Shape S> struct ShapeDim { static const unsigned value = (S == LINE ? 1 : (S == TRI || S == QUAD) ? 2 : 3); };

template <unsigned DIM> struct Vector { typedef std::array<double, DIM> Type; };
template <unsigned DIM> typename Vector<DIM>::Type constantVector(double v) { typename Vector<DIM>::Type r; r.fill(v); return r; }

// one engine per process and device, shared by mesh, fields and solver
class Engine {
public:
    static isl_handle get(int device = 0) {
        static isl_handle h = nullptr;
        if (!h) ISL_VERIFY(isl_engine_create(device, &h));
        return h;
    }
};

// ---- mesh --------------------------------------------------------------------------------------------------
template <Shape SHAPE, unsigned GDEG, unsigned DIM = ShapeDim<SHAPE>::value>
class Unstructured {
public:
    static const Shape shape = SHAPE;
    static const unsigned geomDegree = GDEG;
    static const unsigned dim = DIM;
    struct Node { static const unsigned dim = DIM; typedef typename Vector<DIM>::Type VecDim; };

    //! flat replacement of io::smf::readMesh: coordinates [n][dim], connectivity [n][npe] in hierarchic order
    void set(std::vector<double> coords, std::vector<int32_t> conn) {
        coords_ = std::move(coords); conn_ = std::move(conn);
        npe_ = isl_shape_nfun(SHAPE, GDEG);
        ISL_VERIFY(isl_mesh_set(Engine::get(), SHAPE, GDEG, DIM, (int64_t)numNodes(), coords_.data(), (int64_t)numElements(),
                                conn_.data()));
    }
    std::size_t numNodes() const { return coords_.size() / DIM; }
    std::size_t numElements() const { return npe_ ? conn_.size() / npe_ : 0; }
    const std::vector<double>& coordinates() const { return coords_; }
    const std::vector<int32_t>& connectivity() const { return conn_; }

private:
    std::vector<double> coords_;
    std::vector<int32_t> conn_;
    int npe_ = 0;
};

namespace fe {
template <Shape SHAPE, unsigned DEG> struct Basis { static const Shape shape = SHAPE; static const unsigned degree = DEG; };
}  // namespace fe

namespace dof {
enum DoFStatus { ACTIVE = ISL_ACTIVE, CONSTRAINED = ISL_CONSTRAINED, INACTIVE = ISL_INACTIVE };
}

// ---- field -------------------------------------------------------------------------------------------------
template <typename FEBASIS, unsigned DOFSIZE, unsigned NHIST = 0>
class Field {
public:
    typedef FEBASIS FEBasis;
    static const unsigned dofSize = DOFSIZE;

    //! proxy with the interface of base::dof::DegreeOfFreedom used by Dirichlet functors
    class DegreeOfFreedom {
    public:
        static const unsigned size = DOFSIZE;
        DegreeOfFreedom(Field* f, std::size_t id) : f_(f), id_(id) {}
        std::size_t getID() const { return id_; }
        void constrainValue(unsigned which, number value) {  // DegreeOfFreedom.hpp:232-242
            f_->status_[id_ * DOFSIZE + which] = dof::CONSTRAINED;
            f_->prescribed_[id_ * DOFSIZE + which] = value;
        }
        bool isActive(unsigned which) const { return f_->status_[id_ * DOFSIZE + which] == dof::ACTIVE; }
        bool isConstrained(unsigned which) const { return f_->status_[id_ * DOFSIZE + which] == dof::CONSTRAINED; }
        number getValue(unsigned which) const { return f_->values_[id_ * DOFSIZE + which]; }
        void setValue(unsigned which, number v) { f_->values_[id_ * DOFSIZE + which] = v; }
        std::size_t getIndex(unsigned which) const { return (std::size_t)f_->eqn_[id_ * DOFSIZE + which]; }

    private:
        Field* f_;
        std::size_t id_;
    };

    std::size_t numDoFs() const { return nObj_; }
    DegreeOfFreedom doF(std::size_t id) { return DegreeOfFreedom(this, id); }

    // flat storage (the reference keeps heap DegreeOfFreedom objects, fe/Field.hpp:50-163)
    std::vector<int32_t> elemDof_;
    std::vector<int64_t> eqn_;
    std::vector<uint8_t> status_;
    std::vector<double> prescribed_, values_;
    std::size_t nObj_ = 0;
    int id_ = -1;          //!< engine field slot, assigned by the FieldBinder
    bool uploaded_ = false;

    void allocate(std::size_t nObj) {
        nObj_ = nObj;
        eqn_.assign(nObj * DOFSIZE, -1); status_.assign(nObj * DOFSIZE, dof::ACTIVE);
        prescribed_.assign(nObj * DOFSIZE, 0.); values_.assign(nObj * DOFSIZE, 0.);
    }
    void upload() {
        ISL_VERIFY(isl_field_set(Engine::get(), id_, FEBASIS::degree, DOFSIZE, (int64_t)nObj_, elemDof_.data(), eqn_.data(),
                                 status_.data(), prescribed_.data(), values_.data()));
        uploaded_ = true;
    }
    //! new Newton state / Dirichlet values with unchanged numbering
    void pushValues() { ISL_VERIFY(isl_field_update(Engine::get(), id_, prescribed_.data(), values_.data())); }
};

template <unsigned DEG, Shape SHAPE>
class Quadrature {
public:
    static const unsigned degree = DEG;
    static const Shape shape = SHAPE;
    typedef typename Vector<ShapeDim<SHAPE>::value>::Type VecDim;
    Quadrature() {
        const int n = isl_quadrature(SHAPE, DEG, nullptr, nullptr);
        if (n < 0) ISL_VERIFY(1);
        w_.resize(n); p_.resize((std::size_t)n * ShapeDim<SHAPE>::value);
        isl_quadrature(SHAPE, DEG, w_.data(), p_.data());
    }
    std::size_t numPoints() const { return w_.size(); }
    double weight(std::size_t q) const { return w_[q]; }
    const double* point(std::size_t q) const { return &p_[q * ShapeDim<SHAPE>::value]; }

private:
    std::vector<double> w_, p_;
};

// ---- mesh boundary + DoF handling -----------------------------------------------------------------------------
namespace mesh {
class MeshBoundary {
public:
    typedef std::vector<std::pair<std::size_t, unsigned>> BoundaryElementContainer;
    typedef BoundaryElementContainer::const_iterator BoundConstIter;
    template <typename MESH>
    void create(const MESH& m) {
        int64_t n = 0;
        ISL_VERIFY(isl_mesh_boundary(MESH::shape, MESH::geomDegree, (int64_t)m.numElements(), m.connectivity().data(), nullptr, &n));
        std::vector<int64_t> pairs((std::size_t)n * 2);
        ISL_VERIFY(isl_mesh_boundary(MESH::shape, MESH::geomDegree, (int64_t)m.numElements(), m.connectivity().data(), pairs.data(), &n));
        b_.clear();
        for (int64_t k = 0; k < n; k++) b_.push_back(std::make_pair((std::size_t)pairs[2 * k], (unsigned)pairs[2 * k + 1]));
    }
    BoundConstIter begin() const { return b_.begin(); }
    BoundConstIter end() const { return b_.end(); }

private:
    BoundaryElementContainer b_;
};
}  // namespace mesh

namespace dof {

//! base::dof::generate<FEBasis>(mesh, field)
template <typename FEBASIS, typename MESH, typename FIELD>
void generate(const MESH& m, FIELD& field) {
    const int ndpe = isl_ndpe(MESH::shape, FEBASIS::degree);
    field.elemDof_.assign(m.numElements() * (std::size_t)ndpe, 0);
    int64_t nObj = 0;
    ISL_VERIFY(isl_dof_generate(MESH::shape, MESH::geomDegree, (int64_t)m.numElements(), m.connectivity().data(), FEBASIS::degree,
                                field.elemDof_.data(), &nObj));
    field.allocate((std::size_t)nObj);
}

//! base::dof::constrainBoundary<FEBasis>(first, last, mesh, field, diriFun); diriFun(x, DoF*)
template <typename FEBASIS, typename BITER, typename MESH, typename FIELD, typename DIRIFUN>
void constrainBoundary(BITER first, BITER last, const MESH& m, FIELD& field, DIRIFUN diriFun) {
    std::vector<int64_t> pairs;
    for (BITER it = first; it != last; ++it) { pairs.push_back((int64_t)it->first); pairs.push_back((int64_t)it->second); }
    const int64_t np = (int64_t)pairs.size() / 2;
    int64_t n = 0;
    ISL_VERIFY(isl_boundary_dofs(MESH::shape, MESH::geomDegree, MESH::dim, m.coordinates().data(), (int64_t)m.numElements(),
                                 m.connectivity().data(), FEBASIS::degree, field.elemDof_.data(), np, pairs.data(), nullptr, nullptr, &n));
    std::vector<int32_t> obj((std::size_t)n);
    std::vector<double> x((std::size_t)n * MESH::dim);
    ISL_VERIFY(isl_boundary_dofs(MESH::shape, MESH::geomDegree, MESH::dim, m.coordinates().data(), (int64_t)m.numElements(),
                                 m.connectivity().data(), FEBASIS::degree, field.elemDof_.data(), np, pairs.data(), obj.data(), x.data(), &n));
    for (int64_t k = 0; k < n; k++) {
        typename MESH::Node::VecDim xk;
        for (unsigned d = 0; d < MESH::dim; d++) xk[d] = x[(std::size_t)k * MESH::dim + d];
        typename FIELD::DegreeOfFreedom doFProxy = field.doF((std::size_t)obj[k]);
        diriFun(xk, &doFProxy);
    }
}

//! base::dof::numberDoFsConsecutively(first, last, init) on the whole field
template <typename FIELD>
std::size_t numberDoFsConsecutively(FIELD& field, std::size_t init = 0) {
    int64_t n = 0;
    ISL_VERIFY(isl_number_dofs((int64_t)field.nObj_, FIELD::dofSize, field.status_.data(), (int64_t)init, field.eqn_.data(), &n));
    return (std::size_t)n;
}

//! base::dof::setDoFsFromSolver(solver, field) (base/dof/Distribute.hpp:35-41): ACTIVE <- solution, CONSTRAINED <- prescribed
template <typename FIELD>
void setDoFsFromVector(const std::vector<double>& x, FIELD& field) {
    for (std::size_t k = 0; k < field.status_.size(); k++) {
        if (field.status_[k] == ACTIVE) field.values_[k] = x[(std::size_t)field.eqn_[k]];
        else if (field.status_[k] == CONSTRAINED) field.values_[k] = field.prescribed_[k];
    }
}
}  // namespace dof

// ---- field binder ---------------------------------------------------------------------------------------------
namespace asmb {

struct NoField { static const unsigned dofSize = 0; };

template <typename MESH, typename F1 = NoField, typename F2 = NoField, typename F3 = NoField>
class FieldBinder {
public:
    typedef MESH Mesh;
    FieldBinder(MESH& mesh, F1& f1) : mesh_(mesh), f1_(&f1), f2_(nullptr), f3_(nullptr) { bind(); }
    FieldBinder(MESH& mesh, F1& f1, F2& f2) : mesh_(mesh), f1_(&f1), f2_(&f2), f3_(nullptr) { bind(); }
    FieldBinder(MESH& mesh, F1& f1, F2& f2, F3& f3) : mesh_(mesh), f1_(&f1), f2_(&f2), f3_(&f3) { bind(); }

    //! TupleBinder<I,J>::Type with 1-based field indices: I = test, J = trial (FieldBinder.hpp:132-138)
    template <int I, int J = I>
    struct TupleBinder {
        struct Type {
            static const int test = I - 1, trial = J - 1;
            struct Tuple { static const int test = I - 1, trial = J - 1; };
        };
    };
    //! make sure the bound fields are on the device (after numbering / constraint changes call again)
    void upload() const {
        if (f1_) up(*f1_, 0);
        if (f2_) up(*f2_, 1);
        if (f3_) up(*f3_, 2);
    }

private:
    template <typename F> void up(F& f, int id) const { f.id_ = id; f.upload(); }
    void up(NoField&, int) const {}
    void bind() { if (f1_) setId(*f1_, 0); if (f2_) setId(*f2_, 1); if (f3_) setId(*f3_, 2); }
    template <typename F> void setId(F& f, int id) { f.id_ = id; }
    void setId(NoField&, int) {}
    MESH& mesh_;
    F1* f1_; F2* f2_; F3* f3_;
};
}  // namespace asmb

// ---- solver -------------------------------------------------------------------------------------------------
namespace solver {
//! device-resident replacement of base::solver::Eigen3 (insert / register / finish / getValue / norm)
class B200 {
public:
    explicit B200(std::size_t size) : n_(size) { ISL_VERIFY(isl_system_create(Engine::get(), (int64_t)size)); }

    template <typename MATRIX, typename RDOFS, typename CDOFS>
    void insertToLHS(const MATRIX& matrix, const RDOFS& rowDoFs, const CDOFS& colDoFs) {
        std::vector<double> m(rowDoFs.size() * colDoFs.size());
        std::vector<int64_t> r(rowDoFs.begin(), rowDoFs.end()), c(colDoFs.begin(), colDoFs.end());
        for (std::size_t i = 0; i < r.size(); i++) for (std::size_t j = 0; j < c.size(); j++) m[i * c.size() + j] = matrix(i, j);
        ISL_VERIFY(isl_insert_lhs(Engine::get(), m.data(), r.data(), (int)r.size(), c.data(), (int)c.size()));
    }
    template <typename VECTOR, typename DOFS>
    void insertToRHS(const VECTOR& vector, const DOFS& dofs) {
        std::vector<double> v(dofs.size()); std::vector<int64_t> r(dofs.begin(), dofs.end());
        for (std::size_t i = 0; i < r.size(); i++) v[i] = vector[i];
        ISL_VERIFY(isl_insert_rhs(Engine::get(), v.data(), r.data(), (int)r.size()));
    }
    template <typename FIELDTUPLEBINDER, typename FIELDBINDER>
    void registerFields(const FIELDBINDER&) {
        ISL_VERIFY(isl_pattern_register(Engine::get(), FIELDTUPLEBINDER::test, FIELDTUPLEBINDER::trial));
    }
    void finishAssembly(bool = true) { ISL_VERIFY(isl_finish(Engine::get(), nullptr, &nnz_)); }
    number getValue(std::size_t index) const { double v; ISL_VERIFY(isl_rhs_value(Engine::get(), (int64_t)index, &v)); return v; }
    double norm() const { double v; ISL_VERIFY(isl_rhs_norm(Engine::get(), &v)); return v; }
    void systemInfo(std::ostream& out) const;
    std::size_t size() const { return n_; }
    int64_t nonZeros() const { return nnz_; }
    //! canonical CSR + rhs on the host (hand-off to a host solver)
    void getCSR(std::vector<int64_t>& rowptr, std::vector<int32_t>& col, std::vector<double>& val, std::vector<double>& rhs) {
        finishAssembly();
        rowptr.resize(n_ + 1); col.resize((std::size_t)nnz_); val.resize((std::size_t)nnz_); rhs.resize(n_);
        ISL_VERIFY(isl_get_csr(Engine::get(), rowptr.data(), col.data(), val.data(), rhs.data()));
    }
    //! zero-copy hand-off to a device solver
    void getDeviceCSR(int64_t** rowptr, int32_t** col, double** val, double** rhs) {
        ISL_VERIFY(isl_get_device_csr(Engine::get(), rowptr, col, val, rhs));
    }

private:
    std::size_t n_;
    int64_t nnz_ = 0;
};
}  // namespace solver

// ---- kernel traits: which kernel objects the engine recognises ----------------------------------------------------
template <typename KERNEL> struct KernelTraits { static const bool supported = false; };

namespace asmb {

template <typename FIELDTUPLEBINDER, typename QUADRATURE, typename SOLVER, typename FIELDBINDER, typename KERNEL>
void stiffnessMatrixComputation(const QUADRATURE&, SOLVER&, const FIELDBINDER&, const KERNEL& kernelObj, const bool incremental = true) {
    static_assert(KernelTraits<KERNEL>::supported, "kernel type not supported by the B200 assembly engine (no CPU fallback)");
    double params[4] = {0, 0, 0, 0};
    KernelTraits<KERNEL>::params(kernelObj, params);
    ISL_VERIFY(isl_assemble_matrix(Engine::get(), KernelTraits<KERNEL>::id, params, QUADRATURE::degree, FIELDTUPLEBINDER::test,
                                   FIELDTUPLEBINDER::trial, incremental ? 1 : 0));
}

template <typename FIELDTUPLEBINDER, typename QUADRATURE, typename SOLVER, typename FIELDBINDER, typename KERNEL>
void computeResidualForces(const QUADRATURE&, SOLVER&, const FIELDBINDER&, const KERNEL& kernelObj) {
    static_assert(KernelTraits<KERNEL>::supported, "kernel type not supported by the B200 assembly engine (no CPU fallback)");
    double params[4] = {0, 0, 0, 0};
    KernelTraits<KERNEL>::params(kernelObj, params);
    // the reference moves the forces to the right-hand side with factor -1 (ForceIntegrator.hpp:55)
    ISL_VERIFY(isl_assemble_residual(Engine::get(), KernelTraits<KERNEL>::id, params, QUADRATURE::degree, FIELDTUPLEBINDER::test,
                                     FIELDTUPLEBINDER::trial, -1.0));
}

//! body force with a constant force vector (arbitrary f(x) callbacks cannot run on the device)
template <typename FIELDTUPLEBINDER, typename QUADRATURE, typename SOLVER, typename FIELDBINDER, typename VEC>
void bodyForceComputation(const QUADRATURE&, SOLVER&, const FIELDBINDER&, const VEC& constantForce) {
    double f[3] = {0, 0, 0};
    for (std::size_t d = 0; d < constantForce.size() && d < 3; d++) f[d] = constantForce[d];
    ISL_VERIFY(isl_assemble_bodyforce(Engine::get(), f, QUADRATURE::degree, FIELDTUPLEBINDER::test));
}
}  // namespace asmb
}  // namespace base

// ---- physics kernels (same class names and constructors as the reference) --------------------------------------------
namespace mat {
struct Lame {  // mat/Lame.hpp:24-51
    static double lambda(double E, double nu) { return E * nu / (1. + nu) / (1. - 2. * nu); }
    static double mu(double E, double nu) { return E / 2. / (1. + nu); }
    static double bulk(double E, double nu) { return E / 3. / (1. - 2. * nu); }
};
namespace hypel {
struct StVenant { StVenant(double lambda, double mu) : lambda_(lambda), mu_(mu) {} double lambda_, mu_; };
struct NeoHookeanCompressible { NeoHookeanCompressible(double lambda, double mu) : lambda_(lambda), mu_(mu) {} double lambda_, mu_; };
}  // namespace hypel
}  // namespace mat

namespace heat { template <typename TUPLE> struct Laplace { explicit Laplace(double k) : kappa(k) {} double kappa; }; }
namespace fluid {
template <typename TUPLE> struct VectorLaplace { explicit VectorLaplace(double v) : viscosity(v) {} double viscosity; };
template <typename TUPLE> struct PressureGradient {};
template <typename TUPLE> struct VelocityDivergence { explicit VelocityDivergence(bool c = false) : changeSign(c) {} bool changeSign; };
}  // namespace fluid
namespace solid { template <typename MATERIAL, typename TUPLE> struct HyperElastic { explicit HyperElastic(const MATERIAL& m) : material(m) {} MATERIAL material; }; }

namespace base {
template <typename T> struct KernelTraits<heat::Laplace<T>> {
    static const bool supported = true; static const int id = ISL_K_LAPLACE;
    static void params(const heat::Laplace<T>& k, double* p) { p[0] = k.kappa; }
};
template <typename T> struct KernelTraits<fluid::VectorLaplace<T>> {
    static const bool supported = true; static const int id = ISL_K_VECTOR_LAPLACE;
    static void params(const fluid::VectorLaplace<T>& k, double* p) { p[0] = k.viscosity; }
};
template <typename T> struct KernelTraits<fluid::PressureGradient<T>> {
    static const bool supported = true; static const int id = ISL_K_PRESSURE_GRADIENT;
    static void params(const fluid::PressureGradient<T>&, double*) {}
};
template <typename T> struct KernelTraits<fluid::VelocityDivergence<T>> {
    static const bool supported = true; static const int id = ISL_K_VELOCITY_DIVERGENCE;
    static void params(const fluid::VelocityDivergence<T>& k, double* p) { p[0] = k.changeSign ? 1. : 0.; }
};
template <typename T> struct KernelTraits<solid::HyperElastic<mat::hypel::StVenant, T>> {
    static const bool supported = true; static const int id = ISL_K_HYPEL_STVENANT;
    static void params(const solid::HyperElastic<mat::hypel::StVenant, T>& k, double* p) { p[0] = k.material.lambda_; p[1] = k.material.mu_; }
};
template <typename T> struct KernelTraits<solid::HyperElastic<mat::hypel::NeoHookeanCompressible, T>> {
    static const bool supported = true; static const int id = ISL_K_HYPEL_NEOHOOKE;
    static void params(const solid::HyperElastic<mat::hypel::NeoHookeanCompressible, T>& k, double* p) { p[0] = k.material.lambda_; p[1] = k.material.mu_; }
};
}  // namespace base

#endif  // INSILICO_B200_HPP
