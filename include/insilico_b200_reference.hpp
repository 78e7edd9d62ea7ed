// =============================================================================
// insilico_b200_reference.hpp -- the binding a maintainer of thrueberg/inSilico adds to the reference tree
// (it would live there as base/solver/B200.hpp).  It is compiled TOGETHER WITH THE UNMODIFIED REFERENCE HEADERS
// and routes the assembly hot path to the B200 engine through the C ABI of insilico_b200.h:
//
//   * base::solver::B200            same members as base::solver::Eigen3 (base/solver/Eigen3.hpp:71-341); storage
//                                   = device CSR + device rhs of the engine
//   * base::asmb::stiffnessMatrixComputation<FTB> / computeResidualForces<FTB> / bodyForceComputation<FTB>
//                                   overloads for SOLVER = base::solver::B200 (base/asmb/StiffnessMatrix.hpp:49-87,
//                                   ForceIntegrator.hpp:37-71, BodyForce.hpp:65-84).  SOLVER is a template parameter
//                                   of the reference's functions, so the overloads win by partial ordering and no
//                                   reference file changes; an application switches one typedef
//                                   (e.g. reference/04-heat/dirichlet.cpp:150).
//   * flattening                    the reference's heap objects (Node, Element, DegreeOfFreedom; reached through
//                                   FieldBinder::elementsBegin()/End(), base/asmb/FieldBinder.hpp:150-181) become the
//                                   flat arrays of isl_mesh_set / isl_field_set: topology once per binder, DoF state
//                                   (status, equation numbers, prescribed and current values, linear constraints
//                                   with their master DoFs and weights, node coordinates) re-read at every
//                                   assembly call and uploaded only when it changed.
//   * kernel objects                recognised at compile time (B200KernelTraits); a kernel type the engine does not
//                                   implement is a compile error -- there is no CPU fallback.  The reference keeps
//                                   material constants private, so they are recovered by probing the kernel object
//                                   through its public tangentStiffness on one element (all supported integrands are
//                                   linear in their constants); a maintainer would add accessors instead.
//
// Errors follow the reference: VERIFY_MSG (message on stderr + abort, base/verify.hpp:139-149).
// =============================================================================
#ifndef INSILICO_B200_REFERENCE_HPP
#define INSILICO_B200_REFERENCE_HPP

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <ostream>
#include <vector>

#include <insilico_b200.h>

#include <base/verify.hpp>
#include <base/linearAlgebra.hpp>
#include <base/shape.hpp>
#include <base/Quadrature.hpp>
#include <base/geometry.hpp>
#include <base/asmb/FieldBinder.hpp>
#include <base/asmb/FieldTupleBinder.hpp>
#include <base/asmb/StiffnessMatrix.hpp>
#include <base/asmb/ForceIntegrator.hpp>
#include <base/asmb/BodyForce.hpp>
#include <base/asmb/NeumannForce.hpp>
#include <base/kernel/Mass.hpp>
#include <fluid/Convection.hpp>
#include <heat/Laplace.hpp>
#include <heat/Static.hpp>
#include <mat/thermal/IsotropicConstant.hpp>
#include <fluid/VectorLaplace.hpp>
#include <fluid/PressureGradient.hpp>
#include <fluid/VelocityDivergence.hpp>
#include <solid/HyperElastic.hpp>
#include <mat/hypel/StVenant.hpp>
#include <mat/hypel/NeoHookeanCompressible.hpp>

namespace base {
namespace solver {

namespace b200_detail {

inline void check(int rc) { VERIFY_MSG(rc == 0, std::string("B200 engine: ") + isl_last_error()); }

//! The GPU the calling host thread works on (SURVEY 8(b) "one host thread or process drives 1-8 GPUs").  Like
//! cudaSetDevice it is a per-thread selection: solvers constructed and assembly calls made while device d is selected run
//! on the engine of device d, every device keeps its own engine and its own flattened copy of the binder behind it.  So
//! one thread can loop over the GPUs of a node (selectDevice(d); assemble the d-th element block; ...) and an OpenMP
//! team can drive one GPU per thread.  Default: ISL_B200_DEVICE, else the local rank a launcher exported (torchrun
//! LOCAL_RANK, Open MPI, MVAPICH, Slurm), else 0 -- one process per GPU needs no code at all.
inline int defaultDevice() {
    static const char* const names[] = {"ISL_B200_DEVICE", "LOCAL_RANK", "OMPI_COMM_WORLD_LOCAL_RANK", "MV2_COMM_WORLD_LOCAL_RANK", "SLURM_LOCALID"};
    for (std::size_t i = 0; i < sizeof(names) / sizeof(names[0]); i++) {
        const char* v = std::getenv(names[i]);
        if (v != NULL && *v != 0) return std::atoi(v);
    }
    return 0;
}
inline int& currentDevice() {
    static thread_local int d = defaultDevice();
    return d;
}
struct PerDevice;
inline PerDevice& perDevice();   // the engine and the binder state of the selected device (defined below BinderState)

inline isl_handle engine();

template <typename T>
bool assignIfChanged(std::vector<T>& cache, const std::vector<T>& fresh) {
    if (cache.size() == fresh.size() && (fresh.empty() || std::memcmp(&cache[0], &fresh[0], fresh.size() * sizeof(T)) == 0))
        return false;
    cache = fresh;
    return true;
}

//! meshes below this size are scanned by one thread (measured: 10 ms for 262 k hexahedra; waking the OpenMP team costs more)
#ifndef ISL_B200_PARALLEL_SCAN_MIN
#define ISL_B200_PARALLEL_SCAN_MIN 400000
#endif

//! store into an array that several scanning threads may write with the SAME value (shared nodes / DoFs)
template <typename T>
inline void storeShared(T& dst, T v) { __atomic_store(&dst, &v, __ATOMIC_RELAXED); }

struct FieldState {
    bool bound = false;
    int feDeg = 0, dofSize = 0;
    int64_t nObj = 0;
    std::vector<int32_t> elemDof;
    std::vector<int64_t> eqn;
    std::vector<uint8_t> status;
    std::vector<double> prescribed, values;
    // linear constraints with master DoFs (base/dof/Constraint.hpp): flat form of isl_field_set_constraints
    std::vector<int64_t> conDof, conPtr, masterEqn;
    std::vector<double> weight;
};

struct BinderState {
    const void* key = NULL;            // first geometry element behind the binder
    const void* fieldKey[5] = {NULL, NULL, NULL, NULL, NULL};   // first element of every bound field
    std::size_t numElements = 0;
    int shape = 0, geomDeg = 0, dim = 0;
    int64_t nNodes = 0;
    std::vector<int32_t> conn;
    std::vector<double> coords;
    FieldState field[5];
};

struct PerDevice {
    isl_handle handle = NULL;
    BinderState binder;
    unsigned long scannedForSolver = 0, currentSolver = 0, latestSolver = 0;   // the engine of a device holds ONE system at a time
};
inline PerDevice& perDevice() {
    static std::map<int, PerDevice> all;   // node-based container: references stay valid while other devices are added
#ifdef _OPENMP
    PerDevice* p;
#pragma omp critical(isl_b200_per_device)
    p = &all[currentDevice()];
    return *p;
#else
    return all[currentDevice()];
#endif
}
inline isl_handle engine() {
    PerDevice& d = perDevice();
    if (d.handle == NULL) {
        // ISL_B200_PHYSICAL_DEVICES=n folds the selected indices onto n GPUs (tests of the multi-device logic on a box
        // with fewer GPUs: several engines then share one GPU)
        int physical = currentDevice();
        const char* fold = std::getenv("ISL_B200_PHYSICAL_DEVICES");
        if (fold != NULL && std::atoi(fold) > 0) physical %= std::atoi(fold);
        check(isl_engine_create(physical, &d.handle));
    }
    return d.handle;
}
inline BinderState& state() { return perDevice().binder; }

// ---- flatten one field of the binder (slot N = 1..5 of FieldBinder::ElementPtrTuple) -------------------------------
template <typename ELEMENTPTR>
struct IsDummy {
    static const bool value = boost::is_same<ELEMENTPTR, base::asmb::detail_::DummyElementPtr>::value;
};

template <int N, typename FIELDBINDER, bool DUMMY>
struct FlattenField {
    static void apply(const FIELDBINDER&, BinderState&, bool) {}   // slot not bound by this binder: the engine keeps what it has
};

template <int N, typename FIELDBINDER>
struct FlattenField<N, FIELDBINDER, false> {
    typedef typename FIELDBINDER::ElementPtrTuple EPT;
    typedef typename base::TypeReduction<typename EPT::template Binder<N>::Type>::Type Element;
    typedef typename Element::DegreeOfFreedom DoF;

    static void apply(const FIELDBINDER& fb, BinderState& s, bool topologyChanged) {
        FieldState& f = s.field[N - 1];
        const int ds = static_cast<int>(DoF::size), ndpe = static_cast<int>(Element::numDoFs);
        const typename FIELDBINDER::FieldIterator it = fb.elementsBegin();
        const void* firstElement = static_cast<const void*>((*it).template get<N>());
        bool redefine = topologyChanged || !f.bound || s.fieldKey[N - 1] != firstElement;
        s.fieldKey[N - 1] = firstElement;
        const long numE = static_cast<long>(s.numElements);
        if (redefine) {
            f.feDeg = static_cast<int>(Element::FEFun::degree);
            f.dofSize = ds;
            f.elemDof.assign(s.numElements * ndpe, 0);
            long maxId = 0;
#pragma omp parallel for schedule(static) reduction(max : maxId) if (numE > ISL_B200_PARALLEL_SCAN_MIN)
            for (long e = 0; e < numE; e++) {
                const Element* ep = (*(it + e)).template get<N>();
                int k = 0;
                for (typename Element::DoFPtrConstIter d = ep->doFsBegin(); d != ep->doFsEnd(); ++d, ++k) {
                    const long id = static_cast<long>((*d)->getID());
                    f.elemDof[e * ndpe + k] = static_cast<int32_t>(id);
                    if (id > maxId) maxId = id;
                }
            }
            f.nObj = static_cast<int64_t>(maxId) + 1;
        }
        // DoF state, read through the element -> DoF pointers by all host threads (shared DoFs are simply visited more
        // than once and written with the same values)
        const std::size_t n = static_cast<std::size_t>(f.nObj) * ds;
        std::vector<int64_t> eqn(n, -1);
        std::vector<uint8_t> status(n, ISL_INACTIVE);
        std::vector<double> prescribed(n, 0.), values(n, 0.);
        std::vector<uint8_t> seen(n, 0);
        std::vector<uint8_t> visited(static_cast<std::size_t>(f.nObj), 0);   // a DoF object is shared by all elements around it: read once
        struct Slave { int64_t dof; std::vector<std::pair<base::number, std::size_t> > masters; };
        std::vector<Slave> slaves;
#pragma omp parallel for schedule(static) if (numE > ISL_B200_PARALLEL_SCAN_MIN)
        for (long e = 0; e < numE; e++) {
            const Element* ep = (*(it + e)).template get<N>();
            double pv[DoF::size];
            for (typename Element::DoFPtrConstIter d = ep->doFsBegin(); d != ep->doFsEnd(); ++d) {
                const DoF* doF = *d;
                const std::size_t id = doF->getID();
                if (id >= visited.size()) continue;   // (cannot happen: nObj is the largest id seen when the table was built)
                // (two threads may both find the flag clear: they then store the same values, see storeShared)
                if (__atomic_load_n(&visited[id], __ATOMIC_RELAXED) != 0) continue;
                storeShared(visited[id], static_cast<uint8_t>(1));
                const std::size_t o = id * ds;
                doF->getPrescribedValues(&pv[0], false);
                for (int c = 0; c < ds; c++) {
                    if (doF->isActive(c)) {
                        storeShared(status[o + c], static_cast<uint8_t>(ISL_ACTIVE));
                        storeShared(eqn[o + c], static_cast<int64_t>(doF->getIndex(c)));
                    } else if (doF->isConstrained(c)) {
                        storeShared(status[o + c], static_cast<uint8_t>(ISL_CONSTRAINED));
                        storeShared(prescribed[o + c], static_cast<double>(pv[c]));
                        if (__atomic_exchange_n(&seen[o + c], static_cast<uint8_t>(1), __ATOMIC_RELAXED) == 0) {
                            // masters of a slave DoF (base/dof/Constraint.hpp:118-136), once per DoF
                            Slave sl;
                            sl.dof = static_cast<int64_t>(o + c);
                            const_cast<DoF*>(doF)->getConstraint(c)->getWeightedDoFIDs(sl.masters);
                            if (!sl.masters.empty()) {
#pragma omp critical(isl_b200_slaves)
                                slaves.push_back(sl);
                            }
                        }
                    }
                    storeShared(values[o + c], static_cast<double>(doF->getValue(c)));
                }
            }
        }
        std::sort(slaves.begin(), slaves.end(), [](const Slave& a, const Slave& b) { return a.dof < b.dof; });
        std::vector<int64_t> conDof, conPtr(1, 0), masterEqn;
        std::vector<double> weight;
        for (std::size_t k = 0; k < slaves.size(); k++) {
            conDof.push_back(slaves[k].dof);
            for (std::size_t m = 0; m < slaves[k].masters.size(); m++) {
                weight.push_back(slaves[k].masters[m].first);
                masterEqn.push_back(static_cast<int64_t>(slaves[k].masters[m].second));
            }
            conPtr.push_back(static_cast<int64_t>(masterEqn.size()));
        }
        const bool numberingChanged = assignIfChanged(f.eqn, eqn) | assignIfChanged(f.status, status);
        const bool prescChanged = assignIfChanged(f.prescribed, prescribed);
        const bool valuesChanged = assignIfChanged(f.values, values);
        const bool constraintsChanged = assignIfChanged(f.conDof, conDof) | assignIfChanged(f.conPtr, conPtr) |
                                        assignIfChanged(f.masterEqn, masterEqn) | assignIfChanged(f.weight, weight);
        if (redefine || numberingChanged) {
            check(isl_field_set(engine(), N - 1, f.feDeg, f.dofSize, f.nObj, &f.elemDof[0], &f.eqn[0], &f.status[0],
                                &f.prescribed[0], &f.values[0]));
            if (!f.conDof.empty())
                check(isl_field_set_constraints(engine(), N - 1, static_cast<int64_t>(f.conDof.size()), &f.conDof[0],
                                                &f.conPtr[0], &f.masterEqn[0], &f.weight[0]));
        } else if (constraintsChanged) {
            check(isl_field_set_constraints(engine(), N - 1, static_cast<int64_t>(f.conDof.size()),
                                            f.conDof.empty() ? NULL : &f.conDof[0], &f.conPtr[0],
                                            f.masterEqn.empty() ? NULL : &f.masterEqn[0], f.weight.empty() ? NULL : &f.weight[0]));
            if (prescChanged || valuesChanged)
                check(isl_field_update(engine(), N - 1, prescChanged ? &f.prescribed[0] : NULL,
                                       valuesChanged ? &f.values[0] : NULL));
        } else if (prescChanged || valuesChanged) {
            check(isl_field_update(engine(), N - 1, prescChanged ? &f.prescribed[0] : NULL,
                                   valuesChanged ? &f.values[0] : NULL));
        }
        f.bound = true;
    }
};

template <int N, typename FIELDBINDER>
void flattenField(const FIELDBINDER& fb, BinderState& s, bool topologyChanged) {
    typedef typename FIELDBINDER::ElementPtrTuple EPT;
    FlattenField<N, FIELDBINDER, IsDummy<typename EPT::template Binder<N>::Type>::value>::apply(fb, s, topologyChanged);
}

//! Rescan policy.  The reference's element loop reads its heap objects at every assembly call; re-reading them all for
//! every call costs more than the assembly itself (345 ns per element measured in round 1).  Default here: a FULL scan
//! (by all host threads, OpenMP) at the first assembly call of each solver instance -- the reference's applications
//! build a new solver per Newton iteration and change DoFs, constraints and nodes only between solvers -- and a SAMPLED
//! comparison (every 61st element: connectivity, node coordinates, DoF status / numbering / values) at the other calls,
//! which falls back to the full scan when anything differs.  rescanEveryCall() = true restores the reference's exact
//! semantics; rescanOncePerSolver() = true drops the sampled comparison as well.
inline bool& rescanOncePerSolver() {
    static bool flag = false;
    return flag;
}
inline bool& rescanEveryCall() {
    static bool flag = false;
    return flag;
}
inline unsigned long& scannedForSolver() { return perDevice().scannedForSolver; }
inline unsigned long& currentSolver() { return perDevice().currentSolver; }
inline unsigned long& fullScans() {   // statistics for tests
    static unsigned long n = 0;
    return n;
}

template <int N, typename FIELDBINDER, bool DUMMY>
struct SampleField {
    static bool same(const FIELDBINDER&, const BinderState&, long) { return true; }
};
template <int N, typename FIELDBINDER>
struct SampleField<N, FIELDBINDER, false> {
    typedef typename FIELDBINDER::ElementPtrTuple EPT;
    typedef typename base::TypeReduction<typename EPT::template Binder<N>::Type>::Type Element;
    typedef typename Element::DegreeOfFreedom DoF;
    static bool same(const FIELDBINDER& fb, const BinderState& s, long e) {
        const FieldState& f = s.field[N - 1];
        if (!f.bound) return false;
        const int ds = static_cast<int>(DoF::size);
        const Element* ep = (*(fb.elementsBegin() + e)).template get<N>();
        double pv[DoF::size];
        for (typename Element::DoFPtrConstIter d = ep->doFsBegin(); d != ep->doFsEnd(); ++d) {
            const DoF* doF = *d;
            const std::size_t o = doF->getID() * ds;
            if (o + ds > f.status.size()) return false;
            doF->getPrescribedValues(&pv[0], false);
            for (int c = 0; c < ds; c++) {
                const uint8_t st = doF->isActive(c) ? ISL_ACTIVE : (doF->isConstrained(c) ? ISL_CONSTRAINED : ISL_INACTIVE);
                if (st != f.status[o + c] || doF->getValue(c) != f.values[o + c]) return false;
                if (st == ISL_ACTIVE && static_cast<int64_t>(doF->getIndex(c)) != f.eqn[o + c]) return false;
                if (st == ISL_CONSTRAINED && pv[c] != f.prescribed[o + c]) return false;
            }
        }
        return true;
    }
};

//! sampled comparison of the reference's objects with the flat copies of the last full scan
template <typename FIELDBINDER>
bool sampleUnchanged(const FIELDBINDER& fb, const BinderState& s) {
    typedef typename FIELDBINDER::ElementPtrTuple EPT;
    typedef typename EPT::GeomElement GeomElement;
    typedef typename GeomElement::Node Node;
    const long numE = static_cast<long>(std::distance(fb.elementsBegin(), fb.elementsEnd()));
    if (static_cast<std::size_t>(numE) != s.numElements) return false;
    const int npe = static_cast<int>(GeomElement::numNodes), dim = static_cast<int>(Node::dim);
    const typename FIELDBINDER::FieldIterator it = fb.elementsBegin();
    for (long e = 0; e < numE; e += 61) {
        const GeomElement* gep = (*(it + e)).geomElementPtr();
        int k = 0;
        double x[3];
        for (typename GeomElement::NodePtrConstIter n = gep->nodesBegin(); n != gep->nodesEnd(); ++n, ++k) {
            const std::size_t id = (*n)->getID();
            if (static_cast<int32_t>(id) != s.conn[e * npe + k]) return false;
            (*n)->getX(&x[0]);
            for (int d = 0; d < dim; d++) if (x[d] != s.coords[id * dim + d]) return false;
        }
        if (!SampleField<1, FIELDBINDER, IsDummy<typename EPT::template Binder<1>::Type>::value>::same(fb, s, e)) return false;
        if (!SampleField<2, FIELDBINDER, IsDummy<typename EPT::template Binder<2>::Type>::value>::same(fb, s, e)) return false;
        if (!SampleField<3, FIELDBINDER, IsDummy<typename EPT::template Binder<3>::Type>::value>::same(fb, s, e)) return false;
        if (!SampleField<4, FIELDBINDER, IsDummy<typename EPT::template Binder<4>::Type>::value>::same(fb, s, e)) return false;
        if (!SampleField<5, FIELDBINDER, IsDummy<typename EPT::template Binder<5>::Type>::value>::same(fb, s, e)) return false;
    }
    return true;
}

//! Identity of what a binder binds: the reference builds FieldBinder objects on the fly (a new local one in every method of
//! base::BoundaryValueProblem, BoundaryValueProblem.hpp:238-345), so the binder's own address says nothing; the first
//! geometry element and the first element of every bound field do.
template <int N, typename FIELDBINDER, bool DUMMY>
struct FirstElement {
    static const void* get(const FIELDBINDER&) { return NULL; }
};
template <int N, typename FIELDBINDER>
struct FirstElement<N, FIELDBINDER, false> {
    static const void* get(const FIELDBINDER& fb) { return static_cast<const void*>((*fb.elementsBegin()).template get<N>()); }
};
template <typename FIELDBINDER>
bool sameBinding(const FIELDBINDER& fb, const BinderState& s) {
    typedef typename FIELDBINDER::ElementPtrTuple EPT;
    if (fb.elementsBegin() == fb.elementsEnd()) return false;
    if (s.key != static_cast<const void*>((*fb.elementsBegin()).geomElementPtr())) return false;
    // dummy slots give NULL, and a NULL slot of this binder is not compared: a binder over fewer fields of the same
    // mesh does not redefine anything
    const void* k[5] = {FirstElement<1, FIELDBINDER, IsDummy<typename EPT::template Binder<1>::Type>::value>::get(fb),
                        FirstElement<2, FIELDBINDER, IsDummy<typename EPT::template Binder<2>::Type>::value>::get(fb),
                        FirstElement<3, FIELDBINDER, IsDummy<typename EPT::template Binder<3>::Type>::value>::get(fb),
                        FirstElement<4, FIELDBINDER, IsDummy<typename EPT::template Binder<4>::Type>::value>::get(fb),
                        FirstElement<5, FIELDBINDER, IsDummy<typename EPT::template Binder<5>::Type>::value>::get(fb)};
    for (int i = 0; i < 5; i++)
        if (k[i] != NULL && (!s.field[i].bound || s.fieldKey[i] != k[i])) return false;
    return true;
}
//! Bring the engine's copy of mesh and fields in line with the reference's objects behind this binder.
template <typename FIELDBINDER>
void synchronise(const FIELDBINDER& fb) {
    if (fb.elementsBegin() == fb.elementsEnd()) return;
    const bool known = sameBinding(fb, state());
    if (!rescanEveryCall() && known && scannedForSolver() == currentSolver() &&
        currentSolver() != 0) {
        if (rescanOncePerSolver() || sampleUnchanged(fb, state())) return;
    }
    scannedForSolver() = currentSolver();
    fullScans()++;
    typedef typename FIELDBINDER::ElementPtrTuple EPT;
    typedef typename EPT::GeomElement GeomElement;
    typedef typename GeomElement::Node Node;
    BinderState& s = state();
    const std::size_t numElements = static_cast<std::size_t>(std::distance(fb.elementsBegin(), fb.elementsEnd()));
    const long numE = static_cast<long>(numElements);
    const int npe = static_cast<int>(GeomElement::numNodes), dim = static_cast<int>(Node::dim);
    // another mesh / other fields behind the binder: everything is redefined (whether the same mesh was rebuilt is
    // decided below by comparing the connectivity itself)
    bool topologyChanged = (s.key != static_cast<const void*>((*fb.elementsBegin()).geomElementPtr())) || (s.numElements != numElements);

    // connectivity (also re-read when the binder is known: cheap, and catches a re-meshed binder); all host threads
    std::vector<int32_t> conn(numElements * npe);
    long maxNode = 0;
    const typename FIELDBINDER::FieldIterator it0 = fb.elementsBegin();
#pragma omp parallel for schedule(static) reduction(max : maxNode) if (numE > ISL_B200_PARALLEL_SCAN_MIN)
    for (long e = 0; e < numE; e++) {
        const GeomElement* gep = (*(it0 + e)).geomElementPtr();
        int k = 0;
        for (typename GeomElement::NodePtrConstIter n = gep->nodesBegin(); n != gep->nodesEnd(); ++n, ++k) {
            const long id = static_cast<long>((*n)->getID());
            conn[e * npe + k] = static_cast<int32_t>(id);
            if (id > maxNode) maxNode = id;
        }
    }
    topologyChanged = assignIfChanged(s.conn, conn) || topologyChanged;
    const int64_t nNodes = static_cast<int64_t>(maxNode) + 1;
    std::vector<double> coords(static_cast<std::size_t>(nNodes) * dim, 0.);
    std::vector<uint8_t> nodeVisited(static_cast<std::size_t>(nNodes), 0);   // a node is shared by all elements around it: read once
#pragma omp parallel for schedule(static) if (numE > ISL_B200_PARALLEL_SCAN_MIN)
    for (long e = 0; e < numE; e++) {
        const int32_t* ce = &conn[e * npe];   // node ids of the element, read in the loop above
        const GeomElement* gep = NULL;
        for (int k = 0; k < npe; k++) {
            if (__atomic_load_n(&nodeVisited[ce[k]], __ATOMIC_RELAXED) != 0) continue;
            storeShared(nodeVisited[ce[k]], static_cast<uint8_t>(1));
            if (gep == NULL) gep = (*(it0 + e)).geomElementPtr();
            double x[3];
            (*(gep->nodesBegin() + k))->getX(&x[0]);
            for (int d = 0; d < dim; d++) storeShared(coords[static_cast<std::size_t>(ce[k]) * dim + d], x[d]);
        }
    }
    const bool coordsChanged = assignIfChanged(s.coords, coords);
    if (topologyChanged || s.nNodes != nNodes) {
        s.numElements = numElements;
        s.shape = static_cast<int>(GeomElement::shape);
        s.geomDeg = static_cast<int>(GeomElement::GeomFun::degree);
        s.dim = dim;
        s.nNodes = nNodes;
        check(isl_mesh_set(engine(), s.shape, s.geomDeg, dim, nNodes, &s.coords[0], static_cast<int64_t>(numElements),
                           &s.conn[0]));
        topologyChanged = true;
    } else if (coordsChanged) {
        check(isl_mesh_update_coords(engine(), &s.coords[0]));
    }
    s.key = static_cast<const void*>((*fb.elementsBegin()).geomElementPtr());
    flattenField<1>(fb, s, topologyChanged);
    flattenField<2>(fb, s, topologyChanged);
    flattenField<3>(fb, s, topologyChanged);
    flattenField<4>(fb, s, topologyChanged);
    flattenField<5>(fb, s, topologyChanged);
}

// ---- compile-time facts ----------------------------------------------------------------------------------------
template <typename FTB>
struct TupleIndices;
template <typename EPT, int I, int J, int K, int L, int M>
struct TupleIndices<base::asmb::FieldTupleBinder<EPT, I, J, K, L, M> > {
    static const int test = I - 1, trial = J - 1;  // the engine addresses fields 0-based
    static const int aux = (K > 0 ? K - 1 : -1);    // AuxField1 of the tuple (fluid::Convection: the advection velocity)
};

template <typename QUADRATURE>
struct QuadratureDegree;
template <unsigned DEGREE, base::Shape SHAPE>
struct QuadratureDegree<base::Quadrature<DEGREE, SHAPE> > {
    static const int value = static_cast<int>(DEGREE);
};
template <unsigned DEGREE, base::Shape SHAPE>
struct QuadratureDegree<base::SurfaceQuadrature<DEGREE, SHAPE> > {   // base/Quadrature.hpp:148-151: the rule of the face shape
    static const int value = static_cast<int>(DEGREE);
};

// ---- probing of kernel constants through the public interface ------------------------------------------------------
template <typename KERNEL, typename TUPLE>
base::MatrixD probeTangent(const KERNEL& kernel, const TUPLE& tuple) {
    typedef typename TUPLE::GeomElement GeomElement;
    typedef typename TUPLE::TestElement TestElement;
    typedef typename TUPLE::TrialElement TrialElement;
    const unsigned nr = TestElement::numDoFs * TestElement::DegreeOfFreedom::size;
    const unsigned nc = TrialElement::numDoFs * TrialElement::DegreeOfFreedom::size;
    base::MatrixD K = base::MatrixD::Zero(nr, nc);
    kernel.tangentStiffness(tuple, base::ShapeCentroid<GeomElement::shape>::apply(), 1.0, K);
    return K;
}

inline double maxAbs(const base::MatrixD& A) {
    double m = 0.;
    for (int j = 0; j < A.cols(); j++)
        for (int i = 0; i < A.rows(); i++) m = std::max(m, std::abs(A(i, j)));
    return m;
}

//! K = c * K1 -> c; the candidate that reproduces K bit for bit is preferred over the plain quotient
template <typename MAKEKERNEL, typename TUPLE>
double probeScalar(const base::MatrixD& K, const MAKEKERNEL& make, const TUPLE& tuple) {
    const base::MatrixD K1 = probeTangent(make(1.0), tuple);
    int bi = 0, bj = 0;
    for (int j = 0; j < K1.cols(); j++)
        for (int i = 0; i < K1.rows(); i++)
            if (std::abs(K1(i, j)) > std::abs(K1(bi, bj))) { bi = i; bj = j; }
    VERIFY_MSG(K1(bi, bj) != 0., "B200 engine: cannot probe the kernel constant (zero element matrix)");
    const double c = K(bi, bj) / K1(bi, bj);
    double cand[5] = {c, std::nextafter(c, 1e300), std::nextafter(c, -1e300), 0., 0.};
    cand[3] = std::nextafter(cand[1], 1e300);
    cand[4] = std::nextafter(cand[2], -1e300);
    for (int k = 0; k < 5; k++) {
        const base::MatrixD Kc = probeTangent(make(cand[k]), tuple);
        bool same = true;
        for (int j = 0; j < K.cols() && same; j++)
            for (int i = 0; i < K.rows() && same; i++) same = (Kc(i, j) == K(i, j));
        if (same) return cand[k];
    }
    const double scale = maxAbs(K);
    for (int j = 0; j < K.cols(); j++)
        for (int i = 0; i < K.rows(); i++)
            VERIFY_MSG(std::abs(K(i, j) - c * K1(i, j)) <= 1e-12 * scale,
                       "B200 engine: kernel object is not a constant multiple of the unit kernel "
                       "(non-constant material functions are not supported)");
    return c;
}

}  // namespace b200_detail

//------------------------------------------------------------------------------------------------------------------
/** Which engine integrand a reference kernel object is, and with which constants.  Specialise for further kernels;
 *  the primary template is a static_assert (unsupported kernel = compile error, no CPU fallback).                 */
template <typename KERNEL>
struct B200KernelTraits {
    static_assert(sizeof(KERNEL) == 0,
                  "this kernel object has no implementation in the B200 assembly engine (supported: heat::Laplace, base::kernel::Mass, "
                  "heat::Static<mat::thermal::IsotropicConstant>, fluid::VectorLaplace, fluid::PressureGradient, "
                  "fluid::VelocityDivergence, fluid::Convection, solid::HyperElastic<mat::hypel::StVenant | NeoHookeanCompressible>); "
                  "there is no CPU fallback");
};

template <typename TUPLE>
struct B200KernelTraits<heat::Laplace<TUPLE> > {
    struct Make {
        heat::Laplace<TUPLE> operator()(double c) const { return heat::Laplace<TUPLE>(c); }
    };
    static int describe(const heat::Laplace<TUPLE>& k, const TUPLE& t0, const TUPLE& t1, double* p) {
#ifdef ISL_B200_HAVE_KERNEL_ACCESSORS   // the maintainer added `double conductivity() const`
        (void)t0; (void)t1;
        p[0] = k.conductivity();
        return ISL_K_LAPLACE;
#endif
        p[0] = b200_detail::probeScalar(b200_detail::probeTangent(k, t0), Make(), t0);
        const double p1 = b200_detail::probeScalar(b200_detail::probeTangent(k, t1), Make(), t1);
        // a conductivity FUNCTION (setConductivityFunction): the matrix overload samples it per quadrature point
        // (b200_detail::SampledFactor); p[3] carries the verdict of this first look
        if (!(std::abs(p1 - p[0]) <= 1e-13 * std::abs(p[0]))) p[3] = 1.0;
        return ISL_K_LAPLACE;
    }
};

//! heat::Static with the constant isotropic material (heat/Static.hpp:60-118, mat/thermal/IsotropicConstant.hpp) is the
//! Laplace operator with conductivity kappa; used by heat::PoissonDriver
template <typename TUPLE>
struct B200KernelTraits<heat::Static<mat::thermal::IsotropicConstant, TUPLE> > {
    struct Unit {   // kernel + the material it refers to
        mat::thermal::IsotropicConstant material;
        heat::Static<mat::thermal::IsotropicConstant, TUPLE> kernel;
        explicit Unit(double c) : material(c), kernel(material) {}
        Unit(const Unit& o) : material(o.material), kernel(material) {}
        template <typename XI>
        void tangentStiffness(const TUPLE& t, const XI& xi, double w, base::MatrixD& K) const { kernel.tangentStiffness(t, xi, w, K); }
    };
    struct Make {
        Unit operator()(double c) const { return Unit(c); }
    };
    static int describe(const heat::Static<mat::thermal::IsotropicConstant, TUPLE>& k, const TUPLE& t0, const TUPLE&, double* p) {
        p[0] = b200_detail::probeScalar(b200_detail::probeTangent(k, t0), Make(), t0);
        return ISL_K_LAPLACE;
    }
};

template <typename TUPLE>
struct B200KernelTraits<fluid::VectorLaplace<TUPLE> > {
    struct Make {
        fluid::VectorLaplace<TUPLE> operator()(double c) const { return fluid::VectorLaplace<TUPLE>(c); }
    };
    static int describe(const fluid::VectorLaplace<TUPLE>& k, const TUPLE& t0, const TUPLE&, double* p) {
#ifdef ISL_B200_HAVE_KERNEL_ACCESSORS   // `double viscosity() const`
        (void)t0;
        p[0] = k.viscosity();
        return ISL_K_VECTOR_LAPLACE;
#endif
        p[0] = b200_detail::probeScalar(b200_detail::probeTangent(k, t0), Make(), t0);
        return ISL_K_VECTOR_LAPLACE;
    }
};

//! base::kernel::Mass (base/kernel/Mass.hpp:88-138): factor * detJ * w * phi_M * psi_N on every DoF component
template <typename TUPLE>
struct B200KernelTraits<base::kernel::Mass<TUPLE> > {
    struct Make {
        base::kernel::Mass<TUPLE> operator()(double c) const { return base::kernel::Mass<TUPLE>(c); }
    };
    static int describe(const base::kernel::Mass<TUPLE>& k, const TUPLE& t0, const TUPLE&, double* p) {
        p[0] = b200_detail::probeScalar(b200_detail::probeTangent(k, t0), Make(), t0);
        return ISL_K_MASS;
    }
};

//! fluid::Convection (fluid/Convection.hpp:88-220, Picard form): rho * (terms of the current velocity state).  The density
//! is private: K = rho * K(1) on an element whose velocity state is not zero.  p[2] = 1 reports that both probed elements
//! are at rest; b200_detail::LateProbe then looks further (and skips the launch if the whole field is at rest).
template <typename TUPLE>
struct B200KernelTraits<fluid::Convection<TUPLE> > {
    struct Make {
        fluid::Convection<TUPLE> operator()(double c) const { return fluid::Convection<TUPLE>(c); }
    };
    static bool probe(const fluid::Convection<TUPLE>& k, const TUPLE& t, double* p) {
        if (b200_detail::maxAbs(b200_detail::probeTangent(Make()(1.0), t)) == 0.) return false;
        p[0] = b200_detail::probeScalar(b200_detail::probeTangent(k, t), Make(), t);
        return true;
    }
    static int describe(const fluid::Convection<TUPLE>& k, const TUPLE& t0, const TUPLE& t1, double* p) {
#ifdef ISL_B200_HAVE_KERNEL_ACCESSORS   // `double density() const`
        (void)t0; (void)t1;
        p[0] = k.density();
        return ISL_K_CONVECTION;
#endif
        if (!probe(k, t0, p) && !probe(k, t1, p)) p[2] = 1.0;
        return ISL_K_CONVECTION;
    }
};

template <typename TUPLE>
struct B200KernelTraits<fluid::PressureGradient<TUPLE> > {
    static int describe(const fluid::PressureGradient<TUPLE>&, const TUPLE&, const TUPLE&, double* p) {
        p[0] = 0.;
        return ISL_K_PRESSURE_GRADIENT;
    }
};

template <typename TUPLE>
struct B200KernelTraits<fluid::VelocityDivergence<TUPLE> > {
    static int describe(const fluid::VelocityDivergence<TUPLE>& k, const TUPLE& t0, const TUPLE&, double* p) {
        const base::MatrixD K = b200_detail::probeTangent(k, t0);
        const base::MatrixD K1 = b200_detail::probeTangent(fluid::VelocityDivergence<TUPLE>(false), t0);
        bool same = true, flipped = true;
        for (int j = 0; j < K.cols(); j++)
            for (int i = 0; i < K.rows(); i++) {
                same = same && (K(i, j) == K1(i, j));
                flipped = flipped && (K(i, j) == -K1(i, j));
            }
        VERIFY_MSG(same || flipped, "B200 engine: unexpected fluid::VelocityDivergence behaviour");
        p[0] = (same ? 0. : 1.);  // changeSign
        return ISL_K_VELOCITY_DIVERGENCE;
    }
};

namespace b200_detail {
//! K = lambda * K(1,0) + mu * K(0,1) for solid::HyperElastic with a two-constant material (fixed displacement state)
template <typename MATERIAL, typename TUPLE>
void probeLame(const solid::HyperElastic<MATERIAL, TUPLE>& k, const TUPLE& t0, double* p) {
#ifdef ISL_B200_HAVE_KERNEL_ACCESSORS   // `const MATERIAL& material() const`, `double lambda() const`, `double mu() const`
    (void)t0;
    p[0] = k.material().lambda(); p[1] = k.material().mu();
    return;
#endif
    const MATERIAL m10(1., 0.), m01(0., 1.);
    const solid::HyperElastic<MATERIAL, TUPLE> k10(m10), k01(m01);
    const base::MatrixD K = probeTangent(k, t0), A = probeTangent(k10, t0), B = probeTangent(k01, t0);
    // normal equations of the two-parameter fit over all entries
    double aa = 0., ab = 0., bb = 0., ak = 0., bk = 0.;
    for (int j = 0; j < K.cols(); j++)
        for (int i = 0; i < K.rows(); i++) {
            aa += A(i, j) * A(i, j); ab += A(i, j) * B(i, j); bb += B(i, j) * B(i, j);
            ak += A(i, j) * K(i, j); bk += B(i, j) * K(i, j);
        }
    const double det = aa * bb - ab * ab;
    VERIFY_MSG(det > 1e-8 * aa * bb, "B200 engine: cannot separate the two material constants on the probe element");
    p[0] = (ak * bb - bk * ab) / det;
    p[1] = (bk * aa - ak * ab) / det;
    // polish: constants entered by the application are usually short decimals of E, nu; keep the fit otherwise
    const double scale = maxAbs(K);
    for (int j = 0; j < K.cols(); j++)
        for (int i = 0; i < K.rows(); i++)
            VERIFY_MSG(std::abs(K(i, j) - (p[0] * A(i, j) + p[1] * B(i, j))) <= 1e-11 * scale,
                       "B200 engine: solid::HyperElastic tangent is not linear in the material constants");
}
}  // namespace b200_detail

template <typename TUPLE>
struct B200KernelTraits<solid::HyperElastic<mat::hypel::StVenant, TUPLE> > {
    static int describe(const solid::HyperElastic<mat::hypel::StVenant, TUPLE>& k, const TUPLE& t0, const TUPLE&, double* p) {
        b200_detail::probeLame(k, t0, p);
        return ISL_K_HYPEL_STVENANT;
    }
};

template <typename TUPLE>
struct B200KernelTraits<solid::HyperElastic<mat::hypel::NeoHookeanCompressible, TUPLE> > {
    static int describe(const solid::HyperElastic<mat::hypel::NeoHookeanCompressible, TUPLE>& k, const TUPLE& t0,
                        const TUPLE&, double* p) {
        b200_detail::probeLame(k, t0, p);
        return ISL_K_HYPEL_NEOHOOKE;
    }
};

//------------------------------------------------------------------------------------------------------------------
/** Linear-system storage on the B200: same interface as base::solver::Eigen3 (base/solver/Eigen3.hpp:62-341).
 *  One system exists per engine at a time, exactly like the reference applications use their solver (a fresh
 *  `Solver solver(n)` per Newton iteration, reference/06-elastic/compressible.cpp:263-266).                        */
class B200 {
public:
    //! Hook for the linear solves, which are outside the assembly path: given the finished CSR system
    //! it overwrites rhs with the solution and returns an iteration count.  NULL = calling a solve is an error.
    typedef int (*SolveHook)(const char* method, std::size_t n, const std::vector<int64_t>& rowptr,
                             const std::vector<int32_t>& col, const std::vector<double>& val, std::vector<double>& rhs);
    static SolveHook& solveHook() {
        static SolveHook hook = NULL;
        return hook;
    }

    //! Constructor with the size N of matrix and vector (Eigen3.hpp:71-77)
    B200(const std::size_t size) : size_(size), solved_(false), id_(nextId_()), device_(b200_detail::currentDevice()) {
        b200_detail::perDevice().latestSolver = id_;
        b200_detail::currentSolver() = id_;
        b200_detail::check(isl_system_create(b200_detail::engine(), static_cast<int64_t>(size)));
    }

    //! Select the GPU for the solvers constructed and the assembly calls made by THIS host thread from now on (see
    //! b200_detail::currentDevice; the reference's only parallel construct, the OpenMP element loop of
    //! base/auxi/parallel.hpp:25-60, becomes one element block per GPU: selectDevice(d), bind block d, assemble)
    static void selectDevice(const int device) {
        VERIFY_MSG(device >= 0, "base::solver::B200::selectDevice: negative device index");
        b200_detail::currentDevice() = device;
    }
    static int selectedDevice() { return b200_detail::currentDevice(); }
    //! the device this solver's system lives on
    int device() const { return device_; }

    //! The engine of a device holds ONE system at a time: using a solver after a newer one was constructed on its
    //! device, or while another device is selected, is an error
    void verifyCurrent() const {
        VERIFY_MSG(device_ == b200_detail::currentDevice(),
                   "base::solver::B200: this solver lives on another device than the one selected (B200::selectDevice)");
        VERIFY_MSG(id_ == b200_detail::perDevice().latestSolver,
                   "base::solver::B200: a newer solver object exists; the engine holds one system at a time");
    }

    //! Insert numbers to matrix storage (Eigen3.hpp:81-108): host-side odd contributions
    template <typename MATRIX, typename RDOFS, typename CDOFS>
    void insertToLHS(const MATRIX& matrix, const RDOFS& rowDoFs, const CDOFS& colDoFs) {
        this->verifyCurrent();
        const std::size_t nr = rowDoFs.size(), nc = colDoFs.size();
        if (nr == 0 || nc == 0) return;
        std::vector<int64_t> r(nr), c(nc);
        std::vector<double> m(nr * nc);
        for (std::size_t i = 0; i < nr; i++) {
            VERIFY_MSG(static_cast<std::size_t>(rowDoFs[i]) < size_, "Row index out of bound: " + x2s(rowDoFs[i]));
            r[i] = static_cast<int64_t>(rowDoFs[i]);
        }
        for (std::size_t j = 0; j < nc; j++) {
            VERIFY_MSG(static_cast<std::size_t>(colDoFs[j]) < size_, "Col index out of bound: " + x2s(colDoFs[j]));
            c[j] = static_cast<int64_t>(colDoFs[j]);
        }
        for (std::size_t i = 0; i < nr; i++)
            for (std::size_t j = 0; j < nc; j++) m[i * nc + j] = matrix(i, j);
        b200_detail::check(isl_insert_lhs(b200_detail::engine(), &m[0], &r[0], static_cast<int>(nr), &c[0], static_cast<int>(nc)));
    }

    //! Insert numbers to RHS vector (Eigen3.hpp:112-124)
    template <typename VECTOR, typename DOFS>
    void insertToRHS(const VECTOR& vector, const DOFS& dofs) {
        this->verifyCurrent();
        const std::size_t n = dofs.size();
        if (n == 0) return;
        std::vector<int64_t> r(n);
        std::vector<double> v(n);
        for (std::size_t i = 0; i < n; i++) {
            VERIFY_MSG(static_cast<std::size_t>(dofs[i]) < size_, x2s(dofs[i]) + " out of bound");
            r[i] = static_cast<int64_t>(dofs[i]);
            v[i] = vector[i];
        }
        b200_detail::check(isl_insert_rhs(b200_detail::engine(), &v[0], &r[0], static_cast<int>(n)));
    }

    //! Pre-determine the non-zero pattern (Eigen3.hpp:329-336, TripletContainer.hpp:158-301)
    template <typename FIELDTUPLEBINDER, typename FIELDBINDER>
    void registerFields(const FIELDBINDER& fieldBinder) {
        this->verifyCurrent();
        b200_detail::synchronise(fieldBinder);
        typedef b200_detail::TupleIndices<FIELDTUPLEBINDER> TI;
        b200_detail::check(isl_pattern_register(b200_detail::engine(), TI::test, TI::trial));
    }

    //! Finish the assembly (Eigen3.hpp:142-153): waits for the device
    void finishAssembly(const bool = true) {
        this->verifyCurrent();
        int64_t n = 0, nnz = 0;
        b200_detail::check(isl_finish(b200_detail::engine(), &n, &nnz));
        nnz_ = static_cast<std::size_t>(nnz);
    }

    //! Norm of the rhs/solution vector; like the reference divided by the length (Eigen3.hpp:128-138)
    double norm() const {
        if (solved_) return this->norm(0, size_);
        this->verifyCurrent();
        double v = 0.;
        b200_detail::check(isl_rhs_norm(b200_detail::engine(), &v));
        return v;
    }
    double norm(const std::size_t first, const std::size_t last) const {
        const std::vector<double>& b = this->hostRhs_();
        double s = 0.;
        for (std::size_t i = first; i < last; i++) s += b[i] * b[i];
        return std::sqrt(s) / static_cast<double>(last - first);
    }

    //! Direct access to an entry in the RHS/solution vector (Eigen3.hpp:293-296)
    number getValue(const std::size_t index) const {
        if (solved_) return x_[index];
        this->verifyCurrent();
        double v = 0.;
        b200_detail::check(isl_rhs_value(b200_detail::engine(), static_cast<int64_t>(index), &v));
        return v;
    }

    //! @name Hand-off of the finished system (canonical CSR: rows and columns ascending, explicit zeros kept)
    //@{
    void getCSR(std::vector<int64_t>& rowptr, std::vector<int32_t>& col, std::vector<double>& val,
                std::vector<double>& rhs) const {
        this->verifyCurrent();
        int64_t n = 0, nnz = 0;
        b200_detail::check(isl_finish(b200_detail::engine(), &n, &nnz));
        rowptr.assign(static_cast<std::size_t>(n) + 1, 0);
        col.assign(static_cast<std::size_t>(nnz), 0);
        val.assign(static_cast<std::size_t>(nnz), 0.);
        rhs.assign(static_cast<std::size_t>(n), 0.);
        b200_detail::check(isl_get_csr(b200_detail::engine(), &rowptr[0], nnz ? &col[0] : NULL, nnz ? &val[0] : NULL,
                                       n ? &rhs[0] : NULL));
    }
    //! device pointers for a device-resident solver (valid until the next solver is constructed)
    void getDeviceCSR(int64_t** rowptr, int32_t** col, double** val, double** rhs) const {
        b200_detail::check(isl_get_device_csr(b200_detail::engine(), rowptr, col, val, rhs));
    }
    //@}

    //! @name Linear solves: outside the assembly path, delegated to the hook (host) -- Eigen3.hpp:157-289
    //@{
    //! cgSolve() runs on the device (isl_solve_cg) unless a solve hook is installed and nativeCG() is false
    static bool& nativeCG() {
        static bool flag = true;
        return flag;
    }
    void choleskySolve() { this->solve_("cholesky"); }
    void luSolve() { this->solve_("lu"); }
    void superLUSolve() { this->solve_("superlu"); }
    int cgSolve() { return this->solve_("cg"); }
    int biCGStabSolve() { return this->solve_("bicgstab"); }
    //@}

    //! @name Debug routines for printing (Eigen3.hpp:298-327); the matrix is written column by column like Eigen's
    //@{
    void systemInfo(std::ostream& out) const {
        int64_t n = 0, nnz = 0;
        b200_detail::check(isl_finish(b200_detail::engine(), &n, &nnz));
        out << n << " X " << n << " sparse matrix with " << nnz << " non-zero entries \n";
    }
    void debugLHS(std::ostream& out) const {
        std::vector<int64_t> rowptr;
        std::vector<int32_t> col;
        std::vector<double> val, rhs;
        this->getCSR(rowptr, col, val, rhs);
        const std::size_t n = rhs.size();
        std::vector<std::size_t> start(n + 1, 0);
        for (std::size_t k = 0; k < col.size(); k++) start[col[k] + 1]++;
        for (std::size_t j = 0; j < n; j++) start[j + 1] += start[j];
        std::vector<int64_t> r(col.size());
        std::vector<double> v(col.size());
        std::vector<std::size_t> pos(start.begin(), start.end() - 1);
        for (std::size_t i = 0; i < n; i++)
            for (int64_t k = rowptr[i]; k < rowptr[i + 1]; k++) {
                const std::size_t q = pos[col[k]]++;
                r[q] = static_cast<int64_t>(i);
                v[q] = val[k];
            }
        for (std::size_t j = 0; j < n; j++)
            for (std::size_t q = start[j]; q < start[j + 1]; q++) out << r[q] << " " << j << " " << v[q] << "\n";
    }
    void debugRHS(std::ostream& out) const {
        const std::vector<double>& b = this->hostRhs_();
        for (std::size_t i = 0; i < b.size(); i++) out << i << " " << b[i] << "\n";
    }
    //! Eigen3.hpp:324-327 / TripletContainer.hpp:390-400: "row col value" sorted by (row, col)
    std::ostream& debugTriplet(std::ostream& out) const {
        std::vector<int64_t> rowptr;
        std::vector<int32_t> col;
        std::vector<double> val, rhs;
        this->getCSR(rowptr, col, val, rhs);
        for (std::size_t i = 0; i + 1 < rowptr.size(); i++)
            for (int64_t k = rowptr[i]; k < rowptr[i + 1]; k++) out << i << " " << col[k] << " " << val[k] << std::endl;
        return out;
    }
    //@}

private:
    const std::vector<double>& hostRhs_() const {
        if (!solved_) {
            this->verifyCurrent();
            x_.assign(size_, 0.);
            if (size_) b200_detail::check(isl_get_csr(b200_detail::engine(), NULL, NULL, NULL, &x_[0]));
        }
        return x_;
    }
    int solve_(const char* method) {
        if (std::string(method) == "cg" && (nativeCG() || solveHook() == NULL)) {
            this->verifyCurrent();
            int64_t iterations = 0;
            double error = 0.;
            b200_detail::check(isl_solve_cg(b200_detail::engine(), 0., 0, &iterations, &error));
            x_.assign(size_, 0.);
            if (size_) b200_detail::check(isl_get_csr(b200_detail::engine(), NULL, NULL, NULL, &x_[0]));
            solved_ = true;
            return static_cast<int>(iterations);
        }
        VERIFY_MSG(solveHook() != NULL, std::string("B200 solver: linear solve '") + method +
                                            "' requested but no solve hook installed (the engine covers the assembly path)");
        std::vector<int64_t> rowptr;
        std::vector<int32_t> col;
        std::vector<double> val;
        this->getCSR(rowptr, col, val, x_);
        const int it = solveHook()(method, size_, rowptr, col, val, x_);
        solved_ = true;
        return it;
    }

    static unsigned long nextId_() {
        static unsigned long n = 0;
        unsigned long id;
#ifdef _OPENMP
#pragma omp critical(isl_b200_solver_id)
#endif
        id = ++n;
        return id;
    }

    std::size_t size_, nnz_ = 0;
    mutable std::vector<double> x_;  //!< host copy of rhs / the solution after a solve
    bool solved_;
    unsigned long id_;
    int device_;
};

}  // namespace solver

//------------------------------------------------------------------------------------------------------------------
namespace asmb {

/** Guard against a silent host path.  The overloads below must be VISIBLE WHERE THE CALL IS WRITTEN: the reference
 *  calls `base::asmb::stiffnessMatrixComputation<FTB>(...)` with a qualified name, so inside its own templates
 *  (base/BoundaryValueProblem.hpp:271, the drivers) only overloads declared before that header are considered.  If this
 *  binding is included too late, the reference's generic version would run its element loop on the host and feed the
 *  engine through insertToLHS -- correct numbers, but a CPU assembly.  The generic version instantiates this functor, so
 *  the misuse is a compile error instead: include this header first (e.g. `g++ -include insilico_b200_reference.hpp`). */
template <typename QUAD, typename FIELDTUPLE>
class StiffnessMatrix<QUAD, base::solver::B200, FIELDTUPLE> {
    static_assert(sizeof(QUAD) == 0,
                  "base::asmb::StiffnessMatrix<..., base::solver::B200, ...>: the reference's host element loop was selected for "
                  "the B200 solver. Include insilico_b200_reference.hpp BEFORE the reference headers that call "
                  "stiffnessMatrixComputation (compile with -include insilico_b200_reference.hpp); the engine has no CPU fallback.");
};

namespace b200_detail {
template <typename FIELDTUPLEBINDER, typename FIELDBINDER>
typename FIELDTUPLEBINDER::Tuple probeTuple(const FIELDBINDER& fb, bool last) {
    typename FIELDBINDER::FieldIterator it = fb.elementsBegin();
    if (last) std::advance(it, std::distance(fb.elementsBegin(), fb.elementsEnd()) - 1);
    return FIELDTUPLEBINDER::makeTuple(*it);
}
}  // namespace b200_detail

namespace b200_detail {
//! Material factor that varies in space: heat::Laplace with setConductivityFunction (heat/Laplace.hpp:85-126) evaluates a
//! boost::function per quadrature point.  The function object is private, so kappa(e, q) is recovered through the public
//! interface: K = kappa * K1 with K1 from the unit-conductivity kernel at the same point (two local matrices per point,
//! all host threads; with ISL_B200_HAVE_KERNEL_ACCESSORS a maintainer's `conductivityFunction()` accessor is called
//! directly).  `varies` first looks at 16 elements spread over the mesh x all points.
template <typename KERNEL>
struct SampledFactor {
    template <typename FTB, typename QUADRATURE, typename FIELDBINDER>
    static bool varies(const KERNEL&, const QUADRATURE&, const FIELDBINDER&, double) { return false; }
    template <typename FTB, typename QUADRATURE, typename FIELDBINDER>
    static void sample(const KERNEL&, const QUADRATURE&, const FIELDBINDER&, std::vector<double>&) {}
};
template <typename TUPLE>
struct SampledFactor<heat::Laplace<TUPLE> > {
    typedef heat::Laplace<TUPLE> Kernel;
    template <typename FTB, typename XI>
    static double at(const Kernel& k, const Kernel& unit, const typename FTB::Tuple& tuple, const XI& xi) {
#ifdef ISL_B200_HAVE_KERNEL_ACCESSORS
        (void)unit;
        return k.conductivityFunction()(tuple.geomElementPtr(), xi);
#else
        typedef typename TUPLE::TestElement TestElement;
        typedef typename TUPLE::TrialElement TrialElement;
        const unsigned nr = TestElement::numDoFs * TestElement::DegreeOfFreedom::size;
        const unsigned nc = TrialElement::numDoFs * TrialElement::DegreeOfFreedom::size;
        base::MatrixD K = base::MatrixD::Zero(nr, nc), K1 = base::MatrixD::Zero(nr, nc);
        k.tangentStiffness(tuple, xi, 1.0, K);
        unit.tangentStiffness(tuple, xi, 1.0, K1);
        unsigned bi = 0;
        for (unsigned i = 1; i < std::min(nr, nc); i++) if (std::abs(K1(i, i)) > std::abs(K1(bi, bi))) bi = i;
        return K(bi, bi) / K1(bi, bi);
#endif
    }
    template <typename FTB, typename QUADRATURE, typename FIELDBINDER>
    static bool varies(const Kernel& k, const QUADRATURE& quadrature, const FIELDBINDER& fb, double kappa0) {
        const Kernel unit(1.0);
        const long numE = static_cast<long>(std::distance(fb.elementsBegin(), fb.elementsEnd()));
        const long step = std::max(1L, numE / 16);
        for (long e = 0; e < numE; e += step)
            for (typename QUADRATURE::Iter q = quadrature.begin(); q != quadrature.end(); ++q) {
                const double v = at<FTB>(k, unit, FTB::makeTuple(*(fb.elementsBegin() + e)), q->second);
                if (!(std::abs(v - kappa0) <= 1e-12 * std::abs(kappa0))) return true;
            }
        return false;
    }
    template <typename FTB, typename QUADRATURE, typename FIELDBINDER>
    static void sample(const Kernel& k, const QUADRATURE& quadrature, const FIELDBINDER& fb, std::vector<double>& values) {
        const Kernel unit(1.0);
        const long numE = static_cast<long>(std::distance(fb.elementsBegin(), fb.elementsEnd()));
        std::vector<typename QUADRATURE::Iter> qpts;
        for (typename QUADRATURE::Iter q = quadrature.begin(); q != quadrature.end(); ++q) qpts.push_back(q);
        const std::size_t nq = qpts.size();
        values.assign(static_cast<std::size_t>(numE) * nq, 0.);
        const typename FIELDBINDER::FieldIterator it0 = fb.elementsBegin();
#pragma omp parallel for schedule(static)
        for (long e = 0; e < numE; e++) {
            const typename FTB::Tuple tuple = FTB::makeTuple(*(it0 + e));
            for (std::size_t q = 0; q < nq; q++) values[e * nq + q] = at<FTB>(k, unit, tuple, qpts[q]->second);
        }
    }
};
}  // namespace b200_detail

namespace b200_detail {
//! kernels whose constant cannot be probed on an element at rest (fluid::Convection): look at more elements.
//! resolve() returns false when the contribution is identically zero (every element at rest): nothing to launch.
template <typename KERNEL>
struct LateProbe {
    template <typename FTB, typename FIELDBINDER>
    static bool resolve(const KERNEL&, const FIELDBINDER&, double*) { return true; }
};
template <typename TUPLE>
struct LateProbe<fluid::Convection<TUPLE> > {
    template <typename FTB, typename FIELDBINDER>
    static bool resolve(const fluid::Convection<TUPLE>& k, const FIELDBINDER& fb, double* p) {
        if (p[2] == 0.) return true;
        p[2] = 0.;
        for (typename FIELDBINDER::FieldIterator it = fb.elementsBegin(); it != fb.elementsEnd(); ++it)
            if (base::solver::B200KernelTraits<fluid::Convection<TUPLE> >::probe(k, FTB::makeTuple(*it), p)) return true;
        return false;
    }
};
}  // namespace b200_detail

//! base/asmb/StiffnessMatrix.hpp:49-87 for SOLVER = base::solver::B200
template <typename FIELDTUPLEBINDER, typename QUADRATURE, typename FIELDBINDER, typename KERNEL>
void stiffnessMatrixComputation(const QUADRATURE& quadrature, base::solver::B200& solver, const FIELDBINDER& fieldBinder,
                                const KERNEL& kernelObj, const bool incremental = true) {
    namespace D = base::solver::b200_detail;
    solver.verifyCurrent();
    if (fieldBinder.elementsBegin() == fieldBinder.elementsEnd()) return;
    D::synchronise(fieldBinder);
    double params[4] = {0., 0., 0., 0.};
    const int id = base::solver::B200KernelTraits<KERNEL>::describe(
        kernelObj, b200_detail::probeTuple<FIELDTUPLEBINDER>(fieldBinder, false),
        b200_detail::probeTuple<FIELDTUPLEBINDER>(fieldBinder, true), params);
    if (!b200_detail::LateProbe<KERNEL>::template resolve<FIELDTUPLEBINDER>(kernelObj, fieldBinder, params)) return;
    typedef D::TupleIndices<FIELDTUPLEBINDER> TI;
    typedef b200_detail::SampledFactor<KERNEL> Sampled;
    if (params[3] != 0. || Sampled::template varies<FIELDTUPLEBINDER>(kernelObj, quadrature, fieldBinder, params[0])) {
        std::vector<double> values;   // the caller's material function, evaluated on the host at every quadrature point
        Sampled::template sample<FIELDTUPLEBINDER>(kernelObj, quadrature, fieldBinder, values);
        VERIFY_MSG(!values.empty(), "B200 engine: this kernel object has a material factor that varies in space and no sampled form");
        D::check(isl_assemble_matrix_sampled(D::engine(), id, &values[0], D::QuadratureDegree<QUADRATURE>::value, TI::test,
                                             TI::trial, incremental ? 1 : 0));
        return;
    }
    D::check(isl_assemble_matrix_aux(D::engine(), id, params, D::QuadratureDegree<QUADRATURE>::value, TI::test, TI::trial,
                                     TI::aux, incremental ? 1 : 0));
}

//! base/asmb/ForceIntegrator.hpp:37-71 for SOLVER = base::solver::B200 (forces enter the rhs with factor -1)
template <typename FIELDTUPLEBINDER, typename QUADRATURE, typename FIELDBINDER, typename KERNEL>
void computeResidualForces(const QUADRATURE&, base::solver::B200& solver, const FIELDBINDER& fieldBinder,
                           const KERNEL& kernelObj) {
    namespace D = base::solver::b200_detail;
    solver.verifyCurrent();
    if (fieldBinder.elementsBegin() == fieldBinder.elementsEnd()) return;
    D::synchronise(fieldBinder);
    double params[4] = {0., 0., 0., 0.};
    const int id = base::solver::B200KernelTraits<KERNEL>::describe(
        kernelObj, b200_detail::probeTuple<FIELDTUPLEBINDER>(fieldBinder, false),
        b200_detail::probeTuple<FIELDTUPLEBINDER>(fieldBinder, true), params);
    VERIFY_MSG(params[3] == 0., "B200 engine: residual forces with a material factor that varies in space are not supported");
    if (!b200_detail::LateProbe<KERNEL>::template resolve<FIELDTUPLEBINDER>(kernelObj, fieldBinder, params)) return;
    typedef D::TupleIndices<FIELDTUPLEBINDER> TI;
    D::check(isl_assemble_residual_aux(D::engine(), id, params, D::QuadratureDegree<QUADRATURE>::value, TI::test, TI::trial,
                                       TI::aux, -1.0));
}

//! base/asmb/BodyForce.hpp:65-84 for SOLVER = base::solver::B200.  The caller's function f(x) runs on the host, once per
//! element and quadrature point (as in BodyForce.hpp:172-205); the integration runs on the device.  A force that turns
//! out to be the same at every point takes the constant entry point (fused with the stiffness launch on the Q1 path).
namespace b200_detail {
//! shared by bodyForceComputation (f(x)) and bodyForceComputation2 (f(element, xi)): EVAL(gep, xi) -> force vector
template <typename FIELDTUPLEBINDER, typename QUADRATURE, typename FIELDBINDER, typename EVAL>
void sampledBodyForce(const QUADRATURE& quadrature, base::solver::B200& solver, const FIELDBINDER& fieldBinder, const EVAL& eval) {
    namespace D = base::solver::b200_detail;
    solver.verifyCurrent();
    typedef typename FIELDTUPLEBINDER::Tuple Tuple;
    typedef typename Tuple::GeomElement GeomElement;
    typedef typename Tuple::TestElement TestElement;
    if (fieldBinder.elementsBegin() == fieldBinder.elementsEnd()) return;
    D::synchronise(fieldBinder);
    const unsigned ds = TestElement::DegreeOfFreedom::size;
    const std::size_t n = static_cast<std::size_t>(std::distance(fieldBinder.elementsBegin(), fieldBinder.elementsEnd()));
    const std::size_t nq = static_cast<std::size_t>(std::distance(quadrature.begin(), quadrature.end()));
    // the caller's function is evaluated at every quadrature point of every element (it is taken by const reference and,
    // on large meshes, called concurrently by all host threads, like the kernel objects in the reference's own OpenMP
    // element loop, base/auxi/parallel.hpp:37-57)
    std::vector<double> values;
    const typename FIELDBINDER::FieldIterator it0 = fieldBinder.elementsBegin();
    const long numE = static_cast<long>(n);
    std::vector<typename QUADRATURE::Iter> qpts;
    for (typename QUADRATURE::Iter qIter = quadrature.begin(); qIter != quadrature.end(); ++qIter) qpts.push_back(qIter);
    bool constant = true;
    double first[8] = {0., 0., 0., 0., 0., 0., 0., 0.};
    VERIFY_MSG(ds <= 8, "B200 engine: body force with more than eight components");
    if (numE > ISL_B200_PARALLEL_SCAN_MIN) {
        values.resize(n * nq * ds);
#pragma omp parallel for schedule(static)
        for (long e = 0; e < numE; e++) {
            const GeomElement* gep = FIELDTUPLEBINDER::makeTuple(*(it0 + e)).geomElementPtr();
            for (std::size_t q = 0; q < nq; q++) {
                const typename EVAL::result_type v = eval(gep, qpts[q]->second);
                for (unsigned d = 0; d < ds; d++) values[(e * nq + q) * ds + d] = v[d];
            }
        }
        for (std::size_t k = ds; k < values.size() && constant; k++) constant = (values[k] == values[k % ds]);
        for (unsigned d = 0; d < ds; d++) first[d] = values[d];
    } else {
        // one thread: the table of values is materialised only once the force turns out to vary, so a constant force (the
        // common case) costs its evaluations and nothing else (64^3 hexahedra: 7 -> 3 ms of host time per call)
        typename FIELDBINDER::FieldIterator it = it0;
        for (long e = 0; e < numE; e++, ++it) {
            const GeomElement* gep = FIELDTUPLEBINDER::makeTuple(*it).geomElementPtr();
            for (std::size_t q = 0; q < nq; q++) {
                const typename EVAL::result_type v = eval(gep, qpts[q]->second);
                if (e == 0 && q == 0)
                    for (unsigned d = 0; d < ds; d++) first[d] = v[d];
                if (constant) {
                    bool same = true;
                    for (unsigned d = 0; d < ds; d++) same = same && (v[d] == first[d]);
                    if (same) continue;
                    constant = false;
                    values.resize(n * nq * ds);
                    for (std::size_t k = 0; k < (e * nq + q) * ds; k++) values[k] = first[k % ds];
                }
                for (unsigned d = 0; d < ds; d++) values[(e * nq + q) * ds + d] = v[d];
            }
        }
    }
    typedef D::TupleIndices<FIELDTUPLEBINDER> TI;
    if (constant) {
        double f[3] = {0., 0., 0.};
        for (unsigned d = 0; d < ds && d < 3; d++) f[d] = first[d];
        D::check(isl_assemble_bodyforce(D::engine(), f, D::QuadratureDegree<QUADRATURE>::value, TI::test));
    } else {
        D::check(isl_assemble_bodyforce_sampled(D::engine(), &values[0], D::QuadratureDegree<QUADRATURE>::value, TI::test));
    }
}
template <typename GEOMELEMENT, typename FUN>
struct AtPhysicalPoint {   // f(x), x = x(xi) like base::auxi::EvaluateDirectly (base/auxi/FunEvaluationPolicy.hpp:47-59)
    typedef typename FUN::result_type result_type;
    const FUN& fun;
    template <typename XI>
    result_type operator()(const GEOMELEMENT* gep, const XI& xi) const { return fun(base::Geometry<GEOMELEMENT>()(gep, xi)); }
};
template <typename GEOMELEMENT, typename FUN>
struct AtElementPoint {    // f(element, xi) like base::auxi::EvaluateViaElement (FunEvaluationPolicy.hpp:74-86)
    typedef typename FUN::result_type result_type;
    const FUN& fun;
    template <typename XI>
    result_type operator()(const GEOMELEMENT* gep, const XI& xi) const { return fun(gep, xi); }
};
}  // namespace b200_detail

template <typename FIELDTUPLEBINDER, typename QUADRATURE, typename FIELDBINDER, typename FUN>
void bodyForceComputation(const QUADRATURE& quadrature, base::solver::B200& solver, const FIELDBINDER& fieldBinder,
                          const FUN& forceFun) {
    typedef typename FIELDTUPLEBINDER::Tuple::GeomElement GeomElement;
    const b200_detail::AtPhysicalPoint<GeomElement, FUN> eval = {forceFun};
    b200_detail::sampledBodyForce<FIELDTUPLEBINDER>(quadrature, solver, fieldBinder, eval);
}

//! base/asmb/BodyForce.hpp:92-111 (force function of an element pointer and a local coordinate) for SOLVER = B200
template <typename FIELDTUPLEBINDER, typename QUADRATURE, typename FIELDBINDER, typename FUN>
void bodyForceComputation2(const QUADRATURE& quadrature, base::solver::B200& solver, const FIELDBINDER& fieldBinder,
                           const FUN& forceFun) {
    typedef typename FIELDTUPLEBINDER::Tuple::GeomElement GeomElement;
    const b200_detail::AtElementPoint<GeomElement, FUN> eval = {forceFun};
    b200_detail::sampledBodyForce<FIELDTUPLEBINDER>(quadrature, solver, fieldBinder, eval);
}

//! base/asmb/NeumannForce.hpp:33-66 for SOLVER = base::solver::B200: the terms of an applied surface force over ALL surface
//! elements of a SurfaceFieldBinder in ONE device launch (isl_assemble_neumann_rows), instead of one insertToRHS per
//! surface element.  Read here from the reference's own objects, per surface element (base/mesh/SurfaceElement.hpp): its
//! nodes, their coordinates in the parameter space of the domain element, the equation numbers of the test element's
//! DoFs; the caller's function f(x, normal) is evaluated on the host with the reference's own Geometry / SurfaceNormal
//! functors (NeumannForce.hpp:152-163), the shape functions, the metric, the weighting and the scatter run on the device.
//! Surface elements whose test DoFs are slaves of master DoFs take the reference's generic loop.
namespace b200_detail {
template <typename FIELDTUPLEBINDER, typename SURFACEQUADRATURE, typename FIELDBINDER, typename FUN>
void neumannForces(const SURFACEQUADRATURE& surfaceQuadrature, base::solver::B200& solver, const FIELDBINDER& fieldBinder,
                   const FUN& ff) {
    namespace D = base::solver::b200_detail;
    solver.verifyCurrent();
    typedef typename FIELDTUPLEBINDER::Tuple Tuple;
    typedef typename Tuple::GeomElement SurfaceElement;
    typedef typename Tuple::TestElement TestElement;
    typedef typename TestElement::DegreeOfFreedom DoF;
    typedef base::GeomTraits<SurfaceElement> GT;
    typedef typename NeumannForce<Tuple>::ForceFun ForceFun;
    const unsigned dim = GT::globalDim, ds = DoF::size;
    const std::size_t n = static_cast<std::size_t>(std::distance(fieldBinder.elementsBegin(), fieldBinder.elementsEnd()));
    if (n == 0) return;
    const std::size_t nq = static_cast<std::size_t>(std::distance(surfaceQuadrature.begin(), surfaceQuadrature.end()));
    const Tuple first = FIELDTUPLEBINDER::makeTuple(*fieldBinder.elementsBegin());
    const std::size_t P = static_cast<std::size_t>(std::distance(first.geomElementPtr()->nodesBegin(), first.geomElementPtr()->nodesEnd()));
    const std::size_t ndpe = static_cast<std::size_t>(std::distance(first.testElementPtr()->doFsBegin(), first.testElementPtr()->doFsEnd()));
    std::vector<double> sx(n * P * dim), sp(n * P * dim), values(n * nq * ds);
    std::vector<int32_t> rows(n * ndpe * ds);
    bool slaves = false;
    std::size_t k = 0;
    for (typename FIELDBINDER::FieldIterator it = fieldBinder.elementsBegin(); it != fieldBinder.elementsEnd(); ++it, ++k) {
        const Tuple tuple = FIELDTUPLEBINDER::makeTuple(*it);
        const SurfaceElement* surfEp = tuple.geomElementPtr();
        const TestElement* testEp = tuple.testElementPtr();
        std::size_t p = 0;
        typename SurfaceElement::ParamConstIter par = surfEp->parametricBegin();
        for (typename SurfaceElement::NodePtrConstIter nd = surfEp->nodesBegin(); nd != surfEp->nodesEnd(); ++nd, ++par, ++p) {
            double x[3] = {0., 0., 0.};
            (*nd)->getX(&x[0]);
            for (unsigned d = 0; d < dim; d++) { sx[(k * P + p) * dim + d] = x[d]; sp[(k * P + p) * dim + d] = (*par)[d]; }
        }
        std::size_t s = 0;
        for (typename TestElement::DoFPtrConstIter dp = testEp->doFsBegin(); dp != testEp->doFsEnd(); ++dp, ++s)
            for (unsigned c = 0; c < ds; c++) {
                rows[(k * ndpe + s) * ds + c] = (*dp)->isActive(c) ? static_cast<int32_t>((*dp)->getIndex(c)) : -1;
                if ((*dp)->isConstrained(c) && !slaves) {
                    std::vector<std::pair<base::number, std::size_t> > masters;
                    const_cast<DoF*>(*dp)->getConstraint(c)->getWeightedDoFIDs(masters);
                    slaves = !masters.empty();
                }
            }
        std::size_t q = 0;
        for (typename SURFACEQUADRATURE::Iter qIter = surfaceQuadrature.begin(); qIter != surfaceQuadrature.end(); ++qIter, ++q) {
            const typename GT::GlobalVecDim x = base::Geometry<SurfaceElement>()(surfEp, qIter->second);
            typename GT::GlobalVecDim normal;
            base::SurfaceNormal<SurfaceElement>()(surfEp, qIter->second, normal);
            const typename ForceFun::result_type f = ff(x, normal);
            for (unsigned c = 0; c < ds; c++) values[(k * nq + q) * ds + c] = f[c];
        }
    }
    if (slaves) {   // few and rare: the reference's own loop (-> insertToRHS per surface element) keeps their master weights
        const ForceFun forceFun = ff;
        base::asmb::neumannForceComputation<FIELDTUPLEBINDER, SURFACEQUADRATURE, base::solver::B200, FIELDBINDER>(surfaceQuadrature, solver, fieldBinder, forceFun);
        return;
    }
    typedef typename SurfaceElement::DomainElement DomainElement;
    D::check(isl_assemble_neumann_rows(D::engine(), static_cast<int>(DomainElement::shape), static_cast<int>(DomainElement::GeomFun::degree),
                                       static_cast<int64_t>(n), &sx[0], &sp[0], D::QuadratureDegree<SURFACEQUADRATURE>::value,
                                       static_cast<int>(TestElement::FEFun::degree), static_cast<int>(ds), &rows[0], ISL_NEUMANN_SAMPLED,
                                       &values[0]));
}
}  // namespace b200_detail

//! The overloads.  The reference's signature takes the force function as NeumannForce<Tuple>::ForceFun (a boost::function) in
//! a non-deduced context, which leaves that template and these unordered; they are therefore chosen by the better
//! conversion of the last argument: a functor or bind expression (reference/05-mixedPoisson/mixedPoisson.cpp:184-187) needs
//! no conversion here, a non-const boost::function lvalue (base/BoundaryValueProblem.hpp:329-331) binds to the less
//! cv-qualified reference.  (A `const` boost::function lvalue of exactly that type stays ambiguous: pass it un-const.)
template <typename FIELDTUPLEBINDER, typename SURFACEQUADRATURE, typename FIELDBINDER, typename FUN>
void neumannForceComputation(const SURFACEQUADRATURE& surfaceQuadrature, base::solver::B200& solver, const FIELDBINDER& fieldBinder,
                             const FUN& ff) {
    b200_detail::neumannForces<FIELDTUPLEBINDER>(surfaceQuadrature, solver, fieldBinder, ff);
}
template <typename FIELDTUPLEBINDER, typename SURFACEQUADRATURE, typename FIELDBINDER, typename FUN>
void neumannForceComputation(const SURFACEQUADRATURE& surfaceQuadrature, base::solver::B200& solver, const FIELDBINDER& fieldBinder,
                             FUN& ff) {
    b200_detail::neumannForces<FIELDTUPLEBINDER>(surfaceQuadrature, solver, fieldBinder, ff);
}

}  // namespace asmb
}  // namespace base

#endif
