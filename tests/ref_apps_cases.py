"""Inputs of the reference applications used by tools/make_ref_app_goldens.py and tests/test_reference_run.py."""
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "tests", "golden", "ref")

# name -> (executable, input files copied into the working directory, arguments)
CASES = {
    "dirichlet_tet6": ("dirichlet", [], ["tet6.smf"]),                       # reference/04-heat/dirichlet.cpp
    "linearElastic2D_quad010": ("linearElastic2D", ["quad.010.smf"], ["quad.010.smf"]),   # 06-elastic/linearElastic.cpp
    "linearElastic3D_cube004": ("linearElastic3D", ["cube.004.smf"], ["cube.004.smf"]),
    "linearElastic3D_cube008": ("linearElastic3D", ["cube.008.smf"], ["cube.008.smf"]),
    "compressible_quad010": ("compressible", ["quad.010.smf", "inputCompRefD.dat"],       # 06-elastic/compressible.cpp
                             ["quad.010.smf", "inputCompRefD.dat"]),
    # the same problem through the reference's driver facade (solid/CompressibleDriver.hpp, base/BoundaryValueProblem.hpp):
    # prints the number of Newton iterations per load step
    "compressibleWithDriver_quad010": ("compressibleWithDriver", ["quad.010.smf", "inputCompRefD.dat"],
                                       ["quad.010.smf", "inputCompRefD.dat"]),
    # reference/05-mixedPoisson: heat::Laplace + a body force f(x) (sampled on the host, integrated on the device) +
    # element-wise force and Neumann terms that reach the solver through insertToRHS; prints L2 / H1 errors
    "mixedPoisson_square020": ("mixedPoisson", ["square_020.smf"], ["square_020.smf"]),
    # the same through heat::PoissonDriver (kernel heat::Static<mat::thermal::IsotropicConstant>); writes test.vtk
    "mixedPoissonWithDriver_square020": ("mixedPoissonWithDriver", ["square_020.smf"], ["square_020.smf"], "test.vtk"),
    # traction controlled: asmb::neumannForceComputation runs in the reference's own code and reaches the solver
    # through insertToRHS (the host-side "odd contribution" interface)
    "compressible_neumann_quad010": ("compressible", ["quad.010.smf", "inputCompRefN.dat"],
                                     ["quad.010.smf", "inputCompRefN.dat"]),
}


# applications that need engine entry points written without GPU minutes (isl_assemble_bodyforce_sampled): their GPU
# test lives in tests/test_zz_linear_constraints.py so that it runs after the established suites
LATE = ("mixedPoisson_square020", "mixedPoissonWithDriver_square020")


def output_file(name):
    """file written by the application whose content belongs to the compared output (or None)"""
    return CASES[name][3] if len(CASES[name]) > 3 else None


def prepare(name, workdir):
    exe, files, args = CASES[name][:3]
    for f in files:
        shutil.copy(os.path.join(REF, f), os.path.join(workdir, f))
    if name == "dirichlet_tet6":
        from insilico_b200 import meshgen
        coords, conn = meshgen.unit_cube_tet(6, 6, 6)
        coords = meshgen.perturb_interior(coords, 1.0 / 6, max_dist=0.15)
        with open(os.path.join(workdir, "tet6.smf"), "w") as f:
            f.write("! elementShape tetrahedron\n! elementNumPoints 4\n%d %d\n" % (len(coords), len(conn)))
            for x in coords:
                f.write("%.17g %.17g %.17g\n" % tuple(x))
            for e in conn:
                f.write(" ".join(str(int(v)) for v in e) + "\n")
    return exe, args


def same_output(a, b, rel=2e-5, noise=1e-7):
    """token-wise comparison: text equal, numbers equal to the printed 6 digits (rel); numbers below `noise` on both
    sides (norms of converged Newton iterates, i.e. rounding noise of the linear solve) count as equal"""
    ta, tb = a.split(), b.split()
    if len(ta) != len(tb):
        return False
    for x, y in zip(ta, tb):
        if x == y:
            continue
        try:
            fx, fy = float(x), float(y)
        except ValueError:
            return False
        if abs(fx) < noise and abs(fy) < noise:
            continue
        if abs(fx - fy) > rel * max(abs(fx), abs(fy)):
            return False
    return True
