"""Slaves of master DoFs in the CUDA engine, checked on the CPU: tests/emu/constraints_emu.cpp replays the engine's
pattern expansion and scatter with the index routines the kernels use (insilico_b200/csrc/isl_constraints.hpp), on
random local matrices; the expectation is an independent restatement of asmb::assembleMatrix / assembleForces
(base/asmb/assembleMatrix.hpp:56-130,212-338, assembleForces.hpp:58-139) in plain Python."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from tests import flows

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "emu", "constraints_emu.cpp")
LIB = os.path.join(ROOT, "tests", "emu", "_build", "libconstraints_emu.so")
DEPS = [SRC, os.path.join(ROOT, "insilico_b200", "csrc", "isl_constraints.hpp")]


@pytest.fixture(scope="module")
def emu():
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in DEPS):
        subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-Wall", "-o", LIB, SRC],
                       check=True)
    lib = ctypes.CDLL(LIB)
    lib.emu_constraints.restype = ctypes.c_int64
    return lib


def dense_tables(f):
    """what isl_field_set_constraints builds: dense cptr over all DoF components, masters as equation numbers"""
    n = f["n_obj"] * f["ds"]
    if not f["linear"]:
        return None, None, None
    con_dof, con_ptr, meq, w = flows.Case.constraint_arrays(f)
    cnt = np.zeros(n + 1, dtype=np.int32)
    cnt[con_dof + 1] = np.diff(con_ptr)
    cptr = np.cumsum(cnt).astype(np.int32)
    cm = np.zeros(len(meq), dtype=np.int32); cw = np.zeros(len(meq))
    for k, d in enumerate(con_dof):
        cm[cptr[d]:cptr[d + 1]] = meq[con_ptr[k]:con_ptr[k + 1]]
        cw[cptr[d]:cptr[d + 1]] = w[con_ptr[k]:con_ptr[k + 1]]
    return cptr, cm, cw


def reference_semantics(ft, fc, K, F, incremental, n_eqn):
    """assembleMatrix / assembleForces per element, entries summed per (row, col) in element order"""
    def targets(f):
        masters = {obj * f["ds"] + comp: [(int(f["eqn"][mo, mc]), wt) for mo, mc, wt in m] for obj, comp, _, m in f["linear"]}
        st, eq = f["status"].reshape(-1), f["eqn"].reshape(-1)
        return lambda k: [(int(eq[k]), 1.0)] if st[k] == 0 else (masters.get(k, []) if st[k] == 1 else [])
    tt, tc = targets(ft), targets(fc)
    A, b = {}, np.zeros(n_eqn)
    ne = ft["elem_dof"].shape[0]
    st_c = fc["status"].reshape(-1); pc = fc["presc"].reshape(-1); vc = fc["values"].reshape(-1)
    for e in range(ne):
        rows = [int(o) * ft["ds"] + s for o in ft["elem_dof"][e] for s in range(ft["ds"])]
        cols = [int(o) * fc["ds"] + s for o in fc["elem_dof"][e] for s in range(fc["ds"])]
        for i, kr in enumerate(rows):
            for rt, wr in tt(kr):
                for j, k in enumerate(cols):
                    v = K[e, i, j]
                    if st_c[k] == 1:
                        g = pc[k] - vc[k] if incremental else pc[k]
                        b[rt] -= g * wr * v
                    for ct, wc in tc(k):
                        A[(rt, ct)] = A.get((rt, ct), 0.0) + wr * wc * v
                if F is not None:
                    b[rt] += wr * F[e, i]
    keys = sorted(A)
    return keys, np.array([A[k] for k in keys]), b


@pytest.mark.parametrize("name,n", [("laplace_q1_hex_linear", 4), ("laplace_q2_hex_linear", 2), ("stvenant_q1_hex_linear", 4),
                                    ("stokes_p2p1_tet_linear", 2), ("laplace_q1_hex", 4)])
@pytest.mark.parametrize("incremental", [0, 1])
def test_engine_constraint_scatter_equals_reference_semantics(emu, name, n, incremental):
    c = flows.build_case(name, n, True, False)
    pairs = [(0, 0)] if len(c.fields) == 1 else [(0, 0), (0, 1), (1, 0)]
    rng = np.random.default_rng(7)
    for t, cc in pairs:
        ft, fc = c.fields[t], c.fields[cc]
        ne = ft["elem_dof"].shape[0]
        nr, ncl = ft["elem_dof"].shape[1] * ft["ds"], fc["elem_dof"].shape[1] * fc["ds"]
        K = rng.standard_normal((ne, nr, ncl)); F = rng.standard_normal((ne, nr))
        keys, vals, b = reference_semantics(ft, fc, K, F, incremental, c.n_eqn)
        P = lambda a: None if a is None else a.ctypes.data_as(ctypes.c_void_p)
        arrs = []
        def A(a, dt):
            a = np.ascontiguousarray(a, dtype=dt); arrs.append(a); return a
        e32 = lambda f: A(np.where(f["status"] == 0, f["eqn"], -1), np.int32)
        tabs_t, tabs_c = dense_tables(ft), dense_tables(fc)
        cap = 4 * len(keys) + 1000
        rowptr = np.zeros(c.n_eqn + 1, dtype=np.int64); col = np.zeros(cap, dtype=np.int32); val = np.zeros(cap); rhs = np.zeros(c.n_eqn)
        nnz = emu.emu_constraints(
            ctypes.c_int64(ne), ctypes.c_int64(c.n_eqn), P(A(ft["elem_dof"], np.int32)), ft["elem_dof"].shape[1], ft["ds"], P(e32(ft)),
            P(A(ft["status"], np.uint8)), P(tabs_t[0]), P(tabs_t[1]), P(tabs_t[2]),
            P(A(fc["elem_dof"], np.int32)), fc["elem_dof"].shape[1], fc["ds"], P(e32(fc)), P(A(fc["status"], np.uint8)),
            P(A(fc["presc"], np.float64)), P(A(fc["values"], np.float64)), P(tabs_c[0]), P(tabs_c[1]), P(tabs_c[2]),
            incremental, P(A(K, np.float64)), P(A(F, np.float64)), ctypes.c_int64(cap), P(rowptr), P(col), P(val), P(rhs))
        assert nnz == len(keys), "pattern size differs"
        rows = np.repeat(np.arange(c.n_eqn), np.diff(rowptr))
        assert [(int(r), int(cl)) for r, cl in zip(rows, col[:nnz])] == keys, "pattern differs"
        scale = max(np.abs(vals).max(), 1.0)
        assert np.abs(val[:nnz] - vals).max() <= 1e-13 * scale
        assert np.abs(rhs - b).max() <= 1e-13 * max(np.abs(b).max(), 1.0)


# ---- register-tiled hyperelastic tangent (isl_tangent_tiled.cuh), per-thread routine replayed on the host ------------
@pytest.mark.parametrize("dim,nq,nt,nc", [(3, 27, 27, 27), (3, 11, 10, 10), (3, 8, 8, 8), (2, 4, 9, 9), (3, 5, 7, 4)])
def test_tiled_hyperelastic_tangent_equals_defining_formula(dim, nq, nt, nc):
    src = os.path.join(ROOT, "tests", "emu", "tangent_tiled_emu.cpp")
    lib_path = os.path.join(ROOT, "tests", "emu", "_build", "libtangent_tiled_emu.so")
    deps = [src, os.path.join(ROOT, "insilico_b200", "csrc", "isl_tangent_tiled.cuh")]
    os.makedirs(os.path.dirname(lib_path), exist_ok=True)
    if not os.path.exists(lib_path) or any(os.path.getmtime(d) > os.path.getmtime(lib_path) for d in deps):
        subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-x", "c++", "-Wall",
                        "-o", lib_path, src], check=True)
    lib = ctypes.CDLL(lib_path)
    rng = np.random.default_rng(dim * 100 + nt)
    Gt, Gc = rng.standard_normal((nq, nt, dim)), rng.standard_normal((nq, nc, dim))
    Q = rng.standard_normal((nq, 3, 3, 3, 3))            # Ceff[i, J, k, L], stride 3 also in 2-D
    det, w = rng.random(nq) + 0.5, rng.random(nq)
    K = np.zeros((nt * dim, nc * dim))
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    lib.emu_hypel_tile(dim, P(Gt), P(Gc), P(np.ascontiguousarray(Q)), P(det), P(w), nq, nt, nc, P(K))
    ref = np.einsum("q,qmJ,qiJkL,qnL->mink", det * w, Gt, Q[:, :dim, :dim, :dim, :dim], Gc).reshape(nt * dim, nc * dim)
    assert np.abs(K - ref).max() <= 1e-12 * np.abs(ref).max()
