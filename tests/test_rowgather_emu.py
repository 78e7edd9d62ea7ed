"""Row-gather formulation of the Q1-hex Laplace assembly (insilico_b200/csrc/isl_rowgather.cuh), checked on the CPU.

tests/emu/rowgather_emu.cpp compiles the kernel's per-thread routines and the host preprocessing with g++ and replays
the two kernels (row tables, assembly) thread by thread, including the per-warp staging.  The result must equal the
oracle's system (reference/04-heat/dirichlet.cpp flow) to 1e-12 with the identical pattern.  This checks the
algorithm, its constant tables and its index arithmetic without a GPU; the CUDA kernel itself is covered by the
`gpu` tests (ISL_Q1_ROWS=1 variant)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from insilico_b200 import engine as E
from insilico_b200 import meshgen
from tests import flows, helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "emu", "rowgather_emu.cpp")
LIB = os.path.join(ROOT, "tests", "emu", "_build", "librowgather_emu.so")
DEPS = [SRC] + [os.path.join(ROOT, "insilico_b200", "csrc", f) for f in
                ("isl_rowgather.cuh", "isl_patch_host.hpp", "isl_tables.hpp")]


@pytest.fixture(scope="module")
def emu():
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in DEPS):
        subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-w", "-pthread",
                        "-o", LIB, SRC], check=True)
    lib = ctypes.CDLL(LIB)
    lib.emu_rowgather.restype = ctypes.c_int
    return lib


def run_emu(lib, c, factor, f0, body, incremental, store_mode, rows_per_patch, nt, ref, general=0, elems=None,
            val=None, rhs=None):
    f = c.fields[0]
    rp, col = ref[0].astype(np.int64), ref[1].astype(np.int32)
    if val is None:
        val = np.zeros(len(col)) if store_mode else np.full(len(col), 0.5)
        if store_mode:
            val[:] = np.nan  # lazily-zeroed matrix: every entry must be WRITTEN
    if rhs is None:
        rhs = np.zeros(c.n_eqn)
    stats = np.zeros(5, dtype=np.int64)
    P = ctypes.c_void_p
    arr = lambda a: a.ctypes.data_as(P)
    coords = np.ascontiguousarray(c.coords)
    conn = np.ascontiguousarray(c.conn if elems is None else c.conn[elems], dtype=np.int32)
    eqn = np.ascontiguousarray(f["eqn"].reshape(-1), dtype=np.int32)
    status = np.ascontiguousarray(f["status"].reshape(-1), dtype=np.uint8)
    presc = np.ascontiguousarray(f["presc"].reshape(-1)); values = np.ascontiguousarray(f["values"].reshape(-1))
    rc = lib.emu_rowgather(ctypes.c_int64(len(coords)), ctypes.c_int64(len(conn)), arr(coords), arr(conn), arr(eqn),
                           arr(status), arr(presc), arr(values), ctypes.c_int64(c.n_eqn), arr(rp), arr(col),
                           ctypes.c_double(factor), ctypes.c_double(f0), ctypes.c_int(body), ctypes.c_int(incremental),
                           ctypes.c_int(store_mode), ctypes.c_int(rows_per_patch), ctypes.c_int(nt), ctypes.c_int(general), arr(val),
                           arr(rhs),
                           arr(stats))
    return rc, val, rhs, stats


def sheared(n, permute):
    """unit cube mapped by a general linear map: every element is a parallelepiped (affine, full D tensor)"""
    coords, conn, _ = meshgen.unit_cube_hex(n, n + 1, n - 1)
    A = np.array([[1.0, 0.3, -0.2], [0.1, 0.8, 0.25], [-0.15, 0.2, 1.3]])
    coords = coords @ A.T
    if permute:
        conn = meshgen.permute_elements(conn)
    c = flows.Case(E.HEX, 1, coords, conn)
    c.add_field(1, 1, dirichlet=lambda x: H.fund_sol_laplace(x, np.full(3, -0.5)), values=lambda x: 0.3 * x[:, :1] + 0.1)
    return c


@pytest.mark.parametrize("n,permute,rows_per_patch,nt", [(5, False, 400, 64), (7, True, 40, 64), (9, False, 100, 128),
                                                          (8, True, 37, 32),
                                                          (20, True, 16, 32)])  # 428 patches: threaded preprocessing
def test_rowgather_equals_oracle(emu, n, permute, rows_per_patch, nt):
    c = flows.build_case("laplace_q1_hex", n, False, permute)   # stiffness (incremental) + body force f = 1
    ref = c.run_oracle()
    for variant in (0,):
        rc, val, rhs, stats = run_emu(emu, c, 1.0, 1.0, 1, 1, 1, rows_per_patch, nt, ref, general=variant)
        assert rc == 0
        assert not np.isnan(val).any()
        assert H.csr_rel_diff(ref[0], ref[2], val) <= 1e-12 and H.vec_rel_diff(ref[3], rhs) <= 1e-12
    assert stats[0] >= (n - 1) ** 3 // rows_per_patch and stats[2] > 0
    # write-out: every segment of consecutive rows leaves as one bulk copy (stats[3]) plus at most two odd elements
    assert stats[4] > 0 and stats[3] > 0


@pytest.mark.parametrize("variant", [0])   # stencil sums (rg_row_affine)
@pytest.mark.parametrize("incremental", [True, False])
def test_rowgather_sheared_mesh_lift_and_accumulate(emu, incremental, variant):
    c = sheared(6, True)
    c.ops = [("matrix", E.K_LAPLACE, [2.5], 3, 0, 0, incremental)]
    ref = c.run_oracle()
    rc, val, rhs, _ = run_emu(emu, c, 2.5, 0.0, 0, int(incremental), 1, 48, 64, ref, general=variant)
    assert rc == 0
    assert H.csr_rel_diff(ref[0], ref[2], val) <= 1e-12 and H.vec_rel_diff(ref[3], rhs) <= 1e-12
    # accumulate into a non-empty matrix (second assembly into the same solver)
    rc, val2, rhs2, _ = run_emu(emu, c, 2.5, 0.0, 0, int(incremental), 0, 48, 64, ref, general=variant)
    assert rc == 0 and H.csr_rel_diff(ref[0], ref[2] + 0.5, val2) <= 1e-12


@pytest.mark.parametrize("n,permute,rows_per_patch", [(6, False, 400), (9, True, 50)])
def test_rowgather_general_elements_equal_oracle(emu, n, permute, rows_per_patch):
    """perturbed mesh (no element affine): phase 1 keeps the full symmetric local matrix per element instance, phase 2
    gathers it (k_q1hex_rows_general)."""
    c = flows.build_case("laplace_q1_hex", n, True, permute)
    ref = c.run_oracle()
    rc, val, rhs, _ = run_emu(emu, c, 1.0, 1.0, 1, 1, 1, rows_per_patch, 64, ref, general=1)
    assert rc == 0 and not np.isnan(val).any()
    assert H.csr_rel_diff(ref[0], ref[2], val) <= 1e-12 and H.vec_rel_diff(ref[3], rhs) <= 1e-12


@pytest.mark.parametrize("general", [0, 1])
def test_rowgather_partial_rows_of_a_partition(emu, general):
    """multi-GPU layout: the pattern holds the columns of the neighbour's (pattern-only) elements as well.  Assembling
    the first part of the elements in store mode into a NaN-filled matrix and the rest in accumulate mode must give the
    oracle's full system: rows at the cut are partial (zero-filled before the scatter), rows without any owned element
    are written as zeros."""
    c = flows.build_case("laplace_q1_hex", 8, bool(general), False)
    ref = c.run_oracle()
    ne = len(c.conn)
    first, second = np.arange(0, (3 * ne) // 8), np.arange((3 * ne) // 8, ne)
    rc, val, rhs, _ = run_emu(emu, c, 1.0, 1.0, 1, 1, 1, 60, 64, ref, general=general, elems=first)
    assert rc == 0 and not np.isnan(val).any()
    assert (val == 0.0).sum() > 0
    rc, val, rhs, _ = run_emu(emu, c, 1.0, 1.0, 1, 1, 0, 60, 64, ref, general=general, elems=second, val=val, rhs=rhs)
    assert rc == 0
    assert H.csr_rel_diff(ref[0], ref[2], val) <= 1e-12 and H.vec_rel_diff(ref[3], rhs) <= 1e-12


def test_rowgather_rejects_non_lattice_connectivity(emu):
    """rotating the local numbering of one element breaks the offset consistency: the tables must say 'not eligible'
    (the engine then keeps the patch kernels)."""
    c = flows.build_case("laplace_q1_hex", 4, False, False)
    ref = c.run_oracle()
    conn = c.conn.copy()
    e = len(conn) // 2
    conn[e] = conn[e][[1, 2, 3, 0, 5, 6, 7, 4]]   # same element, local frame rotated about zeta
    c.conn = conn
    rc, _, _, _ = run_emu(emu, c, 1.0, 0.0, 0, 1, 1, 400, 64, ref)
    assert rc == 1
