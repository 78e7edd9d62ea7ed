"""Multi-GPU host logic on CPU: gloo runs (world 2 and 3) of the slab partition and of the general element-block
partition (any mesh, several fields, Stokes blocks) with their ghost-row exchange, the oracle standing in for the device
assembly.  The owned rows of all ranks must equal the global single-process assembly."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from insilico_b200 import partition
from oracle import oracle as orc


def _fun(x):
    d = np.sqrt(((x + 0.5) ** 2).sum(axis=1))
    return (1.0 / (4.0 * np.pi)) / d


def _local_system(wl):
    """local CSR of one rank: pattern from owned + halo elements, values from the owned elements only"""
    conn = wl["conn"].astype(np.int64)
    full = orc.Problem(orc.HEX, 1, wl["coords"], conn)
    own = orc.Problem(orc.HEX, 1, wl["coords"], conn[:wl["n_owned_elems"]])
    for p, c in ((full, conn), (own, conn[:wl["n_owned_elems"]])):
        p.set_field(0, 1, 1, wl["n_obj"], c, wl["eqn"], wl["status"], wl["presc"], wl["values"])
    s = orc.System(wl["n_eqn_local"])
    s.register_fields(full, 0, 0)
    s.stiffness(own, orc.K_LAPLACE, [1.0], 3, 0, 0, True)
    s.bodyforce(own, [1.0], 3, 0)
    return s.finish()


def _worker(rank, world, port, e, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    wl = partition.structured_laplace_slab(e, e, e * world, rank, world, _fun)
    rp, col, val, rhs = _local_system(wl)
    t = [torch.from_numpy(a) for a in (rp, col, val, rhs)]
    plan = partition.GhostExchange(rank, world, wl).setup(t[0], t[1])
    plan.exchange(t[2], t[3])
    lo, hi = wl["owned_rows"]
    off = wl["eqn_offset"]
    out[rank] = dict(rows=np.arange(lo, hi) + off, rowptr=rp[lo:hi + 1] - rp[lo], col=col[rp[lo]:rp[hi]] + off,
                     val=t[2].numpy()[rp[lo]:rp[hi]].copy(), rhs=t[3].numpy()[lo:hi].copy(),
                     n_global=wl["n_eqn_global"])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,e", [(2, 4), (3, 3)])
def test_slab_partition_matches_global_assembly(world, e):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, e, out), nprocs=world, join=True)
    # global reference: one process, whole mesh
    wl = partition.structured_laplace_slab(e, e, e * world, 0, 1, _fun)
    rp, col, val, rhs = _local_system(wl)
    assert wl["n_eqn_local"] == out[0]["n_global"]
    rows_seen = 0
    for r in range(world):
        o = out[r]
        lo, hi = o["rows"][0], o["rows"][-1] + 1
        assert lo == rows_seen
        rows_seen = hi
        assert np.array_equal(o["rowptr"], rp[lo:hi + 1] - rp[lo])
        assert np.array_equal(o["col"], col[rp[lo]:rp[hi]])
        scale = np.abs(val).max()
        assert np.abs(o["val"] - val[rp[lo]:rp[hi]]).max() <= 1e-13 * scale
        assert np.abs(o["rhs"] - rhs[lo:hi]).max() <= 1e-13 * max(np.abs(rhs).max(), 1e-300)
    assert rows_seen == len(rhs)


# ---- general element-block partition (any mesh, several fields, all-to-all ghost exchange) ---------------------------
def _general_local_system(case, wl):
    """local CSR of one rank with the oracle: pattern from owned + halo elements, values from the owned elements"""
    conn = wl["conn"].astype(np.int64)
    no = wl["n_owned_elems"]
    full = orc.Problem(case.shape, case.geom_deg, wl["coords"], conn)
    own = orc.Problem(case.shape, case.geom_deg, wl["coords"], conn[:no])
    for i, f in enumerate(wl["fields"]):
        ed = f["elem_dof"].astype(np.int64)
        full.set_field(i, f["fe_deg"], f["ds"], f["n_obj"], ed, f["eqn"], f["status"], f["presc"], f["values"])
        own.set_field(i, f["fe_deg"], f["ds"], f["n_obj"], ed[:no], f["eqn"], f["status"], f["presc"], f["values"])
    s = orc.System(wl["n_eqn_local"])
    for op in case.ops:
        if op[0] == "matrix":
            s.register_fields(full, op[4], op[5])
    for op in case.ops:
        if op[0] == "matrix":
            s.stiffness(own, op[1], op[2], op[3], op[4], op[5], incremental=op[6], nthreads=1)
        elif op[0] == "residual":
            s.residual(own, op[1], op[2], op[3], op[4], op[5])
        elif op[0] == "body":
            s.bodyforce(own, op[1], op[2], op[3])
    return s.finish()


def _general_worker(rank, world, port, name, n, out):
    from tests import flows
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    case = flows.build_case(name, n, True, True)     # perturbed nodes, permuted element order
    wl = partition.general_partition(case.coords, case.conn, case.fields, case.n_eqn, rank, world)
    rp, col, val, rhs = _general_local_system(case, wl)
    t = [torch.from_numpy(a) for a in (rp, col, val, rhs)]
    plan = partition.GeneralExchange(rank, world, wl).setup(t[0], t[1])
    plan.exchange(t[2], t[3])
    no, l2g = wl["n_owned_rows"], wl["l2g"]
    out[rank] = dict(rows=l2g[:no].copy(), rowptr=rp[:no + 1].copy(), gcol=l2g[col[:rp[no]]].copy(),
                     val=t[2].numpy()[:rp[no]].copy(), rhs=t[3].numpy()[:no].copy(), n_halo=len(wl["elements"]) - wl["n_owned_elems"],
                     n_ghost=wl["n_ghost_rows"])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,name,n", [(2, "laplace_q1_hex", 4), (3, "laplace_p1_tet", 4), (3, "stvenant_q1_hex", 3),
                                          (2, "stokes_p2p1_tet", 2)])
def test_general_partition_matches_global_assembly(world, name, n):
    from tests import flows
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_general_worker, args=(world, port, name, n, out), nprocs=world, join=True)
    case = flows.build_case(name, n, True, True)
    rp, col, val, rhs = case.run_oracle(register=True)
    seen = np.zeros(case.n_eqn, dtype=int)
    scale, rscale = np.abs(val).max(), max(np.abs(rhs).max(), 1e-300)
    assert sum(out[r]["n_ghost"] for r in range(world)) > 0 and sum(out[r]["n_halo"] for r in range(world)) > 0
    for r in range(world):
        o = out[r]
        for k, g in enumerate(o["rows"]):
            seen[g] += 1
            a, b = o["rowptr"][k], o["rowptr"][k + 1]
            order = np.argsort(o["gcol"][a:b])
            assert np.array_equal(o["gcol"][a:b][order], col[rp[g]:rp[g + 1]]), "pattern of an owned row differs"
            if b > a:  # (rows that only couple to constrained DoFs are empty)
                assert np.abs(o["val"][a:b][order] - val[rp[g]:rp[g + 1]]).max() <= 1e-13 * scale
            assert abs(o["rhs"][k] - rhs[g]) <= 1e-13 * rscale
    assert np.all(seen == 1), "every equation must be owned by exactly one rank"


# ---- driven-cavity slabs generated per rank (no global mesh anywhere) -------------------------------------------------
class _StokesShim:
    """what _general_local_system needs of a case: element type and the operations of the step"""
    shape, geom_deg = 2, 1     # TET, P1 geometry (set properly below)

    def __init__(self):
        from insilico_b200 import engine as E
        self.shape, self.geom_deg = E.TET, 1
        self.ops = [("matrix", E.K_VECTOR_LAPLACE, [1.0], 4, 0, 0, True), ("matrix", E.K_PRESSURE_GRADIENT, [0.0], 4, 0, 1, True),
                    ("matrix", E.K_VELOCITY_DIVERGENCE, [0.0], 4, 1, 0, True), ("body", [1.0, -2.0, 0.5], 4, 0)]


def _stokes_slab_worker(rank, world, port, n, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    wl = partition.structured_stokes_slab(n, rank, world)
    rp, col, val, rhs = _general_local_system(_StokesShim(), wl)
    t = [torch.from_numpy(a) for a in (rp, col, val, rhs)]
    plan = partition.GeneralExchange(rank, world, wl).setup(t[0], t[1])
    plan.exchange(t[2], t[3])
    no, l2g = wl["n_owned_rows"], wl["l2g"]
    out[rank] = dict(rows=l2g[:no].copy(), rowptr=rp[:no + 1].copy(), gcol=l2g[col[:rp[no]]].copy(),
                     val=t[2].numpy()[:rp[no]].copy(), rhs=t[3].numpy()[:no].copy(), n_ghost=wl["n_ghost_rows"],
                     n_halo=len(wl["conn"]) - wl["n_owned_elems"])
    dist.barrier()
    dist.destroy_process_group()


def test_stokes_slab_numbering_is_the_consecutive_numbering():
    """one rank = the whole cube: same mesh as the generator of the parity tests, and the closed-form numbering has the
    size and the pattern size of base::dof::numberDoFsConsecutively on the generated DoF objects"""
    from insilico_b200 import engine as E, meshgen, workloads
    n = 3
    wl = partition.structured_stokes_slab(n, 0, 1, perturb=0.0, permute=False)
    coords, conn = meshgen.unit_cube_tet(n, n, n)
    assert np.array_equal(conn, wl["conn"]) and np.abs(coords - wl["coords"]).max() == 0.0
    w = workloads.build("C5", n, perturb=False, permute=False)
    assert w.n_eqn == wl["n_eqn_local"] == wl["n_eqn_global"]
    u = wl["fields"][0]
    assert int((u["status"] == 0).sum()) == int((w.fields[0]["status"] == 0).sum())
    assert float(u["presc"].sum()) == float(w.fields[0]["presc"].sum())
    # same matrix up to the symmetric permutation between the two numberings: equal nnz and equal sorted row sums
    shim = _StokesShim()
    a = _general_local_system(shim, wl)
    full = orc.Problem(E.TET, 1, w.coords, w.conn.astype(np.int64))
    for i, f in enumerate(w.fields):
        full.set_field(i, f["fe_deg"], f["ds"], f["n_obj"], f["elem_dof"].astype(np.int64), f["eqn"], f["status"], f["presc"], f["values"])
    s = orc.System(w.n_eqn)
    for op in shim.ops[:3]:
        s.register_fields(full, op[4], op[5])
    for op in shim.ops[:3]:
        s.stiffness(full, op[1], op[2], op[3], op[4], op[5], incremental=True, nthreads=1)
    s.bodyforce(full, *shim.ops[3][1:])
    b = s.finish()
    assert len(a[1]) == len(b[1])
    assert np.allclose(np.sort(np.abs(a[2])), np.sort(np.abs(b[2])), rtol=0, atol=1e-12)
    assert np.allclose(np.sort(a[3]), np.sort(b[3]), rtol=0, atol=1e-13)


@pytest.mark.parametrize("world,n", [(2, 3), (3, 3)])
def test_stokes_slabs_match_global_assembly(world, n):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_stokes_slab_worker, args=(world, port, n, out), nprocs=world, join=True)
    glob = partition.structured_stokes_slab(n, 0, 1, permute=False)
    rp, col, val, rhs = _general_local_system(_StokesShim(), glob)
    assert np.array_equal(glob["l2g"], np.arange(glob["n_eqn_global"]))
    seen = np.zeros(glob["n_eqn_global"], dtype=int)
    scale, rscale = np.abs(val).max(), np.abs(rhs).max()
    assert all(out[r]["n_ghost"] > 0 for r in range(1, world)) and all(out[r]["n_halo"] > 0 for r in range(world - 1))
    for r in range(world):
        o = out[r]
        for k, g in enumerate(o["rows"]):
            seen[g] += 1
            a, b = o["rowptr"][k], o["rowptr"][k + 1]
            order = np.argsort(o["gcol"][a:b])
            assert np.array_equal(o["gcol"][a:b][order], col[rp[g]:rp[g + 1]]), "pattern of an owned row differs"
            if b > a:
                assert np.abs(o["val"][a:b][order] - val[rp[g]:rp[g + 1]]).max() <= 1e-13 * scale
            assert abs(o["rhs"][k] - rhs[g]) <= 1e-13 * rscale
    assert np.all(seen == 1), "every equation must be owned by exactly one rank"
