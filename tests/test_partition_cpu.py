"""Multi-GPU host logic on CPU: world_size-2 gloo run of the slab partition + ghost-row exchange, with the oracle
standing in for the device assembly.  The owned rows of both ranks must equal the global single-process assembly."""
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
