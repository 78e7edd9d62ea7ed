"""Launched under torchrun by tests/test_multigpu.py (and by hand): slab-partitioned assembly on N GPUs with the
NCCL ghost-row exchange, checked on rank 0 against a single-GPU assembly of the whole mesh."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from insilico_b200 import engine as E  # noqa: E402
from insilico_b200 import partition  # noqa: E402


def fun(x):
    d = np.sqrt(((x + 0.5) ** 2).sum(axis=1))
    return (1.0 / (4.0 * np.pi)) / d


MATRIX_ONLY = len(sys.argv) > 2 and sys.argv[2] == "matrix_only"   # the Q1 stiffness launch is the LAST call before the exchange


def assemble(eng, wl, part=None):
    eng.set_mesh(E.HEX, 1, wl["coords"], wl["conn"])
    if part is not None:
        part.setup_fields()
    else:
        eng.set_field(0, 1, 1, wl["n_obj"], wl["elem_dof"], wl["eqn"], wl["status"], wl["presc"], wl["values"])
    eng.new_solver(wl["n_eqn_local"])
    eng.register_fields(0, 0)
    if part is not None:
        part.setup_exchange()
    for _ in range(2):  # second pass exercises the cached plan
        eng.new_solver(wl["n_eqn_local"])
        if MATRIX_ONLY:
            eng.body_force_computation([1.0], 3, 0)
            eng.stiffness_matrix_computation(E.K_LAPLACE, [1.0], 3, 0, 0, True)   # deferred inside the engine
        else:
            eng.stiffness_matrix_computation(E.K_LAPLACE, [1.0], 3, 0, 0, True)
            eng.body_force_computation([1.0], 3, 0)
        if part is not None:
            part.exchange()
    return eng.get_csr()


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    e = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    backend = os.environ.get("ISL_DIST_BACKEND", "nccl")
    if os.environ.get("ISL_DIST_SAME_GPU"):   # two processes on ONE GPU (gloo carries the exchange through the host)
        local = 0
    torch.cuda.set_device(local)
    if backend == "nccl":
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist.init_process_group(backend)
    wl = partition.structured_laplace_slab(e, e, e * world, rank, world, fun)
    eng = E.Engine(local)
    part = partition.DistributedAssembly(eng, wl, rank, world)
    rp, col, val, rhs = assemble(eng, wl, part)
    lo, hi = wl["owned_rows"]
    off = wl["eqn_offset"]
    mine = dict(lo=lo + off, rowptr=rp[lo:hi + 1] - rp[lo], col=col[rp[lo]:rp[hi]] + off, val=val[rp[lo]:rp[hi]],
                rhs=rhs[lo:hi])
    gathered = [None] * world
    dist.gather_object(mine, gathered if rank == 0 else None, dst=0)
    ok = True
    if rank == 0:
        wg = partition.structured_laplace_slab(e, e, e * world, 0, 1, fun)
        eng1 = E.Engine(local)
        rp1, col1, val1, rhs1 = assemble(eng1, wg)
        seen = 0
        for g in gathered:
            a, n = g["lo"], len(g["rhs"])
            ok &= a == seen
            seen = a + n
            ok &= np.array_equal(g["rowptr"], rp1[a:a + n + 1] - rp1[a])
            ok &= np.array_equal(g["col"], col1[rp1[a]:rp1[a + n]])
            scale = np.abs(val1).max()
            ok &= bool(np.abs(g["val"] - val1[rp1[a]:rp1[a + n]]).max() <= 1e-12 * scale)
            ok &= bool(np.abs(g["rhs"] - rhs1[a:a + n]).max() <= 1e-12 * np.abs(rhs1).max())
        ok &= seen == len(rhs1)
        print("DIST_CHECK", "OK" if ok else "FAILED", "world", world, "rows", seen)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
