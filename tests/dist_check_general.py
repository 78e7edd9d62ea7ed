"""Launched under torchrun by tests/test_multigpu.py (and by hand): general element-block partition
(insilico_b200.partition.general_partition) on N GPUs with the NCCL ghost-row exchange; the owned rows of all ranks are
checked on rank 0 against the CPU oracle's assembly of the whole problem.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/dist_check_general.py stokes_p2p1_tet 4
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from insilico_b200 import engine as E  # noqa: E402
from insilico_b200 import partition  # noqa: E402
from tests import flows  # noqa: E402


class _SlabCase:
    """the whole cube from the same generator (one rank, identity numbering) assembled by the oracle"""

    def __init__(self, n):
        self.shape, self.geom_deg, self.n = E.TET, 1, n
        self.ops = [("matrix", E.K_VECTOR_LAPLACE, [1.0], 4, 0, 0, True), ("matrix", E.K_PRESSURE_GRADIENT, [0.0], 4, 0, 1, True),
                    ("matrix", E.K_VELOCITY_DIVERGENCE, [0.0], 4, 1, 0, True), ("body", [1.0, -2.0, 0.5], 4, 0)]
        self.n_eqn = 3 * (2 * n - 1) ** 3 + (n + 1) ** 3 - 1

    def run_oracle(self, register=True):
        from oracle import oracle as orc
        wl = partition.structured_stokes_slab(self.n, 0, 1, permute=False)
        p = orc.Problem(self.shape, self.geom_deg, wl["coords"], wl["conn"].astype(np.int64))
        for i, f in enumerate(wl["fields"]):
            p.set_field(i, f["fe_deg"], f["ds"], f["n_obj"], f["elem_dof"].astype(np.int64), f["eqn"], f["status"], f["presc"], f["values"])
        s = orc.System(wl["n_eqn_local"])
        for op in self.ops[:3]:
            s.register_fields(p, op[4], op[5])
        for op in self.ops[:3]:
            s.stiffness(p, op[1], op[2], op[3], op[4], op[5], incremental=op[6])
        s.bodyforce(p, *self.ops[3][1:])
        return s.finish()


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    name = sys.argv[1] if len(sys.argv) > 1 else "stokes_p2p1_tet"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    backend = os.environ.get("ISL_DIST_BACKEND", "nccl")
    if os.environ.get("ISL_DIST_SAME_GPU"):   # two processes on ONE GPU (gloo carries the exchange through the host)
        local = 0
    torch.cuda.set_device(local)
    if backend == "nccl":
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist.init_process_group(backend)
    if name == "stokes_slab":     # driven-cavity slabs generated per rank (partition.structured_stokes_slab)
        case = _SlabCase(n)
        wl = partition.structured_stokes_slab(n, rank, world)
    else:
        case = flows.build_case(name, n, True, True)
        wl = partition.general_partition(case.coords, case.conn, case.fields, case.n_eqn, rank, world)
    eng = E.Engine(local)
    part = partition.GeneralDistributedAssembly(eng, wl, rank, world, case.shape, case.geom_deg)
    eng.new_solver(wl["n_eqn_local"])
    for op in case.ops:
        if op[0] == "matrix":
            eng.register_fields(op[4], op[5])
    part.setup_exchange()
    for _ in range(2):  # second pass exercises the cached plan
        eng.new_solver(wl["n_eqn_local"])
        for op in case.ops:
            if op[0] == "matrix":
                eng.register_fields(op[4], op[5])
        for op in case.ops:
            if op[0] == "matrix":
                eng.stiffness_matrix_computation(op[1], op[2], op[3], op[4], op[5], incremental=op[6])
            elif op[0] == "residual":
                eng.compute_residual_forces(op[1], op[2], op[3], op[4], op[5])
            elif op[0] == "body":
                eng.body_force_computation(op[1], op[2], op[3])
        part.exchange()
    rp, col, val, rhs = eng.get_csr()
    no, l2g = wl["n_owned_rows"], wl["l2g"]
    mine = dict(rows=l2g[:no], rowptr=rp[:no + 1], gcol=l2g[col[:rp[no]]], val=val[:rp[no]], rhs=rhs[:no])
    gathered = [None] * world
    dist.gather_object(mine, gathered if rank == 0 else None, dst=0)
    ok = True
    if rank == 0:
        rpg, colg, valg, rhsg = case.run_oracle(register=True)
        seen = np.zeros(case.n_eqn, dtype=int)
        scale, rscale = np.abs(valg).max(), max(np.abs(rhsg).max(), 1e-300)
        worst = 0.0
        for o in gathered:
            for k, g in enumerate(o["rows"]):
                seen[g] += 1
                a, b = o["rowptr"][k], o["rowptr"][k + 1]
                order = np.argsort(o["gcol"][a:b])
                ok &= np.array_equal(o["gcol"][a:b][order], colg[rpg[g]:rpg[g + 1]])
                if ok and b > a:
                    worst = max(worst, float(np.abs(o["val"][a:b][order] - valg[rpg[g]:rpg[g + 1]]).max()) / scale)
                worst = max(worst, abs(o["rhs"][k] - rhsg[g]) / rscale)
        ok &= bool(np.all(seen == 1)) and worst <= 1e-12
        print("DIST_CHECK_GENERAL", "OK" if ok else "FAILED", name, "world", world, "rows", int(seen.sum()), "worst", worst)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
