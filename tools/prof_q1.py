#!/usr/bin/env python
"""Run the Q1-hex hot path a few steps on the structured 256^3 mesh and on the same mesh with perturbed coordinates
(update_coords, same patches) -- the command profiled by ncu (tools/gpu_r2_call*.sh).  Prints ms/step of both."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=256)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--perturb-first", action="store_true", help="build the patches on the perturbed mesh")
    args = ap.parse_args()
    import torch
    import bench
    from insilico_b200 import engine as E
    from insilico_b200 import meshgen
    wl = bench.build_workload(args.n)
    pert = meshgen.perturb_interior(wl["coords"], 1.0 / args.n, max_dist=0.1)
    eng = E.Engine(0)
    stream = torch.cuda.ExternalStream(eng.stream, device=0)
    eng.set_mesh(E.HEX, 1, pert if args.perturb_first else wl["coords"], wl["conn"])
    eng.set_field(0, 1, 1, wl["n_obj"], wl["elem_dof"], wl["eqn"], wl["status"], wl["presc"], wl["values"])
    eng.new_solver(wl["n_eqn_local"])
    eng.register_fields(0, 0)

    def step():
        eng.new_solver(wl["n_eqn_local"])
        eng.stiffness_matrix_computation(E.K_LAPLACE, [1.0], 3, 0, 0, True)
        eng.body_force_computation([1.0], 3, 0)
        eng.finish_assembly()

    def timed(tag):
        for _ in range(3):
            step()
        eng.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(args.steps):
            step()
        b.record(stream)
        eng.synchronize()
        torch.cuda.synchronize()
        print(json.dumps({"mesh": tag, "ms_per_step": a.elapsed_time(b) / args.steps}), flush=True)

    if not args.perturb_first:
        timed("structured")
        eng.update_coords(pert)
    timed("perturbed")


if __name__ == "__main__":
    main()
