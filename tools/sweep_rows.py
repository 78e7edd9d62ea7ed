#!/usr/bin/env python
"""Round-2 tuning sweep of the Q1-hex hot path in ONE process (GPU minutes are scarce): for every patch geometry
(ISL_PATCH_ROWS x ISL_PATCH_STRETCH, which need new preprocessing) the engine is created once and the kernel knobs that
need none (row-gather threads / signed-sum form, the all-affine patch kernel as the incumbent) are switched with
isl_engine_set_option between timed runs.  One line per configuration, best first at the end.

    python tools/sweep_rows.py --n 256 --steps 10
"""
import argparse
import itertools
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=256)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--patch-rows", default="192,256,320")
    ap.add_argument("--stretch", default="1,2,4")
    ap.add_argument("--threads", default="256,320,192,128")
    ap.add_argument("--perturbed", action="store_true", help="perturb the mesh (general-element kernels)")
    args = ap.parse_args()
    import torch
    import bench
    from insilico_b200 import engine as E
    from insilico_b200 import meshgen
    wl = bench.build_workload(args.n)
    if args.perturbed:
        wl["coords"] = meshgen.perturb_interior(wl["coords"], 1.0 / args.n, max_dist=0.1)
    results = []
    for pr, st in itertools.product([int(x) for x in args.patch_rows.split(",")], [float(x) for x in args.stretch.split(",")]):
        os.environ.update(ISL_Q1_ROWS="1", ISL_PATCH_ROWS=str(pr), ISL_PATCH_STRETCH=str(st))
        eng = E.Engine(0)
        stream = torch.cuda.ExternalStream(eng.stream, device=0)
        eng.set_mesh(E.HEX, 1, wl["coords"], wl["conn"])
        eng.set_field(0, 1, 1, wl["n_obj"], wl["elem_dof"], wl["eqn"], wl["status"], wl["presc"], wl["values"])
        eng.new_solver(wl["n_eqn_local"])
        t0 = time.perf_counter()
        eng.register_fields(0, 0)
        eng.synchronize()
        t_reg = time.perf_counter() - t0

        def step():
            eng.new_solver(wl["n_eqn_local"])
            eng.stiffness_matrix_computation(E.K_LAPLACE, [1.0], 3, 0, 0, True)
            eng.body_force_computation([1.0], 3, 0)
            eng.finish_assembly()

        configs = [dict(q1_rows=1, rows_threads=t) for t in [int(x) for x in args.threads.split(",")]]
        for cfg in configs:
            for k, v in cfg.items():
                eng.set_option(k, v)
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
            ms = a.elapsed_time(b) / args.steps
            rec = dict(patch_rows=pr, stretch=st, ms_per_step=ms, register_s=t_reg, **cfg)
            results.append(rec)
            print(json.dumps(rec), flush=True)
        eng.close()
    results.sort(key=lambda r: r["ms_per_step"])
    print("BEST", json.dumps(results[:5]))


if __name__ == "__main__":
    main()
