#!/usr/bin/env python
"""Timing of the generic assembly kernels on the other BASELINE configs (C3 hyperelastic Q2 hex, C4
neo-Hookean P2 tets, C5 Taylor-Hood Stokes blocks) and on the perturbed C2 mesh.  Not the headline bench (bench.py):
a helper to find the next kernel to tune.  One JSON line per case.

  python tools/bench_configs.py                       # default list at moderate sizes
  python tools/bench_configs.py --case stvenant_q2_hex --n 32 --steps 5
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

DEFAULT = [("laplace_q1_hex", 96), ("stvenant_q2_hex", 24), ("neohooke_p2_tet", 24), ("stokes_p2p1_tet", 24),
           ("vector_laplace_q1_hex", 64), ("neohooke_q1_hex", 64)]


def run(name, n, steps, warmup, perturb=True):
    import torch
    from insilico_b200 import engine as E
    from tests import flows
    c = flows.build_case(name, n, perturb, False)
    eng = E.Engine(0)
    stream = torch.cuda.ExternalStream(eng.stream, device=0)
    eng.set_mesh(c.shape, c.geom_deg, c.coords, c.conn)
    for i, f in enumerate(c.fields):
        eng.set_field(i, f["fe_deg"], f["ds"], f["n_obj"], f["elem_dof"], f["eqn"], f["status"], f["presc"], f["values"])
    eng.new_solver(c.n_eqn)
    t0 = time.perf_counter()
    for op in c.ops:
        if op[0] == "matrix":
            eng.register_fields(op[4], op[5])
    eng.synchronize()
    t_reg = time.perf_counter() - t0

    def one(op):
        if op[0] == "matrix":
            eng.stiffness_matrix_computation(op[1], op[2], op[3], op[4], op[5], incremental=op[6])
        elif op[0] == "residual":
            eng.compute_residual_forces(op[1], op[2], op[3], op[4], op[5])
        elif op[0] == "body":
            eng.body_force_computation(op[1], op[2], op[3])

    def step(ev=None):
        eng.new_solver(c.n_eqn)
        for k, op in enumerate(c.ops):
            if ev is not None:
                ev[k][0].record(stream)
            one(op)
            if ev is not None:
                ev[k][1].record(stream)
        eng.finish_assembly()   # flushes a deferred launch

    for _ in range(warmup):
        step()
    eng.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    evs = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in c.ops] for _ in range(steps)]
    a.record(stream)
    for s in range(steps):
        step(evs[s])
    b.record(stream)
    eng.synchronize()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    per_op = [sum(evs[s][k][0].elapsed_time(evs[s][k][1]) for s in range(steps)) / steps for k in range(len(c.ops))]
    ne = int(c.conn.shape[0])
    nnz = int(eng.finish_assembly()[1])
    eng.close()
    return {"case": name, "n": n, "perturbed": perturb, "n_elems": ne, "n_eqn": int(c.n_eqn), "nnz": nnz,
            "ms_per_step": ms, "elements_per_s": ne / (ms * 1e-3), "register_fields_ms": t_reg * 1e3,
            "ops": [{"op": op[0], "kernel": (op[1] if op[0] != "body" else "bodyforce"), "ms": t}
                    for op, t in zip(c.ops, per_op)],
            "note": "per-op times are between API calls on the engine stream; a deferred Q1 stiffness launch is "
                    "booked to the following call"}


DRIVER_TYPE = {"laplace_q1_hex": "laplace_q1_hex", "stvenant_q2_hex": "solid_q2_hex", "neohooke_p2_tet": "solid_p2_tet",
               "stokes_p2p1_tet": "stokes_p2p1_tet", "vector_laplace_q1_hex": "vector_laplace_q1_hex",
               "neohooke_q1_hex": "solid_q1_hex", "stvenant_q1_hex": "solid_q1_hex", "laplace_q2_hex": "laplace_q2_hex",
               "laplace_p1_tet": "laplace_p1_tet", "stokes_q2q1_hex": "stokes_q2q1_hex"}


def run_cpu_reference(name, n, steps, perturb=True):
    """the same case assembled by the UNMODIFIED reference on the host cores (oracle/_ref/ref_driver_omp: OpenMP,
    pre-structured triplets; residual / body-force loops are serial in the reference)"""
    import tempfile
    from tests import flows
    from tools import make_ref_goldens as G
    G.DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_driver_omp")
    c = flows.build_case(name, n, perturb, False)
    with tempfile.TemporaryDirectory() as wd:
        log = G.run_reference(c, DRIVER_TYPE[name], True, wd, repeat=steps, dump=False)["log"]
    reps = [l.split() for l in log.splitlines() if l.startswith("rep ")]
    t = [float(r[r.index("assemble") + 1]) for r in reps]
    ne = int(c.conn.shape[0])
    return {"case": name, "n": n, "impl": "reference (unmodified headers + oracle/compat stand-ins)", "cores": os.cpu_count(),
            "n_elems": ne, "n_eqn": int(c.n_eqn), "ms_per_step": 1e3 * sum(t) / len(t), "elements_per_s": ne / (sum(t) / len(t))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cpu-reference", action="store_true", help="time the reference on the host cores instead of the engine")
    ap.add_argument("--case", default=None)
    ap.add_argument("--n", type=int, default=None)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--structured", action="store_true", help="do not perturb the mesh")
    args = ap.parse_args()
    cases = DEFAULT if args.case is None else [(args.case, args.n or 16)]
    for name, n in cases:
        if args.cpu_reference:
            print(json.dumps(run_cpu_reference(name, args.n or n, max(1, min(args.steps, 3)), not args.structured)), flush=True)
        else:
            print(json.dumps(run(name, args.n or n, args.steps, args.warmup, not args.structured)), flush=True)


if __name__ == "__main__":
    main()
