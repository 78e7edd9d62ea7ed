#!/usr/bin/env python
"""Benchmark of the element-assembly hot path (BASELINE.json config 2).

  python bench.py --gpus N --steps K --warmup W            # CUDA engine (this repo)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port) on host cores

Workload (per GPU, weak scaling): 3-D scalar Laplace, Q1 hexahedra, structured 256^3 unit-cube mesh, Dirichlet data
from the Laplace fundamental solution on the whole boundary, constant body force f = 1, quadrature degree 3 (8 points).
A "step" is one complete assembly pass: fresh solver (zero matrix values and rhs), stiffness matrix with Dirichlet
lift, body force.  Prints ONE JSON line (see README / DESIGN.md for the keys).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALGO_BYTES_PER_ELEM = 537.2  # conn 32 + slot map 256 + coords 24.28 + CSR values 216.84 + rhs 8.09
ALGO_FLOPS_PER_ELEM = 4600.0  # 8 points x (J 144 + inverse/det 50 + gradients 144 + scale 24 + symmetric half of K 216)
METRIC = "fp64_elements_assembled_per_sec_into_csr"


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs, torch copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """samples nvidia-smi clocks and throttle reasons during the timed region"""

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu_index, self.samples, self.stop_flag = gpu_index, [], False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 7:
                    self.samples.append(parts)
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(float(s[0]) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[3 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons,
                "samples": len(self.samples), "power_w_max": max(float(s[2]) for s in self.samples)}


def bind_to_gpu_numa(local_rank):
    """pin this process (and with it the first touch of its pinned host buffers) to the NUMA node of its GPU: without it
    every rank's staging memory sits on node 0 and the end-to-end copies of 8 ranks share one socket's memory bandwidth
    (SCALE_r01: e2e 86 -> 337 ms/step from 1 to 8 GPUs)"""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = torch.cuda.get_device_properties(local_rank).pci_domain_id
        dev = torch.cuda.get_device_properties(local_rank).pci_device_id
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/" % (dom, bus, dev)
        node = int(open(path + "numa_node").read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return node
    except Exception:  # noqa: BLE001  (best effort: containers may hide sysfs)
        return None


def fund_sol_laplace(x):
    d = np.sqrt(((x + 0.5) ** 2).sum(axis=1))
    return (1.0 / (4.0 * np.pi)) / d


def build_workload(n, rank=0, world=1, strong=False):
    """structured slab of the (n x n x n*world) mesh (weak scaling; strong: of the n^3 mesh) owned by `rank`, flat field
    arrays with global numbering info."""
    from insilico_b200 import meshgen
    from insilico_b200 import partition
    return partition.structured_laplace_slab(n, n, n if strong else n * world, rank, world, fund_sol_laplace)


# ---------------------------------------------------------------------------------------------------------
def cpu_reference(n_sample, steps, warmup, threads):
    """the reference's CPU assembly (oracle port, oracle/insilico_oracle.cpp) on an n_sample^3 mesh of the same
    workload.  Returns (elements/s pre-structured OpenMP, elements/s dynamic single thread, ms/step)."""
    from oracle import oracle as orc
    from insilico_b200 import meshgen
    from insilico_b200 import engine as E
    coords, conn, _ = meshgen.unit_cube_hex(n_sample, n_sample, n_sample)
    onb = meshgen.boundary_node_mask(coords)
    status = onb.astype(np.uint8)[:, None]
    presc = np.where(onb, fund_sol_laplace(coords), 0.0)[:, None]
    eqn, ndof = orc.number_dofs(status)
    prob = orc.Problem(orc.HEX, 1, coords, conn.astype(np.int64))
    prob.set_field(0, 1, 1, len(coords), conn.astype(np.int64), eqn, status, presc, np.zeros_like(presc))
    ne = conn.shape[0]

    def step(prestructured):
        s = orc.System(ndof)
        t_reg = 0.0
        if prestructured:
            t0 = time.perf_counter(); s.register_fields(prob, 0, 0); t_reg = time.perf_counter() - t0
        t0 = time.perf_counter()
        s.stiffness(prob, orc.K_LAPLACE, [1.0], 3, 0, 0, True, nthreads=(threads if prestructured else 1))
        s.bodyforce(prob, [1.0], 3, 0)
        s.finish()
        return time.perf_counter() - t0, t_reg

    for _ in range(warmup):
        step(True)
    ts = [step(True) for _ in range(steps)]
    t_pre = sum(t[0] for t in ts) / steps
    t_reg = sum(t[1] for t in ts) / steps
    t_dyn = step(False)[0]
    return ne / t_pre, ne / t_dyn, t_pre * 1e3, t_reg * 1e3, ne


REF_DRIVER = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle", "_ref", "ref_driver_omp")


BINDING_DRIVER = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle", "_ref", "apps_b200", "ref_driver")


def cpu_reference_real(n_sample, steps, driver=None):
    """The UNMODIFIED reference (headers of thrueberg/inSilico compiled here against the std-only Boost/Eigen stand-ins
    of oracle/compat, binary oracle/_ref/ref_driver_omp built by `make -C oracle ref` with the reference's release
    flags -O3 -fopenmp -DNDEBUG, NTHREADS=0 = all cores) on an n_sample^3 mesh of the same workload: per step a fresh
    base::solver::Eigen3, registerFields (untimed, like our cached pattern), stiffnessMatrixComputation +
    bodyForceComputation (timed).  Returns None when the binary is not there."""
    driver = driver or REF_DRIVER
    if not os.access(driver, os.X_OK):
        return None
    import subprocess
    import tempfile
    from insilico_b200 import meshgen
    coords, conn, _ = meshgen.unit_cube_hex(n_sample, n_sample, n_sample)
    presc = fund_sol_laplace(coords)
    with tempfile.TemporaryDirectory() as wd:
        smf = os.path.join(wd, "mesh.smf")
        with open(smf, "w") as f:
            f.write("! elementShape hexahedron\n! elementNumPoints 8\n%d %d\n" % (len(coords), len(conn)))
            np.savetxt(f, coords, fmt="%.17g")
            np.savetxt(f, conn, fmt="%d")
        presc.astype(np.float64).tofile(os.path.join(wd, "presc.bin"))
        job = os.path.join(wd, "job.txt")
        with open(job, "w") as f:
            f.write("type laplace_q1_hex\nmesh %s\nout %s/out\nregister 1\nrepeat %d\ndump 0\n"
                    "field 0 1 -1 %s/presc.bin -\nop matrix laplace 0 0 1 1.0\nop body body 0 0 1 1.0\n"
                    % (smf, wd, max(1, steps), wd))
        try:
            out = subprocess.run([driver, job], check=True, capture_output=True, text=True, timeout=1500).stdout
        except Exception as e:  # noqa: BLE001  (fall back to the port, say so in the sample text)
            sys.stderr.write("reference driver failed: %r\n" % (e,))
            return None
    reps = [l.split() for l in out.splitlines() if l.startswith("rep ")]
    if driver != REF_DRIVER and len(reps) > 1:
        reps = reps[1:]   # the first pass uploads the mesh and builds the pattern (untimed for the CPU arm as well)
    t_asm = [float(r[r.index("assemble") + 1]) + (float(r[r.index("finish") + 1]) if driver != REF_DRIVER else 0.0) for r in reps]
    t_reg = [float(r[r.index("register") + 1]) for r in reps]
    ne = conn.shape[0]
    t = sum(t_asm) / len(t_asm)
    return ne / t, t * 1e3, sum(t_reg) / len(t_reg) * 1e3, ne


def reference_api_on_engine(ns, steps):
    """The SAME job as the CPU baseline (reference objects in host memory, the reference's own API calls) with
    base::solver::Eigen3 replaced by the binding's base::solver::B200 (include/insilico_b200_reference.hpp): per step a
    fresh solver, registerFields (untimed, pattern cached), then stiffnessMatrixComputation + bodyForceComputation +
    finishAssembly -- i.e. re-reading the reference's heap objects, host->device copies of what changed, the kernels,
    the wait for the device.  The finished system stays on the device (getDeviceCSR), like a device solver would use it.
    None when the prebuilt application is missing or fails."""
    try:
        r = cpu_reference_real(ns, steps + 1, driver=BINDING_DRIVER)
    except Exception as e:  # noqa: BLE001
        sys.stderr.write("reference API on the engine failed: %r\n" % (e,))
        return None
    if r is None:
        return None
    v, ms, ms_reg, ne = r
    return {"value": v, "unit": "elements/s", "ms_per_step": ms,
            "sample": "%d^3 Q1 hex mesh (%d elements), the CPU baseline's job through the reference API on the B200 "
                      "binding; registerFields %.0f ms excluded" % (ns, ne, ms_reg),
            "what": "unmodified reference headers + include/insilico_b200_reference.hpp + libinsilico_b200.so: host scan "
                    "of the reference's Node/Element/DegreeOfFreedom objects, isl_assemble_matrix + isl_assemble_bodyforce, "
                    "isl_finish; the system stays on the device"}


def reference_baseline(ns, steps, warmup, cores):
    """cpu_baseline dict + ms per step: the real reference when oracle/_ref is there, else the oracle port."""
    real = cpu_reference_real(ns, steps)
    if real is not None:
        v, ms, ms_reg, ne = real
        return {"value": v, "unit": "elements/s", "cores": cores, "kind": "reference",
                "sample": "%d^3 Q1 hex mesh (%d elements) of the same workload per step, UNMODIFIED reference headers "
                          "(stiffnessMatrixComputation + bodyForceComputation into base::solver::Eigen3, pre-structured "
                          "triplets, OpenMP parallel for over %d threads) compiled against std-only Boost/Eigen stand-ins "
                          "(oracle/compat): %.0f ms/step, registerFields %.0f ms excluded" % (ns, ne, cores, ms, ms_reg)}, ms
    v_pre, v_dyn, ms, ms_reg, ne = cpu_reference(ns, steps, warmup, cores)
    return {"value": v_pre, "unit": "elements/s", "cores": cores, "kind": "port",
            "sample": "%d^3 Q1 hex mesh (%d elements) of the same workload per step, oracle port of the reference "
                      "(oracle/_ref not built): pre-structured triplets + OpenMP %d threads %.0f ms/step (registerFields "
                      "%.0f ms excluded); dynamic std::set mode, 1 thread: %.0f elements/s"
                      % (ns, ne, cores, ms, ms_reg, v_dyn)}, ms


# ---------------------------------------------------------------------------------------------------------
# BASELINE configs 1, 3, 4, 5 (insilico_b200/workloads.py): python bench.py --config C3 [--n 64]
DEFAULT_N = {"C1": 32, "C3": 64, "C4": 64, "C5": 64}
CPU_SAMPLE_N = {"C1": 32, "C3": 10, "C4": 14, "C5": 16}
BOUND = {"C1": "hbm", "C3": "fp64", "C4": "fp64", "C5": "hbm"}
DRIVER_TYPE = {"C1": "laplace_q1_hex", "C2": "laplace_q1_hex", "C3": "solid_q2_hex", "C4": "solid_p2_tet", "C5": "stokes_p2p1_tet"}
KERNEL_NAME = {1: "laplace", 2: "stvenant", 3: "neohooke", 4: "pressure_gradient", 5: "velocity_divergence", 6: "vector_laplace"}
SHAPE_NAME = {2: "triangle", 3: "quadrilateral", 4: "tetrahedron", 5: "hexahedron"}


def cpu_reference_config(w, steps):
    """one workload (insilico_b200.workloads.Workload) assembled by the UNMODIFIED reference on the host cores
    (oracle/_ref/ref_driver_omp, OpenMP over all cores, pre-structured triplets; registerFields untimed).
    Returns (elements/s, ms per step) or None."""
    if not os.access(REF_DRIVER, os.X_OK):
        return None
    import tempfile
    with tempfile.TemporaryDirectory() as wd:
        smf = os.path.join(wd, "mesh.smf")
        c3 = np.zeros((len(w.coords), 3)); c3[:, :w.dim] = w.coords
        with open(smf, "w") as f:
            f.write("! elementShape %s\n! elementNumPoints %d\n%d %d\n" % (SHAPE_NAME[w.shape], w.conn.shape[1], len(c3), len(w.conn)))
            np.savetxt(f, c3, fmt="%.17g")
            np.savetxt(f, w.conn, fmt="%d")
        lines = ["type %s" % DRIVER_TYPE[w.name], "mesh %s" % smf, "out %s/out" % wd, "register 1", "repeat %d" % max(1, steps), "dump 0"]
        for i, fl in enumerate(w.fields):
            for nm, arr in (("presc", fl["presc"]), ("values", fl["values"]), ("status", fl["status"].astype(np.float64))):
                np.ascontiguousarray(arr, dtype=np.float64).tofile(os.path.join(wd, "%s%d.bin" % (nm, i)))
            lines.append("field %d 2 -1 %s/presc%d.bin %s/values%d.bin %s/status%d.bin" % (i, wd, i, wd, i, wd, i))
        for op in w.ops:
            if op[0] == "matrix":
                lines.append("op matrix %s %d %d %d %s" % (KERNEL_NAME[op[1]], op[4], op[5], int(op[6]), " ".join("%.17g" % v for v in op[2])))
            elif op[0] == "residual":
                lines.append("op residual %s %d %d 1 %s" % (KERNEL_NAME[op[1]], op[4], op[5], " ".join("%.17g" % v for v in op[2])))
            else:
                lines.append("op body body %d %d 1 %s" % (op[3], op[3], " ".join("%.17g" % v for v in op[1])))
        job = os.path.join(wd, "job.txt")
        with open(job, "w") as f:
            f.write("\n".join(lines) + "\n")
        try:
            out = subprocess.run([REF_DRIVER, job], check=True, capture_output=True, text=True, timeout=1500).stdout
        except Exception as e:  # noqa: BLE001
            sys.stderr.write("reference driver failed: %r\n" % (e,))
            return None
    reps = [l.split() for l in out.splitlines() if l.startswith("rep ")]
    t = [float(r[r.index("assemble") + 1]) for r in reps]
    if not t:
        return None
    dt = sum(t) / len(t)
    return len(w.conn) / dt, dt * 1e3


def config_cpu_baseline(cfg, ns, steps, cores):
    from insilico_b200 import workloads
    ws = workloads.build(cfg, ns)
    r = cpu_reference_config(ws, steps)
    if r is None:
        return None, None
    return {"value": r[0], "unit": "elements/s", "cores": cores, "kind": "reference",
            "sample": "%s at n = %d (%d elements) per step, UNMODIFIED reference headers (stiffnessMatrixComputation / "
                      "computeResidualForces into base::solver::Eigen3, pre-structured triplets, OpenMP parallel for over %d "
                      "threads) compiled against std-only Boost/Eigen stand-ins (oracle/compat): %.0f ms/step, registerFields "
                      "excluded" % (ws.description.split(":")[0], ns, len(ws.conn), cores, r[1])}, r[1]


def bench_config(args, rank, world, local_rank, cores):
    """one JSON line for BASELINE config C1 / C3 / C4 / C5 on the generic kernels (same keys as the C2 line)"""
    import torch
    import torch.distributed as dist
    from insilico_b200 import engine as E
    from insilico_b200 import partition, workloads
    cfg = args.config
    n = args.n if args.n_given else DEFAULT_N[cfg]
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    eng = E.Engine(local_rank)
    stream = torch.cuda.ExternalStream(eng.stream, device=local_rank)
    extra = {}
    if cfg == "C1":   # dirichlet.cpp on unitCube N N N, N in {8, 16, 32}: the smaller ones are timed first
        for ns in (8, 16):
            ws = workloads.build("C1", ns)
            ws.upload(eng); eng.new_solver(ws.n_eqn); ws.register(eng)
            for _ in range(3):
                ws.step(eng)
            eng.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(args.steps):
                ws.step(eng)
            b.record(stream)
            eng.synchronize(); torch.cuda.synchronize()
            extra["ms_per_step_n%d" % ns] = a.elapsed_time(b) / args.steps
    slab = cfg == "C5" and world > 1 and args.partition != "blocks"
    if slab:
        # the mesh is generated per rank in closed form: no rank ever holds the global mesh (50 M tetrahedra would not fit
        # the host-side numbering of workloads.build in reasonable time); the tiny workload only carries the operations
        w = workloads.build(cfg, 2)
        wl = partition.structured_stokes_slab(n, rank, world)
        w.n_eqn, w.n_elems_global, w.n_nodes_global = wl["n_eqn_global"], wl["n_elems_global"], (n + 1) ** 3
        w.description = w.description.replace("6*2^3", "6*%d^3" % n) + "; mesh and numbering generated per rank (z-slabs)"
        ne_global = wl["n_elems_global"]
    else:
        w = workloads.build(cfg, n)
        ne_global = len(w.conn)
    part = None
    if slab:
        part = partition.GeneralDistributedAssembly(eng, wl, rank, world, w.shape, w.geom_deg)
        n_eqn, ne_rank = wl["n_eqn_local"], wl["n_owned_elems"]
    elif world > 1:
        wl = partition.general_partition(w.coords, w.conn, w.fields, w.n_eqn, rank, world)
        part = partition.GeneralDistributedAssembly(eng, wl, rank, world, w.shape, w.geom_deg)
        n_eqn, ne_rank = wl["n_eqn_local"], wl["n_owned_elems"]
    else:
        w.upload(eng)
        n_eqn, ne_rank = w.n_eqn, ne_global
    eng.new_solver(n_eqn)
    t0 = time.perf_counter()
    w.register(eng)
    if part is not None:
        part.setup_exchange()
    eng.synchronize()
    t_register = time.perf_counter() - t0

    def step(ev=None):
        eng.new_solver(n_eqn)
        for k, op in enumerate(w.ops):
            if ev is not None:
                ev[k][0].record(stream)
            if op[0] == "matrix":
                eng.stiffness_matrix_computation(op[1], op[2], op[3], op[4], op[5], incremental=op[6])
            elif op[0] == "residual":
                eng.compute_residual_forces(op[1], op[2], op[3], op[4], op[5])
            else:
                eng.body_force_computation(op[1], op[2], op[3])
            if ev is not None:
                eng.flush()
                ev[k][1].record(stream)
        if part is not None:
            part.exchange()

    def barrier():
        eng.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = eng.kernel_launches
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    evs = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in w.ops] for _ in range(args.steps)]
    a.record(stream)
    for k in range(args.steps):
        step(evs[k])
    b.record(stream)
    barrier()
    launches = eng.kernel_launches - l0
    ms_total = a.elapsed_time(b)
    per_op = [sum(evs[s][k][0].elapsed_time(evs[s][k][1]) for s in range(args.steps)) / args.steps for k in range(len(w.ops))]
    if world > 1:
        t = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = ne_global / (ms_step * 1e-3)
    nnz = eng.finish_assembly()[1]
    fp64_peak = eng.measure_fp64_peak()

    e2e = None
    if not args.no_e2e and world == 1:
        pin = lambda x: torch.from_numpy(np.ascontiguousarray(x)).pin_memory().numpy()
        h_coords = pin(w.coords)
        h_f = [(pin(f["presc"]), pin(f["values"])) for f in w.fields]
        h_val = torch.empty(nnz, dtype=torch.float64).pin_memory().numpy()
        h_rhs = torch.empty(n_eqn, dtype=torch.float64).pin_memory().numpy()

        def e2e_step():
            eng.update_coords(h_coords)
            for i, (hp, hv) in enumerate(h_f):
                eng.update_field(i, prescribed=hp, values=hv)
            step()
            eng.get_csr(None, None, h_val, h_rhs)

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        ke = max(2, min(args.steps, 5))
        for _ in range(ke):
            e2e_step()
        barrier()
        dt = (time.perf_counter() - t0) / ke
        e2e = {"value": ne_global / dt, "unit": "elements/s",
               "h2d_bytes_per_step": int(h_coords.nbytes + sum(x.nbytes + y.nbytes for x, y in h_f)),
               "d2h_bytes_per_step": int(h_val.nbytes + h_rhs.nbytes), "ms_per_step": dt * 1e3,
               "what": "isl_mesh_update_coords + isl_field_update (pinned H2D), isl_system_create, the assembly calls of the step, "
                       "isl_finish, isl_get_csr values+rhs (pinned D2H); pattern cached"}
    if rank == 0:
        sampler.stop_flag = True
        sampler.join(timeout=2)
    if rank != 0:
        dist.destroy_process_group()
        return
    hbm_peak, peak_src = peaks()
    flops = w.algorithmic_flops_per_element()
    nbytes = w.algorithmic_bytes_per_element(nnz if world == 1 else nnz * world)
    k_dom = int(np.argmax(per_op))
    t_asm = sum(per_op) * 1e-3   # device time of the assembly kernels of one step
    hbm = {"achieved": nbytes * ne_rank / t_asm / 1e9, "peak": hbm_peak, "unit": "GB/s", "peak_source": peak_src,
           "algorithmic_bytes_per_element": nbytes}
    hbm["frac"] = hbm["achieved"] / hbm_peak
    f64 = {"achieved": flops * ne_rank / t_asm / 1e12, "peak": fp64_peak, "unit": "TFLOP/s",
           "peak_source": "measured (isl_measure_fp64_peak: DFMA chains, CUDA events, this run)",
           "algorithmic_flops_per_element": flops}
    f64["frac"] = f64["achieved"] / fp64_peak if fp64_peak else None
    # the scatter: FP64 atomic adds per second against the measured rate of the same access pattern (the generic kernels
    # add three neighbouring entries per node pair for vector fields, single entries for scalar fields)
    atomics = None
    if cfg != "C1" and w.atomics_per_element() > 0:
        n_at = w.atomics_per_element()
        pattern = 1 if max(f["ds"] for f in w.fields) > 1 else 2
        red_peak = eng.measure_red_peak(pattern)
        atomics = {"achieved": n_at * ne_rank / t_asm / 1e9, "peak": red_peak, "unit": "G atomic adds/s", "atomics_per_element": n_at,
                   "peak_source": "measured (isl_measure_red_peak pattern %d: RED.ADD.F64 into a 2 GiB array, this run)" % pattern}
        atomics["frac"] = atomics["achieved"] / red_peak if red_peak else None
    main_r = dict(f64 if BOUND[cfg] == "fp64" else hbm)
    traffic = None   # DRAM bytes per step from the committed ncu capture, valid for the default single-GPU size and kernel path only
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
        if world == 1 and tj.get(cfg + "_n") == n and os.environ.get("ISL_GEN_GATHER", "1") != "0":
            traffic = tj.get(cfg + "_bytes_per_step")
    except (OSError, ValueError):
        pass
    op = w.ops[k_dom]
    roof = {"bound": BOUND[cfg], "kernel": "%s %s (k_tangent / k_force family)" % (op[0], KERNEL_NAME.get(op[1], "bodyforce") if op[0] != "body" else "bodyforce"),
            "achieved": main_r["achieved"], "peak": main_r["peak"], "unit": main_r["unit"], "frac": main_r["frac"], "traffic": traffic,
            "kernel_ms": t_asm * 1e3, "per_op_ms": [{"op": o[0], "kernel": KERNEL_NAME.get(o[1], "body") if o[0] != "body" else "body", "ms": t} for o, t in zip(w.ops, per_op)],
            "hbm": hbm, "fp64": f64, "atomics": atomics}
    cfgd = {"workload": w.description, "n": n, "n_elems": int(ne_global), "n_eqn": int(w.n_eqn), "nnz_per_gpu": int(nnz),
            "partition": ("z-slabs of cube layers generated per rank, owned rows, NCCL ghost-row exchange" if slab else
                          "element blocks along a Z-curve, owned rows, NCCL ghost-row exchange") if world > 1 else "single GPU",
            "l2": "arrays touched per step: %.2f GB (%s L2)" % (nbytes * ne_rank / 1e9, "larger than" if nbytes * ne_rank > 126e6 else "inside"),
            "register_fields_ms": t_register * 1e3}
    cfgd.update(extra)
    line = {"metric": METRIC, "value": value, "unit": "elements/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": cfgd, "roofline": roof, "gpu_launches": int(launches),
            "clocks": sampler.summary(), "e2e": e2e}
    if not args.no_cpu_baseline and world == 1:
        try:
            cb, _ = config_cpu_baseline(cfg, args.cpu_sample if args.cpu_sample_given else CPU_SAMPLE_N[cfg], 2, cores)
            if cb is not None:
                line["cpu_baseline"] = cb
        except Exception as e:  # noqa: BLE001
            sys.stderr.write("cpu_baseline failed: %r\n" % (e,))
    emit_line(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


_RESULT_FD = None


def claim_stdout():
    """stdout carries exactly one JSON line.  Libraries loaded later (NCCL prints its version banner on stdout when a
    communicator or a unique id is made) must not get at it: file descriptor 1 is pointed at stderr for the whole run and
    the result line is written to the saved descriptor."""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit_line(text):
    sys.stdout.flush()
    if _RESULT_FD is None:
        sys.stdout.write(text + "\n"); sys.stdout.flush()
    else:
        os.write(_RESULT_FD, (text + "\n").encode())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C2", choices=["C1", "C2", "C3", "C4", "C5"],
                    help="BASELINE.json config; C2 (default) is the headline, the others run the generic kernels")
    ap.add_argument("--n", "--size", dest="n", type=int, default=None,
                    help="elements per direction (C2: per GPU; default 256); use --size under torchrun, whose parser claims --n")
    ap.add_argument("--cpu-sample", type=int, default=None, help="edge length of the CPU-baseline sample mesh")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="C2 on N GPUs: weak = n^3 elements per GPU (default, the driver's scaling run), strong = one n^3 mesh cut into N slabs")
    ap.add_argument("--partition", default="auto", choices=["auto", "blocks", "slab"],
                    help="C5 on N GPUs: slab (default) generates each rank's z-slab directly; blocks cuts one global mesh along a Z-curve")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    claim_stdout()
    args.n_given, args.cpu_sample_given = args.n is not None, args.cpu_sample is not None
    if args.n is None:
        args.n = 256
    if args.cpu_sample is None:
        args.cpu_sample = 64
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cores = os.cpu_count() or 1

    # ------------------------------------------------------------------------------------------------- reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        ns = args.cpu_sample
        if args.config != "C2":
            cb, ms = config_cpu_baseline(args.config, ns if args.cpu_sample_given else CPU_SAMPLE_N[args.config], max(1, min(args.steps, 3)), cores)
            if cb is None:
                emit_line(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ref_driver_omp missing or failed"}))
                return
            emit_line(json.dumps({"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "elements/s", "n_gpus": args.gpus,
                              "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                              "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": {"workload": cb["sample"]},
                              "cpu_baseline": cb,
                              "e2e": {"value": cb["value"], "unit": "elements/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
            return
        cb, ms = reference_baseline(ns, max(1, min(args.steps, 5)), min(args.warmup, 1), cores)
        line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "elements/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "C2: 3D scalar Laplace Q1 hex structured mesh, stiffness + RHS (CPU sample %d^3)" % ns},
                "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": "elements/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit_line(json.dumps(line))
        return

    # ------------------------------------------------------------------------------------------------------ our arm
    if args.config != "C2":
        bench_config(args, rank, world, local_rank, cores)
        return
    import torch
    import torch.distributed as dist
    from insilico_b200 import engine as E
    from insilico_b200 import partition

    torch.cuda.set_device(local_rank)
    numa_node = bind_to_gpu_numa(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n = args.n
    strong = args.scaling == "strong" and world > 1
    wl = build_workload(n, rank, world, strong)
    eng = E.Engine(local_rank)
    part = partition.DistributedAssembly(eng, wl, rank, world) if world > 1 else None
    stream = torch.cuda.ExternalStream(eng.stream, device=local_rank)

    eng.set_mesh(E.HEX, 1, wl["coords"], wl["conn"])
    if world > 1:
        part.setup_fields()
    else:
        eng.set_field(0, 1, 1, wl["n_obj"], wl["elem_dof"], wl["eqn"], wl["status"], wl["presc"], wl["values"])
    n_eqn = wl["n_eqn_local"]
    eng.new_solver(n_eqn)
    t0 = time.perf_counter()
    eng.register_fields(0, 0)
    if world > 1:
        part.setup_exchange()
    eng.synchronize()
    t_register = time.perf_counter() - t0
    n_elems_rank = wl["n_owned_elems"]

    def step(events=None):
        eng.new_solver(n_eqn)
        if events is not None:
            events[0].record(stream)
        # the engine defers the launch of the stiffness kernel by one call: the body force on the same field is fused
        # into the same pass over the elements, so the dominant kernel is timed around both calls
        eng.stiffness_matrix_computation(E.K_LAPLACE, [1.0], 3, 0, 0, True)
        eng.body_force_computation([1.0], 3, 0)
        if events is not None:
            events[1].record(stream)
        if world > 1 and not os.environ.get("BENCH_NO_EXCHANGE"):
            part.exchange()

    def barrier():
        eng.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = eng.kernel_launches
    ev_start, ev_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    ev_start.record(stream)
    for k in range(args.steps):
        step(kev[k])
    ev_end.record(stream)
    barrier()
    launches = eng.kernel_launches - l0
    ms_total = ev_start.elapsed_time(ev_end)
    ms_kernel = sum(a.elapsed_time(b) for a, b in kev) / args.steps
    if os.environ.get("BENCH_ALL_RANKS"):
        print("[bench] rank %d: %.3f ms/step, dominant kernel %.3f ms, launches %d" %
              (rank, ms_total / args.steps, ms_kernel, launches), file=sys.stderr, flush=True)
    if world > 1:
        t = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    n_elems_total = n ** 3 if strong else n_elems_rank * world
    value = n_elems_total / (ms_step * 1e-3)

    # ---- same steps on a randomly perturbed copy of the mesh (no element is affine any more): reported next to the
    # headline so that the affine-element shortcut of the local matrix is visible as what it is
    ms_step_perturbed = None
    if world == 1:
        from insilico_b200 import meshgen
        pert = meshgen.perturb_interior(wl["coords"], 1.0 / n, max_dist=0.1)
        eng.update_coords(pert)
        for _ in range(2):
            step()
        barrier()
        ev_a, ev_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev_a.record(stream)
        for k in range(args.steps):
            step()
        ev_b.record(stream)
        barrier()
        ms_step_perturbed = ev_a.elapsed_time(ev_b) / args.steps
        eng.update_coords(wl["coords"])
        step()
        barrier()
        del pert

    # ---- end to end through the C ABI with host buffers: H2D of coordinates + field state, D2H of values + rhs
    e2e = None
    if not args.no_e2e:
        nnz = eng.finish_assembly()[1]
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
        h_coords, h_presc, h_vals = pin(wl["coords"]), pin(wl["presc"]), pin(wl["values"])
        h_val = torch.empty(nnz, dtype=torch.float64).pin_memory().numpy()
        h_rhs = torch.empty(n_eqn, dtype=torch.float64).pin_memory().numpy()

        def e2e_step():
            eng.update_coords(h_coords)
            eng.update_field(0, prescribed=h_presc, values=h_vals)
            step()
            eng.get_csr(None, None, h_val, h_rhs)   # synchronises

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        ke = max(2, min(args.steps, 5))
        for _ in range(ke):
            e2e_step()
        barrier()
        dt_serial = (time.perf_counter() - t0) / ke
        # the same with the hand-off of step k overlapping the uploads and kernels of step k+1 (isl_get_csr_async: second
        # set of value buffers, copy stream); every step still uploads its inputs and downloads its full result
        h_val2 = torch.empty(nnz, dtype=torch.float64).pin_memory().numpy()
        h_rhs2 = torch.empty(n_eqn, dtype=torch.float64).pin_memory().numpy()
        outs = [(h_val, h_rhs), (h_val2, h_rhs2)]

        def e2e_step_async(k):
            eng.update_coords(h_coords)
            eng.update_field(0, prescribed=h_presc, values=h_vals)
            step()
            eng.get_csr_async(*outs[k % 2])

        for k in range(2):
            e2e_step_async(k)
        eng.copy_wait(); barrier()
        t0 = time.perf_counter()
        for k in range(ke):
            e2e_step_async(k)
        eng.copy_wait()
        barrier()
        dt = (time.perf_counter() - t0) / ke
        eng.new_solver(n_eqn)
        if world > 1:
            t = torch.tensor([dt, dt_serial], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt, dt_serial = float(t[0].item()), float(t[1].item())
        e2e = {"value": n_elems_total / dt, "unit": "elements/s",
               "h2d_bytes_per_step": int(h_coords.nbytes + h_presc.nbytes + h_vals.nbytes),
               "d2h_bytes_per_step": int(h_val.nbytes + h_rhs.nbytes), "ms_per_step": dt * 1e3,
               "ms_per_step_serial": dt_serial * 1e3,
               "what": "per step: isl_mesh_update_coords + isl_field_update (pinned H2D), isl_system_create, isl_assemble_matrix, "
                       "isl_assemble_bodyforce, isl_get_csr_async values+rhs (pinned D2H on a copy stream, overlapping the next "
                       "step); isl_copy_wait at the end; pattern cached.  ms_per_step_serial: the same with the blocking isl_get_csr"}
    # ---- one device-resident Newton-type iteration (solid/CompressibleDriver.hpp:179-210 order of calls): assembly, Jacobi
    # preconditioned CG on the device (isl_solve_cg = Eigen3::cgSolve), update of the field on the device (isl_distribute =
    # dof::addToDoFsFromSolver); the matrix never leaves the GPU, only the updated field values cross PCIe
    e2e_newton = None
    if not args.no_e2e and world == 1:
        h_u = torch.empty(wl["n_obj"], dtype=torch.float64).pin_memory().numpy()
        step(); barrier()
        t0 = time.perf_counter()
        step()
        eng.finish_assembly()
        t1 = time.perf_counter()
        cg_it, cg_err = eng.cg_solve(tol=1e-8, max_iter=400)
        eng.synchronize()
        t2 = time.perf_counter()
        eng.distribute(0, add=True)
        _chk_vals = E.lib().isl_field_get_values(eng.h, 0, h_u.ctypes.data_as(__import__("ctypes").c_void_p))
        t3 = time.perf_counter()
        eng.update_field(0, values=wl["values"])   # back to the benchmark state
        e2e_newton = {"value": n_elems_total / (t3 - t0), "unit": "elements/s", "ms_per_step": (t3 - t0) * 1e3,
                      "assemble_ms": (t1 - t0) * 1e3, "cg_ms": (t2 - t1) * 1e3, "cg_iterations": int(cg_it), "cg_relative_residual": float(cg_err),
                      "cg_tolerance": 1e-8, "cg_max_iter": 400, "update_and_copy_ms": (t3 - t2) * 1e3,
                      "h2d_bytes_per_step": 0, "d2h_bytes_per_step": int(h_u.nbytes),
                      "what": "isl_system_create + isl_assemble_matrix + isl_assemble_bodyforce + isl_finish, isl_solve_cg (capped at 400 "
                              "iterations), isl_distribute(add), isl_field_get_values -> pinned host: the CSR stays on the device"}
    if rank == 0:
        sampler.stop_flag = True
        sampler.join(timeout=2)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = peaks()
    achieved = ALGO_BYTES_PER_ELEM * n_elems_rank / (ms_kernel * 1e-3) / 1e9
    traffic = traffic_general = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
            if n == 256 and not strong:   # measured at this size only (ncu --set full, profiles/r2)
                traffic = tj.get("k_q1hex_rows_affine_bytes_per_launch_256")
                traffic_general = tj.get("k_q1hex_general_two_kernel_bytes_per_step_256")
    except Exception:
        pass
    fp64_peak = eng.measure_fp64_peak()
    f64_ach = ALGO_FLOPS_PER_ELEM * n_elems_rank / (ms_kernel * 1e-3) / 1e12
    roof_general = None
    if ms_step_perturbed is not None:
        ach_g = ALGO_BYTES_PER_ELEM * n_elems_rank / (ms_step_perturbed * 1e-3) / 1e9
        roof_general = {"bound": "hbm", "kernel": "k_q1hex_elemK + k_q1hex_rows_fromK (general, non-affine elements: the same mesh with "
                        "randomly perturbed nodes; element matrices once, then row gather; two launches per step)",
                        "achieved": ach_g, "peak": peak, "unit": "GB/s", "frac": ach_g / peak, "traffic": traffic_general,
                        "kernel_ms": ms_step_perturbed, "algorithmic_bytes_per_element": ALGO_BYTES_PER_ELEM,
                        "fp64": {"achieved": ALGO_FLOPS_PER_ELEM * n_elems_rank / (ms_step_perturbed * 1e-3) / 1e12, "peak": fp64_peak,
                                 "unit": "TFLOP/s", "frac": ALGO_FLOPS_PER_ELEM * n_elems_rank / (ms_step_perturbed * 1e-3) / 1e12 / fp64_peak}}
    line = {"metric": METRIC, "value": value, "unit": "elements/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C2: 3D scalar Laplace Q1 hex, structured %d^3 mesh %s, stiffness + RHS "
                                   "(Dirichlet lift + constant body force), quadrature degree 3" % (n, "cut into %d slabs" % world if strong else "per GPU"),
                       "n_elems_per_gpu": int(n_elems_rank), "n_eqn_per_gpu": int(n_eqn),
                       "nnz_per_gpu": int(eng.finish_assembly()[1]),
                       "partition": "z-slabs by element blocks, owned row ranges" if world > 1 else "single GPU",
                       "l2": "inputs larger than L2: %.1f GB touched per step, no flush needed"
                             % (ALGO_BYTES_PER_ELEM * n_elems_rank / 1e9),
                       "register_fields_ms": t_register * 1e3,
                       "ms_per_step_perturbed_mesh": ms_step_perturbed},
            "achieved_hbm_gbs": ALGO_BYTES_PER_ELEM * value / world / 1e9,
            "roofline": {"bound": "hbm", "kernel": "k_q1hex_rows_affine<128,4> (all-affine mesh: stiffness + Dirichlet lift + body force in one "
                                                   "launch, bulk-copy write-out)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "kernel_ms": ms_kernel, "algorithmic_bytes_per_element": ALGO_BYTES_PER_ELEM,
                         "fp64": {"achieved": f64_ach, "peak": fp64_peak, "unit": "TFLOP/s", "frac": f64_ach / fp64_peak if fp64_peak else None,
                                  "algorithmic_flops_per_element": ALGO_FLOPS_PER_ELEM,
                                  "note": "algorithmic flops of the general element (SURVEY 8d) over the kernel time; the all-affine kernel "
                                          "executes ~215 FP64 operations per matrix row instead (stencil sums), so this can exceed 1; "
                                          "roofline_nonaffine is the like-for-like figure",
                                  "peak_source": "measured (isl_measure_fp64_peak: DFMA chains, CUDA events, this run)"}},
            "roofline_nonaffine": roof_general,
            "gpu_launches": int(launches), "clocks": sampler.summary(), "e2e": e2e}
    if e2e_newton is not None:
        line["e2e_newton"] = e2e_newton
    if numa_node is not None:
        line["config"]["numa_node"] = numa_node
    if not args.no_cpu_baseline and world == 1:  # reported on rank 0 at N=1 only
        try:
            line["cpu_baseline"], _ = reference_baseline(args.cpu_sample, 2, 1, cores)
        except Exception as e:  # noqa: BLE001  (the measured GPU line must not be lost to a failure of the CPU leg)
            sys.stderr.write("cpu_baseline failed: %r\n" % (e,))
        api = reference_api_on_engine(args.cpu_sample, 3)   # separate process with its own engine on the same device
        if api is not None:
            line["e2e_reference_api"] = api
    emit_line(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
