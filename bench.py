#!/usr/bin/env python
"""bench.py -- particle-steps/s of the dsmcCloud::evolve() hot path (move + sort + NTC collide + sample).

Workload (BASELINE.json configs[4], the one its metric "at 1/2/4/8 B200" is quoted on): a
weak-scaling periodic box, one brick of cells per GPU tiled 1x1x1 -> 2x1x1 -> 2x2x1 -> 2x2x2
(processor / processorCyclic patches between bricks, NCCL parcel migration), ~31 parcels per
cell, argon VHS or 5-species air Larsen-Borgnakke, equilibrium at rest.  A "step" is one full
evolve() of every parcel of every brick.

  python bench.py --gpus N --steps K --warmup W            (torchrun launches one rank per GPU for N>1)
  python bench.py --impl reference ...                     CPU oracle arm on the host cores

One JSON line is printed by rank 0 (see README / DESIGN.md for the keys).
"""
from __future__ import annotations

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
# stdout carries the one JSON line only: native libraries (NCCL's version banner, ...) write to file descriptor 1 behind Python's
# back, so fd 1 is pointed at stderr for the whole run and the JSON line goes to the saved descriptor
_REAL_STDOUT = None


def _capture_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()

KB = 1.38065e-23
# algorithmic bytes per parcel-step and stage (BASELINE.md section 3 / SURVEY.md 8d), FP64 state
STAGE_BYTES_AIR = {"move": 96.0, "sort": 168.0, "collide": 71.0, "sample": 40.0}
STAGE_BYTES_AR = {"move": 96.0, "sort": 136.0, "collide": 63.0, "sample": 28.0}   # no ERot / vibLevel / ELevel
SORT_KERNELS = ("scan", "scatterIndex", "segmentSort", "gather", "histogram")
# the engine times stages; these timers bracket more than one kernel (3-kernel scan; lane + big-cell collide kernels)
KERNELS_PER_TIMER = {"scan": 3, "collide": 2, "movePlan": 5}


def species_table(gas):
    from hystrath_b200 import capi

    if gas == "argon":
        return [capi.make_species("Ar", 66.3e-27, 4.17e-10, 0.81)], [0], [1.0]
    sp = [
        capi.make_species("N2", 46.5e-27, 4.17e-10, 0.74, 1.36, 2, (3371,), (52560,), (3371,), 113500),
        capi.make_species("O2", 53.12e-27, 4.07e-10, 0.77, 1.4, 2, (2256,), (17900,), (2256,), 59500),
        capi.make_species("NO", 49.81e-27, 4.2e-10, 0.79, 1.0, 2, (2719,), (1400,), (2719,), 75500),
        capi.make_species("N", 23.25e-27, 3.0e-10, 0.8, 1.0),
        capi.make_species("O", 26.56e-27, 3.0e-10, 0.8, 1.0),
    ]
    return sp, [0, 1, 2, 3, 4], [0.76, 0.2, 0.02, 0.01, 0.01]


def case_parameters(gas, cells, ppc):
    """Equilibrium box: cell size ~ lambda/3, dt ~ tau_c/5."""
    n = 1e20
    T = 300.0 if gas == "argon" else 1000.0
    dx = 4e-3
    L = cells * dx
    fnum = n * dx ** 3 / ppc
    dt = 5e-6 if gas == "argon" else 2.5e-6
    return dict(n=n, T=T, dx=dx, L=L, fnum=fnum, dt=dt)


def host_cores():
    """Cores this process may run on (the cgroup / affinity mask, not the machine's count)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def procs_for(n):
    return {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[n]


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = False

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for s in self.samples for k in range(4) if len(s) > 2 + k and s[2 + k].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.samples)}


def run_reference(args):
    """--impl reference: the reference's algorithm on the host cores.  dsmcFoam+ itself cannot be built here
    (needs OpenFOAM v1706 + MPI, neither vendored nor installed), so this is the oracle port, OpenMP over all cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from hystrath_b200 import capi, meshgen
    from oracle import pyoracle
    from oracle.pyoracle import Oracle

    cores = pyoracle.set_threads(host_cores())   # the team size actually obtained (torchrun exports OMP_NUM_THREADS=1)
    cells = args.ref_cells
    if getattr(args, "c1", False):
        cells = 32
    cp = case_parameters(args.gas, cells, args.ppc)
    sp, tids, frac = species_table(args.gas)
    mesh = meshgen.box_mesh((cells,) * 3, (cp["L"],) * 3)
    if getattr(args, "numbering", "morton") == "morton":
        mesh, _ = meshgen.renumber_cells(mesh, meshgen.morton_order(mesh))
    model = "VariableHardSphere" if args.gas == "argon" else "LarsenBorgnakkeVariableHardSphere"
    md = capi.build_models(model, nEquivalentParticles=cp["fnum"], deltaT=cp["dt"], seed=0xD5C00005)
    o = Oracle()
    o.set_mesh(mesh); o.set_species(sp); o.set_models(md)
    o.mesh_fill(tids, [cp["n"] * f for f in frac], cp["T"], cp["T"], cp["T"])
    n = o.num_parcels()
    o.evolve(args.warmup)
    t0 = time.perf_counter()
    o.evolve(args.steps)
    dt = time.perf_counter() - t0
    value = n * args.steps / dt
    sample = (f"{cells}^3-cell periodic {args.gas} box, {n} parcels, {args.steps} steps after {args.warmup} warm-up: same cell size, density, dt, models and "
              f"cell numbering as the GPU arm's brick, a smaller brick (the oracle is a CPU port; the largest box whose run stays within a few minutes)")
    line = {
        "impl": "reference", "metric": "particle-steps/s (move+sort+NTC collide+sample)", "value": value, "unit": "particle-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, cells_per_gpu=cells ** 3, parcels_per_gpu=n),
        "cpu_baseline": {"value": value, "unit": "particle-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(args, cells_per_gpu, parcels_per_gpu):
    if getattr(args, "workload", "box") == "wedge":
        return {"workload": "5-species air (N2,O2,NO,N,O) Larsen-Borgnakke over a sharp 15-degree wedge at the orion107kmNR free stream "
                            "(6053.4 m/s, 217.63 K, Mach 20; BASELINE configs[2]): %s cells, ~%d parcels per GPU, diffuse 1000 K wall, "
                            "free-stream inflow + deletion, symmetry planes" % (args.wedge, parcels_per_gpu),
                "cells_per_gpu": cells_per_gpu, "parcels_per_gpu": parcels_per_gpu, "parcels_per_cell": args.ppc, "gas": "air5",
                "collision_model": "LarsenBorgnakkeVariableHardSphere", "partition": "%d slab(s) along x (decomposePar simple)" % args.gpus,
                "l2_policy": "inputs larger than L2 (parcel state >> 126 MB), no flush needed"}
    if getattr(args, "workload", "box") == "capsule":
        return {"workload": "3-D re-entry capsule forebody in 5-species air (N2,O2,NO,N,O) Larsen-Borgnakke at the orion107kmNR free stream "
                            "(6053.4 m/s, 217.63 K; BASELINE configs[3]): %d^3 cells and ~%d parcels per GPU, spherical-segment heat shield "
                            "(diffuse 1000 K wall) on the x = max plane, outflow round the shoulder, free-stream inflow + deletion on the "
                            "other five planes" % (args.cells, parcels_per_gpu),
                "cells_per_gpu": cells_per_gpu, "parcels_per_gpu": parcels_per_gpu, "parcels_per_cell": args.ppc, "gas": "air5",
                "collision_model": "LarsenBorgnakkeVariableHardSphere",
                "partition": "x".join(str(v) for v in procs_for(args.gpus)) + " bricks (decomposePar simple), NCCL migration across processor patches",
                "l2_policy": "inputs larger than L2 (parcel state >> 126 MB per GPU), no flush needed"}
    if getattr(args, "workload", "box") == "cylinder":
        return {"workload": "2-D Mach-10 argon flow over a cylinder (BASELINE configs[1], Lofthouse): O-grid %s cells, ~%d parcels, VHS, "
                            "diffuse 500 K wall, free-stream inflow + deletion" % (args.cyl, parcels_per_gpu),
                "cells_per_gpu": cells_per_gpu, "parcels_per_gpu": parcels_per_gpu, "parcels_per_cell": args.ppc, "gas": "argon",
                "collision_model": "VariableHardSphere", "partition": "1 GPU",
                "l2_policy": "inputs larger than L2 (parcel state >> 126 MB), no flush needed"}
    if getattr(args, "c1", False):
        return {"workload": "C1 (BASELINE configs[0]): periodic argon VHS heat-bath box, 32^3 cells, %d parcels, no walls" % parcels_per_gpu,
                "cells_per_gpu": cells_per_gpu, "parcels_per_gpu": parcels_per_gpu, "parcels_per_cell": args.ppc, "gas": "argon",
                "collision_model": "VariableHardSphere", "partition": "1 GPU",
                "l2_policy": "the whole cloud (1 M parcels, 100 MB with both buffers) fits the 126 MB L2: L2-resident by nature of the case, no flush"}
    return {"workload": "weak-scaling periodic box (BASELINE configs[4]), %s, %d cells and ~%d parcels per GPU, %s" % (
        "argon VHS" if args.gas == "argon" else "5-species air (N2,O2,NO,N,O) Larsen-Borgnakke VHS", cells_per_gpu, parcels_per_gpu,
        {"blockMesh": "blockMesh cell order", "engine": "blockMesh cell order, relabelled along a z-order curve inside the library (dsmcb200_set_cell_order)"}.get(
            getattr(args, "numbering", "morton"), "cells renumbered along a z-order curve (renumberMesh equivalent)")),
        "cells_per_gpu": cells_per_gpu, "parcels_per_gpu": parcels_per_gpu, "parcels_per_cell": args.ppc, "gas": args.gas,
        "collision_model": "VariableHardSphere" if args.gas == "argon" else "LarsenBorgnakkeVariableHardSphere",
        "partition": "x".join(str(v) for v in procs_for(args.gpus)) + " bricks",
        "cell_numbering": {"blockMesh": "blockMesh order (x fastest)",
                           "engine": "blockMesh order (x fastest); the library relabels along a z-order curve (dsmcb200_set_cell_order)"}.get(
                               getattr(args, "numbering", "morton"), "cells relabelled along a z-order curve (renumberMesh equivalent)"),
        "l2_policy": "inputs larger than L2 (parcel state >> 126 MB per GPU), no flush needed"}


def cpu_baseline_leg(args):
    """Oracle (port of the reference algorithm) on a bounded sample of the same workload, rank 0 / N=1 only."""
    from hystrath_b200 import capi, meshgen
    from oracle import pyoracle
    from oracle.pyoracle import Oracle

    cores = pyoracle.set_threads(host_cores())
    if getattr(args, "workload", "box") == "wedge":
        from hystrath_b200 import cases

        mesh, sp, md, fill = cases.air_wedge(200, 100, 2, 25)
        o = Oracle()
        o.set_mesh(mesh); o.set_species(sp); o.set_models(md)
        o.mesh_fill(fill["type_ids"], fill["number_densities"], fill["Ttra"], fill["Trot"], fill["Tvib"], 0.0, fill["velocity"])
        n = o.num_parcels()
        o.evolve(1)
        t0 = time.perf_counter()
        o.evolve(args.cpu_steps)
        dt = time.perf_counter() - t0
        return {"value": n * args.cpu_steps / dt, "unit": "particle-steps/s", "cores": cores, "kind": "port",
                "sample": f"200x100x2-cell wedge, {n} parcels, {args.cpu_steps} steps (oracle, OpenMP over {cores} threads)"}
    if getattr(args, "workload", "box") == "capsule":
        from hystrath_b200 import cases

        mesh, sp, md, fill = cases.capsule_forebody((32, 32, 32), 31)
        o = Oracle()
        o.set_mesh(mesh); o.set_species(sp); o.set_models(md)
        o.mesh_fill(fill["type_ids"], fill["number_densities"], fill["Ttra"], fill["Trot"], fill["Tvib"], 0.0, fill["velocity"])
        n = o.num_parcels()
        o.evolve(1)
        t0 = time.perf_counter()
        o.evolve(args.cpu_steps)
        dt = time.perf_counter() - t0
        return {"value": n * args.cpu_steps / dt, "unit": "particle-steps/s", "cores": cores, "kind": "port",
                "sample": f"32^3-cell capsule forebody, {n} parcels, {args.cpu_steps} steps (oracle, OpenMP over {cores} threads)"}
    if getattr(args, "workload", "box") == "cylinder":
        from hystrath_b200 import cases

        mesh, sp, md, fill = cases.lofthouse_cylinder(160, 312, 25)
        o = Oracle()
        o.set_mesh(mesh); o.set_species(sp); o.set_models(md)
        o.mesh_fill(fill["type_ids"], fill["number_densities"], fill["Ttra"], 0.0, 0.0, 0.0, fill["velocity"])
        n = o.num_parcels()
        o.evolve(1)
        t0 = time.perf_counter()
        o.evolve(args.cpu_steps)
        dt = time.perf_counter() - t0
        return {"value": n * args.cpu_steps / dt, "unit": "particle-steps/s", "cores": cores, "kind": "port",
                "sample": f"160x312-cell cylinder O-grid, {n} parcels, {args.cpu_steps} steps (oracle, OpenMP over {cores} threads)"}
    cells = args.cpu_cells
    cp = case_parameters(args.gas, cells, args.ppc)
    sp, tids, frac = species_table(args.gas)
    mesh = meshgen.box_mesh((cells,) * 3, (cp["L"],) * 3)
    if getattr(args, "numbering", "morton") == "morton":
        mesh, _ = meshgen.renumber_cells(mesh, meshgen.morton_order(mesh))
    model = "VariableHardSphere" if args.gas == "argon" else "LarsenBorgnakkeVariableHardSphere"
    md = capi.build_models(model, nEquivalentParticles=cp["fnum"], deltaT=cp["dt"], seed=0xD5C00001 if getattr(args, "c1", False) else 0xD5C00005)
    o = Oracle()
    o.set_mesh(mesh); o.set_species(sp); o.set_models(md)
    o.mesh_fill(tids, [cp["n"] * f for f in frac], cp["T"], cp["T"], cp["T"])
    n = o.num_parcels()
    if getattr(args, "c1", False):
        # the protocol of BASELINE.md section 4: 200 steps after 20 warm-up on all host threads, and the 1-thread figure on a shorter run
        o.evolve(20)
        t0 = time.perf_counter()
        o.evolve(200)
        dt_all = time.perf_counter() - t0
        one = pyoracle.set_threads(1)
        t0 = time.perf_counter()
        o.evolve(args.c1_single_thread_steps)
        dt_one = time.perf_counter() - t0
        pyoracle.set_threads(cores)
        return {"value": n * 200 / dt_all, "unit": "particle-steps/s", "cores": cores, "kind": "port",
                "single_thread": {"value": n * args.c1_single_thread_steps / dt_one, "cores": one, "steps": args.c1_single_thread_steps},
                "sample": f"C1 (BASELINE.md section 4): 32^3-cell periodic argon VHS box, {n} parcels, 200 steps after 20 warm-up "
                          f"(oracle = port of the reference algorithm, OpenMP over {cores} threads; dsmcFoam+ itself cannot be built here)"}
    o.evolve(1)
    steps = args.cpu_steps
    t0 = time.perf_counter()
    o.evolve(steps)
    dt = time.perf_counter() - t0
    return {"value": n * steps / dt, "unit": "particle-steps/s", "cores": cores, "kind": "port",
            "sample": f"{cells}^3-cell periodic {args.gas} box, {n} parcels, {steps} steps (oracle, OpenMP over {cores} threads)"}


def decomposed_parity_check(rank, world, local, dist, torch):
    """world > 1, outside the timed region: a small Larsen-Borgnakke case on the same brick tiling, every rank's engine (NCCL migration)
    against its own instance of the CPU oracle driven through the reference's transfer protocol (Cloud<T>::move: per-neighbour lists,
    rounds until no rank sent).  The cloud must come out in the same list order with the same cells and vibrational levels, the same
    collisions, and the per-neighbour migration counts of every step.  Returns a dict for `invariants`."""
    from hystrath_b200 import capi, meshgen
    from oracle import pyoracle
    from oracle.pyoracle import Oracle

    pyoracle.set_threads(2)
    n_local, l_local, steps = (4, 4, 3), (0.016, 0.016, 0.012), 4
    procs = procs_for(world)
    sp, _, _ = species_table("air5")
    sp = sp[:2]
    mesh = meshgen.decomposed_box(n_local, l_local, procs, rank)
    fnum = 2e21 * l_local[0] * l_local[1] * l_local[2] / (48 * 40)
    md = capi.build_models("LarsenBorgnakkeVariableHardSphere", nEquivalentParticles=fnum, deltaT=6e-6, seed=79)
    ora = Oracle()
    ora.set_mesh(mesh); ora.set_species(sp); ora.set_models(md); ora.set_rank(rank)
    ora.mesh_fill([0, 1], [1.5e21, 0.5e21], 4000.0, 4000.0, 4000.0, 0.0, (300.0, 100.0, -50.0))
    start = ora.download_parcels()
    start.origProc = np.full(start.n, rank, np.int32)
    sig, rem = ora.download_cellstate()
    eng = capi.Engine(local, rank, world)
    eng.set_mesh(mesh); eng.set_species(sp); eng.set_models(md)
    ident = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        ident.copy_(torch.frombuffer(bytearray(capi.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(ident, 0)
    eng.init_comm(bytes(ident.cpu().numpy().tobytes()))
    eng.upload_parcels(start)
    eng.upload_cellstate(sig, rem)
    # arrivals are appended neighbour by neighbour in the order of ProcessorTopology::procNeighbours: first appearance in patch order
    nbr_order = []
    for pt in mesh.patches:
        if pt["type"] in ("processor", "processorCyclic") and pt["neighbProcNo"] not in nbr_order:
            nbr_order.append(pt["neighbProcNo"])
    ok, why = True, ""
    sent_total = 0
    for step in range(steps):
        eng.evolve(1)
        c = eng.counters()
        ora.evolve_begin()
        sent = {}
        while True:
            d, i = ora.outbox()
            out = {dst: (d[i[:, 0] == dst], i[i[:, 0] == dst]) for dst in set(i[:, 0].tolist())}
            for dst, (dd, _) in out.items():
                sent[dst] = sent.get(dst, 0) + len(dd)
            gathered = [None] * world
            dist.all_gather_object(gathered, out)
            if not any(len(v[0]) for g in gathered for v in g.values()):
                break
            for src in nbr_order:
                if rank in gathered[src] and len(gathered[src][rank][0]):
                    ora.receive_and_move(src, *gathered[src][rank])
        ora.evolve_end()
        mine = {int(c.neighbourProc[k]): int(c.migratedTo[k]) for k in range(c.nNeighbours) if c.migratedTo[k]}
        if mine != {k: v for k, v in sent.items() if v}:
            ok, why = False, f"step {step}: migration counts per neighbour {mine} != oracle {sent}"
        sent_total += sum(sent.values())
    g, o = eng.download_parcels(), ora.download_parcels()
    if ok and not (g.n == o.n and np.array_equal(g.origId, o.origId) and np.array_equal(g.origProc, o.origProc)):
        ok, why = False, "cloud list order differs"
    if ok and not (np.array_equal(g.cell, o.cell) and np.array_equal(g.vibLevel, o.vibLevel)):
        ok, why = False, "cells or vibrational levels differ"
    if ok and not (np.allclose(g.U, o.U, rtol=0, atol=1e-8) and np.allclose(g.position, o.position, rtol=0, atol=1e-12)):
        ok, why = False, "velocities or positions differ"
    eng.close()
    flag = torch.tensor([1.0 if ok else 0.0, float(sent_total), float(g.n)], dtype=torch.float64, device="cuda")
    allok = flag[:1].clone()
    dist.all_reduce(allok, op=dist.ReduceOp.MIN)
    dist.all_reduce(flag[1:], op=dist.ReduceOp.SUM)
    whys = [None] * world
    dist.all_gather_object(whys, why)
    return {"ok": bool(allok.item() == 1.0), "ranks": world, "steps": steps, "parcels": int(flag[2].item()), "migrated": int(flag[1].item()),
            "checked": "list order (origProc, origId), cells, vibrational levels, migration counts per neighbour and step: exact; U to 1e-8 m/s and positions to 1e-12 m (libm vs CUDA ulps in the collision model)",
            "failures": [w for w in whys if w]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="dsmcb200", choices=["dsmcb200", "reference"])
    ap.add_argument("--workload", default=os.environ.get("DSMCB200_BENCH_WORKLOAD", "box"), choices=["box", "cylinder", "wedge", "c1", "capsule"],
                    help="box: BASELINE configs[4] weak-scaling periodic box (default, any N); cylinder: configs[1] Mach-10 argon cylinder (N=1); "
                         "wedge: configs[2] 5-species air over a hypersonic wedge (N > 1: the same case cut into N slabs along x, strong scaling)")
    ap.add_argument("--wedge", default="2000x1000x4", help="wedge cells nx x ny x nz")
    ap.add_argument("--c1-single-thread-steps", type=int, default=20, help="workload c1: steps of the 1-thread leg of BASELINE.md section 4")
    ap.add_argument("--cyl", default="640x1250", help="cylinder O-grid cells nr x ntheta")
    ap.add_argument("--gas", default=os.environ.get("DSMCB200_BENCH_GAS", "air5"), choices=["argon", "air5"])
    ap.add_argument("--cells", type=int, default=int(os.environ.get("DSMCB200_BENCH_CELLS", "200")), help="cells per direction per GPU")
    ap.add_argument("--ppc", type=int, default=31)
    ap.add_argument("--numbering", default=os.environ.get("DSMCB200_BENCH_NUMBERING", "morton"), choices=["blockMesh", "morton", "engine"],
                    help="cell labels of the box: blockMesh's x-fastest order, or relabelled along a z-order curve (meshgen.renumber_cells, a renumberMesh equivalent), or blockMesh's order relabelled by the library itself (engine)")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-cells", type=int, default=32)
    ap.add_argument("--cpu-steps", type=int, default=5)
    ap.add_argument("--ref-cells", type=int, default=96, help="cells per direction of the reference arm's sample: the largest box that keeps a 20-step run on 16 host threads within a few minutes")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    _capture_stdout()
    if args.warmup < 3 and args.impl == "dsmcb200":
        args.warmup = max(args.warmup, 0)
    c1 = args.workload == "c1"
    if c1:
        # BASELINE.md section 4 / SURVEY 8d C1: periodic argon VHS heat bath, 32^3 cells, 1 048 576 parcels expected, dt 5e-6, seed 0xD5C00001
        args.workload, args.gas, args.cells, args.ppc, args.numbering = "box", "argon", 32, 32, "blockMesh"
        args.c1 = True
    if args.impl == "reference":
        run_reference(args)
        return

    import torch

    from hystrath_b200 import capi, meshgen

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    t_setup = time.perf_counter()
    if args.workload == "cylinder":
        if world != 1:
            raise SystemExit("--workload cylinder is a single-GPU configuration (BASELINE configs[1])")
        from hystrath_b200 import cases

        nr, nt = (int(v) for v in args.cyl.split("x"))
        args.gas = "argon"
        args.ppc = 25
        mesh, sp, md, fill = cases.lofthouse_cylinder(nr, nt, args.ppc)
        tids, dens, Tfill, vfill = fill["type_ids"], fill["number_densities"], fill["Ttra"], fill["velocity"]
        n_cells_gpu = mesh.n_cells
    elif args.workload == "capsule":
        from hystrath_b200 import cases

        args.gas = "air5"
        mesh, sp, md, fill = cases.capsule_forebody((args.cells,) * 3, args.ppc, procs=procs_for(world), rank=rank)
        tids, dens, Tfill, vfill = fill["type_ids"], fill["number_densities"], fill["Ttra"], fill["velocity"]
        n_cells_gpu = mesh.n_cells
    elif args.workload == "wedge":
        from hystrath_b200 import cases

        wx, wy, wz = (int(v) for v in args.wedge.split("x"))
        args.gas = "air5"
        args.ppc = 34   # per inlet-cell volume; the rows shrink towards the outlet: ~25 parcels per cell on average, ~200 M in 8 M cells
        mesh, sp, md, fill = cases.air_wedge(wx, wy, wz, args.ppc, procs=world, rank=rank)
        tids, dens, Tfill, vfill = fill["type_ids"], fill["number_densities"], fill["Ttra"], fill["velocity"]
        n_cells_gpu = mesh.n_cells
    else:
        if getattr(args, "c1", False) and world != 1:
            raise SystemExit("--workload c1 is a single-GPU configuration (BASELINE configs[0])")
        cp = case_parameters(args.gas, args.cells, args.ppc)
        sp, tids, frac = species_table(args.gas)
        procs = procs_for(world)
        mesh = meshgen.decomposed_box((args.cells,) * 3, (cp["L"],) * 3, procs, rank)
        if args.numbering == "morton":
            mesh, _ = meshgen.renumber_cells(mesh, meshgen.morton_order(mesh))
        model = "VariableHardSphere" if args.gas == "argon" else "LarsenBorgnakkeVariableHardSphere"
        md = capi.build_models(model, nEquivalentParticles=cp["fnum"], deltaT=cp["dt"], seed=0xD5C00005 + rank)
        dens, Tfill, vfill = [cp["n"] * f for f in frac], cp["T"], (0.0, 0.0, 0.0)
        n_cells_gpu = args.cells ** 3
    eng = capi.Engine(local, rank, world)
    if args.numbering == "engine":      # the mesh keeps blockMesh's labels, the library relabels its cells itself (dsmcb200_set_cell_order)
        eng.set_cell_order("z-curve")
    eng.set_mesh(mesh); eng.set_species(sp); eng.set_models(md)
    if world > 1:
        ident = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            ident = torch.frombuffer(bytearray(capi.nccl_unique_id()), dtype=torch.uint8).cuda()
        dist.broadcast(ident, 0)
        eng.init_comm(bytes(ident.cpu().numpy().tobytes()))
    expected = int(n_cells_gpu * args.ppc * 1.05) + 4096
    eng.reserve(expected)
    eng.mesh_fill(tids, dens, Tfill, Tfill, Tfill, 0.0, vfill)
    n_local = eng.num_parcels()
    setup_s = time.perf_counter() - t_setup

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    decomposed = decomposed_parity_check(rank, world, local, dist, torch) if dist is not None else None
    for _ in range(max(args.warmup, 3)):
        eng.evolve(1)
    eng.kernel_times(reset=True)
    c_before = eng.counters()      # dsmcCloud::info sums (mass, energies) before the timed steps: size-independent invariants below
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    eng.timer_start()
    t0 = time.perf_counter()
    n_processed = 0
    for _ in range(args.steps):
        eng.evolve(1)
        n_processed += eng.num_parcels()
    ms = eng.timer_stop()
    barrier()
    wall = time.perf_counter() - t0
    sampler.stop_flag = True
    kt = eng.kernel_times()
    c_after = eng.counters()
    stage_ms = np.array(c_after.stageMs[:8])
    # ---- parity by property at the full size (outside the timed region): a closed periodic box keeps its parcels and its mass, collisions
    # conserve total energy (translational + rotational + vibrational + electronic), the occupancy offsets are a CSR of the cloud
    def energy(c):
        return c.linearKineticEnergy + c.rotationalEnergy + c.vibrationalEnergy + c.electronicEnergy
    occ = eng.occupancy()
    inv_local = torch.tensor([c_before.mass, c_after.mass, energy(c_before), energy(c_after), float(c_before.nParcels), float(c_after.nParcels),
                              float(c_after.collisions)], dtype=torch.float64, device="cuda")
    occ_ok = bool(occ[0] == 0 and occ[-1] == eng.num_parcels() and (np.diff(occ) >= 0).all())
    occ_t = torch.tensor([1.0 if occ_ok else 0.0], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(inv_local, op=dist.ReduceOp.SUM)
        dist.all_reduce(occ_t, op=dist.ReduceOp.MIN)
    iv = inv_local.cpu().numpy()
    invariants = {"occupancy_is_csr_of_cloud": bool(occ_t.item() == 1.0), "collisions_last_step": int(iv[6])}
    if decomposed is not None:
        invariants["decomposed_parity_vs_oracle"] = decomposed
    if args.workload == "box":
        invariants.update({"parcels_before": int(iv[4]), "parcels_after": int(iv[5]), "mass_rel_change": float(iv[1] / iv[0] - 1.0),
                           "total_energy_rel_change": float(iv[3] / iv[2] - 1.0), "steps_between": args.steps})

    if dist is not None:
        # dsmcDynamicLoadBalancing::update (DSMC/dynamicLoadBalancing/dsmcDynamicLoadBalancing.C:100-149): max |n_rank - n_ideal| / n_ideal
        counts = [None] * world
        dist.all_gather_object(counts, int(eng.num_parcels()))
        ideal = sum(counts) / world
        invariants["load_imbalance"] = {"parcels_per_rank": counts, "maximum_imbalance_pct": 100.0 * max(abs(c - ideal) for c in counts) / ideal}
    tmax = torch.tensor([ms], dtype=torch.float64, device="cuda")
    ntot = torch.tensor([float(n_processed)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(ntot, op=dist.ReduceOp.SUM)
    ms_total = float(tmax.item())
    value = float(ntot.item()) / (ms_total * 1e-3)

    # ---- end to end through the C ABI with host buffers: upload cloud, evolve, download cloud + fields, every step
    e2e = None
    if not args.no_e2e:
        host = capi.ParcelData(int(eng.num_parcels() * 1.02) + 4096, eng.max_modes)   # headroom: inflow changes the count
        host = _download_into(eng, host)
        pinned = {}
        for name, _, _ in capi.ParcelData.FIELDS:
            a = getattr(host, name)
            if a is None:
                continue
            t = torch.from_numpy(a).pin_memory()
            pinned[name] = t
            setattr(host, name, t.numpy())
        row = sum(getattr(host, k)[:1].nbytes for k in ("position", "U", "cell", "tetFace", "tetPt", "typeId", "origId"))
        if args.gas != "argon":
            row += host.ERot[:1].nbytes + host.vibLevel[:1].nbytes + host.ELevel[:1].nbytes
        h2d = row * host.n
        ai = eng.accum_info()
        acc_pin = torch.empty((ai.nCells, ai.nSpecies, ai.nQuantities), dtype=torch.float64).pin_memory()
        coll_pin = torch.empty((ai.nCells, 2), dtype=torch.float64).pin_memory()
        acc_np, coll_np = acc_pin.numpy(), coll_pin.numpy()
        acc_bytes = 0
        barrier()
        t1 = time.perf_counter()
        n_e2e = 0
        for _ in range(args.e2e_steps):
            eng.upload_parcels(host)
            eng.evolve(1)
            host = eng.download_parcels(host) if False else _download_into(eng, host)
            acc, coll, _ = eng.accumulators(acc_np, coll_np)
            acc_bytes = acc.nbytes + coll.nbytes
            n_e2e += host.n
        barrier()
        t_e2e = time.perf_counter() - t1
        te = torch.tensor([t_e2e], dtype=torch.float64, device="cuda")
        ne = torch.tensor([float(n_e2e)], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
            dist.all_reduce(ne, op=dist.ReduceOp.SUM)
        e2e = {"value": float(ne.item()) / float(te.item()), "unit": "particle-steps/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(h2d + acc_bytes), "steps": args.e2e_steps,
               "note": "per step: upload the whole cloud from pinned host memory, evolve(), download cloud and field accumulators"}

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        sb = STAGE_BYTES_AR if args.gas == "argon" else STAGE_BYTES_AIR
        per_step = {k: v[0] / max(1, args.steps) for k, v in kt.items()}
        stage_t = {"move": per_step.get("move", 0.0) + per_step.get("movePlan", 0.0), "sort": sum(per_step.get(k, 0.0) for k in SORT_KERNELS),
                   "collide": per_step.get("collide", 0.0), "sample": per_step.get("sample", 0.0)}
        stages = {}
        for k, t in stage_t.items():
            if t > 0:
                gbs = n_local * sb[k] / (t * 1e-3) / 1e9
                stages[k] = {"ms": t, "bytes_per_parcel": sb[k], "achieved_gbs": gbs, "frac": gbs / peak}
        dom = max(stage_t, key=stage_t.get)
        # DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/r02_traffic.json: dram__bytes_read.sum +
        # dram__bytes_write.sum of one launch on this workload), scaled by parcel count when the capture was taken at another size
        traffic, traffic_src = None, None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
            ent = tj.get(f"{args.workload}:{args.gas}", {}).get(dom)
            if ent:
                traffic = ent["dram_bytes"] / ent["parcels"] * n_local
                traffic_src = f"ncu dram__bytes_read.sum+dram__bytes_write.sum of one {dom} launch at {ent['parcels']} parcels ({tj.get('capture', '')}), scaled by parcel count"
        except Exception:
            pass
        roofline = {"bound": "hbm", "kernel": {"move": "moveKernel", "sort": "gatherKernel+scan+scatterIndex+segmentSort",
                                               "collide": "collideKernel", "sample": "sampleKernel"}[dom],
                    "achieved": stages[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s", "frac": stages[dom]["frac"], "traffic": traffic,
                    "traffic_source": traffic_src, "peak_source": peak_src, "algorithmic_bytes_per_parcel": sb[dom], "parcels_per_launch": n_local,
                    "share_of_step": stage_t[dom] / max(1e-9, sum(stage_t.values()))}
        line = {
            "metric": "particle-steps/s (move+sort+NTC collide+sample)", "value": value, "unit": "particle-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "strong" if args.workload == "wedge" else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, n_cells_gpu, n_local),
            "roofline": roofline, "stages": stages,
            "kernel_ms_per_step": per_step, "wall_ms_per_step": 1e3 * wall / args.steps, "setup_s": setup_s,
            "clocks": sampler.summary(), "gpu_launches": int(sum(v[1] * KERNELS_PER_TIMER.get(k, 1) for k, v in kt.items())),
            "hbm_frac_of_step": sum(n_local * sb[k] for k in sb) / (ms_total / args.steps * 1e-3) / 1e9 / peak,
            "invariants": invariants,
        }
        if e2e is not None:
            line["e2e"] = e2e
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_leg(args)
        emit(line)
    eng.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def _download_into(eng, host):
    """dsmcb200_download_parcels into the caller's (pinned) buffers."""
    import ctypes as C

    st = host.as_struct()
    nn = C.c_int64()
    eng._ck(eng.lib.dsmcb200_download_parcels(eng.h, len(host.position), C.byref(nn), C.byref(st)))
    host.n = nn.value
    return host


if __name__ == "__main__":
    main()
