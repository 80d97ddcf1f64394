#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 assembly engine (contract: see DESIGN.md section 6).

Metric (BASELINE.json): assembled cells/s (FP64) of 3D H1P2 Poisson stiffness + RHS assembly on
a synthetic structured simplexgrid (config 2: n=119 -> 10,110,954 tets, 389.5M nnz), with the
fraction of the HBM roofline of the dominant kernel.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--n 119] [--impl ours|reference]

A "step" = one full re-assembly of the stiffness matrix (BilinearOperator([grad(u)])) and the
right-hand side (LinearOperator(f!, [id(u)]), Example301's f) into the device-resident system.
N > 1 (torchrun): every rank assembles its own slab of the global grid (weak scaling in cells);
see DESIGN.md section 5.
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

METRIC = "assembled_cells_per_s_fp64_poisson3d_p2"
UNIT = "cells/s"


def env_int(name, default):
    return int(os.environ.get(name, default))


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons with nvidia-smi during the timed region."""

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = False
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                self.samples.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        sm = [float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples if len(s) >= 6 for i in range(4) if s[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def build_problem(pkg, n, rank=0, world=1):
    """Structured unit-cube grid (6 tets per cube); rank r of a multi-GPU run gets the z-slab
    [r, r+1] of the stacked domain [0,1]^2 x [0,world] (weak scaling, no data-path collective)."""
    X = np.linspace(0.0, 1.0, n + 1)
    Z = np.linspace(float(rank), float(rank + 1), n + 1)
    grid = pkg.simplexgrid(X, X, Z)
    FES = pkg.FESpace(pkg.H1P2(1, 3), grid)
    return grid, FES


def slab_interfaces(pkg, FES, rank, world):
    """Interface plan of the stacked slabs: the top plane of rank r is the bottom plane of rank r+1 (dofs matched by
    their coordinates, ordered lexicographically in (y, x)); the lower rank owns the shared plane."""
    xyz = FES.dof_coordinates()
    z = xyz[:, 2]
    neigh, ptr, rows = [], [0], []
    owned = np.ones(FES.ndofs, dtype=np.uint8)
    for r, zz in ((rank - 1, z.min()), (rank + 1, z.max())):
        if r < 0 or r >= world:
            continue
        idx = np.nonzero(z == zz)[0]
        idx = idx[np.lexsort((xyz[idx, 0], xyz[idx, 1]))]
        neigh.append(r); rows.append(idx.astype(np.int64) + 1); ptr.append(ptr[-1] + idx.size)
        if r < rank:
            owned[idx] = 0
    return pkg.InterfacePlan(rank, world, np.asarray(neigh, dtype=np.int32), np.asarray(ptr, dtype=np.int64),
                             np.concatenate(rows), owned)


def algorithmic_bytes(grid, FES, nnz):
    """coordinates + dof ids read + matrix values written once (+ rhs written)."""
    dim = grid.dim
    per_cell = 8 * dim * (dim + 1) + 4 * FES.ndofs4cell
    return per_cell * grid.ncells + 8 * nnz


def run_ours(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    pkg = g.load_package()
    rank, world = env_int("RANK", 0), env_int("WORLD_SIZE", 1)
    local_rank = env_int("LOCAL_RANK", 0)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    eng = pkg.lib.Engine(local_rank)
    for kv in filter(None, os.environ.get("EXTFEM_OPTIONS", "").split(",")):   # tuning knobs, e.g. template_pool_bytes=57344
        k, v = kv.split("=")
        eng.set_option(k.strip(), int(v))
    t0 = time.time()
    grid, FES = build_problem(pkg, args.n, rank, world)
    t_mesh = time.time() - t0
    t0 = time.time()
    mesh = eng.mesh_set(grid.coords, grid.cellnodes, grid.cellregions, grid.cellvolumes)
    sp = eng.space_set(mesh, FES.fetype.fe_id, 1, FES.celldofs, FES.ndofs)
    pat = eng.pattern_build([sp])
    eng.synchronize()
    t_setup = time.time() - t0
    nrows, ncols, nnz = eng.pattern_dims(pat)
    if world > 1:
        # the slabs of neighbouring ranks share the dofs of one z-plane: interface-row contributions of the rhs are
        # exchanged over NCCL inside every step (grouped send/recv, dist.cuh)
        uid = [pkg.lib.Engine.dist_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        eng.dist_init(rank, world, uid[0])
        eng.dist_set_interfaces(pat, slab_interfaces(pkg, FES, rank, world))
    lap = eng.make_opdesc([(0, 1)], [(0, 1)], kernel_id=pkg.lib.kernel_id("standard"), factor=1.0)
    rhs = eng.make_opdesc([(0, 0)], kernel_id=pkg.lib.kernel_id("sincos301"), params=[1.0])

    def step_resident():
        eng.assemble_bilinear(pat, lap)
        eng.assemble_linear(pat, rhs)
        if world > 1:
            eng.dist_sum_rhs(pat)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        eng.synchronize()

    launches0 = eng.launch_count()
    for _ in range(args.warmup):
        step_resident()
    launches_per_step = (eng.launch_count() - launches0) // max(1, args.warmup)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    barrier()
    eng.event_record(0)
    kern_ms = np.zeros(3)
    rhs_ms = np.zeros(3)
    for i in range(args.steps):
        # device-resident results: the calls return without waiting for the GPU, the host prepares the next operator
        # meanwhile; the per-phase CUDA-event times are read (which waits) in the last step of the timed region only
        eng.assemble_bilinear(pat, lap)
        if i == args.steps - 1:
            kern_ms = np.array(eng.last_timings())
        eng.assemble_linear(pat, rhs)
        if i == args.steps - 1:
            rhs_ms = np.array(eng.last_timings())
        if world > 1:
            eng.dist_sum_rhs(pat)
    eng.event_record(1)
    ms_total = eng.event_elapsed_ms(0, 1)
    barrier()
    clocks = sampler.finish() if sampler else None

    # ---- e2e through the C-ABI with HOST buffers: H2D of the coordinates (pinned), D2H of nzval and b
    coords_h = torch.from_numpy(np.ascontiguousarray(grid.coords)).pin_memory()
    nz_h = torch.empty(nnz, dtype=torch.float64).pin_memory()
    b_h = torch.empty(nrows, dtype=torch.float64).pin_memory()
    vol_h = torch.from_numpy(np.ascontiguousarray(grid.cellvolumes)).pin_memory()
    e2e_steps = max(1, min(args.steps, 3))

    def step_e2e():
        eng.mesh_update_coords(mesh, coords_h, vol_h)
        eng.assemble_bilinear(pat, lap, nzval_out=nz_h)
        if world > 1:
            eng.assemble_linear(pat, rhs)
            eng.dist_sum_rhs(pat)
            eng.values_get(pat, want_nzval=False, b_out=b_h)
        else:
            eng.assemble_linear(pat, rhs, b_out=b_h)

    step_e2e()
    barrier()
    eng.event_record(2)
    for _ in range(e2e_steps):
        step_e2e()
    eng.event_record(3)
    ms_e2e = eng.event_elapsed_ms(2, 3) / e2e_steps
    barrier()
    checksum = float(nz_h.sum())   # stiffness matrix annihilates constants: sum of all entries ~ 0

    ms_step = ms_total / args.steps
    t = torch.tensor([ms_step, ms_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step, ms_e2e = float(t[0]), float(t[1])
    cells_total = grid.ncells * world
    value = cells_total / (ms_step * 1e-3)
    e2e_value = cells_total / (ms_e2e * 1e-3)

    out = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = peaks.get("hbm_gbs", 6650.0)
        alg = algorithmic_bytes(grid, FES, nnz)
        traffic = None      # ncu dram__bytes_read.sum + dram__bytes_write.sum of the stiffness kernels (committed capture)
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r01_f_traffic.json")))
            if tj.get("workload") == f"n={args.n}":
                traffic = tj["stiffness_dram_bytes_per_assembly"]
        except Exception:
            pass
        dom_ms = kern_ms[0] + kern_ms[1]          # local + gather kernels of the stiffness assembly
        achieved = alg / (dom_ms * 1e-3) / 1e9
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"Example301-like 3D H1P2 Poisson stiffness+RHS on structured simplexgrid n={args.n} "
                                   f"({grid.ncells} tets, {nrows} dofs, {nnz} nnz per GPU)",
                       "l2_policy": "inputs+outputs (>= 4 GB per step) are larger than the 126 MB L2",
                       "parallelism": (f"cell slabs x{world}; interface rows of the rhs exchanged over NCCL send/recv every step"
                                       if world > 1 else "1 GPU")},
            "nnz_per_s": nnz * world / (ms_step * 1e-3),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": "ncu capture profiles/r01_f_final_ncu_summary.txt" if traffic else None,
                         "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback",
                         "kernel": "stiffness assembly kernels (local + gather)", "kernel_ms": dom_ms,
                         "algorithmic_bytes": alg, "frac_of_nominal_8TBs": achieved / 8000.0},
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": int(coords_h.numel() * 8 + vol_h.numel() * 8),
                    "d2h_bytes_per_step": int(nz_h.numel() * 8 + b_h.numel() * 8)},
            "gpu_launches": int(launches_per_step * args.steps),
            "clocks": clocks,
            "setup_s": {"mesh_host": t_mesh, "upload_adjacency_pattern": t_setup},
            "checksum_sum_nzval": checksum,
            "phase_ms": {"stiffness_geo": kern_ms[0], "stiffness_gather": kern_ms[1], "rhs_cell": rhs_ms[0], "rhs_gather": rhs_ms[1]},
            "plan": eng.plan_stats(pat, 0),
        }
        if not args.no_cpu_baseline and world == 1:
            out["cpu_baseline"] = cpu_baseline(pkg, args)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    eng.close()
    return out


def cpu_baseline(pkg, args, n_sample=None, steps=1, threads=None):
    """The CPU restatement (oracle, a port of the reference's loops; Julia is not installed) timed on a bounded sample of
    the same workload with all host threads: the sample grid is cut into one contiguous cell range per thread and every
    thread assembles its range into its own matrix / vector (insertion into an existing CSC pattern) -- the scheme of the
    reference's `parallel = true` path (thread-private parts per partition, bilinear_operator.jl:969-979)."""
    import threading
    from oracle import oracle as ora
    ora.build()
    n = n_sample or args.cpu_n
    T = threads or max(1, len(os.sched_getaffinity(0)))
    X = np.linspace(0, 1, n + 1)
    grid = pkg.simplexgrid(X, X, X)
    FES = pkg.FESpace(pkg.H1P2(1, 3), grid)
    parts = []
    for lo, hi in pkg.cell_ranges(grid.ncells, T):
        sh = pkg.Shard(grid.coords, grid.cellnodes, grid.cellregions, FES.celldofs, lo, hi, order=2)
        om = ora.Mesh(sh.coords, sh.cellnodes, sh.cellregions, np.ascontiguousarray(grid.cellvolumes[lo:hi]))
        gr = ora.OraArg(sh.celldofs, 1, 2, ora.OP_GRAD)
        idu = ora.OraArg(sh.celldofs, 1, 2, ora.OP_ID)
        colptr, rowval = ora.structural_pattern([gr], [gr], (sh.ndofs, sh.ndofs))
        parts.append((om, gr, idu, colptr, rowval, sh.ndofs))

    def work(p):
        om, gr, idu, colptr, rowval, nd = p
        ora.assemble_bilinear(om, [gr], [gr], "standard", csc=(colptr, rowval))
        b = np.zeros(nd)
        ora.assemble_linear(om, [idu], b, "sincos301", params=[1.0])

    times = []
    for _ in range(steps):
        th = [threading.Thread(target=work, args=(p,)) for p in parts]
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        times.append(time.perf_counter() - t0)
    dt = float(np.mean(times))
    return {"value": grid.ncells / dt, "unit": UNIT, "cores": T, "kind": "port",
            "sample": f"same operators on simplexgrid n={n} ({grid.ncells} tets) cut into {T} cell ranges, one thread each, C "
                      f"restatement of the reference loops (gcc -O2), {dt:.2f} s per step; the Julia reference is not installable offline",
            "seconds_per_step": dt}


def run_reference(args):
    """--impl reference: the reference's CPU path.  Julia and its three un-vendored dependencies are
    not available offline, so this arm times the oracle port (kind 'port') on the host cores."""
    rank = env_int("RANK", 0)
    if rank != 0:
        return None
    import __graft_entry__ as g
    pkg = g.load_package()
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_baseline(pkg, args, n_sample=8)
    cb = cpu_baseline(pkg, args, steps=max(1, min(args.steps, 3)))
    return {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": env_int("WORLD_SIZE", 1),
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["seconds_per_step"] * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"Example301-like 3D H1P2 Poisson stiffness+RHS, bounded sample n={args.cpu_n}"},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def main():
    # rank 0 must print exactly ONE JSON line on stdout: route everything libraries write to file descriptor 1 (NCCL prints
    # its version banner there) to stderr for the duration of the run and restore stdout for the result line
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--n", type=int, default=119, help="cubes per axis (119 -> 10.1M tets, config 2)")
    ap.add_argument("--cpu-n", type=int, default=40, help="grid size of the bounded CPU sample")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    out = run_reference(args) if args.impl == "reference" else run_ours(args)
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    os.close(saved_stdout)
    if out is not None:
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
