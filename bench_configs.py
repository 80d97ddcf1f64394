"""Full-size runs of BASELINE.json configs 3, 4 and 5 (parity-test cases, not bench lines).
  3 | 4: Newton Jacobian + residual assembly on the generic path, device-resident: ms per assembly with the per-kernel phase times,
         both rooflines (HBM bytes and FP64 flops as SURVEY.md 8(d) counts them, peaks from MEASURED_PEAKS.json and
         profiles/fp64_peaks.json), ENTRYWISE parity of the full-size matrix columns and rhs rows near the bottom of the mesh
         against the CPU oracle run on that part of the mesh, and size-independent checks.
  5    : Poisson stiffness assembly + Jacobi-CG to convergence on the owned-row distributed system (torchrun for N > 1).
Usage: python bench_configs.py 3|4 [n]  |  [torchrun ...] bench_configs.py 5 [n] [order] [maxit]"""
import json
import os
import sys
import time

import numpy as np

import __graft_entry__ as g

ID, GRAD = 0, 1
HERE = os.path.dirname(os.path.abspath(__file__))
# tuning sweeps only (scripts/gpu_nl_opts.sh): skip the oracle comparison, the size-independent checks still run
NO_PARITY = {"matrix_max_rel": 0.0, "rhs_max_rel": 0.0, "skipped": True} if os.environ.get("EXTFEM_NO_PARITY") else None


def timed(eng, fn, reps=3):
    fn()
    ms = []
    for _ in range(reps):
        eng.event_record(0); fn(); eng.event_record(1)
        ms.append(eng.event_elapsed_ms(0, 1))
    return float(np.median(ms)), eng.last_timings()


def peaks():
    hbm, src = 6553.3, "fallback (round-1 measurement)"
    try:
        hbm = float(json.load(open(os.path.join(HERE, "MEASURED_PEAKS.json")))["hbm_gbs"]); src = "MEASURED_PEAKS.json"
    except Exception:
        pass
    fp = json.load(open(os.path.join(HERE, "profiles", "fp64_peaks.json")))
    return hbm, src, float(fp["fp64_dfma_tflops"]), float(fp["fp64_dmma_m8n8k4_tflops"])


def rooflines(ms, ncells, nnz, nrows, dim, nq, nin, nout, dofs_per_cell, nzB):
    """SURVEY.md 8(d) for a NonlinearOperator: bytes = coordinates + dof ids + solution coefficients read, matrix values and rhs
    written once; flops = the reference's two contractions per quadrature point with the sparse operator evaluations
    (nonlinear_operator.jl:372-401: J B over the nonzero operator entries of every ansatz dof, then the dot with those of every
    test dof) -- the kernel function itself and the geometry are not counted."""
    hbm, src, dfma, dmma = peaks()
    bytes_ = (8 * dim * (dim + 1) + 12 * dofs_per_cell) * ncells + 8 * nnz + 8 * nrows
    flops = ncells * nq * (2 * nout * nzB + 2 * dofs_per_cell * nzB + 2 * nin * nzB)
    a_b, a_f = bytes_ / ms / 1e6, flops / ms / 1e9
    return {"bound": "hbm" if a_b / hbm >= a_f / dmma else "fp64",
            "hbm": {"achieved": a_b, "peak": hbm, "unit": "GB/s", "frac": a_b / hbm, "algorithmic_bytes": int(bytes_), "peak_source": src},
            "fp64": {"achieved": a_f, "peak": dmma, "unit": "TFLOP/s", "frac": a_f / dmma, "algorithmic_flops": int(flops),
                     "peak_source": "profiles/fp64_peaks.json (mma.sync.m8n8k4.f64; DFMA peak %.2f)" % dfma},
            "note": "whole assembly (all kernels of one Newton Jacobian + residual) against either ceiling"}


def block_keys(FES_list, n):
    """Lattice id of every dof of a block system [per block: per component: scalar dofs] and the doubled index of its last
    coordinate.  Ascending key == ascending dof index inside every (block, component, dof class)."""
    keys, last = [], []
    base = 0
    for F in FES_list:
        pts = F.dof_coordinates()
        q = np.rint(pts * 2 * n).astype(np.int64)
        w = 2 * n + 1
        k = q[:, 0]
        for d in range(1, q.shape[1]):
            k = k * w + q[:, d]          # any injective code of the lattice point
        span = w ** q.shape[1]
        for c in range(F.fetype.ncomponents):
            keys.append(base + c * span + k); last.append(q[:, -1])
        base += F.fetype.ncomponents * span
    return np.concatenate(keys), np.concatenate(last)


def parity_submesh(pkg, eng, pat, nzval, bvec, n, fetypes, FES_full, args, kernel, params, solfun, coupling=None, layers=2):
    """Entrywise parity at full size: the bottom `layers` layers of the benchmark mesh are re-meshed alone, the CPU oracle
    assembles the operator there (solution = the same function), and every matrix column / rhs row of a dof strictly below the
    top of that part (all its cells are inside) is compared with the engine's full-size result."""
    from oracle import oracle as ora
    ora.build()
    dim = FES_full[0].xgrid.coords.shape[1]
    X = np.linspace(0.0, 1.0, n + 1)
    Z = X[:layers + 1]
    sub = pkg.simplexgrid(X, Z) if dim == 2 else pkg.simplexgrid(X, X, Z)
    FS = [pkg.FESpace(ft, sub) for ft in fetypes]
    kf, _ = block_keys(FES_full, n)
    ks, last = block_keys(FS, n)
    order = np.argsort(kf)
    pos = np.searchsorted(kf[order], ks)
    assert (pos < kf.size).all() and np.array_equal(kf[order][pos], ks), "a dof of the sub-mesh is missing in the full mesh"
    s2f = order[pos]                                                       # sub-mesh dof -> full-mesh dof
    offs = np.concatenate([[0], np.cumsum([F.ndofs for F in FS])])
    om = ora.Mesh(sub.coords, sub.cellnodes, sub.cellregions, sub.cellvolumes)
    oargs = [ora.OraArg(FS[b].celldofs, FS[b].fetype.ncomponents, FS[b].fetype.order, op, int(offs[b])) for b, op in args]
    N = int(offs[-1])
    if coupling is not None:                       # block coupling of the engine's pattern -> per-argument coupling of the oracle's
        cm = np.asarray(coupling).reshape(len(FS), len(FS))
        coupling = [[int(cm[ba][bt]) for bt, _ in args] for ba, _ in args]
    cp, rv = ora.structural_pattern(oargs, oargs, (N, N), coupling=coupling)
    sol = solfun(FS)
    bref = np.zeros(N)
    nzref, bref = ora.assemble_nonlinear(om, oargs, oargs, sol, bref, kernel, params=params, csc=(cp, rv))
    colptr = eng.pattern_colptr(pat)
    cols = np.nonzero(last <= 2 * (layers - 1))[0]
    worst, scale, count = 0.0, 0.0, 0
    for c in cols:
        a = nzref[cp[c] - 1:cp[c + 1] - 1]
        fc = s2f[c]
        b = nzval[colptr[fc] - 1:colptr[fc + 1] - 1]
        assert a.size == b.size, "pattern of a full-size column differs from the oracle's"
        # rows ascend in both; the sub-mesh numbering is an order-preserving restriction of the full one
        if a.size:
            worst = max(worst, float(np.abs(a - b).max())); scale = max(scale, float(np.abs(a).max())); count += a.size
    rows_f = s2f[cols]
    errb = float(np.abs(bvec[rows_f] - bref[cols]).max())
    return {"matrix_max_rel": worst / scale, "rhs_max_rel": errb / float(np.abs(bref).max()), "entries_checked": int(count),
            "columns_checked": int(cols.size), "against": f"CPU oracle (oracle/assembly_ref.c) on the bottom {layers} layers of the mesh"}


def config3(pkg, eng, n):
    X = np.linspace(0, 1, n + 1)
    t0 = time.time()
    grid = pkg.simplexgrid(X, X)
    FU, FP = pkg.FESpace(pkg.H1P2(2, 2), grid), pkg.FESpace(pkg.H1P1(1), grid)
    t_host = time.time() - t0
    mesh = eng.mesh_set(grid.coords, grid.cellnodes, grid.cellregions, grid.cellvolumes)
    su = eng.space_set(mesh, 2, 2, FU.celldofs, FU.ndofs)
    sp = eng.space_set(mesh, 1, 1, FP.celldofs, FP.ndofs)
    pat = eng.pattern_build([su, sp], None, np.array([1, 1, 1, 0], dtype=np.uint8))   # no p-p block
    nrows, ncols, nnz = eng.pattern_dims(pat)
    xu = FU.dof_coordinates(); xp = FP.dof_coordinates()
    sol = np.concatenate([xu[:, 0] ** 2, xu[:, 0] + xu[:, 1], xp[:, 1] ** 2])    # test_nonlinear_operator.jl:38-39
    args = [(0, ID), (0, GRAD), (1, ID)]
    desc = eng.make_opdesc(args, args=args, kernel_id=pkg.lib.kernel_id("nse2d"), params=[0.05])
    import torch
    sol_d = torch.from_numpy(sol).cuda()
    ms, ph = timed(eng, lambda: eng.assemble_nonlinear(pat, desc, sol_d))
    # size-independent checks: determinism; the divergence rows annihilate constant velocities; residual identity
    nz1, b1 = eng.values_get(pat)
    eng.assemble_nonlinear(pat, desc, sol_d)
    nz2, b2 = eng.values_get(pat)
    cu = np.concatenate([np.ones(FU.coffset), 2 * np.ones(FU.coffset), np.zeros(FP.ndofs)])
    y = eng.spmv(pat, cu)
    div_rows = np.abs(y[FU.ndofs:]).max()
    res = eng.residual(pat, sol)
    ok = bool(np.array_equal(nz1, nz2) and np.array_equal(b1, b2) and div_rows < 1e-12 * np.abs(nz1).max() and np.isfinite(res).all())
    solfun = lambda FS: np.concatenate([FS[0].dof_coordinates()[:, 0] ** 2, FS[0].dof_coordinates().sum(axis=1),   # noqa: E731
                                        FS[1].dof_coordinates()[:, 1] ** 2])
    parity = NO_PARITY or parity_submesh(pkg, eng, pat, nz1, b1, n, [pkg.H1P2(2, 2), pkg.H1P1(1)], [FU, FP], args, "nse2d", [0.05], solfun,
                                         coupling=np.array([1, 1, 1, 0], dtype=np.uint8))
    from oracle import fetables
    roof = rooflines(ms, grid.ncells, nnz, nrows, 2, fetables.quadrature_rule(2, 4)[1].size, 7, 7, 15, 12 * 3 + 3)   # 12 velocity dofs: id + 2 gradient entries; 3 pressure dofs: id
    ok = ok and parity["matrix_max_rel"] <= 1e-12 and parity["rhs_max_rel"] <= 1e-12
    return {"config": 3, "workload": f"Example250-like 2D P2-P1 NSE Newton Jacobian+residual, n={n}", "cells": int(grid.ncells),
            "dofs": int(nrows), "nnz": int(nnz), "ms": ms, "phase_ms": {"local": ph[0], "gather": ph[1]},
            "cells_per_s": grid.ncells / ms * 1e3, "nnz_per_s": nnz / ms * 1e3, "host_mesh_s": t_host, "roofline": roof, "parity": parity,
            "checks": {"deterministic": bool(np.array_equal(nz1, nz2)), "max_divergence_row_of_constant_velocity": float(div_rows),
                       "ok": ok}}


def config4(pkg, eng, n):
    X = np.linspace(0, 1, n + 1)
    t0 = time.time()
    grid = pkg.simplexgrid(X, X, X)
    FU = pkg.FESpace(pkg.H1P2(3, 3), grid)
    t_host = time.time() - t0
    mesh = eng.mesh_set(grid.coords, grid.cellnodes, grid.cellregions, grid.cellvolumes)
    su = eng.space_set(mesh, 2, 3, FU.celldofs, FU.ndofs)
    pat = eng.pattern_build([su])
    nrows, ncols, nnz = eng.pattern_dims(pat)
    x = FU.dof_coordinates()
    sol = 0.1 * np.concatenate([x[:, 0] ** 2, x[:, 0] + x[:, 1], x[:, 1] * x[:, 2]])
    E, nu = 10.0, 0.3
    mu, la = E / (2 * (1 + nu)), E * nu / ((1 + nu) * (1 - 2 * nu))
    args = [(0, GRAD)]
    desc = eng.make_opdesc(args, args=args, kernel_id=pkg.lib.kernel_id("neohooke3d"), params=[mu, la])
    import torch
    sol_d = torch.from_numpy(sol).cuda()
    ms, ph = timed(eng, lambda: eng.assemble_nonlinear(pat, desc, sol_d))
    # checks: determinism, the tangent annihilates rigid translations, symmetry x'(Ay) == y'(Ax) (hyperelastic tangent)
    nz1, _ = eng.values_get(pat, want_b=False)
    eng.assemble_nonlinear(pat, desc, sol_d)
    nz2, _ = eng.values_get(pat, want_b=False)
    tr = np.concatenate([np.ones(FU.coffset), np.zeros(2 * FU.coffset)])
    rigid = np.abs(eng.spmv(pat, tr)).max()
    rng = np.random.default_rng(0)
    xa, ya = rng.standard_normal(nrows), rng.standard_normal(nrows)
    sym = abs(float(xa @ eng.spmv(pat, ya)) - float(ya @ eng.spmv(pat, xa)))
    scale = np.abs(nz1).max()
    ok = bool(np.array_equal(nz1, nz2) and rigid < 1e-10 * scale and sym < 1e-9 * scale * nrows ** 0.5)
    _, b1 = eng.values_get(pat)

    def solfun(FS):
        xs = FS[0].dof_coordinates()
        return 0.1 * np.concatenate([xs[:, 0] ** 2, xs[:, 0] + xs[:, 1], xs[:, 1] * xs[:, 2]])
    parity = NO_PARITY or parity_submesh(pkg, eng, pat, nz1, b1, n, [pkg.H1P2(3, 3)], [FU], args, "neohooke3d", [mu, la], solfun)
    from oracle import fetables
    roof = rooflines(ms, grid.ncells, nnz, nrows, 3, fetables.quadrature_rule(3, 2)[1].size, 9, 9, 30, 30 * 3)      # 30 dofs with 3 gradient entries each
    ok = ok and parity["matrix_max_rel"] <= 1e-12 and parity["rhs_max_rel"] <= 1e-12
    return {"config": 4, "workload": f"Example330-like 3D P2 Neo-Hooke Newton Jacobian+residual, n={n}", "cells": int(grid.ncells),
            "dofs": int(nrows), "nnz": int(nnz), "ms": ms, "phase_ms": {"local": ph[0], "gather": ph[1]},
            "cells_per_s": grid.ncells / ms * 1e3, "nnz_per_s": nnz / ms * 1e3, "host_mesh_s": t_host, "roofline": roof, "parity": parity,
            "checks": {"deterministic": bool(np.array_equal(nz1, nz2)), "max_row_of_rigid_translation": float(rigid),
                       "symmetry_defect": sym, "ok": ok}}


def config5(pkg, n, order, maxit):
    """Config 5: Poisson stiffness + rhs assembly, reduction to the owned-row form and Jacobi-CG TO CONVERGENCE (relative
    residual 1e-10) on the distributed system.  One z-slab of n cube layers of the stacked domain [0,1]^2 x [0,N] per rank
    (weak scaling, host/dist.py: SlabShard); run under torchrun for N > 1."""
    import torch
    import torch.distributed as dist
    rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    eng = pkg.lib.Engine(lr)
    sh = pkg.SlabShard(pkg, n, n * world, pkg.layer_ranges(n * world, world), rank, order)
    grid, FES = sh.grid, sh.FES
    mesh = eng.mesh_set(grid.coords, grid.cellnodes, grid.cellregions, sh.cellvolumes)
    sp = eng.space_set(mesh, order, 1, FES.celldofs, FES.ndofs)
    pat = eng.pattern_build([sp])
    nrows, _, nnz = eng.pattern_dims(pat)
    if world > 1:
        uid = [pkg.lib.Engine.dist_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        eng.dist_init(rank, world, uid[0])
        eng.dist_set_owned(pat, sh.owned_plan())
    lap = eng.make_opdesc([(0, 1)], [(0, 1)])
    rhs = eng.make_opdesc([(0, 0)], kernel_id=pkg.lib.kernel_id("sincos301"), params=[1.0])

    def assemble():
        eng.assemble_bilinear(pat, lap)
        eng.assemble_linear(pat, rhs)
        if world > 1:
            eng.dist_reduce_system(pat, True, True)
    assemble()
    asm_ms = []
    for _ in range(3):
        eng.event_record(0); assemble(); eng.event_record(1)
        asm_ms.append(eng.event_elapsed_ms(0, 1))
    # homogeneous Dirichlet data on the outer boundary of the stacked domain, on every rank that holds the dof
    xyz = FES.dof_coordinates()
    onb = (xyz[:, 0] == 0) | (xyz[:, 0] == 1) | (xyz[:, 1] == 0) | (xyz[:, 1] == 1) | (xyz[:, 2] == 0) | (np.abs(xyz[:, 2] - float(world)) < 1e-12)
    eng.apply_penalties(pat, np.nonzero(onb)[0] + 1, None, 1e30)

    def timed_cg(k, rtol):
        eng.synchronize(); torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        if world > 1:
            x, it, rr = eng.dist_cg_owned(pat, rtol=rtol, maxit=k)
        else:
            x, it, rr = eng.cg(pat, rtol=rtol, maxit=k)
        eng.synchronize()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), it, rr, x
    timed_cg(3, 1e-30)                            # warm-up
    # a short run of fixed length measures the fixed cost (x to and from the host); the converged run minus that is the solve
    d1, it1, _, _ = timed_cg(5, 1e-30)
    d2, it, rr, x = timed_cg(maxit, 1e-10)
    per_it = (d2 - d1) / max(1, it - it1)
    own = sh.owned == 1
    xmax = torch.tensor([float(np.abs(x[own]).max())], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(xmax, op=dist.ReduceOp.MAX)
    out = None
    if rank == 0:
        out = {"config": 5, "workload": f"3D H1P{order} Poisson + Jacobi-CG, slab n={n} per GPU ({sh.ncells_owned} tets, {nrows} local dofs, {nnz} local nnz)",
               "n_gpus": world, "form": "owned rows (ghost layer, reduce-to-owner over NCCL, halo exchange per SpMV)" if world > 1 else "single GPU",
               "assembly_and_reduction_ms": float(np.median(asm_ms)),
               "cells_per_s": sh.ncells_owned * world / float(np.median(asm_ms)) * 1e3,
               "cg_iterations": int(it), "relres": float(rr), "converged": bool(rr <= 1e-10), "solve_s": d2, "ms_per_iteration": per_it * 1e3,
               "spmv_algorithmic_GBs_per_gpu": (12.0 * nnz + 16.0 * nrows) / per_it / 1e9, "max_abs_solution": float(xmax[0]),
               "note": "CG stops at |b - A x| <= 1e-10 |b - A x0|; per iteration one SpMV + halo exchange (ncclSend/Recv) + 3 dot products "
                       "(ncclAllReduce) + 2 vector updates; ms_per_iteration from the difference to a 5-iteration run"}
    if world > 1:
        dist.barrier(); dist.destroy_process_group()
    eng.close()
    return out


if __name__ == "__main__":
    pkg = g.load_package()
    which = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    if which == 5:
        out = config5(pkg, int(sys.argv[2]) if len(sys.argv) > 2 else 119, int(sys.argv[3]) if len(sys.argv) > 3 else 2,
                      int(sys.argv[4]) if len(sys.argv) > 4 else 5000)
        if out is not None:
            print(json.dumps(out))
        sys.exit(0)
    eng = pkg.lib.Engine(0)
    if which == 3:
        out = config3(pkg, eng, int(sys.argv[2]) if len(sys.argv) > 2 else 1414)
    else:
        out = config4(pkg, eng, int(sys.argv[2]) if len(sys.argv) > 2 else 70)
    print(json.dumps(out))
    eng.close()
