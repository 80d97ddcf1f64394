"""Full-size runs of BASELINE.json configs 3 and 4 (parity-test cases, not bench lines): Newton Jacobian + residual
assembly on the generic path, device-resident, with size-independent checks.  Usage: python bench_configs.py [3|4] [n]"""
import json
import sys
import time

import numpy as np

import __graft_entry__ as g

ID, GRAD = 0, 1


def timed(eng, fn, reps=3):
    fn()
    ms = []
    for _ in range(reps):
        eng.event_record(0); fn(); eng.event_record(1)
        ms.append(eng.event_elapsed_ms(0, 1))
    return float(np.median(ms)), eng.last_timings()


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
    return {"config": 3, "workload": f"Example250-like 2D P2-P1 NSE Newton Jacobian+residual, n={n}", "cells": int(grid.ncells),
            "dofs": int(nrows), "nnz": int(nnz), "ms": ms, "phase_ms": {"local": ph[0], "gather": ph[1]},
            "cells_per_s": grid.ncells / ms * 1e3, "nnz_per_s": nnz / ms * 1e3, "host_mesh_s": t_host,
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
    return {"config": 4, "workload": f"Example330-like 3D P2 Neo-Hooke Newton Jacobian+residual, n={n}", "cells": int(grid.ncells),
            "dofs": int(nrows), "nnz": int(nnz), "ms": ms, "phase_ms": {"local": ph[0], "gather": ph[1]},
            "cells_per_s": grid.ncells / ms * 1e3, "nnz_per_s": nnz / ms * 1e3, "host_mesh_s": t_host,
            "checks": {"deterministic": bool(np.array_equal(nz1, nz2)), "max_row_of_rigid_translation": float(rigid),
                       "symmetry_defect": sym, "ok": ok}}


if __name__ == "__main__":
    pkg = g.load_package()
    which = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    eng = pkg.lib.Engine(0)
    if which == 3:
        out = config3(pkg, eng, int(sys.argv[2]) if len(sys.argv) > 2 else 1414)
    else:
        out = config4(pkg, eng, int(sys.argv[2]) if len(sys.argv) > 2 else 70)
    print(json.dumps(out))
    eng.close()
