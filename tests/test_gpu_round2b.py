"""GPU parity tests of the second half of round 2: the cell-local form of the fast right-hand side (`rhs_local`) against
the point-value form and the oracle, and the flat write-out of the walk kernel (`template_flat_writeout`) in overwrite and
accumulate mode.  CUDA path through the C-ABI against the CPU oracle on the same inputs."""
import numpy as np
import pytest

from util import System, check_values, check_values_entrywise

pytestmark = pytest.mark.gpu

ID, GRAD = 0, 1


def _grid(pkg, dim, n):
    X = np.linspace(0, 1, n + 1)
    if dim == 1:
        return pkg.simplexgrid(np.linspace(0, 1, 8 * n + 1) ** 1.5)
    if dim == 2:
        return pkg.simplexgrid(X, X ** 2)
    return pkg.simplexgrid(X, X ** 1.3, X)


@pytest.mark.parametrize("dim,order,kernel,params,quadorder", [
    (3, 2, "sincos301", [1.3], -1),      # config 2's right-hand side: 4-point rule, specialised kernel
    (3, 2, "exp2x", [], 1),              # 1-point rule
    (3, 1, "sincos301", [0.7], 2),       # P1 with the 4-point rule
    (3, 2, "xy", [], 3),                 # 8 points
    (2, 2, "xy", [], -1),                # 3-point rule
    (2, 2, "sincos301", [1.0], 4),       # 9 points
    (2, 1, "exp2x", [], 3),              # 4 points
    (1, 2, "step105", [], 5),            # 3-point Gauss rule in 1D
    (1, 2, "exp2x", [], 3),              # 2 points
    (1, 1, "constant_one", [], -1),
    (2, 2, "xy", [], 6),                 # 16 points: no compile-time kernel, the point-value form serves it
])
def test_rhs_local_form(pkg, ora, engine, dim, order, kernel, params, quadorder):
    """Fast RHS, cell-local form (linear_operator.jl:618-633: local vector per cell, then one add per dof) with and without
    templates, regions, accumulate; must agree with the oracle entrywise and with the point-value form to rounding."""
    g = _grid(pkg, dim, 6 if dim == 3 else 12)
    g.cellregions[::3] = 2
    S = System(pkg, ora, engine, g, [pkg.H1Pk(1, dim, order)])
    kw = dict(kernel_id=pkg.lib.kernel_id(kernel), params=params, factor=1.9)
    if quadorder >= 0:
        kw["quadorder"] = quadorder
    okw = dict(params=params, factor=1.9)
    if quadorder >= 0:
        okw["quadorder"] = quadorder
    got = {}
    try:
        for mincols in (2, 1 << 30):         # template plan | every column on the adjacency-list kernel
            engine.set_option("template_min_cols", mincols)
            S = System(pkg, ora, engine, g, [pkg.H1Pk(1, dim, order)])
            for regions in ((), (2,)):
                ref = np.zeros(S.N); sc = np.zeros(S.N)
                ora.assemble_linear(S.omesh, S.oargs([(0, ID)]), ref, kernel, regions=list(regions), **okw)
                with ora.abs_accumulate():
                    ora.assemble_linear(S.omesh, S.oargs([(0, ID)]), sc, kernel, regions=list(regions), **okw)
                # (cell-local form, column groups a warp of its gather serves at once, occupancy variant of the one-group gather)
                for local, groups, ctas in ((1, 1, 8), (1, 2, 8), (1, 4, 8), (1, 1, 6), (1, 1, 5), (0, 1, 8)):
                    engine.set_option("rhs_local", local)
                    engine.set_option("rhs_groups", groups)
                    engine.set_option("rhs_gather_ctas", ctas)
                    d = engine.make_opdesc([(0, ID)], regions=regions, **kw)
                    b = np.empty(S.N)
                    engine.assemble_linear(S.pat, d, b_out=b)
                    check_values(b, ref, what=f"rhs local={local} mincols={mincols} regions={regions}")
                    check_values_entrywise(b, ref, sc, what=f"rhs entrywise local={local}")
                    engine.assemble_linear(S.pat, d, accumulate=True, b_out=b)
                    check_values(b, 2 * ref, what=f"rhs accumulate local={local}")
                    got[(local, groups, ctas)] = b
                check_values(got[(1, 1, 8)], got[(0, 1, 8)], what="local vs point-value form")
                for key in ((1, 2, 8), (1, 4, 8), (1, 1, 6), (1, 1, 5)):
                    assert np.array_equal(got[(1, 1, 8)], got[key]), key
    finally:
        engine.set_option("rhs_local", 1)
        engine.set_option("rhs_groups", 1)
        engine.set_option("rhs_gather_ctas", 8)
        engine.set_option("template_min_cols", 24)


def test_rhs_local_tabulated(pkg, ora, engine):
    """Host-tabulated integrand through the cell-local form (values indexed by the ORIGINAL cell number in the transposed order)."""
    X = np.linspace(0, 1, 8)
    g = pkg.simplexgrid(X, X, X)
    engine.set_option("template_min_cols", 2)
    try:
        S = System(pkg, ora, engine, g, [pkg.H1P2(1, 3)])
        desc = engine.make_opdesc([(0, ID)], kernel_id=pkg.lib.kernel_id("tabulated"))
        xq = engine.quadrature_points_x(S.pat, desc, g.ncells, 3)
        vals = (np.exp(xq[..., 0]) * xq[..., 1] + np.sin(3 * xq[..., 2]))[..., None]
        desc = engine.make_opdesc([(0, ID)], kernel_id=pkg.lib.kernel_id("tabulated"), tabulated=vals)
        ref = np.zeros(S.N)
        ora.assemble_linear(S.omesh, S.oargs([(0, ID)]), ref, "tabulated", tabulated=vals)
        for local in (1, 0):
            engine.set_option("rhs_local", local)
            b = np.empty(S.N)
            engine.assemble_linear(S.pat, desc, b_out=b)
            check_values(b, ref, what=f"tabulated rhs local={local}")
    finally:
        engine.set_option("rhs_local", 1)
        engine.set_option("template_min_cols", 24)


@pytest.mark.parametrize("n", [10, 13])
def test_walk_flat_writeout(pkg, ora, engine, n):
    """Walk kernel with the flat write-out (skewed accumulator columns): overwrite (first-touch) and accumulate mode against the
    oracle, bit-identical to the position-per-lane write-out."""
    X = np.linspace(0, 1, n + 1)
    g = pkg.simplexgrid(X, X ** 1.2, X)
    S = System(pkg, ora, engine, g, [pkg.H1P2(1, 3)])
    desc = engine.make_opdesc([(0, GRAD)], [(0, GRAD)], factor=1.3)
    ref = ora.assemble_bilinear(S.omesh, S.oargs([(0, GRAD)]), S.oargs([(0, GRAD)]), factor=1.3, csc=(S.colptr, S.rowval))
    res = {}
    try:
        for flat in (0, 20, 1000):
            engine.set_option("template_flat_writeout", flat)
            nz = np.empty(S.rowval.size)
            engine.assemble_bilinear(S.pat, desc, nzval_out=nz)
            stats = engine.plan_stats(S.pat, 0)
            assert stats["templates"] > 0
            check_values(nz, ref, what=f"walk kernel flat={flat}")
            res[flat] = nz.copy()
            # accumulate on top of known values: the kernel first loads the column segments into its accumulators
            base = np.linspace(-1.0, 1.0, S.rowval.size)
            engine.values_set(S.pat, nzval=base, b=np.zeros(S.N))
            engine.assemble_bilinear(S.pat, desc, accumulate=True, nzval_out=nz)
            check_values(nz - base, ref, what=f"walk kernel accumulate flat={flat}")
            engine.values_zero(S.pat, True, True)
            engine.assemble_bilinear(S.pat, desc, accumulate=True, nzval_out=nz)
            check_values(nz, ref, what=f"walk kernel zero + accumulate flat={flat}")
        assert np.array_equal(res[0], res[20]) and np.array_equal(res[0], res[1000])
    finally:
        engine.set_option("template_flat_writeout", 0)


@pytest.mark.parametrize("scale,shift", [(1.0, 0.0), (130.0, -70.0), (3.0, 2.0e6)])
def test_rhs_fast_trig_ranges(pkg, ora, engine, scale, shift):
    """sin / cos of the cell kernel (tp_sin / tp_cos: Cody-Waite reduction + fdlibm kernels) against the oracle's libm on meshes
    whose coordinates give arguments of a few units, of several hundred (many quadrants, both signs) and beyond 2^20 (library
    fallback); and against the library functions on the device (option rhs_fast_trig = 0)."""
    X = np.linspace(0, 1, 8)
    g = pkg.simplexgrid(X * scale + shift, X ** 1.1 * scale + shift, X)
    S = System(pkg, ora, engine, g, [pkg.H1P2(1, 3)])
    ref = np.zeros(S.N); sc = np.zeros(S.N)
    ora.assemble_linear(S.omesh, S.oargs([(0, ID)]), ref, "sincos301", params=[0.8])
    with ora.abs_accumulate():
        ora.assemble_linear(S.omesh, S.oargs([(0, ID)]), sc, "sincos301", params=[0.8])
    d = engine.make_opdesc([(0, ID)], kernel_id=pkg.lib.kernel_id("sincos301"), params=[0.8])
    got = {}
    # the oracle evaluates x_q in another order: an ulp of the coordinates changes the argument of sin / cos by 3.9 ulp(|x|), so the
    # comparison with the oracle is scaled by the coordinate size; device library vs tp_sin / tp_cos see identical arguments
    loose = 1e-12 * max(1.0, 40.0 * float(np.abs(g.coords).max()))
    try:
        for ft in (1, 0):
            engine.set_option("rhs_fast_trig", ft)
            b = np.empty(S.N)
            engine.assemble_linear(S.pat, d, b_out=b)
            check_values_entrywise(b, ref, sc, rtol=loose, what=f"rhs fast_trig={ft} scale={scale} shift={shift}")
            got[ft] = b
        check_values_entrywise(got[1], got[0], sc, rtol=1e-12, what="tp_sin/tp_cos vs library sin/cos")
    finally:
        engine.set_option("rhs_fast_trig", 1)


def test_lower_triangle_export_pipelined(pkg, engine):
    """extfem_values_get_lower on a system large enough for the pipelined export (more than 2^22 lower entries: eight column chunks,
    packed on the main stream and copied on the exchange stream): bit-identical to scipy's tril of the full copy-back; twice, into
    pinned and pageable host arrays."""
    import scipy.sparse as sp
    import torch
    X = np.linspace(0, 1, 36)
    g = pkg.simplexgrid(X, X ** 1.1, X)
    F = pkg.FESpace(pkg.H1P2(1, 3), g)
    mesh = engine.mesh_set(g.coords, g.cellnodes, g.cellregions, g.cellvolumes)
    pat = engine.pattern_build([engine.space_set(mesh, F.fetype.fe_id, 1, F.celldofs, F.ndofs)])
    colptr, rowval = engine.pattern_get(pat)
    engine.assemble_bilinear(pat, engine.make_opdesc([(0, GRAD)], [(0, GRAD)], factor=0.7))
    engine.assemble_linear(pat, engine.make_opdesc([(0, ID)], kernel_id=pkg.lib.kernel_id("sincos301"), params=[1.0]))
    nz, b = engine.values_get(pat)
    L = sp.tril(sp.csc_matrix((nz, rowval - 1, colptr - 1), shape=(F.ndofs, F.ndofs)), format="csc")
    n, cp, rv = engine.pattern_get_lower(pat)
    assert n == L.nnz > (1 << 22) and np.array_equal(cp - 1, L.indptr) and np.array_equal(rv - 1, L.indices)
    lz, lb = engine.values_get_lower(pat)
    assert np.array_equal(lz, L.data) and np.array_equal(lb, b)
    lz_pin = torch.empty(n, dtype=torch.float64).pin_memory(); lb_pin = torch.empty(F.ndofs, dtype=torch.float64).pin_memory()
    engine.values_get_lower(pat, nzval_out=lz_pin, b_out=lb_pin)
    assert np.array_equal(lz_pin.numpy(), L.data) and np.array_equal(lb_pin.numpy(), b)


def _nl_case(pkg, name):
    """(grid, fetypes, args, registry name, oracle kernel name, params, state) of a Newton assembly on the tensor-core path."""
    if name == "neohooke_p2":
        E, nu = 10.0, 0.3
        X = np.linspace(0, 1, 4)
        return (pkg.simplexgrid(X, X ** 1.2, X), [pkg.H1P2(3, 3)], [(0, GRAD)], "neohooke3d", "neohooke3d",
                [E / (2 * (1 + nu)), E * nu / ((1 + nu) * (1 - 2 * nu))],
                [lambda x: 0.1 * np.stack([x[:, 0] ** 2, x[:, 0] + x[:, 1], x[:, 1] * x[:, 2]], axis=1)])
    if name == "neohooke_p1":
        E, nu = 10.0, 0.3
        X = np.linspace(0, 1, 5)
        return (pkg.simplexgrid(X, X, X ** 0.9), [pkg.H1Pk(3, 3, 1)], [(0, GRAD)], "neohooke3d", "neohooke3d",
                [E / (2 * (1 + nu)), E * nu / ((1 + nu) * (1 - 2 * nu))],
                [lambda x: 0.1 * np.stack([x[:, 0] * x[:, 2], x[:, 0] - x[:, 1], x[:, 1] ** 2], axis=1)])
    if name == "nse2d":
        return (pkg.uniform_refine(pkg.grid_unitsquare(), 3), [pkg.H1P2(2, 2), pkg.H1P1(1)], [(0, 0), (0, GRAD), (1, 0)], "nse2d", "nse2d",
                [0.05], [lambda x: np.stack([x[:, 0] ** 2, x[:, 0] + x[:, 1]], axis=1), lambda x: x[:, 1] ** 2])
    X = np.linspace(0, 1, 13)
    return (pkg.simplexgrid(X, X ** 1.3), [pkg.H1P2(1, 2)], [(0, 0), (0, GRAD)], "rcd", "rcd", [],
            [lambda x: np.sin(2 * x[:, 0]) + x[:, 1] ** 2])


@pytest.mark.parametrize("case", ["neohooke_p2", "neohooke_p1", "nse2d", "rcd"])
def test_nonlinear_point_kernel_variants(pkg, ora, engine, case):
    """nl_point_kernel with and without the shared-memory cache of the physical basis values (`nonlinear_point_cache`) and, for
    Neo-Hooke, with the Jacobian produced row by row (`nonlinear_rowwise`): all against the oracle, and bit-identical to each
    other (same formulas in the same order)."""
    g, fet, args, kern, okern, params, state = _nl_case(pkg, case)
    S = System(pkg, ora, engine, g, fet)
    u = pkg.FEVector(S.FES)
    for blk, f in enumerate(state):
        pkg.interpolate(u[blk], f)
    sol = u.entries
    d = engine.make_opdesc(args, args=args, kernel_id=pkg.lib.kernel_id(kern), params=params)
    nzref, bref = ora.assemble_nonlinear(S.omesh, S.oargs(args), S.oargs(args), sol, np.zeros(S.N), okern, params=params,
                                         csc=(S.colptr, S.rowval))
    got = {}
    try:
        for cache, row in ((1, 1), (0, 1), (1, 0), (0, 0)):
            engine.set_option("nonlinear_point_cache", cache)
            engine.set_option("nonlinear_rowwise", row)
            nz = np.empty(S.rowval.size); b = np.empty(S.N)
            engine.assemble_nonlinear(S.pat, d, sol, nzval_out=nz, b_out=b)
            check_values(nz, nzref, what=f"jacobian cache={cache} rowwise={row}")
            check_values(b, bref, scale=np.abs(bref).max() + np.abs(nzref).max() * np.abs(sol).max(), what=f"newton rhs cache={cache} rowwise={row}")
            got[(cache, row)] = (nz, b)
        for key in ((0, 1), (1, 0), (0, 0)):
            assert np.array_equal(got[(1, 1)][0], got[key][0]) and np.array_equal(got[(1, 1)][1], got[key][1]), key
    finally:
        engine.set_option("nonlinear_point_cache", 1)
        engine.set_option("nonlinear_rowwise", 1)
