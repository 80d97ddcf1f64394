"""GPU parity tests: the CUDA path (through the C-ABI) against the CPU oracle on the same inputs.

Bar (BASELINE.json north_star): pattern bit-exact (the reference's value-dependent pattern must be
contained in the structural one, and the structural one must equal the oracle's structural pattern
integer for integer); values, rhs and Newton residuals within 1e-12 (see tests/util.py)."""
import numpy as np
import pytest

from util import System, check_values, csc_subset

pytestmark = pytest.mark.gpu

ID, GRAD, DIV, SYMGRAD = 0, 1, 2, 3


def grids(pkg, dim, n=3):
    X = np.linspace(0, 1, n + 1)
    if dim == 1:
        return pkg.simplexgrid(np.linspace(0, 1, 4 * n + 1) ** 1.5)
    if dim == 2:
        return pkg.uniform_refine(pkg.grid_unitsquare(), 2) if n % 2 else pkg.simplexgrid(X, X ** 2)
    return pkg.simplexgrid(X, X ** 1.3, X)


@pytest.mark.parametrize("dim,order", [(1, 1), (1, 2), (2, 1), (2, 2), (3, 1), (3, 2)])
@pytest.mark.parametrize("op", [ID, GRAD])
def test_standard_bilinear(pkg, ora, engine, dim, order, op):
    """Laplace ([grad u]) and mass ([id u]) matrices: pattern + values."""
    g = grids(pkg, dim, 3)
    S = System(pkg, ora, engine, g, [pkg.H1Pk(1, dim, order)])
    # pattern: bit-exact against the oracle's structural pattern
    cp, rv = ora.structural_pattern(S.oargs([(0, op)]), S.oargs([(0, op)]), (S.N, S.N))
    assert np.array_equal(cp, S.colptr) and np.array_equal(rv, S.rowval)
    # the reference's value-dependent pattern is contained in it
    rcp, rrv, rnz = ora.assemble_bilinear(S.omesh, S.oargs([(0, op)]), S.oargs([(0, op)]), "standard", factor=0.7, shape=(S.N, S.N))
    assert csc_subset(rcp, rrv, S.colptr, S.rowval)
    nz = np.empty(S.rowval.size)
    engine.assemble_bilinear(S.pat, engine.make_opdesc([(0, op)], [(0, op)], factor=0.7), nzval_out=nz)
    ref = ora.assemble_bilinear(S.omesh, S.oargs([(0, op)]), S.oargs([(0, op)]), "standard", factor=0.7, csc=(S.colptr, S.rowval))
    check_values(nz, ref)
    # device-resident copy equals what was returned
    nz2, _ = engine.values_get(S.pat)
    assert np.array_equal(nz, nz2)


@pytest.mark.parametrize("dim,order", [(2, 1), (2, 2), (3, 1), (3, 2)])
def test_fastpath_equals_generic(pkg, ora, engine, dim, order):
    """The fused owner-computes kernel and the generic two-phase path are the same sums reassociated."""
    g = grids(pkg, dim, 4)
    g.cellregions[1::4] = 3
    S = System(pkg, ora, engine, g, [pkg.H1Pk(1, dim, order)])
    for op in (GRAD, ID):
        for regions in ((), (1,)):
            desc = engine.make_opdesc([(0, op)], [(0, op)], factor=1.25, regions=regions)
            a = np.empty(S.rowval.size); b = np.empty(S.rowval.size)
            engine.assemble_bilinear(S.pat, desc, nzval_out=a)
            engine.set_option("fastpath", 0)
            try:
                engine.assemble_bilinear(S.pat, desc, nzval_out=b)
            finally:
                engine.set_option("fastpath", 1)
            check_values(a, b, rtol=1e-13, what="fast vs generic")
            ref = ora.assemble_bilinear(S.omesh, S.oargs([(0, op)]), S.oargs([(0, op)]), factor=1.25, regions=list(regions),
                                        csc=(S.colptr, S.rowval))
            check_values(a, ref, what="fast vs oracle")
            # accumulate on top
            engine.assemble_bilinear(S.pat, desc, accumulate=True, nzval_out=b)
            check_values(b, 2 * ref, what="fast accumulate")
            # table evaluator (closed form switched off) gives the same sums
            engine.set_option("fastpath_closed_form", 0)
            try:
                engine.assemble_bilinear(S.pat, desc, nzval_out=b)
            finally:
                engine.set_option("fastpath_closed_form", 1)
            check_values(b, ref, what="fast (table evaluator) vs oracle")


@pytest.mark.parametrize("dim,order,n,permute", [(3, 2, 7, False), (3, 1, 8, False), (2, 2, 12, False), (2, 1, 12, False),
                                                 (3, 2, 7, True), (2, 2, 12, True)])
def test_template_path(pkg, ora, engine, dim, order, n, permute):
    """Template (dictionary-compressed scatter map) kernels: structured grid, templates forced for small groups so that
    nearly every column runs on the template kernel; first-touch and accumulate modes; against the record kernel
    (templates off), the generic path and the oracle; the fast RHS on the same plan."""
    X = np.linspace(0, 1, n + 1)
    g = pkg.simplexgrid(*([X] * dim))
    g.cellregions[1::5] = 2
    if permute:
        # even permutations of the local vertex order (orientation kept) that differ between neighbouring cells: the local
        # edge end points (a, b) of a P2 edge dof then map to its two global vertices in either order (register-row swap)
        cn = g.cellnodes.copy()
        for k, p in ((1, [1, 2, 0]), (2, [2, 0, 1])):
            sel = np.arange(g.ncells) % 3 == k
            cn[sel, :3] = g.cellnodes[sel][:, p]
        g.cellnodes[:] = cn
        g._cache.clear()
    engine.set_option("template_min_cols", 2)
    try:
        S = System(pkg, ora, engine, g, [pkg.H1Pk(1, dim, order)])
        for op in (GRAD, ID):
            for regions in ((), (1,)):
                desc = engine.make_opdesc([(0, op)], [(0, op)], factor=0.75, regions=regions)
                a = np.empty(S.rowval.size); b = np.empty(S.rowval.size)
                engine.assemble_bilinear(S.pat, desc, nzval_out=a)
                st = engine.plan_stats(S.pat, 0)
                assert st["templates"] > 0 and st["template_warps"] > 0, st
                assert st["record_columns"] < st["columns"] // 2, st
                ref = ora.assemble_bilinear(S.omesh, S.oargs([(0, op)]), S.oargs([(0, op)]), factor=0.75, regions=list(regions),
                                            csc=(S.colptr, S.rowval))
                check_values(a, ref, what="template vs oracle")
                engine.assemble_bilinear(S.pat, desc, accumulate=True, nzval_out=b)
                check_values(b, 2 * ref, what="template accumulate")
                engine.set_option("fastpath", 0)
                try:
                    engine.assemble_bilinear(S.pat, desc, nzval_out=b)
                finally:
                    engine.set_option("fastpath", 1)
                check_values(a, b, rtol=1e-13, what="template vs generic")
        # moving mesh: new coordinates (x2) refresh the transposed-order volume copy; mass scales by 2^dim
        dmass = engine.make_opdesc([(0, ID)], [(0, ID)], factor=0.75)
        refm = ora.assemble_bilinear(S.omesh, S.oargs([(0, ID)]), S.oargs([(0, ID)]), factor=0.75, csc=(S.colptr, S.rowval))
        engine.mesh_update_coords(S.mesh, np.ascontiguousarray(2.0 * g.coords))
        engine.assemble_bilinear(S.pat, dmass, nzval_out=a)
        check_values(a, 2.0 ** dim * refm, what="template mass after coordinate update")
        engine.mesh_update_coords(S.mesh, g.coords, g.cellvolumes)
        engine.assemble_bilinear(S.pat, dmass, nzval_out=a)
        check_values(a, refm, what="template mass after restoring coordinates")
        # the same system with templates switched off: record kernel only, identical sums
        engine.set_option("fastpath_templates", 0)
        try:
            S2 = System(pkg, ora, engine, g, [pkg.H1Pk(1, dim, order)])
            desc = engine.make_opdesc([(0, GRAD)], [(0, GRAD)], factor=0.75)
            c = np.empty(S2.rowval.size)
            engine.assemble_bilinear(S2.pat, desc, nzval_out=c)
            assert engine.plan_stats(S2.pat, 0)["templates"] == 0
        finally:
            engine.set_option("fastpath_templates", 1)
        engine.assemble_bilinear(S.pat, desc, nzval_out=a)
        check_values(a, c, rtol=1e-13, what="template vs record kernel")
        # fast RHS on the template plan
        for kernel, params in (("xy", []), ("sincos301", [1.3]), ("constant_one", [])):
            d = engine.make_opdesc([(0, ID)], kernel_id=pkg.lib.kernel_id(kernel), params=params, factor=2.0, regions=(1,))
            rb = np.empty(S.N)
            engine.assemble_linear(S.pat, d, b_out=rb)
            ref = np.zeros(S.N)
            ora.assemble_linear(S.omesh, S.oargs([(0, ID)]), ref, kernel, params=params, factor=2.0, regions=[1])
            check_values(rb, ref, what="template rhs")
            engine.assemble_linear(S.pat, d, accumulate=True, b_out=rb)
            check_values(rb, 2 * ref, what="template rhs accumulate")
            engine.set_option("fastpath", 0)
            try:
                rg = np.empty(S.N)
                engine.assemble_linear(S.pat, d, b_out=rg)
            finally:
                engine.set_option("fastpath", 1)
            check_values(rg, ref, what="generic rhs")
    finally:
        engine.set_option("template_min_cols", 24)


@pytest.mark.parametrize("dim,order,n,permute", [(3, 2, 7, False), (3, 2, 6, True), (3, 1, 8, False), (2, 2, 12, True), (2, 1, 12, False)])
def test_template_specialised_kernels(pkg, ora, engine, dim, order, n, permute):
    """Plan-time specialisation of the templates (NVRTC): same sums as the static template kernel and the oracle; overwrite,
    accumulate, regions."""
    X = np.linspace(0, 1, n + 1)
    g = pkg.simplexgrid(*([X] * dim))
    g.cellregions[1::5] = 2
    if permute:
        cn = g.cellnodes.copy()
        for k, p in ((1, [1, 2, 0]), (2, [2, 0, 1])):
            sel = np.arange(g.ncells) % 3 == k
            cn[sel, :3] = g.cellnodes[sel][:, p]
        g.cellnodes[:] = cn
        g._cache.clear()
    engine.set_option("template_min_cols", 2)
    engine.set_option("template_jit_min_cols", 0)
    engine.set_option("template_jit", 1)
    try:
        S = System(pkg, ora, engine, g, [pkg.H1Pk(1, dim, order)])
        for regions in ((), (1,)):
            desc = engine.make_opdesc([(0, GRAD)], [(0, GRAD)], factor=0.75, regions=regions)
            a = np.empty(S.rowval.size); b = np.empty(S.rowval.size)
            engine.assemble_bilinear(S.pat, desc, nzval_out=a)
            assert engine.plan_jit_status(S.pat, 0) == 1, "specialised kernels were not built"
            ref = ora.assemble_bilinear(S.omesh, S.oargs([(0, GRAD)]), S.oargs([(0, GRAD)]), factor=0.75, regions=list(regions),
                                        csc=(S.colptr, S.rowval))
            check_values(a, ref, what="specialised vs oracle")
            engine.assemble_bilinear(S.pat, desc, accumulate=True, nzval_out=b)
            check_values(b, 2 * ref, what="specialised accumulate")
            engine.set_option("template_jit", 0)
            try:
                engine.assemble_bilinear(S.pat, desc, nzval_out=b)
            finally:
                engine.set_option("template_jit", 1)
            check_values(a, b, rtol=1e-13, what="specialised vs static template kernel")
    finally:
        engine.set_option("template_jit", 0)
        engine.set_option("template_min_cols", 24)
        engine.set_option("template_jit_min_cols", 200000)


def test_template_path_block_system(pkg, ora, engine):
    """Two-block pattern (u, p): the Laplace fast path on block 0 must leave the other blocks' rows zero (no first-touch
    stores there) and the fast RHS must zero the other row blocks."""
    X = np.linspace(0, 1, 9)
    g = pkg.simplexgrid(X, X)
    engine.set_option("template_min_cols", 2)
    try:
        S = System(pkg, ora, engine, g, [pkg.H1Pk(1, 2, 2), pkg.H1Pk(1, 2, 1)])
        desc = engine.make_opdesc([(0, GRAD)], [(0, GRAD)], factor=1.5)
        a = np.empty(S.rowval.size)
        engine.values_set(S.pat, nzval=np.full(S.rowval.size, 7.0), b=np.full(S.N, 7.0))
        engine.assemble_bilinear(S.pat, desc, nzval_out=a)
        assert engine.plan_stats(S.pat, 0)["templates"] > 0
        ref = ora.assemble_bilinear(S.omesh, S.oargs([(0, GRAD)]), S.oargs([(0, GRAD)]), factor=1.5, csc=(S.colptr, S.rowval))
        check_values(a, ref, what="block system template")
        d = engine.make_opdesc([(1, ID)], kernel_id=pkg.lib.kernel_id("xy"))
        rb = np.empty(S.N)
        engine.assemble_linear(S.pat, d, b_out=rb)
        rref = np.zeros(S.N)
        ora.assemble_linear(S.omesh, S.oargs([(1, ID)]), rref, "xy")
        check_values(rb, rref, what="block system rhs")
    finally:
        engine.set_option("template_min_cols", 24)


def test_fastpath_large_unstructured_like(pkg, ora, engine):
    """Larger perturbed 3D P2 grid: many chunks of every shared-memory class, ragged warps, boundary columns."""
    X = np.linspace(0, 1, 10)
    g = pkg.simplexgrid(X ** 1.1, X, X ** 0.9)
    rng = np.random.default_rng(3)
    interior = np.all((g.coords > 1e-9) & (g.coords < 1 - 1e-9), axis=1)
    g.coords[interior] += 0.02 * (rng.random((int(interior.sum()), 3)) - 0.5)
    g._cache.clear()
    S = System(pkg, ora, engine, g, [pkg.H1P2(1, 3)])
    nz = np.empty(S.rowval.size)
    engine.assemble_bilinear(S.pat, engine.make_opdesc([(0, GRAD)], [(0, GRAD)], factor=2.0), nzval_out=nz)
    ref = ora.assemble_bilinear(S.omesh, S.oargs([(0, GRAD)]), S.oargs([(0, GRAD)]), factor=2.0, csc=(S.colptr, S.rowval))
    check_values(nz, ref)


@pytest.mark.parametrize("dim,order,kernel,params", [
    (2, 2, "xy", []), (3, 2, "sincos301", [1.3]), (3, 1, "constant_one", []), (2, 1, "xy", []), (1, 2, "constant_one", [])])
def test_linear_operator(pkg, ora, engine, dim, order, kernel, params):
    g = grids(pkg, dim, 3)
    S = System(pkg, ora, engine, g, [pkg.H1Pk(1, dim, order)])
    b = np.empty(S.N)
    engine.assemble_linear(S.pat, engine.make_opdesc([(0, ID)], kernel_id=pkg.lib.kernel_id(kernel), params=params, factor=2.0), b_out=b)
    ref = np.zeros(S.N)
    ora.assemble_linear(S.omesh, S.oargs([(0, ID)]), ref, kernel, params=params, factor=2.0)
    check_values(b, ref, what="rhs")
    # accumulate=True adds to the device-resident vector
    engine.assemble_linear(S.pat, engine.make_opdesc([(0, ID)], kernel_id=pkg.lib.kernel_id(kernel), params=params, factor=2.0),
                           accumulate=True, b_out=b)
    check_values(b, 2 * ref, what="rhs accumulate")


def test_linear_vector_valued_and_tabulated(pkg, ora, engine):
    g = grids(pkg, 3, 2)
    S = System(pkg, ora, engine, g, [pkg.H1P2(3, 3)])
    f = [0.0, -0.5, 0.25]
    b = np.empty(S.N)
    engine.assemble_linear(S.pat, engine.make_opdesc([(0, ID)], kernel_id=pkg.lib.kernel_id("constant_params"), params=f), b_out=b)
    ref = np.zeros(S.N)
    ora.assemble_linear(S.omesh, S.oargs([(0, ID)]), ref, "constant_params", params=f)
    check_values(b, ref, what="rhs")
    # tabulated: host evaluates an arbitrary closure at the quadrature points
    desc = engine.make_opdesc([(0, ID)], kernel_id=pkg.lib.kernel_id("tabulated"))
    xq = engine.quadrature_points_x(S.pat, desc, g.ncells, 3)
    vals = np.stack([np.exp(xq[..., 0]) * xq[..., 1], xq[..., 2] ** 2, np.sin(xq[..., 0])], axis=-1)
    desc = engine.make_opdesc([(0, ID)], kernel_id=pkg.lib.kernel_id("tabulated"), tabulated=vals)
    engine.assemble_linear(S.pat, desc, b_out=b)
    ref = np.zeros(S.N)
    ora.assemble_linear(S.omesh, S.oargs([(0, ID)]), ref, "tabulated", tabulated=vals)
    check_values(b, ref, what="rhs tabulated")


def _sol_252(S):
    """u = (x^2, x+y), p = y^2 interpolated (test/test_nonlinear_operator.jl:38-39)."""
    u = S.pkg.FEVector(S.FES)
    S.pkg.interpolate(u[0], lambda x: np.stack([x[:, 0] ** 2, x[:, 0] + x[:, 1]], axis=1))
    S.pkg.interpolate(u[1], lambda x: x[:, 1] ** 2)
    return u.entries


@pytest.mark.parametrize("kernel,params", [("linnse7", [0.1, 2.0]), ("nse2d", [0.05])])
def test_nonlinear_2d_p2p1(pkg, ora, engine, kernel, params):
    """NonlinearOperator([id(u), grad(u), id(p)]) on P2xP1: Jacobian + (J u - F) rhs."""
    g = pkg.uniform_refine(pkg.grid_unitsquare(), 2)
    S = System(pkg, ora, engine, g, [pkg.H1P2(2, 2), pkg.H1P1(1)])
    sol = _sol_252(S)
    args = [(0, ID), (0, GRAD), (1, ID)]
    nlk = {"linnse7": "nl_linnse7", "nse2d": "nse2d"}[kernel]
    nz = np.empty(S.rowval.size); b = np.empty(S.N)
    engine.assemble_nonlinear(S.pat, engine.make_opdesc(args, args=args, kernel_id=pkg.lib.kernel_id(nlk), params=params),
                              sol, nzval_out=nz, b_out=b)
    bref = np.zeros(S.N)
    nzref, bref = ora.assemble_nonlinear(S.omesh, S.oargs(args), S.oargs(args), sol, bref, kernel, params=params,
                                         csc=(S.colptr, S.rowval))
    check_values(nz, nzref, what="jacobian")
    # J*u - F(u) cancels to exactly 0 for a linear kernel: compare against the size of the terms
    check_values(b, bref, scale=max(np.abs(bref).max(), np.abs(nzref).max() * np.abs(sol).max()), what="newton rhs")
    # Newton residual b - A*sol (src/solvers.jl:38-43) on the device-resident system
    res = engine.residual(S.pat, sol)
    import scipy.sparse as sp
    A = sp.csc_matrix((nzref, S.rowval - 1, S.colptr - 1), shape=(S.N, S.N))
    check_values(res, bref - A @ sol, scale=max(np.abs(bref).max(), np.abs(nzref).max() * np.abs(sol).max()), what="residual")


@pytest.mark.parametrize("case", ["nse2d", "neohooke3d"])
def test_nonlinear_local_kernel_versions_agree(pkg, ora, engine, case):
    """The staged-contraction local kernel (v2, default) and the entry-wise one are the same sums reassociated."""
    if case == "nse2d":
        g = pkg.uniform_refine(pkg.grid_unitsquare(), 3)
        S = System(pkg, ora, engine, g, [pkg.H1P2(2, 2), pkg.H1P1(1)])
        sol = _sol_252(S)
        args = [(0, ID), (0, GRAD), (1, ID)]
        desc = engine.make_opdesc(args, args=args, kernel_id=pkg.lib.kernel_id("nse2d"), params=[0.05], regions=(1,))
    else:
        g = grids(pkg, 3, 3)
        g.cellregions[::3] = 2
        S = System(pkg, ora, engine, g, [pkg.H1P2(3, 3)])
        u = pkg.FEVector(S.FES)
        pkg.interpolate(u[0], lambda x: 0.1 * np.stack([x[:, 0] ** 2, x[:, 0] + x[:, 1], x[:, 1] * x[:, 2]], axis=1))
        sol = u.entries
        args = [(0, GRAD)]
        desc = engine.make_opdesc(args, args=args, kernel_id=pkg.lib.kernel_id("neohooke3d"), params=[3.8, 5.7], regions=(1,))
    res = {}
    try:
        for ver in (1, 2, 3, 4):     # entry-wise | staged per block | warp per cell | warp per cell on FP64 tensor cores (default)
            engine.set_option("nonlinear_kernel", ver)
            a = np.empty(S.rowval.size); b = np.empty(S.N)
            engine.assemble_nonlinear(S.pat, desc, sol, nzval_out=a, b_out=b)
            res[ver] = (a, b)
    finally:
        engine.set_option("nonlinear_kernel", 4)
    a1, b1 = res[1]
    for ver in (2, 3, 4):
        check_values(res[ver][0], a1, rtol=1e-13, what=f"jacobian v{ver} vs v1")
        check_values(res[ver][1], b1, rtol=1e-13, scale=max(np.abs(b1).max(), np.abs(a1).max() * np.abs(sol).max()), what=f"rhs v{ver} vs v1")


def test_nonlinear_equals_bilinear_for_linear_kernel(pkg, ora, engine):
    """test/test_nonlinear_operator.jl:29-49: Jacobian of a linear kernel == BilinearOperator matrix, < 1e-14."""
    g = pkg.uniform_refine(pkg.grid_unitsquare(), 2)
    S = System(pkg, ora, engine, g, [pkg.H1P2(2, 2), pkg.H1P1(1)])
    sol = _sol_252(S)
    args = [(0, ID), (0, GRAD), (1, ID)]
    params = [0.1, 2.0]
    A1 = np.empty(S.rowval.size); A2 = np.empty(S.rowval.size)
    engine.assemble_nonlinear(S.pat, engine.make_opdesc(args, args=args, kernel_id=pkg.lib.kernel_id("nl_linnse7"), params=params),
                              sol, nzval_out=A1)
    engine.assemble_bilinear(S.pat, engine.make_opdesc(args, args, kernel_id=pkg.lib.kernel_id("linnse7"), params=params), nzval_out=A2)
    assert np.linalg.norm(A1 - A2) < 1e-14 * max(1.0, np.linalg.norm(A2))


@pytest.mark.parametrize("order", [1, 2])
def test_neohooke_3d(pkg, ora, engine, order):
    """Example330: NonlinearOperator(DW, [grad(u)]) with E=10, nu=0.3 (dense 9x9 local Jacobian)."""
    E, nu = 10.0, 0.3
    mu, la = E / (2 * (1 + nu)), E * nu / ((1 + nu) * (1 - 2 * nu))
    g = grids(pkg, 3, 2)
    S = System(pkg, ora, engine, g, [pkg.H1Pk(3, 3, order)])
    u = pkg.FEVector(S.FES)
    pkg.interpolate(u[0], lambda x: 0.1 * np.stack([x[:, 0] ** 2, x[:, 0] + x[:, 1], x[:, 1] * x[:, 2]], axis=1))
    sol = u.entries
    args = [(0, GRAD)]
    nz = np.empty(S.rowval.size); b = np.empty(S.N)
    engine.assemble_nonlinear(S.pat, engine.make_opdesc(args, args=args, kernel_id=pkg.lib.kernel_id("neohooke3d"), params=[mu, la]),
                              sol, nzval_out=nz, b_out=b)
    bref = np.zeros(S.N)
    nzref, bref = ora.assemble_nonlinear(S.omesh, S.oargs(args), S.oargs(args), sol, bref, "neohooke3d", params=[mu, la],
                                         csc=(S.colptr, S.rowval))
    check_values(nz, nzref, what="jacobian")
    check_values(b, bref, what="newton rhs")


@pytest.mark.parametrize("case", ["stokes2d", "stokes3d", "dcr", "hooke_grad3d", "hooke_voigt2d", "linnse7"])
def test_bilinear_kernels(pkg, ora, engine, case):
    if case == "stokes2d":
        g = grids(pkg, 2, 3); fet = [pkg.H1P2(2, 2), pkg.H1P1(1)]
        test = [(0, GRAD), (1, ID)]; kern, params = "stokes", [0.1]
    elif case == "stokes3d":
        g = grids(pkg, 3, 2); fet = [pkg.H1P2(3, 3), pkg.H1P1(1)]
        test = [(0, GRAD), (1, ID)]; kern, params = "stokes", [0.3]
    elif case == "dcr":
        g = grids(pkg, 2, 4); fet = [pkg.H1P2(1, 2)]
        test = [(0, ID), (0, GRAD)]; kern, params = "dcr", [0.01, 1e-5, 1.0, 0.5]
    elif case == "hooke_grad3d":
        g = grids(pkg, 3, 2); fet = [pkg.H1P2(3, 3)]
        test = [(0, GRAD)]; kern, params = "hooke_grad", [3.8, 5.7]
    elif case == "hooke_voigt2d":
        g = grids(pkg, 2, 3); fet = [pkg.H1P2(2, 2)]
        test = [(0, SYMGRAD)]; kern = "hooke_voigt"
        params = np.array([[3.0, 1.0, 0.0], [1.0, 3.0, 0.0], [0.0, 0.0, 1.0]]).ravel()
    else:
        g = grids(pkg, 2, 3); fet = [pkg.H1P2(2, 2), pkg.H1P1(1)]
        test = [(0, ID), (0, GRAD), (1, ID)]; kern, params = "linnse7", [0.1, 2.0]
    S = System(pkg, ora, engine, g, fet)
    nz = np.empty(S.rowval.size)
    engine.assemble_bilinear(S.pat, engine.make_opdesc(test, test, kernel_id=pkg.lib.kernel_id(kern), params=params, factor=1.5),
                             nzval_out=nz)
    ref = ora.assemble_bilinear(S.omesh, S.oargs(test), S.oargs(test), kern, params=params, factor=1.5, csc=(S.colptr, S.rowval))
    check_values(nz, ref, what=case)


def test_bilinear_options(pkg, ora, engine):
    """regions, lump, transposed_copy, bonus_quadorder / explicit quadorder, accumulate."""
    g = grids(pkg, 2, 4)
    g.cellregions[::3] = 2
    S = System(pkg, ora, engine, g, [pkg.H1P2(2, 2), pkg.H1P1(1)])
    nz = np.empty(S.rowval.size)
    # regions
    engine.assemble_bilinear(S.pat, engine.make_opdesc([(0, GRAD)], [(0, GRAD)], regions=[2]), nzval_out=nz)
    ref = ora.assemble_bilinear(S.omesh, S.oargs([(0, GRAD)]), S.oargs([(0, GRAD)]), regions=[2], csc=(S.colptr, S.rowval))
    check_values(nz, ref, what="regions")
    # lumped mass
    for lump in (1, 2):
        engine.assemble_bilinear(S.pat, engine.make_opdesc([(1, ID)], [(1, ID)], lump=lump), nzval_out=nz)
        ref = ora.assemble_bilinear(S.omesh, S.oargs([(1, ID)]), S.oargs([(1, ID)]), lump=lump, csc=(S.colptr, S.rowval))
        check_values(nz, ref, what=f"lump{lump}")
    # divergence constraint with transposed copy: b(u,q) = -(div u, q) and its transpose
    for tc in (1, -1):
        engine.assemble_bilinear(S.pat, engine.make_opdesc([(1, ID)], [(0, DIV)], factor=-1.0, transposed_copy=tc), nzval_out=nz)
        ref = ora.assemble_bilinear(S.omesh, S.oargs([(1, ID)]), S.oargs([(0, DIV)]), factor=-1.0, transposed_copy=tc,
                                    csc=(S.colptr, S.rowval))
        check_values(nz, ref, what=f"transposed_copy{tc}")
    # quadrature order overrides + accumulate on top of the previous matrix
    prev = nz.copy()
    engine.assemble_bilinear(S.pat, engine.make_opdesc([(0, ID)], [(0, ID)], quadorder=5, bonus_quadorder=1), accumulate=True, nzval_out=nz)
    ref2 = ora.assemble_bilinear(S.omesh, S.oargs([(0, ID)]), S.oargs([(0, ID)]), quadorder=5, bonus_quadorder=1, csc=(S.colptr, S.rowval))
    check_values(nz, prev + ref2, what="accumulate")


def test_bilinear_with_args(pkg, ora, engine):
    """BilinearOperator with args (bilinear_operator.jl:451-596): (beta . grad) u with beta = current solution."""
    g = grids(pkg, 2, 3)
    S = System(pkg, ora, engine, g, [pkg.H1P2(2, 2)])
    u = pkg.FEVector(S.FES)
    pkg.interpolate(u[0], lambda x: np.stack([x[:, 0] ** 2 + 1, x[:, 0] - x[:, 1]], axis=1))
    nz = np.empty(S.rowval.size)
    desc = engine.make_opdesc([(0, ID)], [(0, GRAD)], args=[(0, ID)], kernel_id=pkg.lib.kernel_id("convect_args"), quadorder=4)
    engine.assemble_bilinear(S.pat, desc, sol=u.entries, nzval_out=nz)
    ref = ora.assemble_bilinear(S.omesh, S.oargs([(0, ID)]), S.oargs([(0, GRAD)]), "convect_args", quadorder=4,
                                args=S.oargs([(0, ID)]), sol=u.entries, args_sol_offsets=[0], csc=(S.colptr, S.rowval))
    check_values(nz, ref, what="bilinear with args")


def test_block_coupling_pattern(pkg, ora, engine):
    """Stokes with use_sparsity_pattern: the p-p block is not part of the pattern."""
    g = grids(pkg, 2, 3)
    coupling = np.array([[1, 1], [1, 0]], np.uint8)  # [col][row]
    S = System(pkg, ora, engine, g, [pkg.H1P2(2, 2), pkg.H1P1(1)], block_coupling=coupling)
    t = S.oargs([(0, GRAD), (1, ID)])
    cp, rv = ora.structural_pattern(t, t, (S.N, S.N), coupling=coupling)
    assert np.array_equal(cp, S.colptr) and np.array_equal(rv, S.rowval)
    nz = np.empty(S.rowval.size)
    test = [(0, GRAD), (1, ID)]
    engine.assemble_bilinear(S.pat, engine.make_opdesc(test, test, kernel_id=pkg.lib.kernel_id("stokes"), params=[0.1],
                                                       coupling=coupling), nzval_out=nz)
    ref = ora.assemble_bilinear(S.omesh, t, t, "stokes", params=[0.1], coupling=coupling, csc=(S.colptr, S.rowval))
    check_values(nz, ref, what="stokes without p-p block")


def test_errors(pkg, engine):
    g = grids(pkg, 2, 3)
    F = pkg.FESpace(pkg.H1P1(1), g)
    mesh = engine.mesh_set(g.coords, g.cellnodes)
    with pytest.raises(pkg.lib.ExtFEMError) as e:
        engine.space_set(mesh, 77, 1, F.celldofs, F.ndofs)
    assert e.value.code == -2
    sp = engine.space_set(mesh, 1, 1, F.celldofs, F.ndofs)
    pat = engine.pattern_build([sp])
    with pytest.raises(pkg.lib.ExtFEMError) as e:
        engine.assemble_bilinear(pat, engine.make_opdesc([(0, GRAD)], [(0, GRAD)], kernel_id=999))
    assert e.value.code == -1          # unregistered kernel is rejected with an error (north_star)
    with pytest.raises(pkg.lib.ExtFEMError):
        pkg.lib.kernel_id("my_julia_closure")


def test_penalties_and_cg_poisson(pkg, ora, engine):
    """Example201 end to end on the device: assemble, penalties, CG; golden value of the reference
    (examples/Example201_PoissonProblem.jl:80): sum(sol) = 1.1140313632246377."""
    g = pkg.uniform_refine(pkg.grid_unitsquare(), 2)
    S = System(pkg, ora, engine, g, [pkg.H1Pk(1, 2, 2)])
    engine.assemble_bilinear(S.pat, engine.make_opdesc([(0, GRAD)], [(0, GRAD)]))
    engine.assemble_linear(S.pat, engine.make_opdesc([(0, ID)], kernel_id=pkg.lib.kernel_id("xy")))
    bd = np.unique(S.FES[0].bfacedofs)
    engine.apply_penalties(S.pat, bd, None, 1e30)
    x, it, rr = engine.cg(S.pat, rtol=1e-14, maxit=2000)
    assert rr <= 1e-13
    assert abs(x.sum() - 1.1140313632246377) < 1e-10 * 1.1140313632246377
