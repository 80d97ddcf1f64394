"""GPU parity tests of the round-2 rows: LinearOperator with args (a7), the rcd kernel, host-tabulated elements (H1Pk order 3,
BASELINE config 1), ON_BFACES assembly (f3), ItemIntegrator (f2), entrywise (backward-error) comparison, two contexts on one
device, index range checks.  CUDA path through the C-ABI against the CPU oracle on the same inputs."""
import numpy as np
import pytest

from util import System, check_values, check_values_entrywise

pytestmark = pytest.mark.gpu

ID, GRAD, DIV, SYMGRAD = 0, 1, 2, 3


def grids(pkg, dim, n=3):
    X = np.linspace(0, 1, n + 1)
    if dim == 1:
        return pkg.simplexgrid(np.linspace(0, 1, 4 * n + 1) ** 1.5)
    if dim == 2:
        return pkg.uniform_refine(pkg.grid_unitsquare(), 2) if n % 2 else pkg.simplexgrid(X, X ** 2)
    return pkg.simplexgrid(X, X ** 1.3, X)


def _smooth(S, block=0, scale=1.0):
    u = S.pkg.FEVector(S.FES)
    for j, F in enumerate(S.FES):
        nc = F.fetype.ncomponents
        S.pkg.interpolate(u[j], lambda x: scale * np.stack([np.sin(1.3 * x[:, 0] + 0.4 * c) + 0.3 * (x[:, -1] ** 2) for c in range(nc)], axis=1))
    return u.entries


@pytest.mark.parametrize("dim,order,ncomp,op", [(1, 2, 1, ID), (2, 2, 1, ID), (2, 1, 2, ID), (3, 2, 1, ID), (3, 1, 3, ID), (2, 2, 1, GRAD)])
def test_linear_operator_with_args(pkg, ora, engine, dim, order, ncomp, op):
    """LinearOperator(oa_test, oa_args) (linear_operator.jl:241-480, loop :359-438): b = (input_args(sol), test), the form
    `LinearOperator(M-free mass term, [id(u)], [id(u)])` time steppers use; quadorder = polyorder_test + polyorder_args."""
    g = grids(pkg, dim, 3)
    g.cellregions[::4] = 2
    S = System(pkg, ora, engine, g, [pkg.H1Pk(ncomp, dim, order)])
    sol = _smooth(S)
    for regions in ((), (1,)):
        d = engine.make_opdesc([(0, op)], args=[(0, op)], kernel_id=pkg.lib.kernel_id("standard"), factor=1.7, regions=regions)
        b = np.empty(S.N)
        engine.assemble_linear(S.pat, d, sol=sol, b_out=b)
        ref = np.zeros(S.N)
        ora.assemble_linear(S.omesh, S.oargs([(0, op)]), ref, "standard", args=S.oargs([(0, op)]), sol=sol, args_sol_offsets=[0],
                            factor=1.7, regions=list(regions))
        check_values(b, ref, what="linear with args")
        with ora.abs_accumulate():
            sc = np.zeros(S.N)
            ora.assemble_linear(S.omesh, S.oargs([(0, op)]), sc, "standard", args=S.oargs([(0, op)]), sol=sol, args_sol_offsets=[0],
                                factor=1.7, regions=list(regions))
        check_values_entrywise(b, ref, sc, what="linear with args (entrywise)")
        engine.assemble_linear(S.pat, d, sol=sol, accumulate=True, b_out=b)
        check_values(b, 2 * ref, what="linear with args, accumulate")
    # mass matrix times solution is the same vector (the identity time steppers rely on)
    if op == ID:
        import scipy.sparse as sp
        nz = ora.assemble_bilinear(S.omesh, S.oargs([(0, ID)]), S.oargs([(0, ID)]), factor=1.7, regions=[1], csc=(S.colptr, S.rowval))
        M = sp.csc_matrix((nz, S.rowval - 1, S.colptr - 1), shape=(S.N, S.N))
        check_values(b, 2 * (M @ sol), rtol=1e-11, what="(u, v) == M u")


@pytest.mark.parametrize("dim,order", [(1, 1), (1, 2), (2, 2)])
def test_rcd_kernel(pkg, ora, engine, dim, order):
    """NonlinearOperator(nonlinear_kernel!, [id(u), grad(u)]) of Example108:40-45 in 1D (and 2D): Jacobian, rhs, residual."""
    g = grids(pkg, dim, 3)
    S = System(pkg, ora, engine, g, [pkg.H1Pk(1, dim, order)])
    sol = _smooth(S)
    args = [(0, ID), (0, GRAD)]
    for kernel, params in (("rcd", []), ("nlpoisson105", [1e-3])):
        nz = np.empty(S.rowval.size); b = np.empty(S.N)
        engine.assemble_nonlinear(S.pat, engine.make_opdesc(args, args=args, kernel_id=pkg.lib.kernel_id(kernel), params=params),
                                  sol, nzval_out=nz, b_out=b)
        nzref, bref = ora.assemble_nonlinear(S.omesh, S.oargs(args), S.oargs(args), sol, np.zeros(S.N), kernel, params=params,
                                             csc=(S.colptr, S.rowval))
        check_values(nz, nzref, what=f"{kernel} jacobian")
        scale = max(np.abs(bref).max(), np.abs(nzref).max() * np.abs(sol).max())
        check_values(b, bref, scale=scale, what=f"{kernel} rhs")
        with ora.abs_accumulate():
            nzs, bs = ora.assemble_nonlinear(S.omesh, S.oargs(args), S.oargs(args), sol, np.zeros(S.N), kernel, params=params,
                                             csc=(S.colptr, S.rowval))
        check_values_entrywise(nz, nzref, nzs, what=f"{kernel} jacobian (entrywise)")
        check_values_entrywise(b, bref, bs, rtol=2e-12, what=f"{kernel} rhs (entrywise)")
        import scipy.sparse as sp
        A = sp.csc_matrix((nzref, S.rowval - 1, S.colptr - 1), shape=(S.N, S.N))
        check_values(engine.residual(S.pat, sol), bref - A @ sol, scale=scale, what=f"{kernel} residual")


def test_stvenant230_kernel(pkg, ora, engine):
    """Example230:39-72: region-dependent material inside the kernel (qpinfo.region)."""
    g = pkg.simplexgrid(np.linspace(-2, 2, 9), np.linspace(0, 1, 5))
    mid = g.coords[g.cellnodes.astype(np.int64) - 1].mean(axis=1)
    g.cellregions[:] = np.where(mid[:, 1] < 0.5, 1, 2)
    S = System(pkg, ora, engine, g, [pkg.H1Pk(2, 2, 2)])
    sol = _smooth(S, scale=0.05)
    params = [2.0, 1.2, 0.6, 0.8, 0.4, 7.5e-3, 0.14]
    args = [(0, GRAD)]
    nz = np.empty(S.rowval.size); b = np.empty(S.N)
    engine.assemble_nonlinear(S.pat, engine.make_opdesc(args, args=args, kernel_id=pkg.lib.kernel_id("stvenant230"), params=params),
                              sol, nzval_out=nz, b_out=b)
    nzref, bref = ora.assemble_nonlinear(S.omesh, S.oargs(args), S.oargs(args), sol, np.zeros(S.N), "stvenant230", params=params,
                                         csc=(S.colptr, S.rowval))
    check_values(nz, nzref, what="stvenant230 jacobian")
    check_values(b, bref, what="stvenant230 rhs")
    for ver in (1, 2, 3):
        engine.set_option("nonlinear_kernel", ver)
        try:
            a2 = np.empty(S.rowval.size)
            engine.assemble_nonlinear(S.pat, engine.make_opdesc(args, args=args, kernel_id=pkg.lib.kernel_id("stvenant230"), params=params),
                                      sol, nzval_out=a2)
        finally:
            engine.set_option("nonlinear_kernel", 4)
        check_values(a2, nzref, what=f"stvenant230 jacobian, local kernel v{ver}")


@pytest.mark.parametrize("dim", [1, 2])
def test_h1pk_order3_tabulated(pkg, ora, engine, dim):
    """BASELINE config 1 (README.md:26-52, Example201:66): H1Pk order 3 through EXTFEM_FE_TABULATED -- the host supplies the cubic
    reference basis as polynomial coefficients (extfem_space_set_tables), the engine evaluates it at its quadrature points.
    Stiffness (factor 1e-3), mass, right-hand side f = x*y against the oracle's closed-form cubic basis; then the README problem
    end to end: the discrete solution reproduces a cubic exactly."""
    h = 0.1
    X = np.arange(0, 1 + h / 2, h)
    g = pkg.simplexgrid(X) if dim == 1 else pkg.simplexgrid(X, X)
    S = System(pkg, ora, engine, g, [pkg.H1Pk(1, dim, 3)])
    assert S.FES[0].fetype.fe_id == pkg.EXTFEM_FE_TABULATED
    cp, rv = ora.structural_pattern(S.oargs([(0, ID)]), S.oargs([(0, ID)]), (S.N, S.N))
    assert np.array_equal(cp, S.colptr) and np.array_equal(rv, S.rowval)
    for op, factor in ((GRAD, 1e-3), (ID, 1.0)):
        nz = np.empty(S.rowval.size)
        engine.assemble_bilinear(S.pat, engine.make_opdesc([(0, op)], [(0, op)], factor=factor), nzval_out=nz)
        ref = ora.assemble_bilinear(S.omesh, S.oargs([(0, op)]), S.oargs([(0, op)]), factor=factor, csc=(S.colptr, S.rowval))
        check_values(nz, ref, what="P3 matrix")
        with ora.abs_accumulate():
            sc = ora.assemble_bilinear(S.omesh, S.oargs([(0, op)]), S.oargs([(0, op)]), factor=factor, csc=(S.colptr, S.rowval))
        check_values_entrywise(nz, ref, sc, what="P3 matrix (entrywise)")
    kern = "xy" if dim == 2 else "exp2x"
    b = np.empty(S.N)
    engine.assemble_linear(S.pat, engine.make_opdesc([(0, ID)], kernel_id=pkg.lib.kernel_id(kern)), b_out=b)
    bref = np.zeros(S.N)
    ora.assemble_linear(S.omesh, S.oargs([(0, ID)]), bref, kern)
    check_values(b, bref, what="P3 rhs")
    # exactness: -Laplace u = f with u cubic (tabulated rhs f at the engine's own quadrature points), Dirichlet data = u
    pts = S.FES[0].dof_coordinates()
    if dim == 1:
        uex = lambda x: x[:, 0] ** 3 - 0.5 * x[:, 0]                       # noqa: E731
        fex = lambda x: -6.0 * x[..., 0]                                    # noqa: E731
    else:
        uex = lambda x: x[:, 0] ** 3 - 2 * x[:, 0] * x[:, 1] ** 2 + x[:, 1]   # noqa: E731
        fex = lambda x: -(6.0 * x[..., 0] - 4.0 * x[..., 0])                 # noqa: E731
    d0 = engine.make_opdesc([(0, ID)], kernel_id=pkg.lib.kernel_id("tabulated"), quadorder=4)
    xq = engine.quadrature_points_x(S.pat, d0, g.ncells, dim)
    engine.assemble_bilinear(S.pat, engine.make_opdesc([(0, GRAD)], [(0, GRAD)]))
    engine.assemble_linear(S.pat, engine.make_opdesc([(0, ID)], kernel_id=pkg.lib.kernel_id("tabulated"), quadorder=4,
                                                     tabulated=fex(xq)[..., None]))
    bd = np.unique(S.FES[0].bfacedofs)
    engine.apply_penalties(S.pat, bd, uex(pts)[bd - 1], 1e30)
    import scipy.sparse as sp, scipy.sparse.linalg as spla
    nz, bb = engine.values_get(S.pat)
    x = spla.spsolve(sp.csc_matrix((nz, S.rowval - 1, S.colptr - 1), shape=(S.N, S.N)), bb)
    assert np.abs(x - uex(pts)).max() < 1e-11


@pytest.mark.parametrize("dim,order,ncomp", [(1, 2, 1), (2, 1, 1), (2, 2, 1), (2, 2, 2), (3, 1, 1), (3, 2, 1), (3, 2, 3), (2, 3, 1)])
def test_on_bfaces(pkg, ora, engine, dim, order, ncomp):
    """entities = ON_BFACES (bilinear_operator.jl:707-714, linear_operator.jl:531-544): boundary mass / Robin matrices and
    boundary load vectors (traction, Example330:105) with regions, added into the cell pattern."""
    g = grids(pkg, dim, 2 if dim == 3 else 3)
    S = System(pkg, ora, engine, g, [pkg.H1Pk(ncomp, dim, order)])
    F = S.FES[0]
    engine.mesh_set_bfaces(S.mesh, g.bfacenodes, g.bfaceregions, g.bfacevolumes)
    if F.fetype.fe_id == pkg.EXTFEM_FE_TABULATED:
        engine.space_set_tables(S.spaces[0], 3, F.ref_coeffs, F.ref_coeffs_bface)
    engine.space_set_bfacedofs(S.spaces[0], F.bfacedofs)
    bmesh = ora.Mesh(g.coords, g.bfacenodes, g.bfaceregions, g.bfacevolumes)
    barg = lambda op: [ora.OraArg(F.bfacedofs, ncomp, order, op, 0)]      # noqa: E731
    regs = sorted(set(g.bfaceregions.tolist()))
    for kernel, params, regions in (("standard", [], ()), ("robin108", [2.0], tuple(regs[:1])), ("standard", [], tuple(regs[-2:]))):
        d = engine.make_opdesc([(0, ID)], [(0, ID)], kernel_id=pkg.lib.kernel_id(kernel), params=params, factor=0.6, regions=regions,
                               entities=pkg.lib.ON_BFACES)
        nz = np.empty(S.rowval.size)
        engine.assemble_bilinear(S.pat, d, nzval_out=nz)
        ref = ora.assemble_bilinear(bmesh, barg(ID), barg(ID), kernel, params=params, factor=0.6, regions=list(regions),
                                    csc=(S.colptr, S.rowval))
        assert np.abs(ref).max() > 0
        check_values(nz, ref, what=f"bface {kernel}")
        # on top of a cell operator (accumulate), like assemble_system! does
        engine.assemble_bilinear(S.pat, engine.make_opdesc([(0, GRAD)], [(0, GRAD)]))
        engine.assemble_bilinear(S.pat, d, accumulate=True, nzval_out=nz)
        refc = ora.assemble_bilinear(S.omesh, S.oargs([(0, GRAD)]), S.oargs([(0, GRAD)]), csc=(S.colptr, S.rowval))
        check_values(nz, refc + ref, what=f"cells + bface {kernel}")
    f = [0.3, -1.0, 0.5][:ncomp]
    for kernel, params, regions in (("constant_params", f, ()), ("constant_params", f, tuple(regs[:1])), ("constant_one", [], tuple(regs[-1:]))):
        d = engine.make_opdesc([(0, ID)], kernel_id=pkg.lib.kernel_id(kernel), params=params, factor=1.5, regions=regions, entities=pkg.lib.ON_BFACES)
        b = np.empty(S.N)
        engine.assemble_linear(S.pat, d, b_out=b)
        ref = np.zeros(S.N)
        ora.assemble_linear(bmesh, barg(ID), ref, kernel, params=params, factor=1.5, regions=list(regions))
        check_values(b, ref, what=f"bface rhs {kernel}")
        engine.assemble_linear(S.pat, d, accumulate=True, b_out=b)
        check_values(b, 2 * ref, what=f"bface rhs {kernel}, accumulate")
    # the boundary measure: sum of the boundary mass matrix = |boundary| * ncomp
    d = engine.make_opdesc([(0, ID)], [(0, ID)], entities=pkg.lib.ON_BFACES)
    nz = np.empty(S.rowval.size)
    engine.assemble_bilinear(S.pat, d, nzval_out=nz)
    assert abs(nz.sum() - ncomp * g.bfacevolumes.sum()) < 1e-12 * max(1.0, g.bfacevolumes.sum())
    # gradients on faces are outside the path: rejected, not silently wrong
    with pytest.raises(pkg.lib.ExtFEMError) as e:
        engine.assemble_bilinear(S.pat, engine.make_opdesc([(0, GRAD)], [(0, GRAD)], entities=pkg.lib.ON_BFACES))
    assert e.value.code == -2


@pytest.mark.parametrize("dim,order,ncomp", [(1, 2, 1), (2, 2, 2), (3, 2, 1), (3, 1, 3), (2, 3, 1)])
def test_item_integrator(pkg, ora, engine, dim, order, ncomp):
    """ItemIntegrator (item_integrator.jl:191-249): piecewise [resultdim, ncells] and global sums; standard / l2norm kernels on
    [id(u)] and [grad(u)]; an exact_error!-type closure through host-evaluated reference values (l2diff_tabulated); regions."""
    g = grids(pkg, dim, 3)
    g.cellregions[1::3] = 2
    S = System(pkg, ora, engine, g, [pkg.H1Pk(ncomp, dim, order)])
    sol = _smooth(S)
    for kernel, op, regions, qo in (("ii_standard", ID, (), -1), ("l2norm", GRAD, (1,), -1), ("l2norm", ID, (), 6)):
        rd = ncomp if op == ID else ncomp * dim
        d = engine.make_opdesc([], args=[(0, op)], kernel_id=pkg.lib.kernel_id(kernel), factor=1.2, regions=regions, quadorder=qo)
        out = engine.integrate(S.pat, d, sol, resultdim=rd, piecewise=True, nitems=g.ncells)
        ref = ora.integrate(S.omesh, S.oargs([(0, op)]), sol, kernel, factor=1.2, regions=list(regions),
                            quadorder="auto" if qo < 0 else qo)
        check_values(out, ref, what=f"piecewise {kernel}")
        tot = engine.integrate(S.pat, d, sol, resultdim=rd, piecewise=False)
        check_values(tot, ref.sum(axis=0), rtol=1e-12, scale=np.abs(ref).sum(axis=0).max(), what=f"global {kernel}")
    # exact-error closure: (u_exact - u_h)^2 with u_exact evaluated by the host at the engine's quadrature points
    qo = 5
    xq = ora.quadrature_points(S.omesh, qo)
    tab = np.stack([np.cos(xq[..., 0] + 0.2 * c) for c in range(ncomp)], axis=-1)
    d = engine.make_opdesc([], args=[(0, ID)], kernel_id=pkg.lib.kernel_id("l2diff_tabulated"), quadorder=qo, tabulated=tab)
    xq2 = engine.quadrature_points_x(S.pat, engine.make_opdesc([(0, ID)], kernel_id=pkg.lib.kernel_id("tabulated"), quadorder=qo), g.ncells, dim)
    assert np.abs(xq - xq2).max() < 1e-14
    out = engine.integrate(S.pat, d, sol, resultdim=ncomp, piecewise=True, nitems=g.ncells)
    ref = ora.integrate(S.omesh, S.oargs([(0, ID)]), sol, "l2diff_tabulated", quadorder=qo, tabulated=tab)
    check_values(out, ref, what="l2diff_tabulated")
    # mass identity: integral of id(u) over the domain == 1^T M u
    if ncomp == 1:
        nz = ora.assemble_bilinear(S.omesh, S.oargs([(0, ID)]), S.oargs([(0, ID)]), csc=(S.colptr, S.rowval))
        import scipy.sparse as sp
        M = sp.csc_matrix((nz, S.rowval - 1, S.colptr - 1), shape=(S.N, S.N))
        d = engine.make_opdesc([], args=[(0, ID)], kernel_id=pkg.lib.kernel_id("ii_standard"), quadorder=2 * order)
        tot = engine.integrate(S.pat, d, sol, resultdim=1, piecewise=False)
        assert abs(tot[0] - np.ones(S.N) @ (M @ sol)) < 1e-12 * np.abs(M @ sol).sum()


def test_entrywise_backward_error_3d_p2(pkg, ora, engine):
    """The headline operator on a GRADED 3D P2 mesh, entrywise: every matrix entry within 1e-12 of the oracle relative to the sum of
    its absolute cell contributions (not merely relative to the largest entry of the matrix); fast path and generic path."""
    X = np.linspace(0, 1, 7)
    g = pkg.simplexgrid(X ** 3, X, X ** 0.5)
    S = System(pkg, ora, engine, g, [pkg.H1P2(1, 3)])
    ref = ora.assemble_bilinear(S.omesh, S.oargs([(0, GRAD)]), S.oargs([(0, GRAD)]), csc=(S.colptr, S.rowval))
    with ora.abs_accumulate():
        sc = ora.assemble_bilinear(S.omesh, S.oargs([(0, GRAD)]), S.oargs([(0, GRAD)]), csc=(S.colptr, S.rowval))
    assert sc.min() >= 0 and (np.abs(ref) <= sc * (1 + 1e-14)).all()
    nz = np.empty(S.rowval.size)
    for fast in (1, 0):
        engine.set_option("fastpath", fast)
        try:
            engine.assemble_bilinear(S.pat, engine.make_opdesc([(0, GRAD)], [(0, GRAD)]), nzval_out=nz)
        finally:
            engine.set_option("fastpath", 1)
        worst = check_values_entrywise(nz, ref, sc, what=f"3D P2 Laplace entrywise (fastpath={fast})")
        assert worst < 1e-12
    b = np.empty(S.N)
    engine.assemble_linear(S.pat, engine.make_opdesc([(0, ID)], kernel_id=pkg.lib.kernel_id("sincos301"), params=[1.0]), b_out=b)
    bref = np.zeros(S.N); bs = np.zeros(S.N)
    ora.assemble_linear(S.omesh, S.oargs([(0, ID)]), bref, "sincos301", params=[1.0])
    with ora.abs_accumulate():
        ora.assemble_linear(S.omesh, S.oargs([(0, ID)]), bs, "sincos301", params=[1.0])
    check_values_entrywise(b, bref, bs, what="3D P2 rhs entrywise")


def test_two_contexts_on_one_device(pkg, ora):
    """Two contexts on cuda:0 used alternately WITHOUT extfem_synchronize: they share the device's __constant__ tables (templates,
    reference tables); uploads of one are ordered behind the other's kernels (event wait), so neither matrix is corrupted."""
    e1, e2 = pkg.lib.Engine(0), pkg.lib.Engine(0)
    try:
        for e in (e1, e2):
            e.set_option("template_min_cols", 2)
        X = np.linspace(0, 1, 13)
        gA = pkg.simplexgrid(X, X, X)                       # 3D P2: many template rounds
        gB = pkg.simplexgrid(np.linspace(0, 1, 40), np.linspace(0, 2, 33))
        SA = System(pkg, ora, e1, gA, [pkg.H1P2(1, 3)])
        SB = System(pkg, ora, e2, gB, [pkg.H1Pk(1, 2, 2)])
        dA = e1.make_opdesc([(0, GRAD)], [(0, GRAD)], factor=1.1)
        dB = e2.make_opdesc([(0, GRAD)], [(0, GRAD)], factor=0.9)
        rA = e1.make_opdesc([(0, ID)], kernel_id=pkg.lib.kernel_id("xy"))
        rB = e2.make_opdesc([(0, ID)], kernel_id=pkg.lib.kernel_id("xy"))
        refA = ora.assemble_bilinear(SA.omesh, SA.oargs([(0, GRAD)]), SA.oargs([(0, GRAD)]), factor=1.1, csc=(SA.colptr, SA.rowval))
        refB = ora.assemble_bilinear(SB.omesh, SB.oargs([(0, GRAD)]), SB.oargs([(0, GRAD)]), factor=0.9, csc=(SB.colptr, SB.rowval))
        bA = np.zeros(SA.N); bB = np.zeros(SB.N)
        ora.assemble_linear(SA.omesh, SA.oargs([(0, ID)]), bA, "xy")
        ora.assemble_linear(SB.omesh, SB.oargs([(0, ID)]), bB, "xy")
        for _ in range(6):                                   # device-resident calls return before the GPU finishes
            e1.assemble_bilinear(SA.pat, dA); e2.assemble_bilinear(SB.pat, dB)
            e1.assemble_linear(SA.pat, rA); e2.assemble_linear(SB.pat, rB)
        nzA, b1 = e1.values_get(SA.pat)
        nzB, b2 = e2.values_get(SB.pat)
        check_values(nzA, refA, what="context A matrix"); check_values(nzB, refB, what="context B matrix")
        check_values(b1, bA, what="context A rhs"); check_values(b2, bB, what="context B rhs")
    finally:
        e1.close(); e2.close()


def test_index_range_is_checked(pkg, engine):
    """0-based or out-of-range index arrays from the host are rejected (EXTFEM_ERR_BAD_ARGUMENT), not read out of bounds."""
    g = grids(pkg, 2, 3)
    F = pkg.FESpace(pkg.H1P1(1), g)
    with pytest.raises(pkg.lib.ExtFEMError) as e:
        engine.mesh_set(g.coords, g.cellnodes - 1)
    assert e.value.code == -3 and "1-based" in str(e.value)
    mesh = engine.mesh_set(g.coords, g.cellnodes)
    bad = F.celldofs.copy(); bad[3, 1] = F.ndofs + 1
    with pytest.raises(pkg.lib.ExtFEMError) as e:
        engine.space_set(mesh, 1, 1, bad, F.ndofs)
    assert e.value.code == -3


def test_apply_values_and_zero(pkg, ora, engine):
    """extfem_values_zero (fill!(nzval, 0), solvers.jl:130-135) and the assemble_sol leg of apply_penalties!."""
    g = grids(pkg, 2, 3)
    S = System(pkg, ora, engine, g, [pkg.H1Pk(1, 2, 2)])
    engine.assemble_bilinear(S.pat, engine.make_opdesc([(0, GRAD)], [(0, GRAD)]))
    engine.assemble_linear(S.pat, engine.make_opdesc([(0, ID)], kernel_id=pkg.lib.kernel_id("xy")))
    engine.values_zero(S.pat, True, False)
    nz, b = engine.values_get(S.pat)
    assert not nz.any() and b.any()
    engine.values_zero(S.pat, False, True)
    assert not engine.values_get(S.pat)[1].any()
    sol = np.arange(S.N, dtype=np.float64)
    dofs = np.array([1, 5, S.N])
    engine.apply_values(dofs, np.array([7.0, 8.0, 9.0]), sol)
    assert sol[0] == 7.0 and sol[4] == 8.0 and sol[-1] == 9.0 and sol[1] == 1.0


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["p2_3d", "stokes2d"])
def test_lower_triangle_export(pkg, ora, engine, case):
    """extfem_pattern_get_lower / extfem_values_get_lower: the packed lower triangle (pattern Int64 1-based, values, rhs) equals
    scipy's tril of the full matrix -- scalar P2 in 3D and a 2-block (velocity, pressure) system."""
    import scipy.sparse as sp
    if case == "p2_3d":
        X = np.linspace(0, 1, 6)
        S = System(pkg, ora, engine, pkg.simplexgrid(X, X, X), [pkg.H1P2(1, 3)])
        engine.assemble_bilinear(S.pat, engine.make_opdesc([(0, GRAD)], [(0, GRAD)], factor=1.3))
    else:
        X = np.linspace(0, 1, 9)
        S = System(pkg, ora, engine, pkg.simplexgrid(X, X), [pkg.H1P2(2, 2), pkg.H1P1(1)])
        t = [(0, GRAD), (1, ID)]
        engine.assemble_bilinear(S.pat, engine.make_opdesc(t, t, kernel_id=pkg.lib.kernel_id("stokes"), params=[0.1]))
    engine.assemble_linear(S.pat, engine.make_opdesc([(0, ID)], kernel_id=pkg.lib.kernel_id("constant_one")))
    nz, b = engine.values_get(S.pat)
    A = sp.csc_matrix((nz, S.rowval - 1, S.colptr - 1), shape=(S.N, S.N))
    L = sp.tril(A, format="csc")
    # structural zeros of the pattern survive in A (explicit entries), tril keeps them
    n, cp, rv = engine.pattern_get_lower(S.pat)
    assert n == L.nnz and np.array_equal(cp - 1, L.indptr) and np.array_equal(rv - 1, L.indices)
    lz, lb = engine.values_get_lower(S.pat)
    assert np.array_equal(lz, L.data) and np.array_equal(lb, b)
