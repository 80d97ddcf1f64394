"""CPU tests (no GPU): pin the oracle (oracle/assembly_ref.c + oracle/fetables.py) against the
reference's own golden values and invariants.

What the reference's tests hold for this path, and how each is used here:

  * examples/Example201_PoissonProblem.jl:80   sum(sol.entries) ~ 1.1140313632246377
    (2D P2 Poisson, f = x*y, uniform_refine(grid_unitsquare, 2)).  The right-hand side integrand has
    degree 4 but is integrated with the order-2 rule (linear_operator.jl:534-536), so this number
    pins the order-2 triangle rule, the P2 basis, |T| scaling and the penalty treatment.
  * test/test_nonlinear_operator.jl:29-49      Jacobian of a linear kernel via NonlinearOperator ==
    matrix of BilinearOperator, norm < 1e-14.
  * test/test_timedependence.jl:15-72          backward Euler for the heat equation reproduces a
    solution that is quadratic in space / linear in time, error < 1e-14 (stiffness + mass exactness).
  * examples/Example108_RobinBoundaryCondition.jl:95 is NOT reproducible without the ON_BFACES operator
    (out of scope).
"""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from oracle import fetables
from util import System

ID, GRAD = 0, 1


def _csc(cp, rv, nz, N):
    return sp.csc_matrix((nz, rv - 1, cp - 1), shape=(N, N))


def test_example201_golden_value(pkg, ora):
    """examples/Example201_PoissonProblem.jl:79-80 (nrefs=2, order=2, penalty variant :83-84)."""
    grid = pkg.uniform_refine(pkg.grid_unitsquare(), 2)
    S = System(pkg, ora, None, grid, [pkg.H1Pk(1, 2, 2)])
    cp, rv, nz = ora.assemble_bilinear(S.omesh, S.oargs([(0, GRAD)]), S.oargs([(0, GRAD)]), "standard", factor=1.0,
                                       shape=(S.N, S.N))
    b = np.zeros(S.N)
    ora.assemble_linear(S.omesh, S.oargs([(0, ID)]), b, "xy")
    A = _csc(cp, rv, nz, S.N).tolil()
    bd = np.unique(S.FES[0].bfacedofs) - 1
    for d in bd:                       # apply_penalties!: homogeneousdata_operator.jl:186-201
        A[d, d] = 1e30
        b[d] = 0.0
    x = spla.spsolve(A.tocsc(), b)
    # Julia's isapprox default rtol is sqrt(eps); the oracle reproduces the value to rounding
    assert abs(x.sum() - 1.1140313632246377) <= 1e-12 * 1.1140313632246377


def _sol_252(S):
    sol = np.zeros(S.N)
    xu = S.FES[0].dof_coordinates()
    n0 = S.FES[0].coffset
    sol[0:n0] = xu[:, 0] ** 2
    sol[n0:2 * n0] = xu[:, 0] + xu[:, 1]
    xp = S.FES[1].dof_coordinates()
    sol[S.offsets[1]:] = xp[:, 1] ** 2
    return sol


def test_linear_nonlinear_operator_equivalence(pkg, ora):
    """test/test_nonlinear_operator.jl:29-49 with its tolerance (:9)."""
    grid = pkg.uniform_refine(pkg.grid_unitsquare(), 2)
    S = System(pkg, ora, None, grid, [pkg.H1P2(2, 2), pkg.H1P1(1)])
    ops = [(0, ID), (0, GRAD), (1, ID)]
    sol = _sol_252(S)
    b = np.zeros(S.N)
    (cpn, rvn, nzn), _ = ora.assemble_nonlinear(S.omesh, S.oargs(ops), S.oargs(ops), sol, b, "linnse7", params=[0.1, 2.0],
                                                shape=(S.N, S.N))
    cpl, rvl, nzl = ora.assemble_bilinear(S.omesh, S.oargs(ops), S.oargs(ops), "linnse7", params=[0.1, 2.0], shape=(S.N, S.N))
    D = _csc(cpn, rvn, nzn, S.N) - _csc(cpl, rvl, nzl, S.N)
    assert spla.norm(D) < 1e-14
    # the linearised right-hand side J u - F(u) vanishes for a linear kernel
    assert np.abs(b).max() < 1e-13


def test_heat_equation_exactness(pkg, ora):
    """test/test_timedependence.jl:15-72: u = t + (x^2+y^2)/4 solves u_t - lap u = 0 exactly in P2 / backward Euler."""
    grid = pkg.uniform_refine(pkg.grid_unitsquare(scale=(4, 4), shift=(-0.5, -0.5)), 2)
    S = System(pkg, ora, None, grid, [pkg.H1Pk(1, 2, 2)])
    K = _csc(*ora.assemble_bilinear(S.omesh, S.oargs([(0, GRAD)]), S.oargs([(0, GRAD)]), "standard", shape=(S.N, S.N)), S.N)
    M = _csc(*ora.assemble_bilinear(S.omesh, S.oargs([(0, ID)]), S.oargs([(0, ID)]), "standard", shape=(S.N, S.N)), S.N)
    X = S.FES[0].dof_coordinates()
    exact = lambda t: t + (X[:, 0] ** 2 + X[:, 1] ** 2) / 4
    tau, T = 0.5, 2.0
    u, t = exact(0.0), 0.0
    bd = np.unique(S.FES[0].bfacedofs) - 1
    for _ in range(int(T / tau)):
        t += tau
        A = (K + M / tau).tolil()
        b = M @ u / tau
        for d in bd:
            A[d, d] = 1e30
            b[d] = 1e30 * exact(t)[d]
        u = spla.spsolve(A.tocsc(), b)
    assert np.abs(u - exact(T)).max() < 1e-13


@pytest.mark.parametrize("dim,order", [(1, 1), (1, 2), (2, 1), (2, 2), (3, 1), (3, 2)])
def test_stiffness_and_mass_invariants(pkg, ora, dim, order):
    """Constants are in the kernel of the stiffness matrix; the mass matrix sums to |Omega|;
    both are symmetric; the structural pattern contains the value-dependent one."""
    X = np.linspace(0, 1, 4)
    grid = pkg.simplexgrid(*([X ** 1.2] * dim))
    S = System(pkg, ora, None, grid, [pkg.H1Pk(1, dim, order)])
    K = _csc(*ora.assemble_bilinear(S.omesh, S.oargs([(0, GRAD)]), S.oargs([(0, GRAD)]), "standard", shape=(S.N, S.N)), S.N)
    M = _csc(*ora.assemble_bilinear(S.omesh, S.oargs([(0, ID)]), S.oargs([(0, ID)]), "standard", shape=(S.N, S.N)), S.N)
    scale = abs(K).max()
    assert np.abs(K @ np.ones(S.N)).max() < 1e-13 * scale
    assert abs(M.sum() - 1.0) < 1e-14
    assert abs(K - K.T).max() < 1e-13 * scale and abs(M - M.T).max() < 1e-15
    cp, rv = ora.structural_pattern(S.oargs([(0, GRAD)]), S.oargs([(0, GRAD)]), (S.N, S.N))
    nzs = ora.assemble_bilinear(S.omesh, S.oargs([(0, GRAD)]), S.oargs([(0, GRAD)]), "standard", csc=(cp, rv))
    assert abs(_csc(cp, rv, nzs, S.N) - K).max() < 1e-14 * scale


@pytest.mark.parametrize("n", [2, 3, 5])
def test_structured_grid_counts(pkg, ora, n):
    """ncells = 6n^3, nnodes = (n+1)^3, nedges = 7n^3+9n^2+3n, nnz_P1 = 15n^3+21n^2+9n+1,
    nnz_P2 = 230n^3+138n^2+24n+1 (3D); 2D analogues."""
    X = np.linspace(0, 1, n + 1)
    g3 = pkg.simplexgrid(X, X, X)
    assert g3.ncells == 6 * n ** 3 and g3.nnodes == (n + 1) ** 3 and g3.nedges == 7 * n ** 3 + 9 * n ** 2 + 3 * n
    assert abs(g3.cellvolumes.sum() - 1.0) < 1e-14
    for order, nnz in ((1, 15 * n ** 3 + 21 * n ** 2 + 9 * n + 1), (2, 230 * n ** 3 + 138 * n ** 2 + 24 * n + 1)):
        S = System(pkg, ora, None, g3, [pkg.H1Pk(1, 3, order)])
        cp, rv = ora.structural_pattern(S.oargs([(0, GRAD)]), S.oargs([(0, GRAD)]), (S.N, S.N))
        assert rv.size == nnz
    g2 = pkg.simplexgrid(X, X)
    assert g2.ncells == 2 * n ** 2 and g2.nedges == 3 * n ** 2 + 2 * n
    for order, nnz in ((1, 7 * n ** 2 + 6 * n + 1), (2, 46 * n ** 2 + 16 * n + 1)):
        S = System(pkg, ora, None, g2, [pkg.H1Pk(1, 2, order)])
        cp, rv = ora.structural_pattern(S.oargs([(0, GRAD)]), S.oargs([(0, GRAD)]), (S.N, S.N))
        assert rv.size == nnz
    S = System(pkg, ora, None, g2, [pkg.H1P2(1, 2), pkg.H1P1(1)])
    cp, rv = ora.structural_pattern(S.oargs([(0, ID)]), S.oargs([(1, ID)]), (S.N, S.N))
    assert rv.size == 19 * n ** 2 + 10 * n + 1


@pytest.mark.parametrize("dim", [1, 2, 3])
@pytest.mark.parametrize("order", [0, 1, 2, 3, 4, 5, 6, 8])
def test_quadrature_exactness(dim, order):
    """Rules integrate all monomials up to their order exactly on the reference simplex (weights sum to 1,
    as the reference multiplies by |T| once: bilinear_operator.jl:892,920)."""
    from math import factorial
    x, w = fetables.quadrature_rule(dim, order)
    assert abs(w.sum() - 1.0) < 1e-14
    import itertools
    for e in itertools.product(range(order + 1), repeat=dim):
        if sum(e) > order:
            continue
        exact = factorial(dim) * np.prod([factorial(k) for k in e]) / factorial(sum(e) + dim)
        got = (w * np.prod(x ** np.array(e), axis=1)).sum()
        assert abs(got - exact) < 1e-13, (e, got, exact)


@pytest.mark.parametrize("dim,order", [(1, 1), (1, 2), (2, 1), (2, 2), (3, 1), (3, 2)])
def test_reference_basis(dim, order):
    """Lagrange property at the dof points, partition of unity, gradients by finite differences."""
    pts = [np.zeros(dim)] + [np.eye(dim)[i] for i in range(dim)]
    if order == 2:
        edges = {1: [(0, 1)], 2: [(0, 1), (1, 2), (2, 0)], 3: [(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)]}[dim]
        pts += [0.5 * (pts[a] + pts[b]) for a, b in edges]
    P = np.array(pts)
    vals, grads = fetables.ref_basis(order, P)
    assert np.allclose(vals, np.eye(len(pts)), atol=1e-15)
    rng = np.random.default_rng(0)
    Q = rng.random((5, dim)) / (dim + 1)
    v, g = fetables.ref_basis(order, Q)
    assert np.allclose(v.sum(axis=1), 1.0) and np.allclose(g.sum(axis=1), 0.0, atol=1e-14)
    h = 1e-6
    for d in range(dim):
        E = np.zeros(dim); E[d] = h
        fd = (fetables.ref_basis(order, Q + E)[0] - fetables.ref_basis(order, Q - E)[0]) / (2 * h)
        assert np.allclose(fd, g[:, :, d], atol=1e-8)


@pytest.mark.parametrize("kernel,dim,nin,params", [("nse2d", 2, 7, [0.05]), ("linnse7", 2, 7, [0.1, 2.0]),
                                                   ("neohooke3d", 3, 9, [3.8, 5.7]), ("rcd", 1, 2, [])])
def test_oracle_jacobians_against_finite_differences(ora, kernel, dim, nin, params):
    """The oracle differentiates kernels by complex step (the reference: ForwardDiff,
    nonlinear_operator.jl:358-365); both are exact to rounding.  Cross-check with central differences."""
    rng = np.random.default_rng(1)
    x = 0.1 * rng.standard_normal(nin)
    val, jac = ora.nl_value_and_jacobian(kernel, dim, x, nin, params)
    h = 1e-6
    for j in range(nin):
        e = np.zeros(nin); e[j] = h
        fd = (ora.nl_value_and_jacobian(kernel, dim, x + e, nin, params)[0]
              - ora.nl_value_and_jacobian(kernel, dim, x - e, nin, params)[0]) / (2 * h)
        assert np.allclose(fd, jac[:, j], atol=1e-7 * max(1.0, np.abs(jac).max()))
    if kernel == "neohooke3d":  # DW is the gradient of the energy W (Example330:49-57)
        for j in range(9):
            e = np.zeros(9); e[j] = h
            dW = (ora.neohooke_energy(x + e, *params) - ora.neohooke_energy(x - e, *params)) / (2 * h)
            assert abs(dW - val[j]) < 1e-7 * max(1.0, np.abs(val).max())
        assert np.allclose(jac, jac.T, atol=1e-12 * np.abs(jac).max())


def test_value_dependent_pattern_drops_exact_zeros(pkg, ora):
    """bilinear_operator.jl:925: entries with |Aloc| <= entry_tol are never inserted, so on a structured
    grid the reference pattern is a strict subset of the structural one."""
    X = np.linspace(0, 1, 4)
    grid = pkg.simplexgrid(X, X)
    S = System(pkg, ora, None, grid, [pkg.H1P1(1)])
    cp, rv, nz = ora.assemble_bilinear(S.omesh, S.oargs([(0, GRAD)]), S.oargs([(0, GRAD)]), "standard", shape=(S.N, S.N))
    scp, srv = ora.structural_pattern(S.oargs([(0, GRAD)]), S.oargs([(0, GRAD)]), (S.N, S.N))
    assert rv.size < srv.size
    from util import csc_subset
    assert csc_subset(cp, rv, scp, srv)
