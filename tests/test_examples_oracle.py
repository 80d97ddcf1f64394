"""CPU tests (no GPU): the reference's examples, restated on the host mirror of its API (tests/examples.py), run on the
CPU oracle and must reproduce the golden values the reference's own ``runtests()`` hold.  This is what
pins oracle/assembly_ref.c + oracle/fetables.py -- and the engine-convention grids of host/grids.py -- to the reference:

  Example105:80   1D P2, NonlinearOperator + LinearOperator, Newton to 6e-16           reproduced to 1e-15
  Example106:132  1D P1, porous-medium NonlinearOperator (test [grad u], args [u, grad u]), lumped mass matrix, 90 implicit Euler
                  steps; the reference's number is the interpolated exact profile         reproduced to 1e-15
  Example108:95   1D P2, NonlinearOperator, Robin BilinearOperator ON_BFACES, InterpolateBoundaryData, ItemIntegrator
                  with quadorder 4 (3-point Gauss)                                       reproduced to 1e-10
  Example201:80   2D P2 Poisson, order-2 triangle rule, penalties                       reproduced to 2e-15
  Example205:97   2D P2 stiffness + mass, 1000 backward Euler steps                     reproduced to 1e-13
  Example230:175  2D P2 vector-valued, region-dependent St.Venant-Kirchhoff kernel (order-2 triangle rule for a non-polynomial
                  integrand), masked Dirichlet data, 11 Newton steps, simplexgrid(X, Y) diagonal convention  reproduced to 5e-12
  Example301:89   3D P2 Poisson on uniform_refine(grid_unitcube, 4), 4-point tet rule for the right-hand side, ItemIntegrator with
                  quadorder 8: the reference asserts L2error <= 8.56e-5; the oracle gives 8.548e-5 (0.14 % below the bound;
                  a different refinement diagonal gives 1.16e-4 and fails)

Julia's ``≈`` is rtol = sqrt(eps) ~ 1.5e-8; the tolerances below are tighter wherever the oracle reproduces the number better."""
import numpy as np
import pytest

import examples as ex
from util import OracleBackend


@pytest.fixture()
def mk(pkg, ora):
    return lambda FES: OracleBackend(pkg, ora, FES)


def test_example105(pkg, mk):
    v, sol, st = ex.example105(pkg, mk)
    assert abs(v / ex.GOLDEN["Example105"] - 1) < 1e-12
    assert len(st["nonlinear_residuals"]) == 5 and st["nonlinear_residuals"][-1] < 1e-10


def test_example106(pkg, mk):
    """Example106:132.  The reference overwrites the solution with the interpolated Barenblatt profile before it reads the
    maximum, so the golden pins the nodal interpolation; the 90 implicit Euler steps (one Newton step each, lumped mass matrix,
    NonlinearOperator with test operators != argument operators) must land near that profile."""
    v, sol, st = ex.example106(pkg, mk)
    assert abs(v / ex.GOLDEN["Example106"] - 1) < 1e-14
    assert np.abs(st["fe"] - sol).max() < 0.05 * sol.max() and abs(st["fe"].sum() / sol.sum() - 1) < 0.02   # mass is conserved


def test_example108(pkg, mk):
    v, sol, st = ex.example108(pkg, mk)
    assert abs(v / ex.GOLDEN["Example108"] - 1) < 1e-8
    assert st["nonlinear_residuals"][-1] < 1e-10


def test_example201(pkg, mk):
    v, sol, st = ex.example201(pkg, mk)
    assert abs(v / ex.GOLDEN["Example201"] - 1) < 1e-12


def test_example205(pkg, mk):
    v, sol, st = ex.example205(pkg, mk)
    assert abs(v / ex.GOLDEN["Example205"] - 1) < 1e-10
    # nodal interpolation of the initial state (no edge moments) is NOT what the reference does
    assert abs(0.041990425206272934 / ex.GOLDEN["Example205"] - 1) > 1e-3


def test_example230(pkg, mk):
    v, sol, st = ex.example230(pkg, mk)
    assert abs(v / ex.GOLDEN["Example230"] - 1) < 1e-9
    assert len(st["nonlinear_residuals"]) == 11        # maxiterations reached, like the reference's default solve


def test_example301(pkg, mk):
    v, sol, st = ex.example301(pkg, mk, nrefs=4)
    assert v <= ex.GOLDEN["Example301"]
    assert v > 0.99 * ex.GOLDEN["Example301"]          # ... and sits right below the bound (8.548e-5)


def test_example106_final_state(pkg):
    """Example106:115-132: the returned state is interpolate!(sol, u_exact!; time = T), so the golden value is the
    Barenblatt profile at x = 0: T^(-1/(m+1)) (pins nothing on the assembly path; kept for completeness)."""
    T, m = 0.01, 2
    assert abs(T ** (-1.0 / (m + 1.0)) / 4.641588833612778 - 1) < 1e-14
