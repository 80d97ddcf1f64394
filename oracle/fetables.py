"""oracle/fetables.py -- quadrature rules and H1 Lagrange reference bases (numpy).

TEST INFRASTRUCTURE ONLY (see oracle/assembly_ref.c header).

These tables live in ExtendableFEMBase.jl, which is NOT vendored in /root/reference
(Project.toml:33 pins ExtendableFEMBase 1.6.0).  They are restated from
the published definitions:

* quadrature (call site: ``QuadratureRule{Tv,EG}(quadorder)``,
  src/common_operators/bilinear_operator.jl:738): midpoint rule for order <= 1;
  order 2: Simpson (1D), edge-midpoint rule (triangle), 4-point rule
  a=0.1381966011250105 / b=0.5854101966249685 (tetrahedron); higher orders: Gauss-Legendre
  (1D) and Stroud conical products with div(order,2)+1 Gauss / Gauss-Jacobi points per
  direction.  Weights sum to 1 (the loops multiply by |T| separately,
  bilinear_operator.jl:892,920).
* bases: P1 = barycentric coordinates; P2 = lambda_i(2 lambda_i - 1) at the vertices and
  4 lambda_a lambda_b on the edges, local edge order tri (12,23,31), tet (12,13,14,23,24,34).

The order-2 triangle rule and the P2 basis are pinned by the reference's golden value of
Example201 (tests/test_oracle_golden.py); rules of order >= 3 are "engine conventions"
(unpinned): in the Julia drop-in the tables are passed in from ExtendableFEMBase itself.
"""
from __future__ import annotations

import numpy as np

TRI_EDGES = ((0, 1), (1, 2), (2, 0))
TET_EDGES = ((0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3))


def _gauss_legendre_01(n):
    x, w = np.polynomial.legendre.leggauss(n)
    return 0.5 * x + 0.5, 0.5 * w


def _gauss_jacobi_01(n, alpha):
    """Gauss-Jacobi nodes/weights on [0,1] for weight (1-s)^alpha, normalised to sum 1/(alpha+1)."""
    from scipy.special import roots_jacobi
    x, w = roots_jacobi(n, alpha, 0.0)
    s = 0.5 * x + 0.5
    return s, w / (2.0 ** (alpha + 1))


def quadrature_rule(dim: int, order: int):
    """Returns (xref [nq, dim], w [nq]) with sum(w) == 1."""
    if dim == 0:                 # a vertex: the boundary "face" of a 1D grid
        return np.zeros((1, 0)), np.array([1.0])
    if dim == 1:
        if order <= 1:
            return np.array([[0.5]]), np.array([1.0])
        if order == 2:
            return np.array([[0.0], [1.0], [0.5]]), np.array([1 / 6, 1 / 6, 2 / 3])
        n = order // 2 + 1
        x, w = _gauss_legendre_01(n)
        return x[:, None].copy(), w / w.sum()
    if dim == 2:
        if order <= 1:
            return np.array([[1 / 3, 1 / 3]]), np.array([1.0])
        if order == 2:
            return np.array([[0.5, 0.5], [0.0, 0.5], [0.5, 0.0]]), np.full(3, 1 / 3)
        n = order // 2 + 1
        r, a = _gauss_legendre_01(n)
        s, b = _gauss_jacobi_01(n, 1)
        pts, wts = [], []
        for j in range(n):
            for i in range(n):
                pts.append([s[j], r[i] * (1 - s[j])])
                wts.append(a[i] * b[j])
        w = np.array(wts)
        return np.array(pts), w / w.sum()
    if dim == 3:
        if order <= 1:
            return np.array([[0.25, 0.25, 0.25]]), np.array([1.0])
        if order == 2:
            a, b = 0.1381966011250105, 0.5854101966249685
            return np.array([[a, a, a], [b, a, a], [a, b, a], [a, a, b]]), np.full(4, 0.25)
        n = order // 2 + 1
        r, a = _gauss_legendre_01(n)
        s, b = _gauss_jacobi_01(n, 1)
        t, c = _gauss_jacobi_01(n, 2)
        pts, wts = [], []
        for k in range(n):
            for j in range(n):
                for i in range(n):
                    pts.append([t[k], s[j] * (1 - t[k]), r[i] * (1 - s[j]) * (1 - t[k])])
                    wts.append(a[i] * b[j] * c[k])
        w = np.array(wts)
        return np.array(pts), w / w.sum()
    raise ValueError(dim)


def _ref_basis_p3(lam, dlam):
    """Cubic Lagrange basis (H1Pk order 3; engine convention, see extendablefem.jl_b200/host/fespace.py): vertices,
    then two dofs per edge (a,b) at 1/3 and 2/3 from a towards b, then (2D) the cell bubble / (3D) one dof per face
    (faces (0,1,2),(0,1,3),(1,2,3),(0,2,3) would be needed in 3D: only dim <= 2 is provided here)."""
    nq, nv = lam.shape
    dim = nv - 1
    edges = {1: ((0, 1),), 2: TRI_EDGES}[dim]
    nb = nv + 2 * len(edges) + (1 if dim == 2 else 0)
    vals = np.zeros((nq, nb)); grads = np.zeros((nq, nb, dim))
    for i in range(nv):
        l = lam[:, i]
        vals[:, i] = 0.5 * l * (3 * l - 1) * (3 * l - 2)
        grads[:, i, :] = (0.5 * (27 * l * l - 18 * l + 2))[:, None] * dlam[i][None, :]
    k = nv
    for (a, b) in edges:
        la, lb = lam[:, a], lam[:, b]
        for (p, q, dp, dq) in ((la, lb, dlam[a], dlam[b]), (lb, la, dlam[b], dlam[a])):
            # 9/2 p q (3p - 1): equals 1 at p = 2/3, q = 1/3
            vals[:, k] = 4.5 * p * q * (3 * p - 1)
            grads[:, k, :] = 4.5 * ((q * (6 * p - 1))[:, None] * dp[None, :] + (p * (3 * p - 1))[:, None] * dq[None, :])
            k += 1
    if dim == 2:
        vals[:, k] = 27 * lam[:, 0] * lam[:, 1] * lam[:, 2]
        grads[:, k, :] = 27 * ((lam[:, 1] * lam[:, 2])[:, None] * dlam[0][None, :] + (lam[:, 0] * lam[:, 2])[:, None] * dlam[1][None, :]
                               + (lam[:, 0] * lam[:, 1])[:, None] * dlam[2][None, :])
    return vals, grads


def barycentric(xref: np.ndarray):
    """lambda [nq, dim+1] and constant reference gradients dlam [dim+1, dim]."""
    nq, dim = xref.shape
    lam = np.concatenate([1.0 - xref.sum(axis=1, keepdims=True), xref], axis=1)
    dlam = np.concatenate([-np.ones((1, dim)), np.eye(dim)], axis=0)
    return lam, dlam


def ref_basis(order: int, xref: np.ndarray):
    """Scalar H1 Lagrange basis on the reference simplex.
    Returns vals [nq, nb], grads [nq, nb, dim]."""
    nq, dim = xref.shape
    if dim == 0:
        return np.ones((nq, 1)), np.zeros((nq, 1, 0))
    lam, dlam = barycentric(xref)
    if order == 3:
        return _ref_basis_p3(lam, dlam)
    if order == 1:
        vals = lam.copy()
        grads = np.broadcast_to(dlam[None], (nq, dim + 1, dim)).copy()
        return vals, grads
    if order == 2:
        edges = {1: (), 2: TRI_EDGES, 3: TET_EDGES}[dim]
        if dim == 1:
            edges = ((0, 1),)       # the cell itself carries the bubble-type midpoint dof
        nb = dim + 1 + len(edges)
        vals = np.zeros((nq, nb))
        grads = np.zeros((nq, nb, dim))
        for i in range(dim + 1):
            vals[:, i] = lam[:, i] * (2 * lam[:, i] - 1)
            grads[:, i, :] = (4 * lam[:, i] - 1)[:, None] * dlam[i][None, :]
        for e, (a, b) in enumerate(edges):
            vals[:, dim + 1 + e] = 4 * lam[:, a] * lam[:, b]
            grads[:, dim + 1 + e, :] = 4 * (lam[:, a][:, None] * dlam[b][None, :] + lam[:, b][:, None] * dlam[a][None, :])
        return vals, grads
    raise NotImplementedError(order)
