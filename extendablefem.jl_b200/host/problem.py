"""Host-side mirror of the reference's problem API for the accelerated path:
``ProblemDescription`` / ``assign_operator!`` / ``solve(PD, FES)`` (src/problemdescription.jl:70-128,
src/solvers.jl:677-806), the operator constructors (bilinear_operator.jl:56-73, linear_operator.jl:34-47,
nonlinear_operator.jl:31-48, homogeneousdata_operator.jl, interpolateboundarydata_operator.jl, item_integrator.jl:14-25)
and ``assemble_system!`` (solvers.jl:124-195).

Julia closures cannot cross the C-ABI, so ``kernel`` is the NAME of a registry kernel (``extfem_kernel_id``); an
unregistered name raises (north_star: no fallback).  The assembly itself is delegated to a *backend*: ``EngineBackend``
below drives libextfem_cuda.so (the product path).  Tests plug the CPU oracle in through the same interface to compare
whole examples; nothing in this package imports it.  The linear solve stays on the host (scipy sparse LU), like UMFPACK in
the reference (solver_config.jl:89): it is not part of the accelerated path.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import lib as _lib
from .fespace import FESpace, FEVector

__all__ = ["Unknown", "ProblemDescription", "assign_unknown", "assign_operator", "BilinearOperator", "LinearOperator",
           "NonlinearOperator", "HomogeneousBoundaryData", "InterpolateBoundaryData", "ItemIntegrator", "EngineBackend",
           "assemble_system", "solve", "evaluate", "id", "grad", "div", "symgrad", "ON_CELLS", "ON_BFACES"]

ON_CELLS, ON_BFACES = _lib.ON_CELLS, _lib.ON_BFACES


@dataclass(frozen=True)
class Unknown:
    name: str


# operator pairs (src/unknowns.jl:126-222)
def id(u): return (u, _lib.OP_ID)            # noqa: A001,E704
def grad(u): return (u, _lib.OP_GRAD)        # noqa: E704
def div(u): return (u, _lib.OP_DIV)          # noqa: E704
def symgrad(u): return (u, _lib.OP_SYMGRAD_VOIGT)   # noqa: E704


@dataclass
class Operator:
    kind: str                       # bilinear | linear | nonlinear | homogeneous_bd | interpolate_bd | integrator
    kernel: str = "standard"
    oa_test: list = field(default_factory=list)
    oa_ansatz: list = field(default_factory=list)
    oa_args: list = field(default_factory=list)
    parameters: dict = field(default_factory=dict)
    bdofs: np.ndarray | None = None     # fixed_dofs(O)
    bvals: np.ndarray | None = None


def _params(kwargs, **defaults):
    p = dict(factor=1.0, params=(), regions=(), entities=ON_CELLS, quadorder="auto", bonus_quadorder=0, name="operator",
             tabulated=None)
    p.update(defaults)
    unknown = set(kwargs) - set(p)
    if unknown:
        raise TypeError(f"unsupported keyword arguments {sorted(unknown)}")
    p.update(kwargs)
    return p


def BilinearOperator(*a, **kw) -> Operator:
    """``BilinearOperator([kernel,] oa_test[, oa_ansatz[, oa_args]]; kwargs...)`` (bilinear_operator.jl:203-260)."""
    a = list(a)
    kernel = a.pop(0) if a and isinstance(a[0], str) else "standard"
    oa_test = a[0]
    oa_ansatz = a[1] if len(a) > 1 else oa_test
    oa_args = a[2] if len(a) > 2 else []
    return Operator("bilinear", kernel, list(oa_test), list(oa_ansatz), list(oa_args),
                    _params(kw, transposed_copy=0, lump=0, store=False, name="BilinearOperator"))


def LinearOperator(*a, **kw) -> Operator:
    """``LinearOperator([kernel,] oa_test[, oa_args]; kwargs...)`` (linear_operator.jl:120-200)."""
    a = list(a)
    kernel = a.pop(0) if a and isinstance(a[0], str) else "constant_one"
    oa_test = a[0]
    oa_args = a[1] if len(a) > 1 else []
    if oa_args and kernel == "constant_one":
        kernel = "standard"
    return Operator("linear", kernel, list(oa_test), [], list(oa_args), _params(kw, store=False, name="LinearOperator"))


def NonlinearOperator(kernel: str, oa_test, oa_args=None, **kw) -> Operator:
    """``NonlinearOperator(kernel, oa_test[, oa_args]; kwargs...)`` (nonlinear_operator.jl:51-110)."""
    return Operator("nonlinear", kernel, list(oa_test), [], list(oa_test if oa_args is None else oa_args),
                    _params(kw, sparse_jacobians=True, name="NonlinearOperator"))


def HomogeneousBoundaryData(u, regions=(), mask=(), value=0.0, penalty=1e30) -> Operator:
    """``HomogeneousBoundaryData(u; regions, mask, value, penalty)`` (homogeneousdata_operator.jl:26-60)."""
    return Operator("homogeneous_bd", parameters=dict(u=u, regions=tuple(regions), mask=tuple(mask), value=float(value),
                                                      penalty=penalty, name="HomogeneousData"))


def InterpolateBoundaryData(u, data=None, regions=(), penalty=1e30) -> Operator:
    """``InterpolateBoundaryData(u, data!; regions, penalty)`` (interpolateboundarydata_operator.jl:62-66).  ``data(x)`` takes
    points [n, dim] and returns [n, ncomp]; nodal Lagrange interpolation at the boundary dofs."""
    return Operator("interpolate_bd", parameters=dict(u=u, data=data, regions=tuple(regions), penalty=penalty, name="BoundaryData"))


def ItemIntegrator(*a, **kw) -> Operator:
    """``ItemIntegrator([kernel,] oa_args; kwargs...)`` (item_integrator.jl:71-81)."""
    a = list(a)
    kernel = a.pop(0) if a and isinstance(a[0], str) else "ii_standard"
    return Operator("integrator", kernel, [], [], list(a[0]), _params(kw, resultdim=0, piecewise=True, name="ItemIntegrator"))


class ProblemDescription:
    def __init__(self, name: str = "My problem"):
        self.name = name
        self.unknowns: list[Unknown] = []
        self.operators: list[Operator] = []


def assign_unknown(PD: ProblemDescription, u: Unknown) -> int:
    if u not in PD.unknowns:
        PD.unknowns.append(u)
    return PD.unknowns.index(u)


def assign_operator(PD: ProblemDescription, op: Operator) -> int:
    PD.operators.append(op)
    return len(PD.operators)


# ------------------------------------------------------------------------------------------------------------------------
class EngineBackend:
    """Assembly through libextfem_cuda.so: one device-resident system per (grid, FESpaces)."""

    def __init__(self, engine, FES: list[FESpace]):
        self.eng, self.FES = engine, FES
        g = FES[0].xgrid
        self.mesh = engine.mesh_set(g.coords, g.cellnodes, g.cellregions, g.cellvolumes)
        engine.mesh_set_bfaces(self.mesh, g.bfacenodes, g.bfaceregions, g.bfacevolumes)
        self.spaces = [engine.fespace_set(self.mesh, F) for F in FES]
        for sp, F in zip(self.spaces, FES):
            engine.space_set_bfacedofs(sp, F.bfacedofs)
        self.pat = engine.pattern_build(self.spaces)
        self.colptr, self.rowval = engine.pattern_get(self.pat)
        self.N = sum(F.ndofs for F in FES)

    def _desc(self, op: Operator, blocks, time=0.0):
        P = op.parameters
        pairs = lambda oa: [(blocks[u], o) for u, o in oa]      # noqa: E731
        return self.eng.make_opdesc(pairs(op.oa_test), pairs(op.oa_ansatz), pairs(op.oa_args), kernel_id=_lib.kernel_id(op.kernel),
                                    params=P["params"], factor=P["factor"], time=time,
                                    quadorder=-1 if P["quadorder"] == "auto" else int(P["quadorder"]),
                                    bonus_quadorder=P["bonus_quadorder"], regions=P["regions"],
                                    transposed_copy=P.get("transposed_copy", 0), lump=P.get("lump", 0), entities=P["entities"],
                                    tabulated=P.get("tabulated"))

    def zero(self):
        self.eng.values_zero(self.pat, True, True)

    def assemble(self, op: Operator, blocks, sol, time=0.0):
        d = self._desc(op, blocks, time)
        if op.kind == "bilinear":
            self.eng.assemble_bilinear(self.pat, d, sol=sol if op.oa_args else None, accumulate=True)
        elif op.kind == "linear":
            self.eng.assemble_linear(self.pat, d, sol=sol if op.oa_args else None, accumulate=True)
        else:
            self.eng.assemble_nonlinear(self.pat, d, sol, accumulate=True)

    def penalties(self, dofs, values, penalty):
        self.eng.apply_penalties(self.pat, dofs, values, penalty)

    def system(self):
        import scipy.sparse as sp
        nz, b = self.eng.values_get(self.pat)
        return sp.csc_matrix((nz, self.rowval - 1, self.colptr - 1), shape=(self.N, self.N)), b

    def residual(self, sol):
        return self.eng.residual(self.pat, sol)

    def integrate(self, op: Operator, blocks, sol, resultdim, time=0.0):
        d = self._desc(op, blocks, time)
        return self.eng.integrate(self.pat, d, sol, resultdim=resultdim, piecewise=op.parameters["piecewise"], nitems=self.FES[0].xgrid.ncells)


def _boundary_dofs(FES: FESpace, offset: int, regions, mask=()):
    """bdofs of the boundary operators (homogeneousdata_operator.jl:100-140, interpolateboundarydata_operator.jl:104-125):
    the BFaceDofs of the faces in ``regions`` (1-based, global), components filtered by ``mask``."""
    g = FES.xgrid
    sel = np.isin(g.bfaceregions, np.asarray(regions, dtype=np.int64)) if len(regions) else np.zeros(g.bfaceregions.size, bool)
    bd = FES.bfacedofs[sel].astype(np.int64)
    nc = FES.fetype.ncomponents
    if len(mask) and nc > 1:
        per = bd.shape[1] // nc
        keep = np.concatenate([np.arange(c * per, (c + 1) * per) for c in range(nc) if mask[c]]) if any(mask) else np.zeros(0, int)
        bd = bd[:, keep]
    return np.unique(bd.ravel()) + offset


def _prepare_boundary(PD, FES, offsets, blocks):
    for op in PD.operators:
        P = op.parameters
        if op.kind == "homogeneous_bd":
            j = blocks[P["u"]]
            op.bdofs = _boundary_dofs(FES[j], int(offsets[j]), P["regions"], P["mask"])
            op.bvals = np.full(op.bdofs.size, P["value"])
        elif op.kind == "interpolate_bd":
            j = blocks[P["u"]]
            op.bdofs = _boundary_dofs(FES[j], int(offsets[j]), P["regions"])
            F = FES[j]
            loc = op.bdofs - int(offsets[j]) - 1
            pts = F.dof_coordinates()[loc % F.coffset]
            vals = np.asarray(P["data"](pts), dtype=np.float64).reshape(pts.shape[0], -1) if op.bdofs.size else np.zeros((0, 1))
            op.bvals = vals[np.arange(loc.size), loc // F.coffset] if op.bdofs.size else np.zeros(0)


def assemble_system(backend, PD: ProblemDescription, sol: np.ndarray, blocks, time=0.0):
    """``assemble_system!`` (src/solvers.jl:124-195): zero, every operator's assemble!, then every operator's apply_penalties!
    (including the assemble_sol leg: sol[bdofs] = value)."""
    backend.zero()
    for op in PD.operators:
        if op.kind in ("bilinear", "linear", "nonlinear"):
            backend.assemble(op, blocks, sol, time)
    for op in PD.operators:
        if op.kind in ("homogeneous_bd", "interpolate_bd") and op.bdofs.size:
            backend.penalties(op.bdofs, op.bvals, op.parameters["penalty"])
            sol[op.bdofs - 1] = op.bvals


def solve(PD: ProblemDescription, FES, backend=None, engine=None, init: np.ndarray | None = None, maxiterations=10,
          target_residual=1e-10, time=0.0, return_stats=False):
    """``solve(PD, FES; maxiterations, target_residual)`` (src/solvers.jl:677-806): Newton loop with full reassembly,
    residual b - A*sol with fixed dofs zeroed (:38-80), A dx = residual, sol += dx (:474-494)."""
    import scipy.sparse.linalg as spla
    FES = list(FES) if isinstance(FES, (list, tuple)) else [FES]
    if backend is None:
        backend = EngineBackend(engine, FES)
    blocks = {u: j for j, u in enumerate(PD.unknowns)}
    offsets = np.concatenate([[0], np.cumsum([F.ndofs for F in FES])])
    sol = np.zeros(int(offsets[-1])) if init is None else np.array(init, dtype=np.float64)
    _prepare_boundary(PD, FES, offsets, blocks)
    is_linear = not any(op.kind == "nonlinear" for op in PD.operators)
    maxits = 0 if is_linear else maxiterations
    fixed = [op.bdofs - 1 for op in PD.operators if op.bdofs is not None and op.bdofs.size]
    stats = dict(nonlinear_residuals=[], linear_residuals=[])
    for j in range(1, maxits + 2):
        assemble_system(backend, PD, sol, blocks, time)
        residual = backend.residual(sol)
        for f in fixed:
            residual[f] = 0.0
        nlres = float(np.linalg.norm(residual))
        if not is_linear:
            stats["nonlinear_residuals"].append(nlres)
        if nlres < target_residual or np.isnan(nlres) or (j == maxits + 1 and not is_linear):
            break
        A, _ = backend.system()
        dx = spla.spsolve(A.tocsc(), residual)
        lin = A @ dx - residual
        for f in fixed:
            lin[f] = 0.0
        stats["linear_residuals"].append(float(np.linalg.norm(lin)))
        sol = sol + dx
    vec = FEVector(FES)
    vec.entries[:] = sol
    return (vec, stats, backend) if return_stats else vec


def evaluate(O: Operator, sol: FEVector, PD: ProblemDescription | None = None, backend=None, engine=None, time=0.0):
    """``evaluate(O::ItemIntegrator, sol)`` (item_integrator.jl:323-352): [resultdim, nitems] like the reference (piecewise)
    or [resultdim]."""
    assert O.kind == "integrator"
    if backend is None:
        backend = EngineBackend(engine, sol.FES)
    blocks = {u: j for j, u in enumerate(PD.unknowns)} if PD is not None else {}
    for u, _ in O.oa_args:
        if u not in blocks:
            blocks[u] = u if isinstance(u, int) else 0
    g = sol.FES[0].xgrid
    nin = 0
    for u, o in O.oa_args:
        nc = sol.FES[blocks[u]].fetype.ncomponents
        nin += {0: nc, 1: nc * g.dim, 2: 1, 3: {1: 1, 2: 3, 3: 6}[g.dim]}[o]
    resultdim = O.parameters["resultdim"] or nin
    out = backend.integrate(O, blocks, sol.entries, resultdim, time)
    return out.T.copy() if O.parameters["piecewise"] else out
