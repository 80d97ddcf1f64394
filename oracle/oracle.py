"""oracle/oracle.py -- Python driver of the CPU restatement (ctypes over assembly_ref.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package never imports it.

Parity status: partially pinned (see assembly_ref.c header and DESIGN.md).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

from . import fetables

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle_ref.so")
_lib = None

OP_ID, OP_GRAD, OP_DIV, OP_SYMGRAD_VOIGT = 0, 1, 2, 3
BLK = dict(standard=1, dcr=2, stokes=3, linnse7=4, hooke_grad=5, hooke_voigt=6, convect_args=7, robin108=8)
LIN = dict(constant_one=1, constant_params=2, xy=3, sincos301=4, tabulated=5, exp2x=6, step105=7)
NL = dict(nse2d=1, linnse7=2, neohooke3d=3, rcd=4, nlpoisson105=5, stvenant230=6, porous106=7)
II = dict(ii_standard=1, l2norm=2, l2diff_tabulated=3, l2err_sincos301=4, l2err_exp108=5)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "assembly_ref.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-std=c99", "-fPIC", "-shared", "-o", _LIB_PATH, src, "-lm"])
    return _LIB_PATH


class _Mesh(C.Structure):
    _fields_ = [("dim", C.c_int), ("ncells", C.c_int64), ("nnodes", C.c_int64),
                ("coords", C.c_void_p), ("cellnodes", C.c_void_p),
                ("cellregions", C.c_void_p), ("cellvolumes", C.c_void_p), ("tdim", C.c_int)]


class _Arg(C.Structure):
    _fields_ = [("ncomp", C.c_int), ("nscalar", C.c_int), ("op", C.c_int), ("offdiag", C.c_double),
                ("celldofs", C.c_void_p), ("offset", C.c_int64),
                ("refvals", C.c_void_p), ("refgrads", C.c_void_p)]


class _Coo(C.Structure):
    _fields_ = [("I", C.c_void_p), ("J", C.c_void_p), ("V", C.c_void_p), ("n", C.c_int64), ("cap", C.c_int64)]


class _Csc(C.Structure):
    _fields_ = [("colptr", C.c_void_p), ("rowval", C.c_void_p), ("nzval", C.c_void_p)]


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.ora_coo_segment_sum.restype = C.c_int64
        _lib.ora_neohooke_energy.restype = C.c_double
        _lib.ora_neohooke_energy.argtypes = [C.c_void_p, C.c_double, C.c_double]
    return _lib


class abs_accumulate:
    """Context manager: inside it every assembly accumulates |contribution| instead of the contribution, i.e. returns the
    entrywise scale of a backward-error comparison (tests/util.py: check_values_entrywise)."""

    def __enter__(self):
        lib().ora_set_abs_accumulate(1)

    def __exit__(self, *exc):
        lib().ora_set_abs_accumulate(0)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


@dataclass
class OraArg:
    """One (FESpace, operator) pair -- the oracle's FEEvaluator."""
    celldofs: np.ndarray   # [ncells, ndofs4cell] int32 1-based (block-local)
    ncomp: int
    order: int
    op: int
    offset: int = 0        # block offset in the global system
    offdiag: float = 1.0


class Mesh:
    """Items of one assembly type: the cells of a grid, or (cellnodes = BFaceNodes etc.) its boundary faces,
    whose topological dimension ``tdim`` is one less than the space dimension (bilinear_operator.jl:693-714)."""

    def __init__(self, coords, cellnodes, cellregions=None, cellvolumes=None):
        self.coords = np.ascontiguousarray(coords, dtype=np.float64)
        self.cellnodes = np.ascontiguousarray(cellnodes, dtype=np.int32)
        nc = self.cellnodes.shape[0]
        self.dim = self.coords.shape[1]
        self.tdim = self.cellnodes.shape[1] - 1
        assert self.tdim in (self.dim, self.dim - 1)
        self.cellregions = np.ascontiguousarray(
            np.ones(nc, np.int32) if cellregions is None else cellregions, dtype=np.int32)
        if cellvolumes is None:
            import math
            x = self.coords[self.cellnodes.astype(np.int64) - 1]
            e = x[:, 1:, :] - x[:, :1, :]
            if self.tdim == self.dim:
                cellvolumes = np.abs(np.linalg.det(e)) / math.factorial(self.dim)
            elif self.tdim == 0:
                cellvolumes = np.ones(nc)
            else:
                gram = np.einsum("nik,njk->nij", e, e)
                cellvolumes = np.sqrt(np.abs(np.linalg.det(gram))) / math.factorial(self.tdim)
        self.cellvolumes = np.ascontiguousarray(cellvolumes, dtype=np.float64)
        self.ncells = nc
        self.c = _Mesh(self.dim, nc, self.coords.shape[0], _ptr(self.coords), _ptr(self.cellnodes),
                       _ptr(self.cellregions), _ptr(self.cellvolumes), self.tdim)


def _make_args(args, xref, mesh=None):
    keep, arr = [], (_Arg * max(1, len(args)))()
    for i, a in enumerate(args):
        if mesh is not None and mesh.tdim < mesh.dim:
            assert a.op == OP_ID, "boundary faces: Identity operators only"
        vals, grads = fetables.ref_basis(a.order, xref)
        vals = np.ascontiguousarray(vals); grads = np.ascontiguousarray(grads)
        cd = np.ascontiguousarray(a.celldofs, dtype=np.int32)
        keep += [vals, grads, cd]
        arr[i] = _Arg(a.ncomp, vals.shape[1], a.op, a.offdiag, _ptr(cd), a.offset, _ptr(vals), _ptr(grads))
    return arr, keep


def oplen(a: OraArg, dim: int) -> int:
    return {OP_ID: a.ncomp, OP_GRAD: a.ncomp * dim, OP_DIV: 1,
            OP_SYMGRAD_VOIGT: {1: 1, 2: 3, 3: 6}[dim]}[a.op]


def polyorder(a: OraArg) -> int:
    return a.order - (0 if a.op == OP_ID else 1)


def coo_to_csc(I, J, V, shape):
    """Sequential (insertion-order) accumulation == repeated rawupdateindex!(A,+,...) + flush!.
    Returns (colptr, rowval, nzval), Int64 1-based, rows sorted per column."""
    n = I.size
    order = np.lexsort((I, J))          # stable: keeps insertion order within equal (J,I)
    I2, J2, V2 = (np.ascontiguousarray(x[order]) for x in (I, J, V))
    oi = np.empty(n, np.int64); oj = np.empty(n, np.int64); ov = np.empty(n, np.float64)
    m = lib().ora_coo_segment_sum(C.c_int64(n), _ptr(I2), _ptr(J2), _ptr(V2), _ptr(oi), _ptr(oj), _ptr(ov))
    oi, oj, ov = oi[:m], oj[:m], ov[:m]
    counts = np.bincount(oj - 1, minlength=shape[1])
    colptr = np.concatenate([[1], 1 + np.cumsum(counts)]).astype(np.int64)
    return colptr, oi.copy(), ov.copy()


def _regions(regions):
    r = np.ascontiguousarray(np.asarray(regions if regions is not None else [], dtype=np.int32))
    return r, r.size


def assemble_bilinear(mesh: Mesh, test, ansatz, kernel="standard", params=(), factor=1.0, quadorder="auto",
                      bonus_quadorder=0, regions=None, transposed_copy=0, lump=0, entry_tol=0.0, coupling=None,
                      args=(), sol=None, args_sol_offsets=(), time=0.0, shape=None, csc=None):
    """BilinearOperator assembly.  Returns CSC (colptr,rowval,nzval) with the reference's
    value-dependent pattern (entries with |Aloc| <= entry_tol are never inserted), or, when
    ``csc=(colptr,rowval)`` is given, the nzval accumulated into that pattern."""
    dim = mesh.dim
    if quadorder == "auto":
        quadorder = max(polyorder(a) for a in ansatz) + max(polyorder(a) for a in test)
    xref, w = fetables.quadrature_rule(mesh.tdim, quadorder + bonus_quadorder)
    xref = np.ascontiguousarray(xref); w = np.ascontiguousarray(w)
    ta, k1 = _make_args(test, xref, mesh); aa, k2 = _make_args(ansatz, xref, mesh); ga, k3 = _make_args(args, xref, mesh)
    if coupling is None:
        coupling = np.ones((len(ansatz), len(test)), np.uint8)
    coupling = np.ascontiguousarray(coupling, dtype=np.uint8)
    p = np.ascontiguousarray(np.asarray(params, dtype=np.float64))
    r, nr = _regions(regions)
    so = np.ascontiguousarray(np.asarray(args_sol_offsets, dtype=np.int64))
    solp = None if sol is None else np.ascontiguousarray(sol, dtype=np.float64)
    nloc = sum(t.celldofs.shape[1] for t in test) * sum(a.celldofs.shape[1] for a in ansatz)
    cap = mesh.ncells * nloc * (2 if transposed_copy else 1)
    common = (C.byref(mesh.c), len(test), ta, len(ansatz), aa, len(args), ga, _ptr(solp), _ptr(so), _ptr(coupling),
              w.size, _ptr(w), _ptr(xref), BLK[kernel], _ptr(p), p.size, C.c_double(factor), C.c_double(time),
              _ptr(r), nr, transposed_copy, lump, C.c_double(entry_tol))
    if csc is not None:
        colptr, rowval = csc
        nz = np.zeros(rowval.size)
        s = _Csc(_ptr(colptr), _ptr(rowval), _ptr(nz))
        rc = lib().ora_assemble_bilinear(*common, None, C.byref(s))
        assert rc == 0, rc
        return nz
    I = np.empty(cap, np.int64); J = np.empty(cap, np.int64); V = np.empty(cap, np.float64)
    coo = _Coo(_ptr(I), _ptr(J), _ptr(V), 0, cap)
    rc = lib().ora_assemble_bilinear(*common, C.byref(coo), None)
    assert rc == 0, rc
    n = coo.n
    return coo_to_csc(I[:n], J[:n], V[:n], shape)


def assemble_linear(mesh: Mesh, test, b, kernel="constant_one", params=(), factor=1.0, quadorder="auto",
                    bonus_quadorder=0, regions=None, args=(), sol=None, args_sol_offsets=(), time=0.0,
                    tabulated=None):
    dim = mesh.dim
    if quadorder == "auto":
        quadorder = max(polyorder(a) for a in test) + (max(polyorder(a) for a in args) if args else 0)
    xref, w = fetables.quadrature_rule(mesh.tdim, quadorder + bonus_quadorder)
    xref = np.ascontiguousarray(xref); w = np.ascontiguousarray(w)
    ta, k1 = _make_args(test, xref, mesh); ga, k3 = _make_args(args, xref, mesh)
    p = np.ascontiguousarray(np.asarray(params, dtype=np.float64))
    r, nr = _regions(regions)
    so = np.ascontiguousarray(np.asarray(args_sol_offsets, dtype=np.int64))
    solp = None if sol is None else np.ascontiguousarray(sol, dtype=np.float64)
    tab = None if tabulated is None else np.ascontiguousarray(tabulated, dtype=np.float64)
    kid = BLK[kernel] if args else LIN[kernel]
    rc = lib().ora_assemble_linear(C.byref(mesh.c), len(test), ta, len(args), ga, _ptr(solp), _ptr(so), w.size,
                                   _ptr(w), _ptr(xref), kid, _ptr(p), p.size, C.c_double(factor), C.c_double(time),
                                   _ptr(r), nr, _ptr(tab), _ptr(b))
    assert rc == 0, rc
    return b


def assemble_nonlinear(mesh: Mesh, test, args, sol, b, kernel, params=(), factor=1.0, quadorder="auto",
                       bonus_quadorder=0, regions=None, entry_tol=0.0, args_sol_offsets=None, time=0.0,
                       shape=None, csc=None):
    dim = mesh.dim
    if quadorder == "auto":
        quadorder = max(polyorder(a) for a in args) + max(polyorder(a) for a in test)
    xref, w = fetables.quadrature_rule(mesh.tdim, quadorder + bonus_quadorder)
    xref = np.ascontiguousarray(xref); w = np.ascontiguousarray(w)
    ta, k1 = _make_args(test, xref, mesh); ga, k3 = _make_args(args, xref, mesh)
    p = np.ascontiguousarray(np.asarray(params, dtype=np.float64))
    r, nr = _regions(regions)
    if args_sol_offsets is None:
        args_sol_offsets = [a.offset for a in args]
    so = np.ascontiguousarray(np.asarray(args_sol_offsets, dtype=np.int64))
    solp = np.ascontiguousarray(sol, dtype=np.float64)
    nloc = sum(t.celldofs.shape[1] for t in test) * sum(a.celldofs.shape[1] for a in args)
    common = (C.byref(mesh.c), len(test), ta, len(args), ga, _ptr(solp), _ptr(so), w.size, _ptr(w), _ptr(xref),
              NL[kernel], _ptr(p), p.size, C.c_double(factor), C.c_double(time), _ptr(r), nr, C.c_double(entry_tol))
    if csc is not None:
        colptr, rowval = csc
        nz = np.zeros(rowval.size)
        s = _Csc(_ptr(colptr), _ptr(rowval), _ptr(nz))
        rc = lib().ora_assemble_nonlinear(*common, None, C.byref(s), _ptr(b))
        assert rc == 0, rc
        return nz, b
    cap = mesh.ncells * nloc
    I = np.empty(cap, np.int64); J = np.empty(cap, np.int64); V = np.empty(cap, np.float64)
    coo = _Coo(_ptr(I), _ptr(J), _ptr(V), 0, cap)
    rc = lib().ora_assemble_nonlinear(*common, C.byref(coo), None, _ptr(b))
    assert rc == 0, rc
    n = coo.n
    return coo_to_csc(I[:n], J[:n], V[:n], shape), b


def integrate(mesh: Mesh, args, sol, kernel="ii_standard", params=(), factor=1.0, quadorder="auto", bonus_quadorder=0,
              regions=None, resultdim=0, args_sol_offsets=None, time=0.0, tabulated=None, piecewise=True):
    """ItemIntegrator ``evaluate`` (item_integrator.jl:323-352): [ncells, resultdim] (piecewise) or [resultdim]."""
    if quadorder == "auto":
        quadorder = max(polyorder(a) for a in args)
    xref, w = fetables.quadrature_rule(mesh.tdim, quadorder + bonus_quadorder)
    xref = np.ascontiguousarray(xref); w = np.ascontiguousarray(w)
    ga, k3 = _make_args(args, xref, mesh)
    nin = sum(oplen(a, mesh.dim) for a in args)
    if resultdim == 0:
        resultdim = nin
    p = np.ascontiguousarray(np.asarray(params, dtype=np.float64))
    r, nr = _regions(regions)
    if args_sol_offsets is None:
        args_sol_offsets = [a.offset for a in args]
    so = np.ascontiguousarray(np.asarray(args_sol_offsets, dtype=np.int64))
    solp = np.ascontiguousarray(sol, dtype=np.float64)
    tab = None if tabulated is None else np.ascontiguousarray(tabulated, dtype=np.float64)
    b = np.zeros((mesh.ncells, resultdim))
    rc = lib().ora_integrate(C.byref(mesh.c), len(args), ga, _ptr(solp), _ptr(so), w.size, _ptr(w), _ptr(xref), II[kernel],
                             _ptr(p), p.size, C.c_double(factor), C.c_double(time), _ptr(r), nr, _ptr(tab), resultdim, _ptr(b))
    assert rc == 0, rc
    if piecewise:
        return b
    out = np.zeros(resultdim)
    for row in b:                      # b .+= result_kernel in item order (item_integrator.jl:241-243)
        out += row
    return out


def quadrature_points(mesh: Mesh, quadorder):
    """x at the quadrature points of every item: [nitems, nq, dim] (what a host evaluates tabulated closures at)."""
    xref, w = fetables.quadrature_rule(mesh.tdim, quadorder)
    lam, _ = fetables.barycentric(xref) if mesh.tdim > 0 else (np.ones((1, 1)), None)
    x = mesh.coords[mesh.cellnodes.astype(np.int64) - 1]          # [n, tdim+1, dim]
    return np.einsum("qv,nvd->nqd", lam, x)


def nl_value_and_jacobian(kernel, dim, x, nout, params=()):
    x = np.ascontiguousarray(x, dtype=np.float64)
    p = np.ascontiguousarray(np.asarray(params, dtype=np.float64))
    val = np.zeros(nout); jac = np.zeros((nout, x.size))
    rc = lib().ora_nl_value_and_jacobian(NL[kernel], dim, x.size, nout, _ptr(x), _ptr(p), p.size, _ptr(val), _ptr(jac))
    assert rc == 0
    return val, jac


def neohooke_energy(gradu, mu, la):
    g = np.ascontiguousarray(gradu, dtype=np.float64)
    return lib().ora_neohooke_energy(_ptr(g), mu, la)


def structural_pattern(test, ansatz, shape, coupling=None):
    """Structural CSC pattern: dofs that share a cell (filtered by the coupling matrix)."""
    import scipy.sparse as sp
    rows, cols = [], []
    for ia, a in enumerate(ansatz):
        for it, t in enumerate(test):
            if coupling is not None and not coupling[ia][it]:
                continue
            tj = t.celldofs.astype(np.int64) + t.offset
            ak = a.celldofs.astype(np.int64) + a.offset
            rows.append(np.repeat(tj, ak.shape[1], axis=1).ravel())
            cols.append(np.tile(ak, (1, tj.shape[1])).ravel())
    r = np.concatenate(rows) - 1
    c = np.concatenate(cols) - 1
    M = sp.csc_matrix((np.ones(r.size, np.int8), (r, c)), shape=shape)
    M.sum_duplicates(); M.sort_indices()
    return M.indptr.astype(np.int64) + 1, M.indices.astype(np.int64) + 1
