"""ctypes binding of libextfem_cuda.so (include/extfem_cuda.h).

This is the Python twin of the ``ccall`` stubs in julia/ExtFEMCuda.jl.  There is NO CPU
fallback: if the shared library is missing or no CUDA device is present, loading / context
creation raises."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("EXTFEM_LIB", os.path.join(os.path.dirname(_HERE), "csrc", "libextfem_cuda.so"))

MAXARGS = 4
OP_ID, OP_GRAD, OP_DIV, OP_SYMGRAD_VOIGT = 0, 1, 2, 3
ON_CELLS, ON_BFACES = 0, 1
FE_TABULATED = 100

EXPORTS = [
    "extfem_ctx_create", "extfem_ctx_destroy", "extfem_last_error", "extfem_kernel_id", "extfem_synchronize", "extfem_set_option",
    "extfem_launch_count", "extfem_last_timings", "extfem_event_record", "extfem_event_elapsed_ms", "extfem_mesh_set", "extfem_mesh_update_coords",
    "extfem_space_set", "extfem_space_set_tables", "extfem_pattern_build", "extfem_pattern_dims",
    "extfem_pattern_get", "extfem_assemble_bilinear", "extfem_assemble_linear", "extfem_assemble_nonlinear",
    "extfem_quadrature_points", "extfem_values_get", "extfem_values_set", "extfem_device_ptrs",
    "extfem_apply_penalties", "extfem_residual", "extfem_spmv", "extfem_cg", "extfem_plan_stats", "extfem_plan_jit_status",
    "extfem_dist_unique_id", "extfem_dist_init", "extfem_dist_set_interfaces", "extfem_dist_sum_rhs", "extfem_dist_spmv", "extfem_dist_cg",
    "extfem_mesh_set_bfaces", "extfem_space_set_bfacedofs", "extfem_integrate", "extfem_values_zero", "extfem_apply_values",
    "extfem_dist_set_owned", "extfem_dist_reduce_system", "extfem_dist_spmv_owned", "extfem_dist_cg_owned",
    "extfem_pattern_get_lower", "extfem_values_get_lower",
]


class ExtFEMError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


class OpDesc(C.Structure):
    _fields_ = [
        ("ntest", C.c_int32), ("test_block", C.c_int32 * MAXARGS), ("test_op", C.c_int32 * MAXARGS),
        ("nansatz", C.c_int32), ("ansatz_block", C.c_int32 * MAXARGS), ("ansatz_op", C.c_int32 * MAXARGS),
        ("nargs", C.c_int32), ("args_block", C.c_int32 * MAXARGS), ("args_op", C.c_int32 * MAXARGS),
        ("kernel_id", C.c_int32), ("nparams", C.c_int32), ("params", C.c_void_p),
        ("factor", C.c_double), ("time", C.c_double), ("symgrad_offdiag", C.c_double),
        ("quadorder", C.c_int32), ("bonus_quadorder", C.c_int32),
        ("nregions", C.c_int32), ("regions", C.c_void_p),
        ("transposed_copy", C.c_int32), ("lump", C.c_int32), ("coupling", C.c_void_p),
        ("nq_custom", C.c_int32), ("qweights", C.c_void_p), ("qpoints", C.c_void_p), ("tabulated", C.c_void_p),
        ("entities", C.c_int32),
    ]


_lib = None


def load_library():
    """Loads libextfem_cuda.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run __graft_entry__.build() (there is no CPU fallback)")
        _lib = C.CDLL(LIB_PATH)
        _lib.extfem_last_error.restype = C.c_char_p
        _lib.extfem_last_error.argtypes = [C.c_void_p]
        _lib.extfem_launch_count.restype = C.c_int64
        _lib.extfem_launch_count.argtypes = [C.c_void_p]
        _lib.extfem_kernel_id.argtypes = [C.c_char_p]
    return _lib


def kernel_id(name: str) -> int:
    kid = load_library().extfem_kernel_id(name.encode())
    if kid < 0:
        raise ExtFEMError(kid, load_library().extfem_last_error(None).decode())
    return kid


def _p(a):
    """pointer of a numpy array / torch tensor / int address / None"""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    raise TypeError(type(a))


class Engine:
    """One extfem context (one GPU)."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        self.ctx = C.c_void_p()
        rc = self.lib.extfem_ctx_create(int(device), C.byref(self.ctx))
        if rc != 0:
            raise ExtFEMError(rc, self.lib.extfem_last_error(None).decode())
        self._keep = []

    def close(self):
        if self.ctx:
            self.lib.extfem_ctx_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise ExtFEMError(rc, self.lib.extfem_last_error(self.ctx).decode())

    # ---- grid / spaces / pattern ------------------------------------------------------------
    def mesh_set(self, coords, cellnodes, cellregions=None, cellvolumes=None) -> int:
        coords = np.ascontiguousarray(coords, dtype=np.float64)
        cellnodes = np.ascontiguousarray(cellnodes)
        assert cellnodes.dtype in (np.int32, np.int64)
        reg = None if cellregions is None else np.ascontiguousarray(cellregions, dtype=np.int32)
        vol = None if cellvolumes is None else np.ascontiguousarray(cellvolumes, dtype=np.float64)
        out = C.c_int()
        self._check(self.lib.extfem_mesh_set(self.ctx, coords.shape[1], C.c_int64(cellnodes.shape[0]),
                                             C.c_int64(coords.shape[0]), _p(coords), _p(cellnodes),
                                             cellnodes.dtype.itemsize, _p(reg), _p(vol), C.byref(out)))
        return out.value

    def mesh_set_bfaces(self, mesh: int, bfacenodes, bfaceregions=None, bfacevolumes=None):
        """xgrid[BFaceNodes], xgrid[BFaceRegions], xgrid[BFaceVolumes]"""
        bn = np.ascontiguousarray(bfacenodes)
        assert bn.dtype in (np.int32, np.int64)
        reg = None if bfaceregions is None else np.ascontiguousarray(bfaceregions, dtype=np.int32)
        vol = None if bfacevolumes is None else np.ascontiguousarray(bfacevolumes, dtype=np.float64)
        self._check(self.lib.extfem_mesh_set_bfaces(self.ctx, mesh, C.c_int64(bn.shape[0]), _p(bn), bn.dtype.itemsize, _p(reg), _p(vol)))

    def mesh_update_coords(self, mesh: int, coords, cellvolumes=None):
        self._check(self.lib.extfem_mesh_update_coords(self.ctx, mesh, _p(coords), _p(cellvolumes)))

    def space_set(self, mesh: int, fetype: int, ncomp: int, celldofs, ndofs: int) -> int:
        celldofs = np.ascontiguousarray(celldofs)
        assert celldofs.dtype in (np.int32, np.int64)
        out = C.c_int()
        self._check(self.lib.extfem_space_set(self.ctx, mesh, fetype, ncomp, _p(celldofs), celldofs.dtype.itemsize,
                                              celldofs.shape[1], C.c_int64(ndofs), C.byref(out)))
        return out.value

    def space_set_tables(self, space: int, order: int, coeffs, bface_coeffs=None):
        """EXTFEM_FE_TABULATED: polynomial reference basis coeffs[nscalar][nmono] (+ its restriction to a boundary face)."""
        cf = np.ascontiguousarray(coeffs, dtype=np.float64)
        bf = None if bface_coeffs is None else np.ascontiguousarray(bface_coeffs, dtype=np.float64)
        self._check(self.lib.extfem_space_set_tables(self.ctx, space, int(order), cf.shape[0], _p(cf),
                                                     0 if bf is None else bf.shape[0], _p(bf)))

    def space_set_bfacedofs(self, space: int, bfacedofs):
        """FES[BFaceDofs]"""
        bd = np.ascontiguousarray(bfacedofs)
        assert bd.dtype in (np.int32, np.int64)
        self._check(self.lib.extfem_space_set_bfacedofs(self.ctx, space, _p(bd), bd.dtype.itemsize, bd.shape[1]))

    def fespace_set(self, mesh: int, FES) -> int:
        """Uploads a host FESpace: cell dof map, boundary-face dof map (when the grid's bfaces were set) and, for elements
        that are not built in, the polynomial reference basis."""
        sp = self.space_set(mesh, FES.fetype.fe_id, FES.fetype.ncomponents, FES.celldofs, FES.ndofs)
        if FES.fetype.fe_id == FE_TABULATED:
            self.space_set_tables(sp, FES.fetype.order, FES.ref_coeffs, FES.ref_coeffs_bface)
        return sp

    def pattern_build(self, rowspaces, colspaces=None, block_coupling=None) -> int:
        colspaces = rowspaces if colspaces is None else colspaces
        rs = (C.c_int * len(rowspaces))(*rowspaces)
        cs = (C.c_int * len(colspaces))(*colspaces)
        bc = None if block_coupling is None else np.ascontiguousarray(block_coupling, dtype=np.uint8)
        out = C.c_int()
        self._check(self.lib.extfem_pattern_build(self.ctx, len(rowspaces), rs, len(colspaces), cs, _p(bc), C.byref(out)))
        return out.value

    def pattern_dims(self, pattern: int):
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        self._check(self.lib.extfem_pattern_dims(self.ctx, pattern, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def pattern_get(self, pattern: int):
        nrows, ncols, nnz = self.pattern_dims(pattern)
        colptr = np.empty(ncols + 1, np.int64)
        rowval = np.empty(nnz, np.int64)
        self._check(self.lib.extfem_pattern_get(self.ctx, pattern, _p(colptr), _p(rowval)))
        return colptr, rowval

    def pattern_colptr(self, pattern: int):
        """colptr only (Int64, 1-based): the row indices of a 10M-cell pattern are 3 GB"""
        _, ncols, _ = self.pattern_dims(pattern)
        colptr = np.empty(ncols + 1, np.int64)
        self._check(self.lib.extfem_pattern_get(self.ctx, pattern, _p(colptr), None))
        return colptr

    # ---- operators -------------------------------------------------------------------------------
    def make_opdesc(self, test, ansatz=(), args=(), kernel_id=1, params=(), factor=1.0, time=0.0, quadorder=-1,
                    bonus_quadorder=0, regions=(), transposed_copy=0, lump=0, coupling=None, offdiag=1.0,
                    qweights=None, qpoints=None, tabulated=None, entities=ON_CELLS):
        """test/ansatz/args: sequences of (block, op)."""
        d = OpDesc()
        keep = []
        for name, lst in (("test", test), ("ansatz", ansatz), ("args", args)):
            setattr(d, "n" + name, len(lst))
            for i, (blk, op) in enumerate(lst):
                getattr(d, name + "_block")[i] = blk
                getattr(d, name + "_op")[i] = op
        d.kernel_id = kernel_id
        p = np.ascontiguousarray(np.asarray(params, dtype=np.float64).ravel())
        r = np.ascontiguousarray(np.asarray(regions, dtype=np.int32).ravel())
        keep += [p, r]
        d.nparams, d.params = p.size, _p(p)
        d.factor, d.time, d.symgrad_offdiag = float(factor), float(time), float(offdiag)
        d.quadorder, d.bonus_quadorder = int(quadorder), int(bonus_quadorder)
        d.nregions, d.regions = r.size, _p(r)
        d.transposed_copy, d.lump = int(transposed_copy), int(lump)
        if coupling is not None:
            cp = np.ascontiguousarray(coupling, dtype=np.uint8)
            keep.append(cp)
            d.coupling = _p(cp)
        if qweights is not None:
            qw = np.ascontiguousarray(qweights, dtype=np.float64)
            qx = np.ascontiguousarray(qpoints, dtype=np.float64)
            keep += [qw, qx]
            d.nq_custom, d.qweights, d.qpoints = qw.size, _p(qw), _p(qx)
        if tabulated is not None:
            if isinstance(tabulated, np.ndarray):
                tabulated = np.ascontiguousarray(tabulated, dtype=np.float64)
            keep.append(tabulated)
            d.tabulated = _p(tabulated)
        d.entities = int(entities)
        d._keep = keep
        return d

    def assemble_bilinear(self, pattern, desc, sol=None, accumulate=False, nzval_out=None):
        self._check(self.lib.extfem_assemble_bilinear(self.ctx, pattern, C.byref(desc), _p(sol), int(accumulate), _p(nzval_out)))

    def assemble_linear(self, pattern, desc, sol=None, accumulate=False, b_out=None):
        self._check(self.lib.extfem_assemble_linear(self.ctx, pattern, C.byref(desc), _p(sol), int(accumulate), _p(b_out)))

    def assemble_nonlinear(self, pattern, desc, sol, accumulate=False, nzval_out=None, b_out=None):
        self._check(self.lib.extfem_assemble_nonlinear(self.ctx, pattern, C.byref(desc), _p(sol), int(accumulate),
                                                       _p(nzval_out), _p(b_out)))

    def quadrature_points(self, pattern, desc, is_linear=True):
        nq = C.c_int()
        self._check(self.lib.extfem_quadrature_points(self.ctx, pattern, C.byref(desc), int(is_linear), C.byref(nq), None))
        return nq.value

    def quadrature_points_x(self, pattern, desc, ncells, dim, is_linear=True):
        nq = self.quadrature_points(pattern, desc, is_linear)
        xq = np.empty((ncells, nq, dim))
        self._check(self.lib.extfem_quadrature_points(self.ctx, pattern, C.byref(desc), int(is_linear), C.byref(C.c_int()), _p(xq)))
        return xq

    def integrate(self, pattern, desc, sol, resultdim=0, piecewise=True, nitems=None):
        """ItemIntegrator ``evaluate``: [nitems, resultdim] (piecewise) or [resultdim]."""
        sol = np.ascontiguousarray(sol, dtype=np.float64)
        if resultdim <= 0:
            raise ValueError("resultdim must be given (the length of the arguments when the reference says 0)")
        out = np.empty((nitems, resultdim)) if piecewise else np.empty(resultdim)
        self._check(self.lib.extfem_integrate(self.ctx, pattern, C.byref(desc), _p(sol), int(resultdim), int(piecewise), _p(out)))
        return out

    # ---- device-resident system ----------------------------------------------------------------
    def values_zero(self, pattern, matrix=True, rhs=True):
        """fill!(nzval, 0) / fill!(b, 0) at the start of assemble_system! (src/solvers.jl:130-135)"""
        self._check(self.lib.extfem_values_zero(self.ctx, pattern, int(matrix), int(rhs)))

    def apply_values(self, dofs, values, sol):
        """sol[dofs] = values (the assemble_sol leg of apply_penalties!); sol is modified in place"""
        dofs = np.ascontiguousarray(dofs, dtype=np.int64)
        vals = None if values is None else np.ascontiguousarray(values, dtype=np.float64)
        assert isinstance(sol, np.ndarray) and sol.dtype == np.float64 and sol.flags.c_contiguous
        self._check(self.lib.extfem_apply_values(self.ctx, C.c_int64(dofs.size), _p(dofs), _p(vals), _p(sol), C.c_int64(sol.size)))

    def values_get(self, pattern, want_nzval=True, want_b=True, nzval_out=None, b_out=None):
        """Device-resident values; ``*_out`` (numpy array / pinned torch tensor / device pointer) receive them in place."""
        nrows, ncols, nnz = self.pattern_dims(pattern)
        nz = nzval_out if nzval_out is not None else (np.empty(nnz) if want_nzval else None)
        b = b_out if b_out is not None else (np.empty(nrows) if want_b else None)
        self._check(self.lib.extfem_values_get(self.ctx, pattern, _p(nz), _p(b)))
        return nz, b

    def pattern_get_lower(self, pattern: int, want_rowval=True):
        """Pattern of the lower triangle (Int64, 1-based): (nnz_lower, colptr, rowval)."""
        _, ncols, _ = self.pattern_dims(pattern)
        n = C.c_int64(0)
        self._check(self.lib.extfem_pattern_get_lower(self.ctx, pattern, C.byref(n), None, None))
        colptr = np.empty(ncols + 1, np.int64)
        rowval = np.empty(n.value, np.int64) if want_rowval else None
        self._check(self.lib.extfem_pattern_get_lower(self.ctx, pattern, None, _p(colptr), _p(rowval)))
        return int(n.value), colptr, rowval

    def values_get_lower(self, pattern, nzval_out=None, b_out=None, want_b=True):
        """Lower triangle of the device-resident (symmetric) matrix, packed in column order, and the rhs."""
        nrows, _, _ = self.pattern_dims(pattern)
        if nzval_out is None:
            n = C.c_int64(0)
            self._check(self.lib.extfem_pattern_get_lower(self.ctx, pattern, C.byref(n), None, None))
            nzval_out = np.empty(n.value)
        b = b_out if b_out is not None else (np.empty(nrows) if want_b else None)
        self._check(self.lib.extfem_values_get_lower(self.ctx, pattern, _p(nzval_out), _p(b)))
        return nzval_out, b

    def values_set(self, pattern, nzval=None, b=None):
        self._check(self.lib.extfem_values_set(self.ctx, pattern, _p(nzval), _p(b)))

    def device_ptrs(self, pattern):
        ptrs = [C.c_void_p() for _ in range(4)]
        self._check(self.lib.extfem_device_ptrs(self.ctx, pattern, *[C.byref(p) for p in ptrs]))
        return [p.value for p in ptrs]

    def apply_penalties(self, pattern, dofs, values=None, penalty=1e30):
        dofs = np.ascontiguousarray(dofs, dtype=np.int64)
        vals = None if values is None else np.ascontiguousarray(values, dtype=np.float64)
        self._check(self.lib.extfem_apply_penalties(self.ctx, pattern, C.c_int64(dofs.size), _p(dofs), _p(vals), C.c_double(penalty)))

    def residual(self, pattern, sol):
        nrows, _, _ = self.pattern_dims(pattern)
        sol = np.ascontiguousarray(sol, dtype=np.float64)
        res = np.empty(nrows)
        self._check(self.lib.extfem_residual(self.ctx, pattern, _p(sol), _p(res)))
        return res

    def spmv(self, pattern, x):
        nrows, _, _ = self.pattern_dims(pattern)
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty(nrows)
        self._check(self.lib.extfem_spmv(self.ctx, pattern, _p(x), _p(y)))
        return y

    def cg(self, pattern, b=None, x0=None, rtol=1e-10, maxit=10000):
        nrows, _, _ = self.pattern_dims(pattern)
        x = np.zeros(nrows) if x0 is None else np.ascontiguousarray(x0, dtype=np.float64).copy()
        bb = None if b is None else np.ascontiguousarray(b, dtype=np.float64)
        it, rr = C.c_int(), C.c_double()
        self._check(self.lib.extfem_cg(self.ctx, pattern, _p(bb), _p(x), C.c_double(rtol), int(maxit), C.byref(it), C.byref(rr)))
        return x, it.value, rr.value

    def event_record(self, slot: int):
        self._check(self.lib.extfem_event_record(self.ctx, slot))

    def event_elapsed_ms(self, a: int, b: int) -> float:
        ms = C.c_double()
        self._check(self.lib.extfem_event_elapsed_ms(self.ctx, a, b, C.byref(ms)))
        return ms.value

    def set_option(self, key: str, value: int):
        self._check(self.lib.extfem_set_option(self.ctx, key.encode(), int(value)))

    def synchronize(self):
        self._check(self.lib.extfem_synchronize(self.ctx))

    def plan_stats(self, pattern: int, block: int = 0) -> dict:
        """Statistics of the scatter-map plans of a diagonal block (after its first fast-path assembly)."""
        st = (C.c_int64 * 8)()
        self._check(self.lib.extfem_plan_stats(self.ctx, pattern, block, st))
        keys = ["period", "templates", "template_warps", "record_columns", "template_ctas", "pool_bytes", "template_rounds", "columns"]
        return dict(zip(keys, [int(v) for v in st]))

    def plan_jit_status(self, pattern: int, block: int = 0) -> int:
        return int(self.lib.extfem_plan_jit_status(self.ctx, pattern, block))

    # ---- multi-GPU (one process per GPU) ------------------------------------------------------------
    @staticmethod
    def dist_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        lib = load_library()
        rc = lib.extfem_dist_unique_id(buf)
        if rc != 0:
            raise ExtFEMError(rc, lib.extfem_last_error(None).decode())
        return buf.raw

    def dist_init(self, rank: int, world: int, unique_id: bytes | None = None):
        self._check(self.lib.extfem_dist_init(self.ctx, int(rank), int(world), unique_id))

    def dist_set_interfaces(self, pattern: int, plan):
        neigh = np.ascontiguousarray(plan.neigh, dtype=np.int32)
        ptr = np.ascontiguousarray(plan.ptr, dtype=np.int64)
        rows = np.ascontiguousarray(plan.rows, dtype=np.int64)
        owned = np.ascontiguousarray(plan.owned, dtype=np.uint8)
        self._check(self.lib.extfem_dist_set_interfaces(self.ctx, pattern, int(neigh.size), _p(neigh), _p(ptr), _p(rows), _p(owned)))

    def dist_sum_rhs(self, pattern: int):
        self._check(self.lib.extfem_dist_sum_rhs(self.ctx, pattern))

    def dist_spmv(self, pattern, x):
        nrows, _, _ = self.pattern_dims(pattern)
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty(nrows)
        self._check(self.lib.extfem_dist_spmv(self.ctx, pattern, _p(x), _p(y)))
        return y

    def dist_cg(self, pattern, b=None, x0=None, rtol=1e-10, maxit=10000):
        nrows, _, _ = self.pattern_dims(pattern)
        x = np.zeros(nrows) if x0 is None else np.ascontiguousarray(x0, dtype=np.float64).copy()
        bb = None if b is None else np.ascontiguousarray(b, dtype=np.float64)
        it, rr = C.c_int(), C.c_double()
        self._check(self.lib.extfem_dist_cg(self.ctx, pattern, _p(bb), _p(x), C.c_double(rtol), int(maxit), C.byref(it), C.byref(rr)))
        return x, it.value, rr.value

    # owned-row form
    def dist_set_owned(self, pattern: int, plan):
        a = lambda v, t: np.ascontiguousarray(v, dtype=t)      # noqa: E731
        arrs = [a(plan.neigh, np.int32), a(plan.red_send_ptr, np.int64), a(plan.red_send, np.int64), a(plan.red_recv_ptr, np.int64),
                a(plan.red_recv, np.int64), a(plan.halo_send_ptr, np.int64), a(plan.halo_send, np.int64), a(plan.halo_recv_ptr, np.int64),
                a(plan.halo_recv, np.int64), a(plan.owned, np.uint8)]
        self._check(self.lib.extfem_dist_set_owned(self.ctx, pattern, int(arrs[0].size), *[_p(x) for x in arrs]))

    def dist_reduce_system(self, pattern: int, matrix=True, rhs=True):
        self._check(self.lib.extfem_dist_reduce_system(self.ctx, pattern, int(matrix), int(rhs)))

    def dist_spmv_owned(self, pattern, x):
        nrows, _, _ = self.pattern_dims(pattern)
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty(nrows)
        self._check(self.lib.extfem_dist_spmv_owned(self.ctx, pattern, _p(x), _p(y)))
        return y

    def dist_cg_owned(self, pattern, b=None, x0=None, rtol=1e-10, maxit=10000):
        nrows, _, _ = self.pattern_dims(pattern)
        x = np.zeros(nrows) if x0 is None else np.ascontiguousarray(x0, dtype=np.float64).copy()
        bb = None if b is None else np.ascontiguousarray(b, dtype=np.float64)
        it, rr = C.c_int(), C.c_double()
        self._check(self.lib.extfem_dist_cg_owned(self.ctx, pattern, _p(bb), _p(x), C.c_double(rtol), int(maxit), C.byref(it), C.byref(rr)))
        return x, it.value, rr.value

    def launch_count(self) -> int:
        return int(self.lib.extfem_launch_count(self.ctx))

    def last_timings(self):
        ms = (C.c_double * 3)()
        self._check(self.lib.extfem_last_timings(self.ctx, ms))
        return list(ms)
