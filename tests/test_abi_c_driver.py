"""The C-ABI driven by a plain C program with Julia-layout arrays (tests/abi_driver.c): no Python between caller and library.
CPU part: the driver compiles and links against libextfem_cuda.so.  GPU part: two consecutive system assemblies of Example108
(NonlinearOperator + BilinearOperator ON_BFACES + LinearOperator + penalties) equal the CPU oracle's."""
import os
import subprocess

import numpy as np
import pytest

import __graft_entry__ as g
import examples as ex
from util import OracleBackend, check_values

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(g.PKG_DIR, "csrc")


def _build(tmp_path):
    if not os.path.exists(os.path.join(CSRC, "libextfem_cuda.so")):
        g.build()
    exe = str(tmp_path / "abi_driver")
    subprocess.check_call(["gcc", "-O1", "-std=c99", "-o", exe, os.path.join(HERE, "abi_driver.c"), "-L" + CSRC, "-lextfem_cuda",
                           "-Wl,-rpath," + CSRC])
    return exe


def test_c_driver_builds_and_links(tmp_path):
    exe = _build(tmp_path)
    out = subprocess.run(["ldd", exe], capture_output=True, text=True).stdout
    assert "libextfem_cuda.so" in out and "not found" not in out


@pytest.mark.gpu
def test_c_driver_two_system_assemblies(pkg, ora, tmp_path):
    exe = _build(tmp_path)
    pr = pkg.problem
    grid = pkg.simplexgrid(np.linspace(0, 1, 12) ** 1.2)
    F = pkg.FESpace(pkg.H1Pk(1, 1, 2), grid)
    x = F.dof_coordinates()[:, 0]
    sols = [0.3 + x ** 2, 1.0 + np.sin(3 * x)]
    bd = (np.unique(F.bfacedofs[grid.bfaceregions == 2]).astype(np.int64))
    bv = np.exp(x[bd - 1])
    inp, outp = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(inp, "wb") as f:
        np.array([1, grid.ncells, grid.nnodes, F.celldofs.shape[1], F.ndofs, grid.bfacenodes.shape[0], F.bfacedofs.shape[1], bd.size],
                 dtype=np.int64).tofile(f)
        grid.coords.astype(np.float64).tofile(f)
        grid.cellnodes.astype(np.int64).tofile(f)
        F.celldofs.astype(np.int64).tofile(f)
        grid.bfacenodes.astype(np.int64).tofile(f)
        grid.bfaceregions.astype(np.int32).tofile(f)
        F.bfacedofs.astype(np.int64).tofile(f)
        bd.tofile(f); bv.tofile(f)
        for s in sols:
            s.astype(np.float64).tofile(f)
    subprocess.check_call([exe, inp, outp])
    raw = open(outp, "rb").read()
    nnz = int(np.frombuffer(raw, np.int64, 1)[0]); o = 8
    colptr = np.frombuffer(raw, np.int64, F.ndofs + 1, o); o += 8 * (F.ndofs + 1)
    rowval = np.frombuffer(raw, np.int64, nnz, o); o += 8 * nnz
    # the same two assemblies on the oracle through the host mirror of assemble_system!
    PD = pr.ProblemDescription()
    u = pr.Unknown("u")
    pr.assign_unknown(PD, u)
    pr.assign_operator(PD, pr.NonlinearOperator("rcd", [pr.id(u), pr.grad(u)]))
    pr.assign_operator(PD, pr.BilinearOperator("robin108", [pr.id(u)], entities=pr.ON_BFACES, regions=[1], params=[2.0]))
    pr.assign_operator(PD, pr.LinearOperator("exp2x", [pr.id(u)]))
    pr.assign_operator(PD, pr.InterpolateBoundaryData(u, lambda p: np.exp(p[:, :1]), regions=[2]))
    cpu = OracleBackend(pkg, ora, [F])
    pr._prepare_boundary(PD, [F], np.array([0, F.ndofs]), {u: 0})
    assert np.array_equal(cpu.colptr, colptr) and np.array_equal(cpu.rowval, rowval)      # pattern: bit-exact
    for s in sols:
        nz = np.frombuffer(raw, np.float64, nnz, o); o += 8 * nnz
        b = np.frombuffer(raw, np.float64, F.ndofs, o); o += 8 * F.ndofs
        res = np.frombuffer(raw, np.float64, F.ndofs, o); o += 8 * F.ndofs
        s2 = s.copy()
        pr.assemble_system(cpu, PD, s2, {u: 0})
        A2, b2 = cpu.system()
        check_values(nz, A2.data, what="C driver: matrix")
        check_values(b, b2, what="C driver: rhs")
        check_values(res, cpu.residual(s2), scale=np.abs(b2).max(), what="C driver: residual")
