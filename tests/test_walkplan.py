"""CPU test of the walk planner (csrc/walkplan.h, host-only C++): for every column shape of structured and permuted 3D P2 grids
the planner must produce a walk program that its symbolic verifier accepts (every (cell, local row) contribution reaches its
position exactly once, in first-touch and in accumulate mode), and the programs must need far fewer shared-memory round trips
than one read-modify-write per contribution."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import __graft_entry__ as g

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def planner(tmp_path_factory):
    so = tmp_path_factory.mktemp("walk") / "libwalkplan.so"
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-o", str(so), os.path.join(HERE, "walkplan_host.cpp")])
    return C.CDLL(str(so))


def column_templates(grid, FES):
    """(cells, local index, positions) of every column, like the engine's adjacency + posmap."""
    cd = FES.celldofs.astype(np.int64) - 1
    ncells, nd = cd.shape
    order = np.lexsort((np.repeat(np.arange(ncells), nd), cd.ravel()))
    dof_sorted = cd.ravel()[order]
    cell_sorted = order // nd
    loc_sorted = order % nd
    starts = np.searchsorted(dof_sorted, np.arange(FES.ndofs + 1))
    for k in range(FES.ndofs):
        cells = cell_sorted[starts[k]:starts[k + 1]]
        locs = loc_sorted[starts[k]:starts[k + 1]]
        rows = np.unique(cd[cells].ravel())
        pos = np.searchsorted(rows, cd[cells])
        yield k, cells, locs, pos, rows.size


@pytest.mark.parametrize("permute", [False, True])
def test_walk_programs_verify(pkg, planner, permute):
    X = np.linspace(0, 1, 5)
    grid = pkg.simplexgrid(X, X, X)
    if permute:
        cn = grid.cellnodes.copy()
        for k, p in ((1, [1, 2, 0]), (2, [2, 0, 1])):
            sel = np.arange(grid.ncells) % 3 == k
            cn[sel, :3] = grid.cellnodes[sel][:, p]
        grid.cellnodes[:] = cn
        grid._cache.clear()
    FES = pkg.FESpace(pkg.H1P2(1, 3), grid)
    nn = grid.nnodes
    cnn = grid.cellnodes.astype(np.int64)
    seen = set()
    nrmw = nld = nst = 0
    for k, cells, locs, pos, L in column_templates(grid, FES):
        m = cells.size
        orient = np.zeros(m, np.int32)
        if k >= nn:     # edge dof: orientation of the local edge relative to round 0 (fastplan.cuh: tp_tmpl_kernel bit 20)
            ea = np.array([0, 0, 0, 1, 1, 2]); eb = np.array([1, 2, 3, 2, 3, 3])
            pa = pos[np.arange(m), ea[locs - 4]]
            orient = (pa != pa[0]).astype(np.int32)
        key = (m, L, tuple(locs), tuple(pos.ravel()), tuple(orient))
        if key in seen:
            continue
        seen.add(key)
        words = np.zeros((m, 12), np.uint32)
        nl, ns = C.c_int(), C.c_int()
        ok = planner.tw_test_plan(m, L, (cells - cells[0]).astype(np.int32).ctypes.data_as(C.c_void_p),
                                  locs.astype(np.int32).ctypes.data_as(C.c_void_p), orient.ctypes.data_as(C.c_void_p),
                                  np.ascontiguousarray(pos, dtype=np.int32).ctypes.data_as(C.c_void_p),
                                  words.ctypes.data_as(C.c_void_p), C.byref(nl), C.byref(ns))
        assert ok == 1, f"column {k}: no verified walk program (m={m}, L={L})"
        nrmw += 10 * m
        nld += nl.value
        nst += ns.value
    assert len(seen) > 20
    # read-modify-write costs a load and a store per contribution; the walk programs need well under half of that
    assert nld + nst < 0.5 * 2 * nrmw, (nld, nst, nrmw)
    print(f"templates {len(seen)}: contributions {nrmw}, walk loads {nld}, stores {nst}")
