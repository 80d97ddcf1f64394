"""Worker of the multi-rank tests (spawned once per rank).  backend 'gloo': CPU, local systems from the oracle, host
exchange -- checks the partition / interface logic.  backend 'nccl': one GPU per rank through the C-ABI
(extfem_dist_*) -- the product path."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

RTOL = 1e-12       # matrix / rhs entries (north_star)
RTOL_SOL = 1e-10   # solutions (north_star)


def problem(pkg, n=4, dim=3, order=2):
    X = np.linspace(0, 1, n + 1)
    grid = pkg.simplexgrid(*([X ** 1.2] + [X] * (dim - 1)))
    FES = pkg.FESpace(pkg.H1Pk(1, dim, order), grid)
    bdofs = np.unique(FES.bfacedofs)
    return grid, FES, bdofs


def run(rank, world, backend, port, out):
    import torch
    import torch.distributed as dist
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl
    import __graft_entry__ as g
    from oracle import oracle as ora
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    if backend == "nccl":
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = g.load_package()
    ora.build()
    grid, FES, bdofs = problem(pkg)
    lo, hi = pkg.cell_ranges(grid.ncells, world)[rank]
    sh = pkg.Shard(grid.coords, grid.cellnodes, grid.cellregions, FES.celldofs, lo, hi, order=2)

    def allgather(obj):
        lst = [None] * world
        dist.all_gather_object(lst, obj)
        return lst

    plan = pkg.build_interface_plan(sh, rank, world, allgather)
    # every global dof is owned exactly once
    owned_g = np.concatenate(allgather(sh.l2g[plan.owned == 1]))
    assert np.array_equal(np.sort(owned_g), np.arange(1, FES.ndofs + 1)), "ownership is not a partition of the dofs"
    assert plan.neigh.size >= 1

    # ---- reference: the global system on one process (oracle)
    om = ora.Mesh(grid.coords, grid.cellnodes, grid.cellregions, grid.cellvolumes)
    gr = ora.OraArg(FES.celldofs, 1, 2, ora.OP_GRAD)
    idu = ora.OraArg(FES.celldofs, 1, 2, ora.OP_ID)
    cpg, rvg = ora.structural_pattern([gr], [gr], (FES.ndofs, FES.ndofs))
    nzg = ora.assemble_bilinear(om, [gr], [gr], "standard", csc=(cpg, rvg))
    bg = np.zeros(FES.ndofs)
    ora.assemble_linear(om, [idu], bg, "sincos301", params=[1.0])
    Ag = sp.csc_matrix((nzg, rvg - 1, cpg - 1), shape=(FES.ndofs, FES.ndofs))

    # ---- local system of this rank
    vol = pkg.host.grids.simplex_volumes(sh.coords, sh.cellnodes)
    if backend == "nccl":
        eng = pkg.lib.Engine(rank)
        uid = [pkg.lib.Engine.dist_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        eng.dist_init(rank, world, uid[0])
        mesh = eng.mesh_set(sh.coords, sh.cellnodes, sh.cellregions, vol)
        spc = eng.space_set(mesh, 2, 1, sh.celldofs, sh.ndofs)
        pat = eng.pattern_build([spc])
        cp, rv = eng.pattern_get(pat)
        eng.dist_set_interfaces(pat, plan)
        nz = np.empty(rv.size)
        eng.assemble_bilinear(pat, eng.make_opdesc([(0, 1)], [(0, 1)]), nzval_out=nz)
        bl = np.empty(sh.ndofs)
        eng.assemble_linear(pat, eng.make_opdesc([(0, 0)], kernel_id=pkg.lib.kernel_id("sincos301"), params=[1.0]), b_out=bl)
    else:
        oml = ora.Mesh(sh.coords, sh.cellnodes, sh.cellregions, vol)
        grl = ora.OraArg(sh.celldofs, 1, 2, ora.OP_GRAD)
        idl = ora.OraArg(sh.celldofs, 1, 2, ora.OP_ID)
        cp, rv = ora.structural_pattern([grl], [grl], (sh.ndofs, sh.ndofs))
        nz = ora.assemble_bilinear(oml, [grl], [grl], "standard", csc=(cp, rv))
        bl = np.zeros(sh.ndofs)
        ora.assemble_linear(oml, [idl], bl, "sincos301", params=[1.0])
    Al = sp.csc_matrix((nz, rv - 1, cp - 1), shape=(sh.ndofs, sh.ndofs))

    # (1) sum of the local matrices == global matrix, entry by entry
    parts = allgather((sh.l2g, Al.tocoo()))
    S = sp.csc_matrix((FES.ndofs, FES.ndofs))
    for l2g, M in parts:
        S = S + sp.csc_matrix((M.data, (l2g[M.row] - 1, l2g[M.col] - 1)), shape=S.shape)
    D = (S - Ag).tocoo()
    err = np.abs(D.data).max() if D.nnz else 0.0
    assert err <= RTOL * np.abs(nzg).max(), f"sum of local matrices differs from the global matrix: {err:.3e}"

    # (2) interface-row sum of the rhs
    if backend == "nccl":
        eng.dist_sum_rhs(pat)
        _, bc = eng.values_get(pat, want_nzval=False)
    else:
        bc = pkg.exchange_add_host(bl, plan, dist)
    eb = np.abs(bc - bg[sh.l2g - 1]).max()
    assert eb <= RTOL * np.abs(bg).max(), f"consistent rhs differs from the global rhs: {eb:.3e}"

    # (3) sharded SpMV == global SpMV
    xg = np.cos(np.arange(FES.ndofs) * 0.37) + 0.1
    if backend == "nccl":
        y = eng.dist_spmv(pat, xg[sh.l2g - 1])
    else:
        y = pkg.exchange_add_host(Al @ xg[sh.l2g - 1], plan, dist)
    ey = np.abs(y - (Ag @ xg)[sh.l2g - 1]).max()
    assert ey <= RTOL * np.abs(Ag @ xg).max() * 10, f"sharded spmv differs: {ey:.3e}"

    # (4) Dirichlet penalties (owner only) + Jacobi CG on the sharded system == global solve
    pen = 1e30
    Agp = Ag.tolil(); bgp = bg.copy()
    for d in bdofs:
        Agp[d - 1, d - 1] = pen; bgp[d - 1] = 0.0
    xref = spl.spsolve(Agp.tocsc(), bgp)
    isb = np.isin(sh.l2g, bdofs)
    mine = np.nonzero(isb & (plan.owned == 1))[0] + 1        # the owner applies the penalty
    if backend == "nccl":
        eng.apply_penalties(pat, mine, None, pen)
        xs, it, rr = eng.dist_cg(pat, rtol=1e-13, maxit=5000)
    else:
        Alp = Al.tolil(); blp = bc.copy()
        for d in mine:
            Alp[d - 1, d - 1] = pen
        blp[isb] = 0.0
        Alp = Alp.tocsr()
        w = plan.owned.astype(np.float64)

        def gdot(a, b):
            t = torch.tensor([float(np.dot(w * a, b))], dtype=torch.float64)
            dist.all_reduce(t)
            return float(t[0])
        dinv = 1.0 / pkg.exchange_add_host(Alp.diagonal(), plan, dist)
        xs = np.zeros(sh.ndofs); r = blp.copy(); z = dinv * r; p = z.copy(); rz = gdot(r, z); bn = np.sqrt(gdot(blp, blp)); it = 0
        while np.sqrt(gdot(r, r)) / bn > 1e-13 and it < 5000:
            q = pkg.exchange_add_host(Alp @ p, plan, dist)
            alpha = rz / gdot(p, q)
            xs += alpha * p; r -= alpha * q; z = dinv * r
            rzn = gdot(r, z); p = z + (rzn / rz) * p; rz = rzn; it += 1
    es = np.abs(xs - xref[sh.l2g - 1]).max()
    assert es <= RTOL_SOL * np.abs(xref).max() * 10, f"sharded CG solution differs from the global solve: {es:.3e} after {it} iterations"
    msg_owned = owned_row_form(pkg, ora, dist, allgather, backend, rank, world, grid, FES, bdofs, Ag, bg, cpg, rvg, nzg, xref,
                               eng if backend == "nccl" else None)
    if backend == "nccl":
        eng.close()
    dist.barrier()
    dist.destroy_process_group()
    with open(out + f".{rank}", "w") as f:
        f.write(f"ok err_matrix={err:.2e} err_rhs={eb:.2e} err_spmv={ey:.2e} err_sol={es:.2e} iters={it} | owned-row: {msg_owned}\n")


def owned_row_form(pkg, ora, dist, allgather, backend, rank, world, grid, FES, bdofs, Ag, bg, cpg, rvg, nzg, xref, eng):
    """Owned-row form (host/dist.py: OwnedShard / build_owned_plan; C-ABI: extfem_dist_set_owned / _reduce_system / _cg_owned):
    after the reduction every owner holds the complete column of each of its dofs -- compared ENTRYWISE with the global
    (single-process) matrix, which is possible because the slab partition does not renumber."""
    import scipy.sparse as sp
    ranges = pkg.cell_ranges(grid.ncells, world)
    sh = pkg.OwnedShard(grid.coords, grid.cellnodes, grid.cellregions, grid.cellvolumes, FES.celldofs, ranges, rank)
    plan = pkg.build_owned_plan(sh, allgather)
    owned_g = np.concatenate(allgather(sh.l2g[sh.owned == 1]))
    assert np.array_equal(np.sort(owned_g), np.arange(1, FES.ndofs + 1)), "ownership is not a partition of the dofs"
    own = np.nonzero(sh.owned == 1)[0]
    # local system (ghost cells have volume 0: they shape the pattern, contribute nothing)
    if backend == "nccl":
        mesh = eng.mesh_set(sh.coords, sh.cellnodes, sh.cellregions, sh.cellvolumes)
        spc = eng.space_set(mesh, 2, 1, sh.celldofs, sh.ndofs)
        pat = eng.pattern_build([spc])
        cp, rv = eng.pattern_get(pat)
        eng.dist_set_owned(pat, plan)
        eng.assemble_bilinear(pat, eng.make_opdesc([(0, 1)], [(0, 1)]))
        eng.assemble_linear(pat, eng.make_opdesc([(0, 0)], kernel_id=pkg.lib.kernel_id("sincos301"), params=[1.0]))
        eng.dist_reduce_system(pat, True, True)
        nz, bl = eng.values_get(pat)
    else:
        oml = ora.Mesh(sh.coords, sh.cellnodes, sh.cellregions, sh.cellvolumes)
        grl = ora.OraArg(sh.celldofs, 1, 2, ora.OP_GRAD)
        idl = ora.OraArg(sh.celldofs, 1, 2, ora.OP_ID)
        cp, rv = ora.structural_pattern([grl], [grl], (sh.ndofs, sh.ndofs))
        nz = ora.assemble_bilinear(oml, [grl], [grl], "standard", csc=(cp, rv))
        bl = np.zeros(sh.ndofs)
        ora.assemble_linear(oml, [idl], bl, "sincos301", params=[1.0])
        bl = pkg.reduce_to_owner_host(bl, plan, dist)
        # host twin of the matrix reduction: whole column segments, shared columns have equal lengths on both ranks
        collen = np.diff(cp)
        k = plan.neigh.size
        sends = [np.concatenate([nz[cp[c - 1] - 1:cp[c] - 1] for c in plan.red_send[plan.red_send_ptr[i]:plan.red_send_ptr[i + 1]]] or [np.zeros(0)])
                 for i in range(k)]
        nrecv = [int(collen[plan.red_recv[plan.red_recv_ptr[i]:plan.red_recv_ptr[i + 1]] - 1].sum()) for i in range(k)]
        from extfem_b200.host.dist import _p2p
        recvs = _p2p(sends, nrecv, plan, dist)
        nz = nz.copy()
        for i in range(k):
            o = 0
            for c in plan.red_recv[plan.red_recv_ptr[i]:plan.red_recv_ptr[i + 1]]:
                n = collen[c - 1]
                nz[cp[c - 1] - 1:cp[c] - 1] += recvs[i][o:o + n]
                o += n
            assert o == recvs[i].size
    # (1) owned columns == global columns: same rows (pattern, bit-exact) and same values
    errA = 0.0
    for c in own:
        gcol = sh.l2g[c] - 1
        rows_l = sh.l2g[rv[cp[c] - 1:cp[c + 1] - 1] - 1]
        rows_g = rvg[cpg[gcol] - 1:cpg[gcol + 1] - 1]
        assert np.array_equal(rows_l, rows_g), f"column {gcol + 1}: local pattern of an owned column differs from the global one"
        errA = max(errA, np.abs(nz[cp[c] - 1:cp[c + 1] - 1] - nzg[cpg[gcol] - 1:cpg[gcol + 1] - 1]).max())
    assert errA <= RTOL * np.abs(nzg).max(), f"owned columns differ from the global matrix: {errA:.3e}"
    # (2) owned rhs rows == global rhs
    errb = np.abs(bl[own] - bg[sh.l2g[own] - 1]).max()
    assert errb <= RTOL * np.abs(bg).max(), f"owned rhs rows differ: {errb:.3e}"
    # (3) SpMV on owned rows after a halo exchange
    xg = np.cos(np.arange(FES.ndofs) * 0.37) + 0.1
    xl = np.where(sh.owned == 1, xg[sh.l2g - 1], 777.0)          # ghost copies start wrong: the halo exchange must fix them
    if backend == "nccl":
        y = eng.dist_spmv_owned(pat, xl)
    else:
        Al = sp.csc_matrix((nz, rv - 1, cp - 1), shape=(sh.ndofs, sh.ndofs))
        y = Al.T @ pkg.halo_host(xl, plan, dist)
    erry = np.abs(y[own] - (Ag @ xg)[sh.l2g[own] - 1]).max()
    assert erry <= RTOL * np.abs(Ag @ xg).max() * 10, f"owned-row spmv differs: {erry:.3e}"
    # (4) penalties with NON-ZERO Dirichlet values on every local boundary dof (also interface ones, on every rank) + CG
    msg = f"err_matrix={errA:.2e} err_rhs={errb:.2e} err_spmv={erry:.2e}"
    if backend == "nccl":
        import scipy.sparse.linalg as spl
        pen = 1e30
        gval = lambda d: 0.25 + 0.5 * np.sin(d * 0.1)                  # noqa: E731
        Agp = Ag.tolil(); bgp = bg.copy()
        for d in bdofs:
            Agp[d - 1, d - 1] = pen; bgp[d - 1] = pen * gval(d)
        xr = spl.spsolve(Agp.tocsc(), bgp)
        isb = np.isin(sh.l2g, bdofs)
        eng.apply_penalties(pat, np.nonzero(isb)[0] + 1, gval(sh.l2g[isb]), pen)
        x0 = np.zeros(sh.ndofs)
        x0[isb] = gval(sh.l2g[isb])                                  # the assemble_sol leg of apply_penalties!
        xs, it, rr = eng.dist_cg_owned(pat, x0=x0, rtol=1e-13, maxit=5000)
        es = np.abs(xs - xr[sh.l2g - 1]).max()
        assert es <= RTOL_SOL * np.abs(xr).max() * 10, f"owned-row CG differs from the global solve: {es:.3e} after {it} iterations"
        msg += f" err_sol={es:.2e} iters={it}"
    return msg


if __name__ == "__main__":
    run(int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], int(sys.argv[4]), sys.argv[5])
