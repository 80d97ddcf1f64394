"""CPU tests of the host-side partition logic of the owned-row form (host/dist.py): the structured-slab construction that the
multi-GPU benchmark uses (SlabShard: no global mesh, no communication) must produce exactly the local meshes, dof maps,
ownership and exchange lists that the generic construction (OwnedShard + build_owned_plan on the global mesh) produces."""
import numpy as np
import pytest


def _generic(pkg, n, world):
    X = np.linspace(0, 1, n + 1)
    grid = pkg.simplexgrid(X, X, X)
    FES = pkg.FESpace(pkg.H1P2(1, 3), grid)
    cranges = [(a * 6 * n * n, b * 6 * n * n) for a, b in pkg.layer_ranges(n, world)]
    shards = [pkg.OwnedShard(grid.coords, grid.cellnodes, grid.cellregions, grid.cellvolumes, FES.celldofs, cranges, r) for r in range(world)]
    objs = []
    for sh in shards:
        nonowned = sh.owner != sh.rank
        d = {}
        for r in np.unique(sh.owner[nonowned]):
            sel = nonowned & (sh.owner == r)
            d[int(r)] = (sh.l2g[sel & sh.touched], sh.l2g[sel])
        objs.append(d)
    return FES, shards, [pkg.build_owned_plan(sh, lambda o: objs) for sh in shards]


@pytest.mark.parametrize("n,world", [(5, 2), (6, 3), (8, 4)])
def test_slab_plan_equals_generic_plan(pkg, n, world):
    FES, shards, plans = _generic(pkg, n, world)
    ranges = pkg.layer_ranges(n, world)
    owned_total = 0
    for r in range(world):
        ss = pkg.SlabShard(pkg, n, n, ranges, r)
        pb = ss.owned_plan()
        assert ss.FES.ndofs == shards[r].ndofs and ss.grid.ncells == shards[r].ncells
        assert np.array_equal(ss.FES.celldofs, shards[r].celldofs)          # local numbering == global numbering restricted
        assert np.array_equal(ss.owned, shards[r].owned)
        assert np.array_equal(ss.cellvolumes == 0.0, ~shards[r].is_owned_cell)
        for f in ("neigh", "red_send_ptr", "red_send", "red_recv_ptr", "red_recv", "halo_send_ptr", "halo_send", "halo_recv_ptr", "halo_recv"):
            assert np.array_equal(getattr(plans[r], f), getattr(pb, f)), (r, f)
        owned_total += int(ss.owned.sum())
    assert owned_total == FES.ndofs                                          # every dof has exactly one owner


def test_send_and_receive_lists_pair_up(pkg):
    """What rank a sends to rank b is what b expects from a (same dofs, same order), for both exchanges."""
    n, world = 6, 3
    FES, shards, plans = _generic(pkg, n, world)
    for a in range(world):
        pa = plans[a]
        for ia, b in enumerate(pa.neigh):
            pb = plans[b]
            ib = list(pb.neigh).index(a)
            for snd, rcv in (("red_send", "red_recv"), ("halo_send", "halo_recv")):
                sa = getattr(pa, snd)[getattr(pa, snd + "_ptr")[ia]:getattr(pa, snd + "_ptr")[ia + 1]]
                rb = getattr(pb, rcv)[getattr(pb, rcv + "_ptr")[ib]:getattr(pb, rcv + "_ptr")[ib + 1]]
                assert np.array_equal(shards[a].l2g[sa - 1], shards[b].l2g[rb - 1]), (a, b, snd)
