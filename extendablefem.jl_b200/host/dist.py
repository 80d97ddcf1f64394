"""Cell partitioning and partition interfaces of the sharded (multi-GPU) system -- host logic.

SURVEY.md 8(e): cells are split into contiguous ranges (``simplexgrid`` numbers cells lexicographically, so a
range is a slab); every rank assembles its cells into a local system over its local dofs and the global system
is the sum of the local ones.  Dofs on a partition interface live on several ranks; the lowest rank owns them.
This module builds, for one rank, the local mesh / dofmap (``Shard``) and the lists of interface rows per
neighbouring rank (``InterfacePlan``) that ``extfem_dist_set_interfaces`` takes.  The reference's analogue is the
coloured partitioning of ``ExtendableGrids.partition`` (Example201:63-65) with thread-private matrix parts that
``flush!`` merges (bilinear_operator.jl:969-993).

Nothing here touches the GPU: the same code drives the engine on the box (NCCL) and the CPU tests (gloo)."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .grids import TET_EDGES, TRI_EDGES

__all__ = ["cell_ranges", "Shard", "InterfacePlan", "build_interface_plan", "exchange_add_host",
           "OwnedShard", "OwnedPlan", "build_owned_plan", "reduce_to_owner_host", "halo_host", "SlabShard", "layer_ranges"]

# local faces of a cell as node subsets (any orientation)
_FACES = {1: [(0,), (1,)], 2: [(0, 1), (1, 2), (2, 0)], 3: [(0, 1, 2), (0, 1, 3), (1, 2, 3), (0, 2, 3)]}
_EDGES = {1: np.zeros((0, 2), dtype=np.int64), 2: TRI_EDGES, 3: TET_EDGES}


def cell_ranges(ncells: int, world: int):
    """Contiguous, balanced cell ranges [lo, hi) per rank."""
    base, rem = divmod(ncells, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


def _face_local_dofs(dim: int, order: int, ncomp: int):
    """For every local face the cell-local scalar dof indices lying on it (vertices, then edges with both ends on it),
    replicated per component (component c is offset by c * nscalar_per_cell)."""
    nv = dim + 1
    edges = _EDGES[dim]
    nscalar = nv if order == 1 else (nv + (1 if dim == 1 else edges.shape[0]))
    out = []
    for f in _FACES[dim]:
        loc = list(f)
        if order == 2 and dim > 1:
            loc += [nv + e for e, (a, b) in enumerate(edges) if a in f and b in f]
        out.append(np.array([c * nscalar + i for c in range(ncomp) for i in loc], dtype=np.int64))
    return out


class Shard:
    """The cells [lo, hi) of a global mesh with nodes and dofs renumbered locally (ascending global id, 1-based)."""

    def __init__(self, coords, cellnodes, cellregions, celldofs, lo: int, hi: int, order: int, ncomp: int = 1):
        self.lo, self.hi = lo, hi
        self.dim = coords.shape[1]
        self.order, self.ncomp = order, ncomp
        cn = np.asarray(cellnodes[lo:hi], dtype=np.int64)
        self.nodes_l2g = np.unique(cn)
        self.cellnodes = (np.searchsorted(self.nodes_l2g, cn) + 1).astype(np.int32)
        self.coords = np.ascontiguousarray(coords[self.nodes_l2g - 1])
        self.cellregions = np.ascontiguousarray(cellregions[lo:hi]).astype(np.int32)
        cd = np.asarray(celldofs[lo:hi], dtype=np.int64)
        self.l2g = np.unique(cd)                      # 1-based global dof ids, ascending
        self.celldofs = (np.searchsorted(self.l2g, cd) + 1).astype(np.int32)
        self.ndofs = int(self.l2g.size)
        self.ncells = hi - lo

    def boundary_dofs_global(self) -> np.ndarray:
        """Global ids of the dofs on the boundary of this rank's cell set (faces that belong to one local cell):
        the only dofs that can be shared with another rank."""
        cn = self.cellnodes.astype(np.int64)
        n1 = int(cn.max()) + 1
        keys = []
        for f in _FACES[self.dim]:
            s = np.sort(cn[:, list(f)], axis=1)
            k = np.zeros(cn.shape[0], dtype=np.int64)
            for j in range(s.shape[1]):
                k = k * n1 + s[:, j]
            keys.append(k)
        keys = np.stack(keys, axis=1)                 # [ncells, nfaces]
        _, inv, cnt = np.unique(keys.ravel(), return_inverse=True, return_counts=True)
        bmask = (cnt[inv] == 1).reshape(keys.shape)
        fl = _face_local_dofs(self.dim, self.order, self.ncomp)
        cd = self.celldofs.astype(np.int64)
        found = [cd[bmask[:, f]][:, fl[f]].ravel() for f in range(len(fl))]
        loc = np.unique(np.concatenate(found)) if found else np.zeros(0, dtype=np.int64)
        return self.l2g[loc - 1]


@dataclass
class InterfacePlan:
    rank: int
    world: int
    neigh: np.ndarray        # int32 neighbour ranks
    ptr: np.ndarray          # int64 [nneigh + 1]
    rows: np.ndarray         # int64 local rows (1-based), per neighbour ascending in global id
    owned: np.ndarray        # uint8 [ndofs]: 1 when this rank is the lowest rank holding the dof


def build_interface_plan(shard: Shard, rank: int, world: int, allgather) -> InterfacePlan:
    """``allgather(obj) -> list`` gathers one Python object per rank (torch.distributed.all_gather_object or a stub)."""
    mine = shard.boundary_dofs_global()
    everyone = allgather(mine)
    neigh, ptr, rows = [], [0], []
    owned = np.ones(shard.ndofs, dtype=np.uint8)
    for r in range(world):
        if r == rank:
            continue
        shared = np.intersect1d(mine, everyone[r], assume_unique=True)
        if shared.size == 0:
            continue
        loc = np.searchsorted(shard.l2g, shared) + 1
        neigh.append(r)
        rows.append(loc.astype(np.int64))
        ptr.append(ptr[-1] + loc.size)
        if r < rank:
            owned[loc - 1] = 0
    return InterfacePlan(rank, world, np.asarray(neigh, dtype=np.int32), np.asarray(ptr, dtype=np.int64),
                         np.concatenate(rows) if rows else np.zeros(0, dtype=np.int64), owned)


def exchange_add_host(v: np.ndarray, plan: InterfacePlan, dist) -> np.ndarray:
    """Host twin of the device interface exchange (dist.cuh: iface_exchange_add): v_i <- sum over sharing ranks.
    ``dist`` is torch.distributed (any backend that moves CPU tensors, i.e. gloo)."""
    import torch
    sends = [torch.from_numpy(np.ascontiguousarray(v[plan.rows[plan.ptr[k]:plan.ptr[k + 1]] - 1])) for k in range(plan.neigh.size)]
    recvs = [torch.empty_like(s) for s in sends]
    ops = []
    for k, r in enumerate(plan.neigh):
        ops.append(dist.P2POp(dist.isend, sends[k], int(r)))
        ops.append(dist.P2POp(dist.irecv, recvs[k], int(r)))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    out = v.copy()
    for k in range(plan.neigh.size):
        out[plan.rows[plan.ptr[k]:plan.ptr[k + 1]] - 1] += recvs[k].numpy()
    return out


# ------------------------------------------------------------------------------------------------------------------------
# Owned-row form (north_star: "owned-row assembly ... NCCL exchanges only interface-row contributions").  Every global dof has
# ONE owner rank; after the reduction the owner holds the complete matrix column / vector row of each of its dofs -- the
# distributed analogue of flush! merging the thread-private partitions into one matrix (bilinear_operator.jl:969-993).
class OwnedShard:
    """The cells [lo, hi) of rank ``rank`` plus one layer of GHOST cells (cells of other ranks that share a dof with an owned
    cell), nodes and dofs renumbered locally in ascending global id.  Ghost cells carry volume 0: they shape the local pattern --
    so that the column of an interface dof has the same rows on every rank that holds it -- but contribute exactly 0."""

    def __init__(self, coords, cellnodes, cellregions, cellvolumes, celldofs, ranges, rank: int):
        world = len(ranges)
        lo, hi = ranges[rank]
        self.rank, self.world, self.lo, self.hi = rank, world, lo, hi
        self.dim = coords.shape[1]
        cd = np.asarray(celldofs, dtype=np.int64)
        ncells, ndofs = cd.shape[0], int(cd.max())
        starts = np.array([r[0] for r in ranges], dtype=np.int64)
        cell_rank = (np.searchsorted(starts, np.arange(ncells), side="right") - 1).astype(np.int32)
        # owner of a dof: the rank of the lowest cell that touches it
        owner = np.full(ndofs, world, dtype=np.int32)
        np.minimum.at(owner, cd - 1, cell_rank[:, None])
        touched = np.zeros(ndofs, dtype=bool)
        touched[cd[lo:hi].ravel() - 1] = True
        local_cells = np.nonzero(touched[cd - 1].any(axis=1))[0]        # owned cells and their ghost layer, ascending
        self.cells = local_cells
        self.is_owned_cell = (local_cells >= lo) & (local_cells < hi)
        cn = np.asarray(cellnodes, dtype=np.int64)[local_cells]
        self.nodes_l2g = np.unique(cn)
        self.cellnodes = (np.searchsorted(self.nodes_l2g, cn) + 1).astype(np.int32)
        self.coords = np.ascontiguousarray(coords[self.nodes_l2g - 1])
        self.cellregions = np.ascontiguousarray(np.asarray(cellregions)[local_cells]).astype(np.int32)
        self.cellvolumes = np.where(self.is_owned_cell, np.asarray(cellvolumes)[local_cells], 0.0)
        cdl = cd[local_cells]
        self.l2g = np.unique(cdl)
        self.celldofs = (np.searchsorted(self.l2g, cdl) + 1).astype(np.int32)
        self.ndofs = int(self.l2g.size)
        self.ncells = int(local_cells.size)
        self.owner = owner[self.l2g - 1]                                 # owner rank of every local dof
        self.owned = (self.owner == rank).astype(np.uint8)
        self.touched = touched[self.l2g - 1]                              # local dofs that receive contributions of owned cells


@dataclass
class OwnedPlan:
    rank: int
    world: int
    neigh: np.ndarray          # int32 neighbour ranks
    red_send_ptr: np.ndarray   # per neighbour: my NON-owned rows (1-based, local) whose contributions go to their owner ...
    red_send: np.ndarray
    red_recv_ptr: np.ndarray   # ... and my owned rows that the neighbour contributes to (same order: ascending global id)
    red_recv: np.ndarray
    halo_send_ptr: np.ndarray  # my owned rows of which the neighbour holds a ghost copy
    halo_send: np.ndarray
    halo_recv_ptr: np.ndarray  # my ghost rows owned by the neighbour
    halo_recv: np.ndarray
    owned: np.ndarray          # uint8 [ndofs]


def build_owned_plan(shard: OwnedShard, allgather) -> OwnedPlan:
    """Exchange lists of the owned-row form.  ``allgather(obj) -> list`` gathers one Python object per rank."""
    rank, world = shard.rank, shard.world
    nonowned = shard.owner != rank
    mine = {}
    for r in np.unique(shard.owner[nonowned]):
        sel = nonowned & (shard.owner == r)
        mine[int(r)] = (shard.l2g[sel & shard.touched], shard.l2g[sel])   # (reduce list, halo list), global ids ascending
    everyone = allgather(mine)
    neigh = sorted(set(mine) | {r for r in range(world) if r != rank and rank in everyone[r]})
    ptrs = {k: [0] for k in ("rs", "rr", "hs", "hr")}
    rows = {k: [] for k in ("rs", "rr", "hs", "hr")}
    loc = lambda g: (np.searchsorted(shard.l2g, g) + 1).astype(np.int64)      # noqa: E731
    empty = np.zeros(0, dtype=np.int64)
    for r in neigh:
        rs, hr = mine.get(r, (empty, empty))                  # what I send to / receive from owner r
        rr, hs = everyone[r].get(rank, (empty, empty))        # what r sends to me / needs from me (I am the owner)
        for k, g in (("rs", rs), ("rr", rr), ("hs", hs), ("hr", hr)):
            rows[k].append(loc(g))
            ptrs[k].append(ptrs[k][-1] + g.size)
    cat = lambda k: np.concatenate(rows[k]) if rows[k] else empty               # noqa: E731
    return OwnedPlan(rank, world, np.asarray(neigh, dtype=np.int32),
                     np.asarray(ptrs["rs"], dtype=np.int64), cat("rs"), np.asarray(ptrs["rr"], dtype=np.int64), cat("rr"),
                     np.asarray(ptrs["hs"], dtype=np.int64), cat("hs"), np.asarray(ptrs["hr"], dtype=np.int64), cat("hr"), shard.owned)


def _p2p(sends, nrecv, plan: OwnedPlan, dist):
    import torch
    st = [torch.from_numpy(np.ascontiguousarray(s)) for s in sends]
    rt = [torch.empty(n, dtype=torch.float64) for n in nrecv]
    ops = []
    for k, r in enumerate(plan.neigh):
        if st[k].numel():
            ops.append(dist.P2POp(dist.isend, st[k], int(r)))
        if rt[k].numel():
            ops.append(dist.P2POp(dist.irecv, rt[k], int(r)))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return [t.numpy() for t in rt]


def reduce_to_owner_host(v: np.ndarray, plan: OwnedPlan, dist) -> np.ndarray:
    """Host twin of extfem_dist_reduce_system for a vector: owned rows += the contributions of the other ranks."""
    k = plan.neigh.size
    sends = [v[plan.red_send[plan.red_send_ptr[i]:plan.red_send_ptr[i + 1]] - 1] for i in range(k)]
    recvs = _p2p(sends, [int(plan.red_recv_ptr[i + 1] - plan.red_recv_ptr[i]) for i in range(k)], plan, dist)
    out = v.copy()
    for i in range(k):
        out[plan.red_recv[plan.red_recv_ptr[i]:plan.red_recv_ptr[i + 1]] - 1] += recvs[i]
    return out


def halo_host(v: np.ndarray, plan: OwnedPlan, dist) -> np.ndarray:
    """Host twin of the halo exchange: ghost rows <- the owner's value."""
    k = plan.neigh.size
    sends = [v[plan.halo_send[plan.halo_send_ptr[i]:plan.halo_send_ptr[i + 1]] - 1] for i in range(k)]
    recvs = _p2p(sends, [int(plan.halo_recv_ptr[i + 1] - plan.halo_recv_ptr[i]) for i in range(k)], plan, dist)
    out = v.copy()
    for i in range(k):
        out[plan.halo_recv[plan.halo_recv_ptr[i]:plan.halo_recv_ptr[i + 1]] - 1] = recvs[i]
    return out


# ------------------------------------------------------------------------------------------------------------------------
def layer_ranges(nlayers: int, world: int):
    """Contiguous, balanced ranges [z0, z1) of cube layers per rank (a cell range of a structured simplexgrid)."""
    return cell_ranges(nlayers, world)


class SlabShard:
    """Rank-local part of the structured grid simplexgrid(0:1/n:1, 0:1/n:1, 0:1/n:nz/n) cut into z-slabs of whole cube layers,
    built WITHOUT the global mesh: the rank's own layers [z0, z1) plus one ghost layer towards each neighbour (volume 0), the
    FESpace on it, and the exchange lists of the owned-row form computed from the integer lattice position of every dof
    (2i, 2j, 2k for nodes, odd entries for edge midpoints) -- the same ownership rule as OwnedShard (the rank of the lowest
    cell that touches the dof), no communication needed.  Local node / edge / dof numbering is the global one restricted to the
    slab, so column segments of the local and of the global matrix can be compared entry by entry."""

    def __init__(self, pkg, n: int, nz: int, ranges, rank: int, order: int = 2):
        self.n, self.nz, self.rank, self.world, self.order = n, nz, rank, len(ranges), order
        self.ranges = ranges
        z0, z1 = ranges[rank]
        g0, g1 = int(rank > 0), int(rank < self.world - 1)
        self.z0, self.z1, self.l0, self.l1 = z0, z1, z0 - g0, z1 + g1
        h = 1.0 / n
        X = np.linspace(0.0, 1.0, n + 1)
        Z = h * np.arange(self.l0, self.l1 + 1)
        self.grid = pkg.simplexgrid(X, X, Z)
        per_layer = 6 * n * n
        layer = np.arange(self.grid.ncells) // per_layer + self.l0
        self.is_owned_cell = (layer >= z0) & (layer < z1)
        self.cellvolumes = np.where(self.is_owned_cell, self.grid.cellvolumes, 0.0)
        self.ncells_owned = int(self.is_owned_cell.sum())
        self.FES = pkg.FESpace(pkg.H1Pk(1, 3, order), self.grid)
        pts = self.FES.dof_coordinates()
        ix = np.rint(pts[:, 0] * 2 * n).astype(np.int64)
        iy = np.rint(pts[:, 1] * 2 * n).astype(np.int64)
        kz = np.rint(pts[:, 2] * 2 * n).astype(np.int64)             # doubled global z index
        self.kz = kz
        w = 2 * n + 1
        self.key = (kz * w + iy) * w + ix                              # global lattice id: ascending == global dof order per class
        # owner: rank of the lowest cube layer whose cells touch the point
        starts = np.array([r[0] for r in ranges], dtype=np.int64)
        low_layer = np.maximum(0, (kz - 1) // 2)
        self.owner = (np.searchsorted(starts, low_layer, side="right") - 1).astype(np.int32)
        self.owned = (self.owner == rank).astype(np.uint8)
        self.touched = (kz >= 2 * z0) & (kz <= 2 * z1)

    def owned_plan(self) -> OwnedPlan:
        rank, world = self.rank, self.world
        neigh = [r for r in (rank - 1, rank + 1) if 0 <= r < world]
        ptr = {k: [0] for k in ("rs", "rr", "hs", "hr")}
        rows = {k: [] for k in ("rs", "rr", "hs", "hr")}

        def add(k, sel):
            idx = np.nonzero(sel)[0]          # ascending local index == ascending global dof id: the order both sides agree on
            rows[k].append(idx.astype(np.int64) + 1)
            ptr[k].append(ptr[k][-1] + idx.size)

        for r in neigh:
            z0r, z1r = self.ranges[r]
            l0r, l1r = z0r - int(r > 0), z1r + int(r < world - 1)
            mine = self.owner == rank
            add("rs", (self.owner == r) & self.touched)                                   # my contributions to rows r owns
            add("rr", mine & (self.kz >= 2 * z0r) & (self.kz <= 2 * z1r))                # rows I own that r's cells touch
            add("hs", mine & (self.kz >= 2 * l0r) & (self.kz <= 2 * l1r))                # my rows of which r holds a copy
            add("hr", self.owner == r)                                                    # my copies of r's rows
        cat = lambda k: np.concatenate(rows[k]) if rows[k] else np.zeros(0, dtype=np.int64)   # noqa: E731
        return OwnedPlan(rank, world, np.asarray(neigh, dtype=np.int32),
                         np.asarray(ptr["rs"], dtype=np.int64), cat("rs"), np.asarray(ptr["rr"], dtype=np.int64), cat("rr"),
                         np.asarray(ptr["hs"], dtype=np.int64), cat("hs"), np.asarray(ptr["hr"], dtype=np.int64), cat("hr"), self.owned)
