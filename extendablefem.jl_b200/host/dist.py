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

__all__ = ["cell_ranges", "Shard", "InterfacePlan", "build_interface_plan", "exchange_add_host"]

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
