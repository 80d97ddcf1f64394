"""Simplex grids for the host-side mirror of the reference API.

The reference obtains its meshes from ExtendableGrids.jl (``simplexgrid``,
``grid_unitsquare``, ``grid_unitcube``, ``uniform_refine``), which is NOT part of
/root/reference.  The generators below are therefore
"engine conventions": they produce the same *geometric* meshes (which is what the
reference's golden solution functionals depend on), with a documented numbering.
In the Julia drop-in the grid arrays come from ExtendableGrids itself, so
numbering conventions cannot diverge there.

All index arrays are 1-based and column-per-item (``[nodes_per_item, nitems]``
Fortran order in Julia == ``[nitems, nodes_per_item]`` C order here), exactly as
they would be handed over the C-ABI by the Julia side.

Reference call sites these arrays feed: ``xgrid[Coordinates]``, ``xgrid[CellNodes]``,
``xgrid[CellVolumes]``, ``xgrid[CellRegions]`` read in
src/common_operators/bilinear_operator.jl:693-695.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

__all__ = [
    "ExtendableGrid", "simplexgrid", "grid_unitsquare", "grid_unitcube",
    "uniform_refine", "TET_EDGES", "TRI_EDGES",
]

# local edge enumeration (ExtendableGrids local_celledgenodes / local_cellfacenodes)
TRI_EDGES = np.array([[0, 1], [1, 2], [2, 0]], dtype=np.int64)
TET_EDGES = np.array([[0, 1], [0, 2], [0, 3], [1, 2], [1, 3], [2, 3]], dtype=np.int64)
EDGE_EDGES = np.zeros((0, 2), dtype=np.int64)
# local boundary faces of a cell (nodes of the face)
TRI_FACES = TRI_EDGES
TET_FACES = np.array([[0, 2, 1], [0, 1, 3], [1, 2, 3], [0, 3, 2]], dtype=np.int64)


@dataclass
class ExtendableGrid:
    """Minimal stand-in for ExtendableGrids.ExtendableGrid (simplices only)."""
    coords: np.ndarray          # [nnodes, dim] float64
    cellnodes: np.ndarray       # [ncells, dim+1] int32, 1-based
    cellregions: np.ndarray     # [ncells] int32
    bfacenodes: np.ndarray      # [nbfaces, dim] int32, 1-based
    bfaceregions: np.ndarray    # [nbfaces] int32
    _cache: dict = field(default_factory=dict, repr=False)

    @property
    def dim(self) -> int:
        return self.coords.shape[1]

    @property
    def ncells(self) -> int:
        return self.cellnodes.shape[0]

    @property
    def nnodes(self) -> int:
        return self.coords.shape[0]

    # --- derived components -------------------------------------------------
    @property
    def cellvolumes(self) -> np.ndarray:
        if "vol" not in self._cache:
            self._cache["vol"] = simplex_volumes(self.coords, self.cellnodes)
        return self._cache["vol"]

    @property
    def bfacevolumes(self) -> np.ndarray:
        if "bvol" not in self._cache:
            self._cache["bvol"] = simplex_volumes(self.coords, self.bfacenodes)
        return self._cache["bvol"]

    def edges(self):
        """(edgenodes [nedges,2] 1-based, celledges [ncells, nlocaledges] 1-based).

        Edge numbering: ascending by (min node, max node); this is an engine
        convention (ExtendableGrids numbers edges by first encounter)."""
        if "edges" not in self._cache:
            self._cache["edges"] = _enumerate_edges(self)
        return self._cache["edges"]

    @property
    def nedges(self) -> int:
        return self.edges()[0].shape[0]


def simplex_volumes(coords: np.ndarray, itemnodes: np.ndarray) -> np.ndarray:
    """|T| for simplices of any codimension (segment length, triangle area, ...)."""
    x = coords[itemnodes.astype(np.int64) - 1]          # [n, k+1, dim]
    k = itemnodes.shape[1] - 1
    if k == 0:
        return np.ones(itemnodes.shape[0])
    e = x[:, 1:, :] - x[:, :1, :]                       # [n, k, dim]
    if k == e.shape[2]:
        return np.abs(np.linalg.det(e)) / math.factorial(k)
    gram = np.einsum("nik,njk->nij", e, e)
    return np.sqrt(np.abs(np.linalg.det(gram))) / math.factorial(k)


def _enumerate_edges(g: ExtendableGrid):
    loc = {1: EDGE_EDGES, 2: TRI_EDGES, 3: TET_EDGES}[g.dim]
    if g.dim == 1:
        return np.zeros((0, 2), np.int32), np.zeros((g.ncells, 0), np.int32)
    cn = g.cellnodes.astype(np.int64)
    a = cn[:, loc[:, 0]]
    b = cn[:, loc[:, 1]]
    lo = np.minimum(a, b)
    hi = np.maximum(a, b)
    key = lo * (g.nnodes + 1) + hi                      # [ncells, nloc]
    uniq, inv = np.unique(key.ravel(), return_inverse=True)
    edgenodes = np.stack([uniq // (g.nnodes + 1), uniq % (g.nnodes + 1)], axis=1)
    celledges = inv.reshape(key.shape) + 1
    return edgenodes.astype(np.int32), celledges.astype(np.int32)


# ---------------------------------------------------------------------------
# structured grids
# ---------------------------------------------------------------------------
# Kuhn split of the unit cube into 6 tets around the diagonal (000)-(111);
# corner index = ix + 2*iy + 4*iz.  Every tet is positively oriented.
_KUHN = np.array([
    [0, 1, 3, 7], [0, 3, 2, 7], [0, 2, 6, 7],
    [0, 6, 4, 7], [0, 4, 5, 7], [0, 5, 1, 7],
], dtype=np.int64)


def simplexgrid(*axes) -> ExtendableGrid:
    """Tensor-product simplex grid, mirror of ExtendableGrids ``simplexgrid(X[,Y[,Z]])``.

    Nodes are numbered x-fastest.  2D: two triangles per rectangle
    (p00,p10,p11),(p11,p01,p00); boundary regions bottom/right/top/left = 1/2/3/4
    (confirmed by test/test_helper_functions.jl:18-31 of the reference).
    3D: six tets per cuboid (Kuhn split, translation invariant); boundary regions
    follow the ExtendableGrids convention 1:bottom(z-) 2:front(y-) 3:right(x+)
    4:back(y+) 5:left(x-) 6:top(z+)  [engine convention, see SURVEY.md 8c].
    """
    axes = [np.asarray(a, dtype=np.float64) for a in axes]
    dim = len(axes)
    if dim == 1:
        X = axes[0]
        n = X.size
        coords = X.reshape(-1, 1).copy()
        cn = np.stack([np.arange(1, n), np.arange(2, n + 1)], axis=1)
        bf = np.array([[1], [n]])
        return ExtendableGrid(coords, cn.astype(np.int32), np.ones(n - 1, np.int32),
                              bf.astype(np.int32), np.array([1, 2], np.int32))
    if dim == 2:
        X, Y = axes
        nx, ny = X.size, Y.size
        xx, yy = np.meshgrid(X, Y, indexing="xy")       # [ny, nx], x fastest when raveled
        coords = np.stack([xx.ravel(), yy.ravel()], axis=1)
        ix, iy = np.meshgrid(np.arange(nx - 1), np.arange(ny - 1), indexing="xy")
        p00 = (ix + iy * nx).ravel() + 1
        p10, p01, p11 = p00 + 1, p00 + nx, p00 + nx + 1
        t1 = np.stack([p00, p10, p11], axis=1)
        t2 = np.stack([p11, p01, p00], axis=1)
        cn = np.empty((2 * p00.size, 3), np.int64)
        cn[0::2], cn[1::2] = t1, t2
        bottom = np.stack([np.arange(1, nx), np.arange(2, nx + 1)], axis=1)
        top = bottom + (ny - 1) * nx
        left = np.stack([np.arange(0, ny - 1) * nx + 1, np.arange(1, ny) * nx + 1], axis=1)
        right = left + (nx - 1)
        bf = np.concatenate([bottom, right, top[:, ::-1], left[:, ::-1]])
        br = np.concatenate([np.full(nx - 1, 1), np.full(ny - 1, 2),
                             np.full(nx - 1, 3), np.full(ny - 1, 4)])
        return ExtendableGrid(coords, cn.astype(np.int32), np.ones(cn.shape[0], np.int32),
                              bf.astype(np.int32), br.astype(np.int32))
    if dim == 3:
        X, Y, Z = axes
        nx, ny, nz = X.size, Y.size, Z.size
        zz, yy, xx = np.meshgrid(Z, Y, X, indexing="ij")
        coords = np.stack([xx.ravel(), yy.ravel(), zz.ravel()], axis=1)
        iz, iy, ix = np.meshgrid(np.arange(nz - 1), np.arange(ny - 1), np.arange(nx - 1), indexing="ij")
        base = (ix + nx * (iy + ny * iz)).ravel() + 1
        corner = np.array([(c & 1) + nx * ((c >> 1) & 1) + nx * ny * ((c >> 2) & 1) for c in range(8)])
        cube = base[:, None] + corner[None, :]           # [ncubes, 8]
        cn = cube[:, _KUHN].reshape(-1, 4)               # 6 consecutive tets per cube
        bf, br = _structured_bfaces_3d(nx, ny, nz)
        return ExtendableGrid(coords, cn.astype(np.int32), np.ones(cn.shape[0], np.int32),
                              bf.astype(np.int32), br.astype(np.int32))
    raise ValueError("simplexgrid supports 1, 2 or 3 axes")


def _structured_bfaces_3d(nx, ny, nz):
    def nid(ix, iy, iz):
        return ix + nx * (iy + ny * iz) + 1
    faces, regs = [], []

    def quad(a, b, c, d, reg):
        # split along the a-c diagonal, consistent with the Kuhn cube diagonal
        faces.append(np.stack([a, b, c], axis=1)); regs.append(np.full(a.size, reg))
        faces.append(np.stack([a, c, d], axis=1)); regs.append(np.full(a.size, reg))

    iy, ix = [v.ravel() for v in np.meshgrid(np.arange(ny - 1), np.arange(nx - 1), indexing="ij")]
    for iz, reg in ((0, 1), (nz - 1, 6)):
        quad(nid(ix, iy, iz), nid(ix + 1, iy, iz), nid(ix + 1, iy + 1, iz), nid(ix, iy + 1, iz), reg)
    iz, ix = [v.ravel() for v in np.meshgrid(np.arange(nz - 1), np.arange(nx - 1), indexing="ij")]
    for iy_, reg in ((0, 2), (ny - 1, 4)):
        quad(nid(ix, iy_, iz), nid(ix + 1, iy_, iz), nid(ix + 1, iy_, iz + 1), nid(ix, iy_, iz + 1), reg)
    iz, iy = [v.ravel() for v in np.meshgrid(np.arange(nz - 1), np.arange(ny - 1), indexing="ij")]
    for ix_, reg in ((0, 5), (nx - 1, 3)):
        quad(nid(ix_, iy, iz), nid(ix_, iy + 1, iz), nid(ix_, iy + 1, iz + 1), nid(ix_, iy, iz + 1), reg)
    return np.concatenate(faces), np.concatenate(regs)


def grid_unitsquare(scale=(1.0, 1.0), shift=(0.0, 0.0)) -> ExtendableGrid:
    """``grid_unitsquare(Triangle2D)``: 5 nodes, 4 triangles meeting in the centre;
    boundary regions bottom/right/top/left = 1/2/3/4."""
    coords = np.array([[0, 0], [1, 0], [1, 1], [0, 1], [0.5, 0.5]], dtype=np.float64)
    coords = (coords + np.asarray(shift, dtype=np.float64)) * np.asarray(scale, dtype=np.float64)
    cn = np.array([[1, 2, 5], [2, 3, 5], [3, 4, 5], [4, 1, 5]], dtype=np.int32)
    bf = np.array([[1, 2], [2, 3], [3, 4], [4, 1]], dtype=np.int32)
    return ExtendableGrid(coords, cn, np.ones(4, np.int32), bf, np.array([1, 2, 3, 4], np.int32))


def grid_unitcube() -> ExtendableGrid:
    """``grid_unitcube(Tetrahedron3D)``: 8 nodes, 6 tets around the diagonal (0,0,0)-(1,1,1), in ExtendableGrids' node and
    cell order as recalled (the package is not in /root/reference): with it and the refinement rule below, Example301's
    acceptance bound ``L2error <= 8.56e-5`` (examples/Example301_PoissonProblem.jl:89-105) is met at 8.548e-5, whereas other
    interior-diagonal choices give 1.16e-4 (tests/test_oracle_golden.py).  Boundary regions 1:z=0 2:y=0 3:x=1 4:y=1 5:x=0 6:z=1."""
    coords = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]], dtype=np.float64)
    cn = np.array([[1, 2, 3, 7], [1, 3, 4, 7], [1, 5, 6, 7], [1, 8, 5, 7], [1, 6, 2, 7], [1, 4, 8, 7]], dtype=np.int32)
    seen = {}
    for c in cn:
        for f in TET_FACES:
            seen.setdefault(tuple(sorted(c[f])), []).append(c[f])
    bf = np.array([v[0] for v in seen.values() if len(v) == 1], dtype=np.int32)
    mid = coords[bf - 1].mean(axis=1)
    reg = np.zeros(bf.shape[0], np.int32)
    for r, (ax, val) in enumerate(((2, 0.0), (1, 0.0), (0, 1.0), (1, 1.0), (0, 0.0), (2, 1.0)), start=1):
        reg[np.isclose(mid[:, ax], val)] = r
    return ExtendableGrid(coords, cn, np.ones(6, np.int32), bf, reg)


# ---------------------------------------------------------------------------
# uniform (red) refinement
# ---------------------------------------------------------------------------
def uniform_refine(g: ExtendableGrid, nrefs: int = 1) -> ExtendableGrid:
    """Red refinement: every simplex is split into 2^dim children.  3D: ExtendableGrids' rule as recalled -- with the edge
    midpoints numbered 5..10 in local-edge order the children are (1 5 6 7) (2 5 9 8) (3 10 6 8) (4 10 9 7) (10 5 8 9)
    (5 10 7 9) (5 10 8 6) (10 5 7 6), i.e. the inner octahedron is cut along the midpoints of the local edges 12 and 34
    (see grid_unitcube for the evidence)."""
    for _ in range(nrefs):
        g = _refine_once(g)
    return g


def _midpoint_ids(pairs: np.ndarray, nnodes: int):
    lo = np.minimum(pairs[:, 0], pairs[:, 1]).astype(np.int64)
    hi = np.maximum(pairs[:, 0], pairs[:, 1]).astype(np.int64)
    return lo * (nnodes + 1) + hi


def _refine_once(g: ExtendableGrid) -> ExtendableGrid:
    dim, nn = g.dim, g.nnodes
    cn = g.cellnodes.astype(np.int64)
    if dim == 1:
        mid = nn + np.arange(1, g.ncells + 1)
        coords = np.concatenate([g.coords, 0.5 * (g.coords[cn[:, 0] - 1] + g.coords[cn[:, 1] - 1])])
        newcn = np.empty((2 * g.ncells, 2), np.int64)
        newcn[0::2] = np.stack([cn[:, 0], mid], axis=1)
        newcn[1::2] = np.stack([mid, cn[:, 1]], axis=1)
        return ExtendableGrid(coords, newcn.astype(np.int32), np.repeat(g.cellregions, 2),
                              g.bfacenodes.copy(), g.bfaceregions.copy())
    edgenodes, celledges = g.edges()
    ne = edgenodes.shape[0]
    en = edgenodes.astype(np.int64)
    coords = np.concatenate([g.coords, 0.5 * (g.coords[en[:, 0] - 1] + g.coords[en[:, 1] - 1])])
    m = celledges.astype(np.int64) + nn                   # midpoint node ids per local edge
    if dim == 2:
        v0, v1, v2 = cn[:, 0], cn[:, 1], cn[:, 2]
        m01, m12, m20 = m[:, 0], m[:, 1], m[:, 2]
        kids = [np.stack(k, axis=1) for k in (
            (v0, m01, m20), (m01, v1, m12), (m20, m12, v2), (m01, m12, m20))]
        nk = 4
    else:
        v0, v1, v2, v3 = cn.T
        m01, m02, m03, m12, m13, m23 = m.T
        kids = [np.stack(k, axis=1) for k in (
            (v0, m01, m02, m03), (v1, m01, m13, m12), (v2, m23, m02, m12), (v3, m23, m13, m03),
            (m23, m01, m12, m13), (m01, m23, m03, m13), (m01, m23, m12, m02), (m23, m01, m03, m02))]
        nk = 8
    newcn = np.empty((nk * g.ncells, dim + 1), np.int64)
    for k, kid in enumerate(kids):
        newcn[k::nk] = kid
    # boundary faces
    bn = g.bfacenodes.astype(np.int64)
    edge_keys = _midpoint_ids(en, nn)                     # sorted ascending by construction
    def mid_of(a, b):
        return np.searchsorted(edge_keys, _midpoint_ids(np.stack([a, b], axis=1), nn)) + nn + 1
    if dim == 2:
        mb = mid_of(bn[:, 0], bn[:, 1])
        newbf = np.empty((2 * bn.shape[0], 2), np.int64)
        newbf[0::2] = np.stack([bn[:, 0], mb], axis=1)
        newbf[1::2] = np.stack([mb, bn[:, 1]], axis=1)
        newbr = np.repeat(g.bfaceregions, 2)
    else:
        a, b, c = bn.T
        mab, mbc, mca = mid_of(a, b), mid_of(b, c), mid_of(c, a)
        newbf = np.empty((4 * bn.shape[0], 3), np.int64)
        newbf[0::4] = np.stack([a, mab, mca], axis=1)
        newbf[1::4] = np.stack([mab, b, mbc], axis=1)
        newbf[2::4] = np.stack([mca, mbc, c], axis=1)
        newbf[3::4] = np.stack([mab, mbc, mca], axis=1)
        newbr = np.repeat(g.bfaceregions, 4)
    out = ExtendableGrid(coords, newcn.astype(np.int32), np.repeat(g.cellregions, nk),
                         newbf.astype(np.int32), newbr.astype(np.int32))
    _fix_orientation(out)
    return out


def _fix_orientation(g: ExtendableGrid) -> None:
    """Make all top-dimensional cells positively oriented (swap last two nodes)."""
    if g.dim == 1:
        return
    x = g.coords[g.cellnodes.astype(np.int64) - 1]
    det = np.linalg.det(x[:, 1:, :] - x[:, :1, :])
    neg = det < 0
    if neg.any():
        cn = g.cellnodes
        cn[neg, -2], cn[neg, -1] = cn[neg, -1].copy(), cn[neg, -2].copy()
