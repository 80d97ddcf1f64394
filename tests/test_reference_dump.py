"""GPU test against dumps of the REAL reference (baseline/run_reference.jl, produced by scripts/probe_reference.sh where a Julia
toolchain and the reference's packages exist).  Skipped when no dump is present -- which is the case in this repository's build
image; the oracle is then pinned by the reference's golden values instead (tests/test_examples_oracle.py)."""
import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
DUMPS = sorted(glob.glob(os.path.join(HERE, "..", "baseline", "_dump", "*.bin")))


def read_dump(path):
    raw = np.fromfile(path, dtype=np.uint8)
    hdr = raw[:48].view(np.int64)
    dim, ncells, nnodes, nd, ndofs, nnz = (int(v) for v in hdr)
    o = 48

    def take(n, dt):
        nonlocal o
        a = raw[o:o + n * 8].view(dt).copy()
        o += n * 8
        return a
    coords = take(dim * nnodes, np.float64).reshape(nnodes, dim)            # Julia [dim, nnodes] column-major == [nnodes][dim]
    cellnodes = take((dim + 1) * ncells, np.int64).reshape(ncells, dim + 1)
    celldofs = take(nd * ncells, np.int64).reshape(ncells, nd)
    colptr = take(ndofs + 1, np.int64); rowval = take(nnz, np.int64); nzval = take(nnz, np.float64); b = take(ndofs, np.float64)
    return dict(dim=dim, coords=coords, cellnodes=cellnodes, celldofs=celldofs, ndofs=ndofs, colptr=colptr, rowval=rowval, nzval=nzval, b=b)


@pytest.mark.skipif(not DUMPS, reason="no dump of the Julia reference (scripts/probe_reference.sh found no julia)")
@pytest.mark.parametrize("path", DUMPS or ["-"])
def test_against_reference_dump(pkg, engine, path):
    import scipy.sparse as sp
    D = read_dump(path)
    name = os.path.basename(path)
    order = 3 if "p3" in name else (1 if "p1" in name else 2)
    mesh = engine.mesh_set(D["coords"], D["cellnodes"], None, None)
    if order == 3:
        pytest.skip("order-3 dumps need the reference's basis coefficients (extfem_space_set_tables); compared through Example201's golden")
    sp_ = engine.space_set(mesh, order, 1, D["celldofs"], D["ndofs"])
    pat = engine.pattern_build([sp_])
    cp, rv = engine.pattern_get(pat)
    engine.assemble_bilinear(pat, engine.make_opdesc([(0, 1)], [(0, 1)]))
    engine.assemble_linear(pat, engine.make_opdesc([(0, 0)], kernel_id=pkg.lib.kernel_id("xy")))
    nz, b = engine.values_get(pat)
    N = D["ndofs"]
    A = sp.csc_matrix((nz, rv - 1, cp - 1), shape=(N, N))
    R = sp.csc_matrix((D["nzval"], D["rowval"] - 1, D["colptr"] - 1), shape=(N, N))
    # the reference drops exact zeros (bilinear_operator.jl:925): its pattern is a subset of the structural one
    assert (abs(A - R)).max() <= 1e-12 * abs(R).max()
    assert np.abs(b - D["b"]).max() <= 1e-12 * np.abs(D["b"]).max()
