"""CPU tests (no GPU, no compute calls): the C-ABI library loads and exports every symbol that
include/extfem_cuda.h declares, the ctypes mirror of extfem_opdesc matches the C layout, the kernel
registry resolves names and rejects unregistered kernels, and the product path fails loudly without a
CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import __graft_entry__ as g

HEADER = os.path.join(g.ROOT, "include", "extfem_cuda.h")


@pytest.fixture(scope="module")
def lib(pkg):
    so = os.path.join(g.PKG_DIR, "csrc", "libextfem_cuda.so")
    if not os.path.exists(so):
        g.build()
    return pkg.lib.load_library()


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(extfem_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(pkg, lib):
    names = _declared_functions()
    assert len(names) >= 28
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/extfem_cuda.h but not exported by libextfem_cuda.so"
    # the ctypes binding lists the same entry points
    assert sorted(pkg.lib.EXPORTS) == names


def test_opdesc_layout_matches_ctypes(pkg, tmp_path):
    fields = [f[0] for f in pkg.lib.OpDesc._fields_]
    prog = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void){",
            'printf("%zu\\n", sizeof(extfem_opdesc));']
    for f in fields:
        prog.append(f'printf("%zu\\n", offsetof(extfem_opdesc, {f}));')
    prog.append("return 0;}")
    src = tmp_path / "layout.c"
    src.write_text("\n".join(prog))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-o", str(exe), str(src)])
    out = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert out[0] == C.sizeof(pkg.lib.OpDesc)
    for f, off in zip(fields, out[1:]):
        assert getattr(pkg.lib.OpDesc, f).offset == off, f


def test_kernel_registry(pkg, lib):
    for name in ["standard", "dcr", "stokes", "linnse7", "hooke_grad", "hooke_voigt", "convect_args", "constant_one",
                 "constant_params", "xy", "sincos301", "tabulated", "nse2d", "nl_linnse7", "neohooke3d", "rcd"]:
        assert pkg.lib.kernel_id(name) > 0
    with pytest.raises(pkg.lib.ExtFEMError) as e:
        pkg.lib.kernel_id("my_julia_closure")
    assert e.value.code == -1 and "not registered" in str(e.value)


def test_no_cpu_fallback(pkg, lib):
    """Without a CUDA device, context creation fails with EXTFEM_ERR_CUDA (there is no CPU path)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(pkg.lib.ExtFEMError) as e:
        pkg.lib.Engine(0)
    assert e.value.code == -4 and "no CPU fallback" in str(e.value)


def test_null_context_is_rejected(lib):
    assert lib.extfem_synchronize(None) == -3
    assert b"ctx is NULL" in lib.extfem_last_error(None)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under the package may reference it."""
    for root, _, files in os.walk(g.PKG_DIR):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl")):
                txt = open(os.path.join(root, f)).read()
                assert "assembly_ref" not in txt and "import oracle" not in txt and "from oracle" not in txt, f


def test_fespace_dofmaps(pkg):
    """Host mirror of FES[CellDofs] (helper_functions.jl:561-567): P2 = nodes then edges, components stacked."""
    X = np.linspace(0, 1, 4)
    grid = pkg.simplexgrid(X, X, X)
    F = pkg.FESpace(pkg.H1P2(3, 3), grid)
    nn, ne = grid.nnodes, grid.nedges
    assert F.ndofs == 3 * (nn + ne) and F.celldofs.shape == (grid.ncells, 30)
    assert np.array_equal(F.celldofs[:, :4], grid.cellnodes)
    assert F.celldofs[:, 4:10].min() == nn + 1 and F.celldofs[:, 4:10].max() == nn + ne
    assert np.array_equal(F.celldofs[:, 10:20], F.celldofs[:, :10] + nn + ne)
    # every edge dof sits at the midpoint of its two vertices
    xd = F.dof_coordinates()
    cn = grid.cellnodes.astype(np.int64) - 1
    cd = F.celldofs.astype(np.int64) - 1
    for e, (a, b) in enumerate(pkg.TET_EDGES):
        assert np.allclose(xd[cd[:, 4 + e]], 0.5 * (grid.coords[cn[:, a]] + grid.coords[cn[:, b]]))
