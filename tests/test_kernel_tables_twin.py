"""CPU twins of small index / table logic inside the CUDA sources (parsed from the sources, so the tests follow the code):
the cofactor-derivative table of the row-wise Neo-Hooke Jacobian against the EXTFEM_D2 list of the array form, and the flat
write-out of the walk kernel (coverage of the (column, position) block and freedom from shared-memory bank conflicts with the
skewed accumulator columns)."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "extendablefem.jl_b200", "csrc")


def test_neo_d2_table_equals_the_array_form_list():
    src = open(os.path.join(CSRC, "kernels_generic.cuh")).read()
    # array form (nl_apply, EXTFEM_NL_NEOHOOKE3D): J[i][j] += s * c * F[k]
    ref = {}
    for i, j, s, k in re.findall(r"EXTFEM_D2\((\d), (\d), (-?1), (\d)\)", src):
        assert (int(i), int(j)) not in ref
        ref[(int(i), int(j))] = int(s) * (int(k) + 1)
    assert len(ref) == 36
    # row-wise form: constexpr table T[9][8] of neo_d2 = four (j, sign * (k + 1)) pairs per row
    body = re.search(r"constexpr int neo_d2\(int i, int j\)\s*\{\s*constexpr int T\[9\]\[8\] = \{(.*?)\};", src, re.S).group(1)
    rows = re.findall(r"\{([^{}]*)\}", body)
    assert len(rows) == 9
    tab = {}
    for i, row in enumerate(rows):
        v = [int(t) for t in row.split(",")]
        assert len(v) == 8
        for q in range(4):
            assert (i, v[2 * q]) not in tab
            tab[(i, v[2 * q])] = v[2 * q + 1]
    assert tab == ref
    # the derivative of a cofactor never involves its own row / column entry, and D2W is symmetric: d(dd_i)/dF_j = d(dd_j)/dF_i
    for (i, j), v in tab.items():
        assert i != j and tab[(j, i)] == v


def _tp_ld():
    src = open(os.path.join(CSRC, "fastplan.cuh")).read()
    return int(re.search(r"constexpr int TP_LD = (\d+);", src).group(1))


@pytest.mark.parametrize("L", [1, 3, 5, 8, 15, 16, 19, 27, 31, 32, 33, 41, 64, 65, 87])
def test_flat_writeout_mapping(L):
    """tw_writeout_flat (fastplan.cuh): lanes run over i = j L + p; (p, j) advance by (dp, dj) with one conditional wrap."""
    LD = _tp_ld()
    assert LD % 16 == 1          # a position step moves one 8-byte bank on
    skew = L if L % 2 else 1
    dj, dp = 32 // L, 32 - (32 // L) * L
    lane = np.arange(32)
    j, p = lane // L, lane - (lane // L) * L
    seen = np.zeros((32, L), int)
    conflict_free = True
    for _ in range(L):
        assert (j < 32).all() and (p < L).all()
        seen[j, p] += 1
        bank = (p * LD + ((j * skew) & 31)) % 16          # 8-byte banks of a half-warp access
        for half in (bank[:16], bank[16:]):
            conflict_free &= len(set(half.tolist())) == 16
        p = p + dp
        j = j + dj
        wrap = p >= L
        p = np.where(wrap, p - L, p)
        j = np.where(wrap, j + 1, j)
    assert (seen == 1).all()                               # every entry of the [32][L] block exactly once
    if L % 2:
        assert conflict_free                               # odd L: the skew j -> (j L) mod 32 keeps the reads conflict-free
        assert sorted(((np.arange(32) * skew) & 31).tolist()) == list(range(32))   # and is a permutation of the columns
