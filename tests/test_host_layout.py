"""Host logic of the closed-form CSR geometry (no GPU): the row pointers the gather kernels / structured pattern
generators compute in closed form, and the row ranges of slabs, against the oracle's restatement of the reference's
pattern builders (tools/sparsity-pattern.hh:34-144)."""
import ctypes as C

import numpy as np
import pytest

from dune_gdt_b200 import descriptors as D

CG, DG = D.SPACE_CG, D.SPACE_DG


def host_rowptr(gdt, gdesc, kind, order, size):
    rp = np.empty(size + 1, dtype=np.int64)
    gdt.capi.check(gdt.capi.lib().gdtb_host_closed_form_rowptr(C.byref(gdesc), kind, order, rp.ctypes.data_as(C.POINTER(C.c_int64))))
    return rp


@pytest.mark.parametrize("kind,order,stencil,n", [
    (CG, 1, D.STENCIL_ELEMENT, [1]), (CG, 1, D.STENCIL_ELEMENT, [9]), (CG, 1, D.STENCIL_ELEMENT, [1, 1]), (CG, 1, D.STENCIL_ELEMENT, [9, 2]),
    (CG, 1, D.STENCIL_ELEMENT, [1, 1, 1]), (CG, 1, D.STENCIL_ELEMENT, [5, 4, 3]), (CG, 1, D.STENCIL_ELEMENT, [2, 1, 7]),
    (CG, 2, D.STENCIL_ELEMENT, [1, 1]), (CG, 2, D.STENCIL_ELEMENT, [9, 2]), (CG, 2, D.STENCIL_ELEMENT, [2, 5]),
    (CG, 2, D.STENCIL_ELEMENT, [1, 1, 1]), (CG, 2, D.STENCIL_ELEMENT, [4, 3, 2]), (CG, 2, D.STENCIL_ELEMENT, [1, 3, 2]), (CG, 2, D.STENCIL_ELEMENT, [2, 2, 5]),
    (DG, 1, D.STENCIL_ELEMENT_AND_INTERSECTION, [1]), (DG, 1, D.STENCIL_ELEMENT_AND_INTERSECTION, [7]), (DG, 2, D.STENCIL_ELEMENT_AND_INTERSECTION, [7]),
    (DG, 1, D.STENCIL_ELEMENT_AND_INTERSECTION, [1, 1]), (DG, 1, D.STENCIL_ELEMENT_AND_INTERSECTION, [6, 5]), (DG, 1, D.STENCIL_ELEMENT_AND_INTERSECTION, [1, 4]),
    (DG, 2, D.STENCIL_ELEMENT_AND_INTERSECTION, [3, 4]), (DG, 1, D.STENCIL_ELEMENT_AND_INTERSECTION, [1, 1, 1]),
    (DG, 1, D.STENCIL_ELEMENT_AND_INTERSECTION, [4, 3, 3]), (DG, 1, D.STENCIL_ELEMENT_AND_INTERSECTION, [2, 1, 3]), (DG, 0, D.STENCIL_ELEMENT_AND_INTERSECTION, [5, 4]),
])
def test_closed_form_rowptr_matches_the_reference_pattern(gdt, oracle, kind, order, stencil, n):
    gdesc = D.grid_desc(0.0, 1.0, n)
    rp, _ = oracle.pattern(gdesc, (kind, order), stencil=stencil)
    got = host_rowptr(gdt, gdesc, kind, order, rp.size - 1)
    assert np.array_equal(got, rp)


@pytest.mark.parametrize("n,periodic", [([7], 1), ([3], 1), ([6, 5], 3), ([6, 5], 1), ([5, 4], 2), ([4, 3, 3], 7),
                                        ([4, 3, 5], 5), ([3, 3, 4], 2)])
@pytest.mark.parametrize("order", [0, 1, 2])
def test_closed_form_rowptr_on_periodic_dg_grids(gdt, oracle, n, periodic, order):
    """periodic directions with >= 3 cells: every element has two neighbours there, the wrap neighbour's block is ordered
    by its element index (tools/sparsity-pattern.hh:74-95 through the periodic view)"""
    gdesc = D.grid_desc(0.0, 1.0, n, periodic)
    rp, _ = oracle.pattern(gdesc, (DG, order), stencil=D.STENCIL_ELEMENT_AND_INTERSECTION)
    got = host_rowptr(gdt, gdesc, DG, order, rp.size - 1)
    assert np.array_equal(got, rp)


def test_closed_form_rowptr_at_the_benchmark_sizes(gdt):
    """nnz of BASELINE's configurations from the closed forms alone: no pattern is ever built"""
    for n, kind, order, rows, nnz in (([128, 128], CG, 1, 16641, 148225), ([64, 64, 64], CG, 1, 65**3, 193**3),
                                      ([32, 32, 32], CG, 2, 65**3, 257**3), ([256, 256], DG, 1, 4 * 256**2, 16 * (256**2 + 4 * 256 * 255))):
        rp = host_rowptr(gdt, D.grid_desc(-1.0, 1.0, n), kind, order, rows)
        assert rp[0] == 0 and rp[-1] == nnz and np.all(np.diff(rp) > 0)


@pytest.mark.parametrize("order,n,cuts", [(1, [6, 5, 8], [0, 3, 8]), (1, [7, 6], [0, 2, 5, 6]), (1, [12], [0, 5, 12]),
                                          (2, [4, 3, 6], [0, 2, 6]), (2, [3, 4, 7], [0, 1, 3, 7]), (2, [5, 6], [0, 2, 6]), (2, [3, 2, 4], [0, 1, 2, 3, 4])])
def test_slab_row_ranges_tile_the_global_matrix(gdt, oracle, order, n, cuts):
    """owner-computes-rows slabs (gdtb_matop_set_slab): the row ranges of all slabs cover every global row exactly once,
    their CSR offsets and counts are the oracle pattern's (Q1: one range per slab, Q2: one per sub-entity group)"""
    gdesc = D.grid_desc(-1.0, 1.0, n)
    rp, _ = oracle.pattern(gdesc, (CG, order))
    lib = gdt.capi.lib()
    seen = np.zeros(rp.size - 1, dtype=int)
    total = 0
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        cnt = C.c_int32()
        gdt.capi.check(lib.gdtb_host_slab_row_ranges(C.byref(gdesc), CG, order, lo, hi, 0, None, None, None, None, C.byref(cnt)))
        assert cnt.value == (1 if order == 1 else 2 ** len(n))
        arrs = [np.zeros(cnt.value, dtype=np.int64) for _ in range(4)]
        ptr = [a.ctypes.data_as(C.POINTER(C.c_int64)) for a in arrs]
        gdt.capi.check(lib.gdtb_host_slab_row_ranges(C.byref(gdesc), CG, order, lo, hi, cnt.value, *ptr, C.byref(cnt)))
        for rb, re_, vo, vc in zip(*arrs):
            assert vo == rp[rb] and vc == rp[re_] - rp[rb]
            seen[rb:re_] += 1
            total += vc
    assert (seen == 1).all() and total == rp[-1]


def test_host_layout_error_conventions(gdt):
    lib = gdt.capi.lib()
    rp = np.zeros(64, dtype=np.int64)
    p = rp.ctypes.data_as(C.POINTER(C.c_int64))
    per = D.grid_desc(0.0, 1.0, [4, 2], 3)
    assert lib.gdtb_host_closed_form_rowptr(C.byref(per), DG, 1, p) == 7  # periodic with < 3 cells: sort-and-unique only
    g1 = D.grid_desc(0.0, 1.0, [4])
    assert lib.gdtb_host_closed_form_rowptr(C.byref(g1), CG, 2, p) == 7  # CG Q2 closed forms are 2D / 3D
    assert lib.gdtb_host_closed_form_rowptr(C.byref(g1), CG, 0, p) == 5  # space_error: CG needs order >= 1
    n = C.c_int32()
    g3 = D.grid_desc(0.0, 1.0, [4, 4, 4])
    assert lib.gdtb_host_slab_row_ranges(C.byref(g3), CG, 1, 3, 2, 0, None, None, None, None, C.byref(n)) == 1
