"""Oracle-vs-GPU parity AT the benchmarked sizes (VERDICT r01 "weak #1"): the CUDA path assembles BASELINE.json's full
configurations, the CPU oracle assembles matching SUB-GRIDS (slabs of a few element layers at the bottom, in the middle
and at the top of the grid, same `lower + i h` coordinates -- h is a power of two for all three configurations, so the
sub-grid's cells are bit-identical to the global grid's), and every row that is complete inside a slab (it does not touch
an artificial cut) is compared: pattern segment bit-exact, values / right-hand side <= 1e-12 (norm-wise).

  C2  3D Q1 Laplace + RHS, 256^3, bench.py's own forms (kappa = 1, cos-product source of declared order 3:
      the sum-factorised `sf3` branch + the separable right-hand-side tables)
  C5  3D Q2 Laplace, 128^3 (sum-factorised Q2 gather, MCMG numbering)
  C3  2D SWIPDG DG-Q1, 2048^2 (factorised DG gather with constant-coefficient tables)
"""
import ctypes as C
import os

import numpy as np
import pytest

from dune_gdt_b200 import descriptors as D
from helpers import TOL, rel_err

pytestmark = pytest.mark.gpu

THREADS = max(1, min(os.cpu_count() or 1, 32))


def _tensor(torch, ptr, n, dtype="<f8"):
    class _Arr:
        __cuda_array_interface__ = {"shape": (n,), "typestr": dtype, "data": (ptr, False), "version": 3}

    return torch.as_tensor(_Arr(), device="cuda")


def _device_csr(gdt, torch, op_h, rows, nnz):
    """device views of the operator's values and of the CSR pattern they follow"""
    lib, check = gdt.capi.lib(), gdt.capi.check
    pv, prp, pci = C.c_void_p(), C.c_void_p(), C.c_void_p()
    check(lib.gdtb_matop_values_device(op_h, C.byref(pv)))
    check(lib.gdtb_matop_pattern_device(op_h, C.byref(prp), C.byref(pci)))
    return (_tensor(torch, pv.value, nnz), _tensor(torch, prp.value, rows + 1, "<i8"), _tensor(torch, pci.value, nnz, "<i4"))


def _slabs(n_last, thickness):
    return [0, (n_last - thickness) // 2, n_last - thickness]


def test_c2_q1_256_cubed_bench_forms_vs_oracle_slabs(gdt, ctx, oracle):
    import torch

    import bench  # the forms of the benchmark itself

    lib, check = gdt.capi.lib(), gdt.capi.check
    n, T = 256, 6
    h = 2.0 / n
    grid = gdt.make_cube_grid(ctx, -1.0, 1.0, [n, n, n])
    space = gdt.make_continuous_lagrange_space(grid, 1)
    op_h, fun_h = C.c_void_p(), C.c_void_p()
    check(lib.gdtb_matop_create(ctx._h, space._h, space._h, None, C.byref(op_h)))
    check(lib.gdtb_vecfun_create(ctx._h, space._h, C.byref(fun_h)))
    lap, rhs = bench.forms()
    check(lib.gdtb_matop_append_element(op_h, C.byref(lap)))
    check(lib.gdtb_vecfun_append_element(fun_h, C.byref(rhs)))
    assert lib.gdtb_matop_plan(op_h).decode() == "q1_gather"
    check(lib.gdtb_assemble(op_h, fun_h, D.ASSEMBLE_OVERWRITE))
    rows, nnz = (n + 1) ** 3, (3 * n + 1) ** 3
    vals, rp_g, ci_g = _device_csr(gdt, torch, op_h, rows, nnz)
    pb = C.c_void_p()
    check(lib.gdtb_vecfun_device(fun_h, C.byref(pb)))
    b_g = _tensor(torch, pb.value, rows)
    layer = (n + 1) ** 2
    for z0 in _slabs(n, T):
        sub = D.grid_desc([-1.0, -1.0, -1.0 + z0 * h], [1.0, 1.0, -1.0 + (z0 + T) * h], [n, n, T])
        rp, ci = oracle.pattern(sub, (D.SPACE_CG, 1))
        ref_v, ref_b = oracle.assemble(sub, D.SPACE_CG, 1, rp, ci, [lap], rhs_forms=[rhs], num_threads=THREADS)
        # vertex layers of the slab that are complete: not on an artificial cut
        l_lo = 0 if z0 == 0 else 1
        l_hi = T + 1 if z0 + T == n else T  # exclusive
        a, b = l_lo * layer, l_hi * layer  # local row range
        off = z0 * layer
        ga, gb = int(rp_g[a + off].item()), int(rp_g[b + off].item())
        assert gb - ga == rp[b] - rp[a]
        assert np.array_equal(rp_g[a + off : b + off + 1].cpu().numpy() - ga, rp[a : b + 1] - rp[a])
        assert np.array_equal(ci_g[ga:gb].cpu().numpy(), ci[rp[a] : rp[b]] + off)
        assert rel_err(vals[ga:gb].cpu().numpy(), ref_v[rp[a] : rp[b]]) <= TOL
        assert rel_err(b_g[a + off : b + off].cpu().numpy(), ref_b[a:b]) <= TOL
    del vals, rp_g, ci_g, b_g
    lib.gdtb_matop_destroy(op_h)
    lib.gdtb_vecfun_destroy(fun_h)
    torch.cuda.empty_cache()


def test_c3_swipdg_2048_squared_vs_oracle_slabs(gdt, ctx, oracle):
    import torch

    lib, check = gdt.capi.lib(), gdt.capi.check
    n, T = 2048, 6
    h = 2.0 / n
    grid = gdt.make_cube_grid(ctx, -1.0, 1.0, [n, n])
    space = gdt.make_discontinuous_lagrange_space(grid, 1)
    pat = gdt.make_sparsity_pattern(space, space, gdt.Stencil.element_and_intersection)
    op = gdt.MatrixOperator(space, space, pat)
    lap = D.form(D.integrand(D.INT_LAPLACE, diffusion=1.0))
    inner = D.form([D.integrand(D.INT_IPDG_INNER_COUPLING, prefactor=1.0, diffusion=1.0, weight=1.0),
                    D.integrand(D.INT_IPDG_INNER_PENALTY, prefactor=8.0, weight=1.0, hI_kind=D.HI_VOLUME)])
    bnd = D.form([D.integrand(D.INT_IPDG_DIRICHLET_COUPLING, prefactor=1.0, diffusion=1.0),
                  D.integrand(D.INT_IPDG_BOUNDARY_PENALTY, prefactor=14.0, weight=1.0, hI_kind=D.HI_VOLUME)])
    check(lib.gdtb_matop_append_element(op._h, C.byref(lap)))
    check(lib.gdtb_matop_append_coupling(op._h, C.byref(inner), D.FILTER_INNER_ONCE))
    check(lib.gdtb_matop_append_boundary(op._h, C.byref(bnd), D.FILTER_ALL_BOUNDARY))
    assert op.plan == "dg_gather"
    check(lib.gdtb_assemble(op._h, None, D.ASSEMBLE_OVERWRITE))
    vals, rp_g, ci_g = _device_csr(gdt, torch, op._h, pat.rows, pat.nnz)
    for y0 in _slabs(n, T):
        sub = D.grid_desc([-1.0, -1.0 + y0 * h], [1.0, -1.0 + (y0 + T) * h], [n, T])
        rp, ci = oracle.pattern(sub, (D.SPACE_DG, 1), stencil=D.STENCIL_ELEMENT_AND_INTERSECTION)
        ref_v, _ = oracle.assemble(sub, D.SPACE_DG, 1, rp, ci, [lap], [inner], [bnd], num_threads=THREADS)
        # element rows whose faces are all real: not adjacent to an artificial cut (there the oracle sees a Dirichlet
        # boundary face instead of an inner face)
        e_lo = 0 if y0 == 0 else 1
        e_hi = T if y0 + T == n else T - 1
        a, b = 4 * n * e_lo, 4 * n * e_hi
        off = 4 * n * y0
        ga, gb = int(rp_g[a + off].item()), int(rp_g[b + off].item())
        assert gb - ga == rp[b] - rp[a]
        assert np.array_equal(rp_g[a + off : b + off + 1].cpu().numpy() - ga, rp[a : b + 1] - rp[a])
        assert np.array_equal(ci_g[ga:gb].cpu().numpy(), ci[rp[a] : rp[b]] + off)
        assert rel_err(vals[ga:gb].cpu().numpy(), ref_v[rp[a] : rp[b]]) <= TOL
    del vals, rp_g, ci_g, op, pat
    torch.cuda.empty_cache()


def test_c5_q2_128_cubed_vs_oracle_slabs(gdt, ctx, oracle):
    import torch

    lib, check = gdt.capi.lib(), gdt.capi.check
    n, T = 128, 3
    h = 2.0 / n
    gdesc = D.grid_desc(-1.0, 1.0, [n, n, n])
    grid = gdt.make_cube_grid(ctx, -1.0, 1.0, [n, n, n])
    space = gdt.make_continuous_lagrange_space(grid, 2)
    op_h = C.c_void_p()
    check(lib.gdtb_matop_create(ctx._h, space._h, space._h, None, C.byref(op_h)))  # pattern-free: closed-form CSR
    lap = D.form(D.integrand(D.INT_LAPLACE, diffusion=1.0))
    check(lib.gdtb_matop_append_element(op_h, C.byref(lap)))
    assert lib.gdtb_matop_plan(op_h).decode() == "q2_gather"
    check(lib.gdtb_assemble(op_h, None, D.ASSEMBLE_OVERWRITE))
    rows, nnz = (2 * n + 1) ** 3, (8 * n + 1) ** 3
    vals, rp_dev, ci_g = _device_csr(gdt, torch, op_h, rows, nnz)
    rp_g = rp_dev.cpu().numpy()
    for z0 in _slabs(n, T):
        sub = D.grid_desc([-1.0, -1.0, -1.0 + z0 * h], [1.0, 1.0, -1.0 + (z0 + T) * h], [n, n, T])
        rp, ci = oracle.pattern(sub, (D.SPACE_CG, 2))
        ref_v, _ = oracle.assemble(sub, D.SPACE_CG, 2, rp, ci, [lap], num_threads=THREADS)
        # DoF map sub-grid -> global grid through the elements (same local DoF of the same element), and the lattice
        # layer of every sub-grid DoF (the MCMG numbering groups DoFs by sub-entity kind: the map is not monotone)
        n_sub = rp.size - 1
        M = np.full(n_sub, -1, dtype=np.int64)
        zl = np.full(n_sub, -1, dtype=np.int64)
        az = np.arange(27) // 9
        for e in range(n * n * T):
            gs = oracle.global_indices(sub, D.SPACE_CG, 2, e)
            gg = oracle.global_indices(gdesc, D.SPACE_CG, 2, e + z0 * n * n)
            M[gs] = gg
            zl[gs] = 2 * (e // (n * n)) + az
        assert (M >= 0).all()
        complete = ((zl > 0) | (z0 == 0)) & ((zl < 2 * T) | (z0 + T == n))
        rows_s = np.nonzero(complete)[0]
        rows_m = M[rows_s]
        order = np.argsort(rows_m, kind="stable")
        rows_s, rows_m = rows_s[order], rows_m[order]
        len_s = rp[rows_s + 1] - rp[rows_s]
        assert np.array_equal(len_s, rp_g[rows_m + 1] - rp_g[rows_m])  # same row lengths
        # entry positions: sub-grid rows in the order of their global rows; global rows ascending
        starts = np.cumsum(len_s) - len_s
        within = np.arange(len_s.sum()) - np.repeat(starts, len_s)
        pos_s = np.repeat(rp[rows_s], len_s) + within
        pos_g = np.repeat(rp_g[rows_m], len_s) + within
        row_id = np.repeat(np.arange(rows_s.size), len_s)
        col_m = M[ci[pos_s]]
        perm = np.lexsort((col_m, row_id))  # sort every sub-grid row by its mapped (global) column index
        idx = torch.from_numpy(pos_g).cuda()
        assert np.array_equal(ci_g[idx].cpu().numpy().astype(np.int64), col_m[perm])
        assert rel_err(vals[idx].cpu().numpy(), ref_v[pos_s[perm]]) <= TOL
        del idx
    del vals, rp_dev, ci_g
    lib.gdtb_matop_destroy(op_h)
    torch.cuda.empty_cache()
