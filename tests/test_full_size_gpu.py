"""BASELINE.json's full sizes on the device (the oracle would take minutes to hours there): size-independent properties
of what the CUDA path produced, evaluated with the library's own CSR mat-vec (gdtb_matop_apply) and device reductions.

  C2  3D Q1 Laplace + RHS, 256^3          constants in the kernel, symmetry, sum(b) = |Omega| for f = 1, bit-identical reruns
  C5  3D Q2 Laplace, 128^3                constants in the kernel, symmetry, pattern-free == pattern-based values
  C3  2D SWIPDG DG-Q1, 2048^2             symmetry (symmetric IP), constants in the kernel away from the Dirichlet boundary
"""
import ctypes as C

import numpy as np
import pytest

from dune_gdt_b200 import descriptors as D

pytestmark = pytest.mark.gpu


def _as_tensor(torch, ptr, n):
    class _Arr:
        __cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 3}

    return torch.as_tensor(_Arr(), device="cuda")


def _matvec(gdt, ctx, op_h, x):
    import torch

    y = torch.empty_like(x)
    torch.cuda.synchronize()
    gdt.capi.check(gdt.capi.lib().gdtb_matop_apply(op_h, C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr())))
    ctx.synchronize()
    return y


def _values(gdt, torch, op_h, nnz):
    p = C.c_void_p()
    gdt.capi.check(gdt.capi.lib().gdtb_matop_values_device(op_h, C.byref(p)))
    return _as_tensor(torch, p.value, nnz)


def _symmetry_and_kernel(gdt, ctx, torch, op_h, rows, scale, interior=None):
    g = torch.Generator(device="cuda").manual_seed(20251017)
    x = torch.rand(rows, dtype=torch.float64, device="cuda", generator=g) - 0.5
    y = torch.rand(rows, dtype=torch.float64, device="cuda", generator=g) - 0.5
    Ax, Ay = _matvec(gdt, ctx, op_h, x), _matvec(gdt, ctx, op_h, y)
    lhs, rhs = torch.dot(Ax, y).item(), torch.dot(Ay, x).item()
    assert abs(lhs - rhs) <= 1e-11 * (Ax.abs().max().item() * rows**0.5), (lhs, rhs)  # (Ax, y) == (x, Ay)
    A1 = _matvec(gdt, ctx, op_h, torch.ones(rows, dtype=torch.float64, device="cuda"))
    if interior is not None:
        A1 = A1[interior]
    assert A1.abs().max().item() <= 1e-12 * scale * 64  # row sums vanish (<= 125 entries per row)


def test_c2_q1_256_cubed(gdt, ctx):
    import torch

    lib, check = gdt.capi.lib(), gdt.capi.check
    n = 256
    grid = gdt.make_cube_grid(ctx, -1.0, 1.0, [n, n, n])
    space = gdt.make_continuous_lagrange_space(grid, 1)
    op_h, fun_h = C.c_void_p(), C.c_void_p()
    check(lib.gdtb_matop_create(ctx._h, space._h, space._h, None, C.byref(op_h)))
    check(lib.gdtb_vecfun_create(ctx._h, space._h, C.byref(fun_h)))
    lap = D.form(D.integrand(D.INT_LAPLACE, diffusion=1.0))
    one = D.form(D.integrand(D.INT_PRODUCT, diffusion=1.0, weight=D.fn_const(1.0)))
    check(lib.gdtb_matop_append_element(op_h, C.byref(lap)))
    check(lib.gdtb_vecfun_append_element(fun_h, C.byref(one)))
    check(lib.gdtb_assemble(op_h, fun_h, D.ASSEMBLE_OVERWRITE))
    rows, nnz = (n + 1) ** 3, (3 * n + 1) ** 3
    assert lib.gdtb_matop_local_nnz(op_h) == nnz
    vals = _values(gdt, torch, op_h, nnz)
    first = vals.clone()
    scale = vals.abs().max().item()
    h = 2.0 / n
    assert scale == pytest.approx(8 * h / 3, rel=1e-12)  # closed-form interior diagonal
    pb = C.c_void_p()
    check(lib.gdtb_vecfun_device(fun_h, C.byref(pb)))
    b = _as_tensor(torch, pb.value, rows)
    assert b.sum().item() == pytest.approx(8.0, rel=1e-12)  # sum_i (1, phi_i) = |Omega|
    _symmetry_and_kernel(gdt, ctx, torch, op_h, rows, scale)
    check(lib.gdtb_assemble(op_h, fun_h, D.ASSEMBLE_OVERWRITE))
    assert torch.equal(_values(gdt, torch, op_h, nnz), first)  # run-to-run bit-identical
    del vals, first, b
    lib.gdtb_matop_destroy(op_h)
    lib.gdtb_vecfun_destroy(fun_h)
    torch.cuda.empty_cache()


def test_c5_q2_128_cubed(gdt, ctx):
    import torch

    lib, check = gdt.capi.lib(), gdt.capi.check
    n = 128
    grid = gdt.make_cube_grid(ctx, -1.0, 1.0, [n, n, n])
    space = gdt.make_continuous_lagrange_space(grid, 2)
    pat = gdt.make_sparsity_pattern(space, space, gdt.Stencil.element)
    assert pat.nnz == (8 * n + 1) ** 3 and pat.rows == (2 * n + 1) ** 3  # per axis: 5 n + 1 (even) + 3 n (odd lattice points)
    op = gdt.MatrixOperator(space, space, pat)
    lap = D.form(D.integrand(D.INT_LAPLACE, diffusion=1.0))
    check(lib.gdtb_matop_append_element(op._h, C.byref(lap)))
    check(lib.gdtb_assemble(op._h, None, D.ASSEMBLE_OVERWRITE))
    assert op.plan == "q2_gather"
    vals = _values(gdt, torch, op._h, pat.nnz)
    scale = vals.abs().max().item()
    _symmetry_and_kernel(gdt, ctx, torch, op._h, pat.rows, scale)
    # the pattern-free operator (closed-form CSR positions only) writes the same values at the same places
    free_h = C.c_void_p()
    check(lib.gdtb_matop_create(ctx._h, space._h, space._h, None, C.byref(free_h)))
    check(lib.gdtb_matop_append_element(free_h, C.byref(lap)))
    check(lib.gdtb_assemble(free_h, None, D.ASSEMBLE_OVERWRITE))
    assert lib.gdtb_matop_local_nnz(free_h) == pat.nnz
    assert torch.equal(_values(gdt, torch, free_h, pat.nnz), vals)
    del vals
    lib.gdtb_matop_destroy(free_h)
    del op, pat
    torch.cuda.empty_cache()


def test_c3_swipdg_2048_squared(gdt, ctx):
    import torch

    lib, check = gdt.capi.lib(), gdt.capi.check
    n = 2048
    grid = gdt.make_cube_grid(ctx, -1.0, 1.0, [n, n])
    space = gdt.make_discontinuous_lagrange_space(grid, 1)
    pat = gdt.make_sparsity_pattern(space, space, gdt.Stencil.element_and_intersection)
    assert pat.nnz == 16 * (n * n + 4 * n * (n - 1)) and pat.rows == 4 * n * n
    op = gdt.MatrixOperator(space, space, pat)
    lap = D.form(D.integrand(D.INT_LAPLACE, diffusion=1.0))
    inner = D.form([D.integrand(D.INT_IPDG_INNER_COUPLING, prefactor=1.0, diffusion=1.0, weight=1.0),
                    D.integrand(D.INT_IPDG_INNER_PENALTY, prefactor=8.0, weight=1.0, hI_kind=D.HI_VOLUME)])
    bnd = D.form([D.integrand(D.INT_IPDG_DIRICHLET_COUPLING, prefactor=1.0, diffusion=1.0),
                  D.integrand(D.INT_IPDG_BOUNDARY_PENALTY, prefactor=14.0, weight=1.0, hI_kind=D.HI_VOLUME)])
    check(lib.gdtb_matop_append_element(op._h, C.byref(lap)))
    check(lib.gdtb_matop_append_coupling(op._h, C.byref(inner), D.FILTER_INNER_ONCE))
    check(lib.gdtb_matop_append_boundary(op._h, C.byref(bnd), D.FILTER_ALL_BOUNDARY))
    check(lib.gdtb_assemble(op._h, None, D.ASSEMBLE_OVERWRITE))
    assert op.plan == "dg_gather"
    vals = _values(gdt, torch, op._h, pat.nnz)
    scale = vals.abs().max().item()
    # rows of elements that do not touch the Dirichlet boundary: constants are in the kernel of the SWIPDG form there
    e = torch.arange(n * n, device="cuda")
    ex, ey = e % n, e // n
    inner_elems = (ex > 0) & (ex < n - 1) & (ey > 0) & (ey < n - 1)
    interior = inner_elems.repeat_interleave(4)
    _symmetry_and_kernel(gdt, ctx, torch, op._h, pat.rows, scale, interior)
    del vals
    del op, pat
    torch.cuda.empty_cache()
