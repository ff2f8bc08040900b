"""Parity of the rows either side of the hot path through the C ABI against the oracle:
DirichletConstraints (tools/dirichlet-constraints.hh), ConstMatrixOperator::apply / apply_inverse
(operators/matrix-based.hh:121-159), BilinearForm::apply2 norms (operators/bilinear-form.hh) and
default_interpolation (interpolations/default.hh).  Index sets and constrained matrices are bit-exact; floating-point
results within 1e-12 relative; the solves reproduce the reference's ESV2007 table and second-order convergence."""
import ctypes as C

import numpy as np
import pytest

from dune_gdt_b200 import descriptors as D
from helpers import TOL, rel_err

pytestmark = pytest.mark.gpu

CG, DG, FV = D.SPACE_CG, D.SPACE_DG, D.SPACE_FV


def laplace(kappa=1.0, **kw):
    return D.form(D.integrand(D.INT_LAPLACE, diffusion=kappa), **kw)


def mass(w=1.0, **kw):
    return D.form(D.integrand(D.INT_PRODUCT, diffusion=w), **kw)


# ---- n1: Dirichlet constraints -----------------------------------------------------------------------------------
@pytest.mark.parametrize("kind,order,n,periodic,mask", [
    (CG, 1, [9], 0, 0x3F), (CG, 1, [7, 5], 0, 0x3F), (CG, 2, [4, 3], 0, 0x3F), (CG, 3, [3, 2], 0, 0x3F),
    (CG, 1, [4, 3, 5], 0, 0x3F), (CG, 2, [3, 2, 2], 0, 0x3F), (CG, 1, [6, 4], 0, 0b0110), (CG, 2, [3, 3, 2], 0, 0b100001),
    (DG, 1, [5, 4], 0, 0x3F), (DG, 2, [3, 3], 0, 0x3F), (DG, 1, [3, 2, 2], 0, 0x3F), (DG, 1, [5, 4], 1, 0x3F),
    (FV, 0, [6, 5], 0, 0x3F), (DG, 0, [4, 4], 0, 0x3F),
])
def test_dirichlet_dofs_match_the_oracle(gdt, ctx, oracle, kind, order, n, periodic, mask):
    g = D.grid_desc(-1.0, 1.0, n, periodic)
    space = gdt.Space(gdt.Grid(ctx, g), kind, order)
    dc = gdt.DirichletConstraints(space, mask)
    ref = oracle.dirichlet_dofs(g, kind, order, mask)
    assert np.array_equal(dc.dirichlet_DoFs(), ref)


@pytest.mark.parametrize("only_clear", [False, True])
@pytest.mark.parametrize("ensure_symmetry", [True, False])
@pytest.mark.parametrize("kind,order,n,stencil", [
    (CG, 1, [6, 5], D.STENCIL_ELEMENT), (CG, 2, [3, 3, 2], D.STENCIL_ELEMENT), (CG, 1, [4, 3, 3], D.STENCIL_ELEMENT),
    (DG, 1, [5, 4], D.STENCIL_ELEMENT_AND_INTERSECTION),
])
def test_dirichlet_apply_matches_the_oracle(gdt, ctx, oracle, kind, order, n, stencil, only_clear, ensure_symmetry):
    g = D.grid_desc(-1.0, 1.0, n)
    space = gdt.Space(gdt.Grid(ctx, g), kind, order)
    op = gdt.make_matrix_operator(space, stencil)
    op.append(gdt.LocalElementIntegralBilinearForm(gdt.LocalLaplaceIntegrand(1.0) + gdt.LocalProductIntegrand(0.5)))
    src = D.fn_builtin(D.BUILTIN_QUADRATIC, 2, 1.0, 0.5)
    fun = gdt.make_vector_functional(space)
    fun.append(gdt.LocalElementIntegralFunctional(gdt.LocalProductIntegrand().with_ansatz(src)))
    op.append(fun)
    op.assemble()
    A, b = op.matrix(), fun.vector()
    dc = gdt.make_dirichlet_constraints(space)
    dc.apply(op, fun, only_clear=only_clear, ensure_symmetry=ensure_symmetry)
    v_ref, b_ref = oracle.dirichlet_apply(A.rowptr, A.colidx, A.values, b, dc.dirichlet_DoFs(), only_clear, ensure_symmetry)
    assert np.array_equal(op.values(), v_ref)  # untouched entries keep their bits, touched ones are exactly 0 / 1
    assert np.array_equal(fun.vector(), b_ref)


def test_dirichlet_apply_on_the_closed_form_q1_operator(gdt, ctx, oracle):
    """the pattern-free CG Q1 operator materialises its CSR pattern on first use"""
    lib = gdt.capi.lib()
    g = D.grid_desc(-1.0, 1.0, [5, 4, 3])
    space = gdt.Space(gdt.Grid(ctx, g), CG, 1)
    op_h = C.c_void_p()
    gdt.capi.check(lib.gdtb_matop_create(ctx._h, space._h, space._h, None, C.byref(op_h)))
    lap = laplace()
    gdt.capi.check(lib.gdtb_matop_append_element(op_h, C.byref(lap)))
    gdt.capi.check(lib.gdtb_assemble(op_h, None, D.ASSEMBLE_OVERWRITE))
    rp, ci = oracle.pattern(g, (CG, 1))
    values = np.empty(len(ci))
    gdt.capi.check(lib.gdtb_matop_values_download(op_h, gdt.capi.dptr(values)))
    dc = gdt.DirichletConstraints(space)
    gdt.capi.check(lib.gdtb_dirichlet_apply(dc._h, op_h, None, 0, 1))
    out = np.empty_like(values)
    gdt.capi.check(lib.gdtb_matop_values_download(op_h, gdt.capi.dptr(out)))
    ref, _ = oracle.dirichlet_apply(rp, ci, values, None, dc.dirichlet_DoFs())
    assert np.array_equal(out, ref)
    lib.gdtb_matop_destroy(op_h)


def test_unit_row_without_diagonal_is_an_error(gdt, ctx):
    """XT::LA unit_row throws when (i, i) is not in the pattern: an intersection-only stencil has no diagonal blocks"""
    g = D.grid_desc(0.0, 1.0, [4, 4])
    space = gdt.Space(gdt.Grid(ctx, g), DG, 1)
    op = gdt.make_matrix_operator(space, gdt.Stencil.intersection)
    dc = gdt.DirichletConstraints(space)
    with pytest.raises(gdt.capi.OperatorError):
        dc.apply(op)
    dc.apply(op, only_clear=True)  # clear_row / clear_col need no diagonal


# ---- n2: mat-vec and solves -----------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind,order,n,stencil", [
    (CG, 1, [33, 17], D.STENCIL_ELEMENT), (CG, 2, [5, 4, 3], D.STENCIL_ELEMENT), (CG, 1, [9, 8, 7], D.STENCIL_ELEMENT),
    (DG, 1, [12, 9], D.STENCIL_ELEMENT_AND_INTERSECTION), (CG, 3, [4, 3], D.STENCIL_ELEMENT), (CG, 1, [40], D.STENCIL_ELEMENT),
])
def test_matrix_operator_apply_matches_the_oracle(gdt, ctx, oracle, kind, order, n, stencil):
    g = D.grid_desc(-1.0, 1.0, n)
    space = gdt.Space(gdt.Grid(ctx, g), kind, order)
    op = gdt.make_matrix_operator(space, stencil)
    op.append(gdt.LocalElementIntegralBilinearForm(gdt.LocalLaplaceIntegrand(1.0) + gdt.LocalProductIntegrand(2.0)))
    op.assemble()
    A = op.matrix()
    x = np.random.default_rng(20251017).random(A.cols) - 0.5
    y = op.apply(x)
    assert rel_err(y, oracle.csr_mv(A.rowptr, A.colidx, A.values, x)) <= TOL
    assert np.array_equal(y, op.apply(x))  # deterministic
    with pytest.raises(gdt.capi.OperatorError):
        op.apply(x[:-1])


def heat_equation(gdt, ctx, n, order=1):
    """examples/stationary-heat-equation.cc:87-110 (d-dimensional): assemble, constrain, solve on the device"""
    d = len(n)
    g = D.grid_desc(-1.0, 1.0, n)
    space = gdt.make_continuous_lagrange_space(gdt.Grid(ctx, g), order)
    op = gdt.make_matrix_operator(space, gdt.Stencil.element)
    op.append(gdt.LocalElementIntegralBilinearForm(gdt.LocalLaplaceIntegrand(1.0)))
    src = D.fn_builtin(D.BUILTIN_COS_PRODUCT, 3, d * np.pi**2 / 4, np.pi / 2)
    fun = gdt.make_vector_functional(space)
    fun.append(gdt.LocalElementIntegralFunctional(gdt.LocalProductIntegrand().with_ansatz(src)))
    op.append(fun)
    op.assemble()
    dc = gdt.make_dirichlet_constraints(space)
    dc.apply(op, fun)
    return g, space, op, fun, dc


@pytest.mark.parametrize("precond", [D.PRECOND_NONE, D.PRECOND_JACOBI])
def test_cg_solve_matches_a_direct_solve(gdt, ctx, precond):
    import scipy.sparse.linalg as spla

    g, space, op, fun, dc = heat_equation(gdt, ctx, [24, 20])
    A, b = op.matrix(), fun.vector()
    u, info = op.apply_inverse(b, D.solver_opts(D.SOLVER_CG, precond, precision=1e-13))
    assert info.converged == 1 and 0 < info.iterations < 400
    ref = spla.spsolve(A.to_scipy().tocsc(), b)
    assert rel_err(u, ref) <= 1e-10
    assert np.all(u[dc.dirichlet_DoFs()] == 0.0)
    u2, info2 = op.apply_inverse(b, D.solver_opts(D.SOLVER_CG, precond, precision=1e-13))
    assert np.array_equal(u, u2) and info.iterations == info2.iterations  # run-to-run bit-identical


def test_heat_equation_example_on_the_device(gdt, ctx, oracle):
    """config 1 end to end (assemble, constrain, solve, norms), errors against cos(pi x/2) cos(pi y/2): O(h), O(h^2);
    the norms agree with the oracle's BilinearForm::apply2 on the same DoF vector"""
    exact = D.fn_builtin(D.BUILTIN_COS_PRODUCT, 3, 1.0, np.pi / 2)
    h1, l2 = [], []
    for N in (16, 32, 64):
        g, space, op, fun, dc = heat_equation(gdt, ctx, [N, N])
        u, info = op.apply_inverse(fun.vector(), D.solver_opts(precision=1e-12))
        assert info.converged
        h1_prod = gdt.make_bilinear_form(space, u, exact)
        h1_prod += gdt.LocalElementIntegralBilinearForm(gdt.LocalLaplaceIntegrand(1.0))
        l2_prod = gdt.make_bilinear_form(space, u, exact)
        l2_prod += gdt.LocalElementIntegralBilinearForm(gdt.LocalProductIntegrand())
        h1.append(np.sqrt(h1_prod.apply2()))
        l2.append(np.sqrt(l2_prod.apply2()))
        assert h1[-1] ** 2 == pytest.approx(oracle.bilinear_form_apply2(g, CG, 1, u, exact, laplace()), rel=1e-12)
        assert l2[-1] ** 2 == pytest.approx(oracle.bilinear_form_apply2(g, CG, 1, u, exact, mass()), rel=1e-12)
    assert h1[0] / h1[1] == pytest.approx(2.0, rel=0.03) and h1[1] / h1[2] == pytest.approx(2.0, rel=0.03)
    assert l2[0] / l2[1] == pytest.approx(4.0, rel=0.03) and l2[1] / l2[2] == pytest.approx(4.0, rel=0.03)


def test_heat_equation_3d_q1_and_q2(gdt, ctx):
    exact = D.fn_builtin(D.BUILTIN_COS_PRODUCT, 3, 1.0, np.pi / 2)
    errs = {}
    for order, N in ((1, 8), (1, 16), (2, 4), (2, 8)):
        g, space, op, fun, dc = heat_equation(gdt, ctx, [N, N, N], order)
        u, info = op.apply_inverse(fun.vector(), D.solver_opts(precision=1e-12))
        assert info.converged
        l2 = gdt.make_bilinear_form(space, u, exact)
        l2 += gdt.LocalElementIntegralBilinearForm(gdt.LocalProductIntegrand())
        errs[(order, N)] = np.sqrt(l2.apply2())
    assert errs[(1, 8)] / errs[(1, 16)] == pytest.approx(4.0, rel=0.1)
    assert errs[(2, 4)] / errs[(2, 8)] == pytest.approx(8.0, rel=0.2)


def swipdg_operator(gdt, ctx, N, symmetry=1.0, sigma_inner=8.0, sigma_dirichlet=14.0):
    """test/stationary-heat-equation/ESV2007.hh:58-112 on the YaspGrid variant, h_I = |I|"""
    g = D.grid_desc(-1.0, 1.0, [N, N])
    space = gdt.make_discontinuous_lagrange_space(gdt.Grid(ctx, g), 1)
    op = gdt.make_matrix_operator(space, gdt.Stencil.element_and_intersection)
    op.append(gdt.LocalElementIntegralBilinearForm(gdt.LocalLaplaceIntegrand(1.0)))
    op.append(gdt.LocalCouplingIntersectionIntegralBilinearForm(
        gdt.LocalLaplaceIPDGIntegrands.InnerCoupling(symmetry, 1.0, 1.0)
        + gdt.LocalIPDGIntegrands.InnerPenalty(sigma_inner, 1.0, D.HI_VOLUME)))
    op.append(gdt.LocalIntersectionIntegralBilinearForm(
        gdt.LocalIPDGIntegrands.BoundaryPenalty(sigma_dirichlet, 1.0, D.HI_VOLUME)
        + gdt.LocalLaplaceIPDGIntegrands.DirichletCoupling(symmetry, 1.0)))
    force = D.fn_builtin(D.BUILTIN_COS_PRODUCT, 2, np.pi**2 / 2, np.pi / 2)
    fun = gdt.make_vector_functional(space)
    fun.append(gdt.LocalElementIntegralFunctional(gdt.LocalProductIntegrand().with_ansatz(force)))
    op.append(fun)
    op.assemble()
    return g, space, op, fun


def test_swipdg_esv2007_table_on_the_device(gdt, ctx):
    """stationary_heat_equation__ESV2007__table_1.mini:31 (YaspGrid, DG-Q1, 8^2 + 2 refinements): norm.H_1_semi =
    [2.52e-01 1.26e-01 6.30e-02] -- assembled, solved (CG) and measured (apply2) on the GPU"""
    exact = D.fn_builtin(D.BUILTIN_COS_PRODUCT, 4, 1.0, np.pi / 2)
    for N, ref in zip((8, 16, 32), [2.52e-01, 1.26e-01, 6.30e-02]):
        g, space, op, fun = swipdg_operator(gdt, ctx, N)
        u, info = op.apply_inverse(fun.vector(), D.solver_opts(D.SOLVER_CG, D.PRECOND_JACOBI, precision=1e-12))
        assert info.converged
        h1 = gdt.make_bilinear_form(space, u, exact)
        h1 += gdt.LocalElementIntegralBilinearForm(gdt.LocalLaplaceIntegrand(1.0))
        assert np.sqrt(h1.apply2()) == pytest.approx(ref, rel=6e-3)


def test_bicgstab_solves_the_unsymmetric_nipdg_system(gdt, ctx):
    import scipy.sparse.linalg as spla

    g, space, op, fun = swipdg_operator(gdt, ctx, 12, symmetry=-1.0)
    A, b = op.matrix(), fun.vector()
    S = A.to_scipy()
    assert abs(S - S.T).max() > 1e-3  # really unsymmetric
    u, info = op.apply_inverse(b, D.solver_opts(D.SOLVER_BICGSTAB, D.PRECOND_JACOBI, precision=1e-13))
    assert info.converged
    assert rel_err(u, spla.spsolve(S.tocsc(), b)) <= 1e-9


def test_solver_failure_is_an_operator_error(gdt, ctx):
    g, space, op, fun, dc = heat_equation(gdt, ctx, [32, 32])
    # (the example's own right-hand side is almost a discrete eigenvector: CG needs a handful of iterations for it)
    b = np.random.default_rng(20251017).random(space.mapper.size)
    with pytest.raises(gdt.capi.OperatorError):
        op.apply_inverse(b, D.solver_opts(precision=1e-14, max_iter=4, check_every=2))
    u, info = op.apply_inverse(b, D.solver_opts(precision=1e-12))
    assert info.converged and info.iterations > 4


# ---- n4: norms and interpolation ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind,order,n", [(CG, 1, [7, 6]), (CG, 2, [4, 3, 2]), (DG, 1, [5, 5]), (CG, 3, [3, 3]), (CG, 1, [10])])
def test_apply2_matches_the_oracle(gdt, ctx, oracle, kind, order, n):
    g = D.grid_desc(-1.0, 1.0, n)
    space = gdt.Space(gdt.Grid(ctx, g), kind, order)
    rng = np.random.default_rng(20251017)
    u = rng.random(space.mapper.size)
    f = D.fn_builtin(D.BUILTIN_COS_PRODUCT, 3, 0.7, 1.3)
    kap = D.fn_const(np.array([[2.0, 0.3, 0.1], [0.3, 1.5, 0.2], [0.1, 0.2, 1.1]])[: len(n), : len(n)])
    for form in (laplace(), mass(), laplace(kap, over_integrate=1), D.form([D.integrand(D.INT_LAPLACE, diffusion=1.0),
                                                                            D.integrand(D.INT_PRODUCT, diffusion=3.0)], scaling=0.5)):
        for dofs, fn in ((u, f), (u, None), (None, f)):
            ref = oracle.bilinear_form_apply2(g, kind, order, dofs, fn, form)
            bf = gdt.make_bilinear_form(space, dofs, fn)
            bf += form
            assert bf.apply2() == pytest.approx(ref, rel=1e-12)


@pytest.mark.parametrize("kind,order,n", [(CG, 1, [9, 7]), (CG, 2, [4, 3, 3]), (DG, 2, [5, 4]), (CG, 3, [6]), (DG, 0, [4, 4])])
def test_lagrange_interpolation_matches_the_oracle(gdt, ctx, oracle, kind, order, n):
    g = D.grid_desc([-1.0] * len(n), [1.0, 0.5, 2.0][: len(n)], n)
    space = gdt.Space(gdt.Grid(ctx, g), kind, order)
    for f in (D.fn_builtin(D.BUILTIN_QUADRATIC, 2, 0.25, 1.5), D.fn_builtin(D.BUILTIN_COS_PRODUCT, 3, 1.0, np.pi / 2),
              D.fn_builtin(D.BUILTIN_INDICATOR, 0, -0.5, 0.25)):
        out = gdt.default_interpolation(f, space)
        assert rel_err(out, oracle.lagrange_interpolate(g, kind, order, f)) <= 1e-14
