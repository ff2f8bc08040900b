"""Parity of the CUDA path (through the C ABI) against the CPU restatement oracle, same seeded inputs.

Bar (BASELINE.json north_star): sparsity pattern and DoF mapping bit-exact; matrix, RHS and FV update values within
1e-12 relative (norm-wise metric of SURVEY.md section 8c); run-to-run bit-identical results.
"""
import ctypes as C

import numpy as np
import pytest

from dune_gdt_b200 import descriptors as D
from helpers import TOL, rel_err

pytestmark = pytest.mark.gpu

CG, DG, FV = D.SPACE_CG, D.SPACE_DG, D.SPACE_FV
SEED = 20251017


def make_space(gdt, ctx, gdesc, kind, order):
    grid = gdt.Grid(ctx, gdesc)
    return gdt.Space(grid, kind, order)


def gpu_assemble(gdt, ctx, gdesc, kind, order, stencil, element=(), coupling=(), boundary=(), rhs=(),
                 pattern_method=D.PATTERN_AUTO, coupling_filter=D.FILTER_INNER_ONCE):
    """raw C-ABI path: returns (rowptr, colidx, values, rhs_vector, plan)"""
    lib = gdt.capi.lib()
    space = make_space(gdt, ctx, gdesc, kind, order)
    pat = gdt.SparsityPattern(space, space, stencil, pattern_method)
    op = gdt.MatrixOperator(space, space, pat)
    for f in element:
        gdt.capi.check(lib.gdtb_matop_append_element(op._h, C.byref(f)))
    for f in coupling:
        gdt.capi.check(lib.gdtb_matop_append_coupling(op._h, C.byref(f), coupling_filter))
    for f in boundary:
        gdt.capi.check(lib.gdtb_matop_append_boundary(op._h, C.byref(f), D.FILTER_ALL_BOUNDARY))
    fun = gdt.VectorBasedFunctional(space)
    for f in rhs:
        gdt.capi.check(lib.gdtb_vecfun_append_element(fun._h, C.byref(f)))
    plan = op.plan
    has_op = bool(element or coupling or boundary)
    gdt.capi.check(lib.gdtb_assemble(op._h if has_op else None, fun._h if rhs else None, D.ASSEMBLE_OVERWRITE))
    rowptr, colidx = pat.download()
    return rowptr, colidx, op.values(), fun.vector(), plan


def laplace(kappa=1.0, **kw):
    return D.form(D.integrand(D.INT_LAPLACE, diffusion=kappa), **kw)


def mass(w=1.0, **kw):
    return D.form(D.integrand(D.INT_PRODUCT, diffusion=w), **kw)


def source(f, w=1.0, **kw):
    return D.form(D.integrand(D.INT_PRODUCT, diffusion=w, weight=f), **kw)


# ------------------------------------------------------------------------------------------------------------------
# DoF mapping and sparsity patterns: bit-exact
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind,order", [(CG, 1), (CG, 2), (DG, 1), (DG, 2), (FV, 0)])
@pytest.mark.parametrize("n", [[7], [5, 4], [4, 3, 2]])
def test_dof_mapping_bit_exact(gdt, ctx, oracle, kind, order, n):
    gdesc = D.grid_desc(0.0, 1.0, n)
    space = make_space(gdt, ctx, gdesc, kind, order)
    assert space.mapper.size == oracle.space_size(gdesc, kind, order)
    for e in range(int(np.prod(n))):
        assert np.array_equal(space.mapper.global_indices(e), oracle.global_indices(gdesc, kind, order, e))


PATTERN_CASES = [
    (CG, 1, D.STENCIL_ELEMENT, [9], 0),
    (CG, 1, D.STENCIL_ELEMENT, [9, 2], 0),
    (CG, 1, D.STENCIL_ELEMENT, [5, 4, 3], 0),
    (CG, 1, D.STENCIL_ELEMENT, [1, 1, 1], 0),
    (CG, 2, D.STENCIL_ELEMENT, [9, 2], 0),
    (CG, 2, D.STENCIL_ELEMENT, [4, 3, 2], 0),
    (DG, 1, D.STENCIL_ELEMENT_AND_INTERSECTION, [6, 5], 0),
    (DG, 1, D.STENCIL_INTERSECTION, [6, 5], 0),
    (DG, 1, D.STENCIL_ELEMENT_AND_INTERSECTION, [4, 3, 3], 0),
    (DG, 2, D.STENCIL_ELEMENT_AND_INTERSECTION, [3, 4], 0),
    (DG, 1, D.STENCIL_ELEMENT_AND_INTERSECTION, [5, 4], 3),
    (FV, 0, D.STENCIL_ELEMENT_AND_INTERSECTION, [8, 8], 3),
    (FV, 0, D.STENCIL_ELEMENT_AND_INTERSECTION, [16], 1),
]


@pytest.mark.parametrize("kind,order,stencil,n,periodic", PATTERN_CASES)
def test_pattern_bit_exact_sort_unique(gdt, ctx, oracle, kind, order, stencil, n, periodic):
    gdesc = D.grid_desc(0.0, 1.0, n, periodic)
    space = make_space(gdt, ctx, gdesc, kind, order)
    pat = gdt.SparsityPattern(space, space, stencil, D.PATTERN_SORT_UNIQUE)
    rowptr, colidx = pat.download()
    rp, ci = oracle.pattern(gdesc, (kind, order), stencil=stencil)
    assert rowptr.dtype == np.int64 and colidx.dtype == np.int32
    assert np.array_equal(rowptr, rp)
    assert np.array_equal(colidx, ci)


@pytest.mark.parametrize("n", [[1], [9], [1, 1], [9, 2], [1, 1, 1], [5, 4, 3], [2, 1, 7]])
def test_pattern_bit_exact_structured_q1(gdt, ctx, oracle, n):
    gdesc = D.grid_desc(0.0, 1.0, n)
    space = make_space(gdt, ctx, gdesc, CG, 1)
    rowptr, colidx = gdt.SparsityPattern(space, space, D.STENCIL_ELEMENT, D.PATTERN_STRUCTURED).download()
    rp, ci = oracle.pattern(gdesc, (CG, 1))
    assert np.array_equal(rowptr, rp) and np.array_equal(colidx, ci)


@pytest.mark.parametrize("n", [[1, 1], [9, 2], [2, 5], [1, 1, 1], [4, 3, 2], [1, 3, 2], [3, 1, 1], [2, 2, 5]])
def test_pattern_bit_exact_structured_q2(gdt, ctx, oracle, n):
    """closed-form CG Q2 element stencil on the MCMG lattice numbering == the reference's insert + sort (oracle)"""
    gdesc = D.grid_desc(0.0, 1.0, n)
    space = make_space(gdt, ctx, gdesc, CG, 2)
    rowptr, colidx = gdt.SparsityPattern(space, space, D.STENCIL_ELEMENT, D.PATTERN_STRUCTURED).download()
    rp, ci = oracle.pattern(gdesc, (CG, 2))
    assert np.array_equal(rowptr, rp) and np.array_equal(colidx, ci)


@pytest.mark.parametrize("n,order", [([1], 1), ([7], 1), ([7], 2), ([1, 1], 1), ([6, 5], 1), ([1, 4], 1), ([3, 4], 2), ([1, 1, 1], 1),
                                     ([4, 3, 3], 1), ([2, 1, 3], 1), ([2, 2, 2], 2), ([5, 4], 0)])
def test_pattern_bit_exact_structured_dg(gdt, ctx, oracle, n, order):
    """closed-form DG element_and_intersection stencil (blocks by ascending neighbour index) == the oracle"""
    gdesc = D.grid_desc(0.0, 1.0, n)
    space = make_space(gdt, ctx, gdesc, DG, order)
    rowptr, colidx = gdt.SparsityPattern(space, space, D.STENCIL_ELEMENT_AND_INTERSECTION, D.PATTERN_STRUCTURED).download()
    rp, ci = oracle.pattern(gdesc, (DG, order), stencil=D.STENCIL_ELEMENT_AND_INTERSECTION)
    assert np.array_equal(rowptr, rp) and np.array_equal(colidx, ci)


def test_structured_pattern_limits(gdt, ctx):
    space = make_space(gdt, ctx, D.grid_desc(0.0, 1.0, [4, 2], 3), DG, 1)
    with pytest.raises(gdt.capi.NotImplementedGdt):  # a periodic direction with < 3 cells goes through sort-and-unique
        gdt.SparsityPattern(space, space, D.STENCIL_ELEMENT_AND_INTERSECTION, D.PATTERN_STRUCTURED)
    gdt.SparsityPattern(space, space, D.STENCIL_ELEMENT_AND_INTERSECTION, D.PATTERN_AUTO)
    # periodic directions with >= 3 cells have a closed form: same CSR as the sort-and-unique builder
    space = make_space(gdt, ctx, D.grid_desc(0.0, 1.0, [4, 5], 3), DG, 1)
    a = gdt.SparsityPattern(space, space, D.STENCIL_ELEMENT_AND_INTERSECTION, D.PATTERN_STRUCTURED)
    b = gdt.SparsityPattern(space, space, D.STENCIL_ELEMENT_AND_INTERSECTION, D.PATTERN_SORT_UNIQUE)
    (rp_a, ci_a), (rp_b, ci_b) = a.download(), b.download()
    assert np.array_equal(rp_a, rp_b) and np.array_equal(ci_a, ci_b)


def test_pattern_c1_size(gdt, ctx):
    # C1: 2D Q1 128^2: 16641 DoFs, 148225 nnz
    space = make_space(gdt, ctx, D.grid_desc(-1.0, 1.0, [128, 128]), CG, 1)
    for method in (D.PATTERN_STRUCTURED, D.PATTERN_SORT_UNIQUE):
        pat = gdt.SparsityPattern(space, space, D.STENCIL_ELEMENT, method)
        assert pat.rows == 16641 and pat.nnz == 148225


# ------------------------------------------------------------------------------------------------------------------
# CG element forms: matrix + RHS values
# ------------------------------------------------------------------------------------------------------------------
def rng_elem(n, lo=0.5, hi=2.0, seed=SEED):
    return np.random.default_rng(seed).uniform(lo, hi, int(np.prod(n)))


def element_cases():
    cases = []
    for n in ([6], [9, 2], [5, 4, 3]):
        d = len(n)
        kt = np.eye(d) + 0.25 * np.arange(d * d).reshape(d, d) / (d * d)
        cases += [
            ("q1-laplace-const", n, 1, [laplace(1.0)], "q1_gather"),
            ("q1-laplace-scaled", n, 1, [laplace(2.5, scaling=0.5)], "q1_gather"),
            ("q1-mass", n, 1, [mass(1.0)], "q1_gather"),
            ("q1-laplace+mass-sum", n, 1, [D.form([D.integrand(D.INT_LAPLACE, diffusion=1.5), D.integrand(D.INT_PRODUCT, diffusion=0.7)])], "q1_gather"),
            ("q1-two-forms", n, 1, [laplace(1.0), mass(3.0, over_integrate=1)], "q1_gather"),
            ("q1-laplace-elem", n, 1, [laplace(D.fn_elem(rng_elem(n)))], "q1_gather"),
            ("q1-laplace-elem+mass-elem", n, 1, [laplace(D.fn_elem(rng_elem(n))), mass(D.fn_elem(rng_elem(n, seed=7)))], "q1_gather"),
            ("q1-laplace-tensor", n, 1, [laplace(D.fn_const(kt))], "q1_gather"),
            ("q1-laplace-builtin", n, 1, [laplace(D.fn_builtin(D.BUILTIN_QUADRATIC, 2, 1.0, 0.5))], "q1_gather_qp"),
            ("q1-laplace-elemtensor", n, 1, [laplace(D.fn_elem(np.tile(kt, (int(np.prod(n)), 1, 1)) * rng_elem(n)[:, None, None]))], "q1_gather"),
            ("q2-laplace-const", n, 2, [laplace(1.0)], "q2_gather" if d > 1 else "generic_coloured"),
            ("q2-laplace-scaled-overint", n, 2, [laplace(2.5, scaling=0.5, over_integrate=1)], "q2_gather" if d > 1 else "generic_coloured"),
            ("q2-mass-elem", n, 2, [mass(D.fn_elem(rng_elem(n, seed=11)))], "q2_gather" if d > 1 else "generic_coloured"),
            ("q2-laplace-tensor", n, 2, [laplace(D.fn_const(kt))], "q2_gather_qp" if d > 1 else "generic_coloured"),
            ("q2-mass-builtin", n, 2, [mass(D.fn_builtin(D.BUILTIN_AFFINE, 1, 1.0, 0.3, 0.2, 0.1))], "q2_gather_qp" if d > 1 else "generic_coloured"),
            ("q2-laplace-elem+mass", n, 2, [D.form([D.integrand(D.INT_LAPLACE, diffusion=D.fn_elem(rng_elem(n))), D.integrand(D.INT_PRODUCT, diffusion=2.0)])], "q2_gather" if d > 1 else "generic_coloured"),
        ]
    cases += [
        ("q3-laplace-2d", [4, 3], 3, [laplace(1.0)], "generic_coloured"),
        ("q3-mass-1d", [5], 3, [mass(1.0)], "generic_coloured"),
        ("q1-underintegrated", [4, 3, 2], 1, [laplace(1.0, over_integrate=-2)], "q1_gather"),
        ("q1-single-element", [1, 1, 1], 1, [laplace(1.0)], "q1_gather"),
        ("q1-anisotropic-cells", [3, 5, 2], 1, [laplace(1.0)], "q1_gather"),
        ("q2-single-element", [1, 1, 1], 2, [laplace(1.0)], "q2_gather"),
        ("q2-thin-grid", [1, 7, 2], 2, [laplace(1.0), mass(0.5)], "q2_gather"),
        ("q2-2d-single-row", [13, 1], 2, [laplace(1.0)], "q2_gather"),
        ("q2-bigger-3d", [17, 9, 12], 2, [laplace(D.fn_elem(rng_elem([17, 9, 12])))], "q2_gather"),
        # lattice lines longer than a work item: the line-based row bookkeeping of k_q2_gather<..., LN = true>
        # (single group / several groups / element-wise coefficients; items that end exactly at a line end, start at one)
        ("q2-long-lines-3d-const", [87, 3, 4], 2, [laplace(1.0)], "q2_gather"),
        ("q2-long-lines-3d-two-forms", [101, 2, 3], 2, [laplace(0.75), mass(0.5)], "q2_gather"),
        ("q2-long-lines-3d-elem", [86, 3, 2], 2, [laplace(D.fn_elem(rng_elem([86, 3, 2])))], "q2_gather"),
        ("q2-long-lines-3d-170", [170, 2, 2], 2, [laplace(1.0)], "q2_gather"),
        ("q2-long-lines-2d-const", [255, 6], 2, [laplace(1.0)], "q2_gather"),
        ("q2-long-lines-2d-elem", [128, 5], 2, [mass(D.fn_elem(rng_elem([128, 5], seed=3)))], "q2_gather"),
    ]
    return cases


@pytest.mark.parametrize("name,n,order,forms,plan", element_cases(), ids=lambda v: v if isinstance(v, str) else None)
def test_cg_matrix_parity(gdt, ctx, oracle, name, n, order, forms, plan):
    lower, upper = ([0.0, -1.0, 0.5][: len(n)], [3.0, 1.0, 2.0][: len(n)])
    gdesc = D.grid_desc(lower, upper, n)
    rowptr, colidx, values, _, got_plan = gpu_assemble(gdt, ctx, gdesc, CG, order, D.STENCIL_ELEMENT, element=forms)
    assert got_plan == plan
    rp, ci = oracle.pattern(gdesc, (CG, order))
    assert np.array_equal(rowptr, rp) and np.array_equal(colidx, ci)
    ref, _ = oracle.assemble(gdesc, CG, order, rp, ci, forms)
    assert rel_err(values, ref) <= TOL


def rhs_cases():
    cos3 = D.fn_builtin(D.BUILTIN_COS_PRODUCT, 3, 0.75 * np.pi**2, 0.5 * np.pi)
    out = []
    for n in ([7], [9, 2], [5, 4, 3]):
        out += [
            ("const", n, 1, [source(D.fn_const(1.0))]),
            ("cos-product", n, 1, [source(cos3)]),
            ("cos-product-weighted", n, 1, [source(cos3, w=2.0, over_integrate=1)]),
            ("gaussian", n, 1, [source(D.fn_builtin(D.BUILTIN_GAUSSIAN, 3, 0.33, 0.4))]),
            ("elem", n, 1, [source(D.fn_elem(rng_elem(n, -1.0, 1.0)))]),
            ("elem+const+cos", n, 1, [source(D.fn_elem(rng_elem(n, -1.0, 1.0))), source(D.fn_const(0.5)), source(cos3)]),
            ("affine-generic", n, 1, [source(D.fn_builtin(D.BUILTIN_AFFINE, 1, 1.0, 0.3, 0.2, 0.1))]),
            ("q2-cos", n, 2, [source(cos3)]),
        ]
    return out


@pytest.mark.parametrize("name,n,order,forms", rhs_cases(), ids=lambda v: v if isinstance(v, str) else None)
def test_cg_rhs_parity(gdt, ctx, oracle, name, n, order, forms):
    gdesc = D.grid_desc(-1.0, 1.0, n)
    _, _, _, b, _ = gpu_assemble(gdt, ctx, gdesc, CG, order, D.STENCIL_ELEMENT, rhs=forms)
    rp, ci = oracle.pattern(gdesc, (CG, order))
    _, ref = oracle.assemble(gdesc, CG, order, rp, ci, rhs_forms=forms)
    assert rel_err(b, ref) <= TOL


def test_fused_matrix_and_rhs_one_walk(gdt, ctx, oracle):
    """examples/stationary-heat-equation.cc:94-106 (config C1, 2D Q1 128^2): one walk for matrix + RHS"""
    n = [128, 128]
    gdesc = D.grid_desc(-1.0, 1.0, n)
    src = D.fn_builtin(D.BUILTIN_COS_PRODUCT, 3, 0.5 * np.pi**2, 0.5 * np.pi)
    before = ctx.launch_count
    rowptr, colidx, values, b, plan = gpu_assemble(gdt, ctx, gdesc, CG, 1, D.STENCIL_ELEMENT, element=[laplace(1.0)],
                                                   rhs=[source(src)], pattern_method=D.PATTERN_STRUCTURED)
    assert plan == "q1_gather"
    # pattern + rhs tables + ONE fused gather kernel (+ the grid's geometry tables, built once per grid)
    assert ctx.launch_count - before <= 4
    rp, ci = oracle.pattern(gdesc, (CG, 1))
    ref_v, ref_b = oracle.assemble(gdesc, CG, 1, rp, ci, [laplace(1.0)], rhs_forms=[source(src)])
    assert np.array_equal(rowptr, rp) and np.array_equal(colidx, ci)
    assert rel_err(values, ref_v) <= TOL and rel_err(b, ref_b) <= TOL


def test_accumulate_mode_adds_to_existing_content(gdt, ctx, oracle):
    lib = gdt.capi.lib()
    gdesc = D.grid_desc(0.0, 1.0, [5, 4, 3])
    space = make_space(gdt, ctx, gdesc, CG, 1)
    pat = gdt.SparsityPattern(space, space, D.STENCIL_ELEMENT)
    rp, ci = oracle.pattern(gdesc, (CG, 1))
    ref_l, _ = oracle.assemble(gdesc, CG, 1, rp, ci, [laplace(1.0)])
    ref_m, _ = oracle.assemble(gdesc, CG, 1, rp, ci, [mass(2.0)])
    for second in (mass(2.0), mass(D.fn_builtin(D.BUILTIN_AFFINE, 0, 2.0, 0.0, 0.0, 0.0))):  # gather / generic path
        op = gdt.MatrixOperator(space, space, pat)
        f1 = laplace(1.0)
        gdt.capi.check(lib.gdtb_matop_append_element(op._h, C.byref(f1)))
        gdt.capi.check(lib.gdtb_assemble(op._h, None, D.ASSEMBLE_OVERWRITE))
        gdt.capi.check(lib.gdtb_matop_clear_forms(op._h))
        gdt.capi.check(lib.gdtb_matop_append_element(op._h, C.byref(second)))
        gdt.capi.check(lib.gdtb_assemble(op._h, None, D.ASSEMBLE_ACCUMULATE))
        assert rel_err(op.values(), ref_l + ref_m) <= TOL


def test_results_are_run_to_run_bit_identical(gdt, ctx):
    gdesc = D.grid_desc(-1.0, 1.0, [17, 9, 5])
    kappa = rng_elem([17, 9, 5])
    runs = []
    for _ in range(3):
        for order in (1, 2):
            _, _, v, b, _ = gpu_assemble(gdt, ctx, gdesc, CG, order, D.STENCIL_ELEMENT, element=[laplace(D.fn_elem(kappa))],
                                         rhs=[source(D.fn_builtin(D.BUILTIN_QUADRATIC, 2, 1.0, 1.0))])
            runs.append((order, v, b))
    for order, v, b in runs[2:]:
        ref = runs[order - 1]
        assert np.array_equal(v, ref[1]) and np.array_equal(b, ref[2])


def test_structural_properties_at_scale(gdt, ctx):
    """size-independent properties at a size the oracle would not finish quickly (3D Q1 96^3):
    constant vectors are in the kernel of the stiffness matrix, the mass matrix sums to the volume,
    and the matrix is symmetric."""
    import scipy.sparse as sp

    n = [96, 96, 96]
    gdesc = D.grid_desc(-1.0, 1.0, n)
    rowptr, colidx, values, b, plan = gpu_assemble(gdt, ctx, gdesc, CG, 1, D.STENCIL_ELEMENT, element=[laplace(1.0)],
                                                   rhs=[source(D.fn_const(1.0))])
    assert plan == "q1_gather"
    A = sp.csr_matrix((values, colidx, rowptr), shape=(97**3, 97**3))
    assert np.abs(A @ np.ones(97**3)).max() <= 1e-12 * np.abs(values).max()
    assert abs(A - A.T).max() <= 1e-15
    assert b.sum() == pytest.approx(8.0, rel=1e-13)
    _, _, mvals, _, _ = gpu_assemble(gdt, ctx, gdesc, CG, 1, D.STENCIL_ELEMENT, element=[mass(1.0)])
    assert mvals.sum() == pytest.approx(8.0, rel=1e-12)
    h = 2.0 / 96  # closed-form interior stencil
    row = 48 + 97 * (48 + 97 * 48)
    vals = values[rowptr[row] : rowptr[row + 1]]
    assert vals[13] == pytest.approx(8 * h / 3, rel=1e-13) and vals[0] == pytest.approx(-h / 12, rel=1e-13)


# ------------------------------------------------------------------------------------------------------------------
# DG / SWIPDG
# ------------------------------------------------------------------------------------------------------------------
def swipdg(kappa=1.0, omega=1.0, sigma_in=8.0, sigma_d=14.0, hI=D.HI_VOLUME, s=1.0):
    element = D.form(D.integrand(D.INT_LAPLACE, diffusion=kappa))
    coupling = D.form([
        D.integrand(D.INT_IPDG_INNER_COUPLING, diffusion=kappa, weight=omega, prefactor=s),
        D.integrand(D.INT_IPDG_INNER_PENALTY, weight=omega, prefactor=sigma_in, hI_kind=hI),
    ])
    boundary = D.form([
        D.integrand(D.INT_IPDG_BOUNDARY_PENALTY, weight=omega, prefactor=sigma_d, hI_kind=hI),
        D.integrand(D.INT_IPDG_DIRICHLET_COUPLING, diffusion=kappa, prefactor=s),
    ])
    return element, coupling, boundary


def swipdg_driver_order(kappa, omega, sigma_in=8.0, sigma_d=14.0, hI=D.HI_VOLUME, s=1.0, scaling=1.0):
    """the SWIPDG operator exactly as the reference's drivers append it (examples/adaptive_elliptic_swipdg.cc:230-251):
    {Laplace}, {inner coupling, inner penalty}, {Dirichlet coupling, boundary penalty} -- the structure the factorised DG
    kernel specialises on when the coefficients are element-wise (DgGatherParams::swip)"""
    element = D.form(D.integrand(D.INT_LAPLACE, diffusion=kappa), scaling=scaling)
    coupling = D.form([
        D.integrand(D.INT_IPDG_INNER_COUPLING, diffusion=kappa, weight=omega, prefactor=s),
        D.integrand(D.INT_IPDG_INNER_PENALTY, weight=omega, prefactor=sigma_in, hI_kind=hI),
    ], scaling=scaling)
    boundary = D.form([
        D.integrand(D.INT_IPDG_DIRICHLET_COUPLING, diffusion=kappa, prefactor=s),
        D.integrand(D.INT_IPDG_BOUNDARY_PENALTY, weight=omega, prefactor=sigma_d, hI_kind=hI),
    ], scaling=scaling)
    return element, coupling, boundary


def swipdg_cases():
    out = []
    for n in ([6], [8, 8], [5, 3], [4, 3, 2]):
        d = len(n)
        kt = np.diag(np.arange(1, d + 1, dtype=float)) + 0.1
        kap = rng_elem(n, seed=21)
        out += [
            # kappa = omega from ONE array (one load per cell), two arrays, a constant omega; both h_I conventions
            ("driver-order-elem-kappa-is-omega", n, 1, swipdg_driver_order(D.fn_elem(kap), D.fn_elem(kap))),
            ("driver-order-elem-kappa-omega", n, 1, swipdg_driver_order(D.fn_elem(kap), D.fn_elem(rng_elem(n, seed=22)), scaling=0.75)),
            ("driver-order-elem-kappa-const-omega-diameter", n, 1,
             swipdg_driver_order(D.fn_elem(kap), 2.0, sigma_in=16.0, sigma_d=16.0, hI=D.HI_DIAMETER, s=-1.0)),
            ("esv2007", n, 1, swipdg()),
            ("example-sigma16-diameter", n, 1, swipdg(sigma_in=16.0, sigma_d=16.0, hI=D.HI_DIAMETER)),
            ("nipdg", n, 1, swipdg(s=-1.0)),
            ("elem-kappa-swip", n, 1, swipdg(kappa=D.fn_elem(rng_elem(n)), omega=D.fn_elem(rng_elem(n)))),
            ("tensor-kappa", n, 1, swipdg(kappa=D.fn_const(kt), omega=D.fn_const(kt))),
            ("builtin-kappa", n, 1, swipdg(kappa=D.fn_builtin(D.BUILTIN_QUADRATIC, 2, 1.0, 0.5))),
            ("q2", n, 2, swipdg()),
            ("const-kappa3-omega2-scaled-mixed-hI", n, 1, _swipdg_mixed(3.0, 2.0)),
            ("elem-kappa-scaled-mixed-hI", n, 1, _swipdg_mixed(D.fn_elem(rng_elem(n)), D.fn_elem(rng_elem(n, seed=5)))),
        ]
    return out


def _swipdg_mixed(kappa, omega):
    """scaled forms, an extra mass term in the element form, both intersection-diameter conventions in one form and a
    sum of a symmetric and a non-symmetric coupling: exercises every accumulation of the factorised DG kernels"""
    element = D.form([D.integrand(D.INT_LAPLACE, diffusion=kappa), D.integrand(D.INT_PRODUCT, diffusion=0.5)], scaling=1.5)
    coupling = D.form([
        D.integrand(D.INT_IPDG_INNER_COUPLING, diffusion=kappa, weight=omega, prefactor=1.0),
        D.integrand(D.INT_IPDG_INNER_PENALTY, weight=omega, prefactor=8.0, hI_kind=D.HI_VOLUME),
        D.integrand(D.INT_IPDG_INNER_PENALTY, weight=omega, prefactor=3.0, hI_kind=D.HI_DIAMETER),
        D.integrand(D.INT_IPDG_INNER_COUPLING, diffusion=kappa, weight=omega, prefactor=-1.0),
    ], scaling=0.75)
    boundary = D.form([
        D.integrand(D.INT_IPDG_BOUNDARY_PENALTY, weight=omega, prefactor=14.0, hI_kind=D.HI_DIAMETER),
        D.integrand(D.INT_IPDG_DIRICHLET_COUPLING, diffusion=kappa, prefactor=1.0),
        D.integrand(D.INT_IPDG_BOUNDARY_PENALTY, weight=omega, prefactor=2.0, hI_kind=D.HI_VOLUME),
    ], scaling=2.0)
    return element, coupling, boundary


@pytest.mark.parametrize("name,n,order,forms", swipdg_cases(), ids=lambda v: v if isinstance(v, str) else None)
def test_swipdg_parity(gdt, ctx, oracle, name, n, order, forms):
    el, co, bo = forms
    gdesc = D.grid_desc(-1.0, 1.0, n)
    force = D.fn_builtin(D.BUILTIN_COS_PRODUCT, 2, 0.5 * np.pi**2, 0.5 * np.pi)
    rowptr, colidx, values, b, plan = gpu_assemble(
        gdt, ctx, gdesc, DG, order, D.STENCIL_ELEMENT_AND_INTERSECTION, element=[el], coupling=[co], boundary=[bo],
        rhs=[source(force)])
    rp, ci = oracle.pattern(gdesc, (DG, order), stencil=D.STENCIL_ELEMENT_AND_INTERSECTION)
    assert np.array_equal(rowptr, rp) and np.array_equal(colidx, ci)
    ref_v, ref_b = oracle.assemble(gdesc, DG, order, rp, ci, [el], [co], [bo], [source(force)])
    assert rel_err(values, ref_v) <= TOL
    assert rel_err(b, ref_b) <= TOL


def test_swipdg_wrong_stencil_is_an_error(gdt, ctx):
    """coupling entries outside an element-only pattern: the reference's add_to_entry throws; so do we"""
    el, co, bo = swipdg()
    with pytest.raises(gdt.capi.ShapesDoNotMatch):
        gpu_assemble(gdt, ctx, D.grid_desc(-1.0, 1.0, [4, 4]), DG, 1, D.STENCIL_ELEMENT, element=[el], coupling=[co])


# ------------------------------------------------------------------------------------------------------------------
# FV operator apply
# ------------------------------------------------------------------------------------------------------------------
def fv_cases():
    out = []
    for n, per in (([16], 1), ([64], 1), ([12, 10], 3), ([12, 10], 0), ([9, 8], 1), ([6, 5, 4], 7), ([6, 5, 4], 2)):
        d = len(n)
        a = [1.0, 0.5, -0.75][:d]
        out += [
            ("linear-upwind", n, per, D.flux(D.FLUX_LINEAR, D.NUMFLUX_UPWIND, a)),
            ("linear-x-upwind", n, per, D.flux(D.FLUX_LINEAR, D.NUMFLUX_UPWIND, [1.0, 0.0, 0.0][:d])),
            ("burgers-upwind", n, per, D.flux(D.FLUX_BURGERS, D.NUMFLUX_UPWIND, [])),
            ("linear-lf", n, per, D.flux(D.FLUX_LINEAR, D.NUMFLUX_LAX_FRIEDRICHS, a)),
            ("burgers-lf", n, per, D.flux(D.FLUX_BURGERS, D.NUMFLUX_LAX_FRIEDRICHS, [])),
        ]
    return out


@pytest.mark.parametrize("name,n,periodic,fl", fv_cases(), ids=lambda v: v if isinstance(v, str) else None)
def test_fv_apply_parity(gdt, ctx, oracle, name, n, periodic, fl):
    gdesc = D.grid_desc(0.0, 1.0, n, periodic)
    space = make_space(gdt, ctx, gdesc, FV, 0)
    num = gdt.NumericalUpwindFlux(fl.kind, list(fl.p))
    num.desc.numflux = fl.numflux
    L = gdt.make_advection_fv_operator(num, space)
    u = np.random.default_rng(SEED).uniform(-1.0, 1.0, int(np.prod(n)))
    out = L.apply(u)
    ref = oracle.fv_apply(gdesc, fl, u)
    assert rel_err(out, ref) <= TOL


@pytest.mark.parametrize("bad", [np.nan, np.inf, -np.inf])
def test_fv_apply_refuses_non_finite_source(gdt, ctx, bad):
    """apply(VectorType source, ...) throws operator_error for a source with inf / nan
    (operators/localizable-operator.hh:383-385); afterwards the operator is still usable"""
    gdesc = D.grid_desc(0.0, 1.0, [33, 17], 3)
    space = make_space(gdt, ctx, gdesc, FV, 0)
    L = gdt.make_advection_fv_operator(gdt.NumericalUpwindFlux(D.FLUX_LINEAR, [1.0, 0.5]), space)
    u = np.random.default_rng(SEED).uniform(-1.0, 1.0, 33 * 17)
    good = L.apply(u)
    v = u.copy()
    v[-1] = bad
    with pytest.raises(gdt.capi.OperatorError):
        L.apply(v)
    assert np.array_equal(L.apply(u), good)


@pytest.mark.parametrize("n,periodic", [([512, 2], 3), ([512, 7], 0), ([512, 23], 1), ([1024, 61], 3), ([1536, 9], 2),
                                        ([1024, 130], 3), ([512, 200], 0)])
@pytest.mark.parametrize("kind,numflux", [(D.FLUX_LINEAR, D.NUMFLUX_UPWIND), (D.FLUX_BURGERS, D.NUMFLUX_UPWIND),
                                          (D.FLUX_LINEAR, D.NUMFLUX_LAX_FRIEDRICHS), (D.FLUX_BURGERS, D.NUMFLUX_LAX_FRIEDRICHS)])
@pytest.mark.parametrize("tma", ["0", "1"])
def test_fv_apply_parity_tma_staged(gdt, ctx, oracle, monkeypatch, n, periodic, kind, numflux, tma):
    """rows of a multiple of 512 cells can take the TMA-staged kernel (fv_tma.cu, GDTB_FV_TMA=1): one / several strips, runs
    that end inside a row group, every periodicity, apply and the fused Euler step; anisotropic cells; the register-marching
    kernel on the same inputs"""
    monkeypatch.setenv("GDTB_FV_TMA", tma)
    gdesc = D.grid_desc([0.0, -1.0], [3.0, 1.0], n, periodic)
    space = make_space(gdt, ctx, gdesc, FV, 0)
    fl = D.flux(kind, numflux, [1.0, -0.5] if kind == D.FLUX_LINEAR else [])
    num = gdt.NumericalUpwindFlux(kind, list(fl.p))
    num.desc.numflux = numflux
    L = gdt.make_advection_fv_operator(num, space)
    u = np.random.default_rng(SEED).uniform(-1.0, 1.0, int(np.prod(n)))
    ref = oracle.fv_apply(gdesc, fl, u)
    assert rel_err(L.apply(u), ref) <= TOL
    dt = 1e-3
    assert rel_err(L.explicit_euler(u, dt, 3), oracle.fv_euler(gdesc, fl, u, dt, 3)) <= TOL


def test_fv_explicit_euler_reference_tables(gdt, ctx, oracle):
    """linear transport with dt = h is an exact shift and conserves mass (linear_transport__1d__explicit__fv.mini)"""
    for N in (16, 32, 64):
        gdesc = D.grid_desc([0.0], [1.0], [N], periodic=1)
        space = make_space(gdt, ctx, gdesc, FV, 0)
        L = gdt.make_advection_fv_operator(gdt.NumericalUpwindFlux(D.FLUX_LINEAR, [1.0]), space)
        u0 = oracle.fv_interpolate(gdesc, D.fn_builtin(D.BUILTIN_INDICATOR, 0, 0.25, 0.5))
        u = L.explicit_euler(u0, 1.0 / N, N + 1)
        np.testing.assert_allclose(u, np.roll(u0, N + 1), atol=1e-13)
        ref = oracle.fv_euler(gdesc, D.flux(D.FLUX_LINEAR, D.NUMFLUX_UPWIND, [1.0]), u0, 1.0 / N, N + 1)
        assert rel_err(u, ref) <= TOL


def test_fv_burgers_euler_parity_and_mass(gdt, ctx, oracle):
    N, dt = 32, 0.0096815612792968738 / 2
    gdesc = D.grid_desc([0.0], [1.0], [N], periodic=1)
    space = make_space(gdt, ctx, gdesc, FV, 0)
    L = gdt.make_advection_fv_operator(gdt.NumericalUpwindFlux(D.FLUX_BURGERS), space)
    u0 = oracle.fv_interpolate(gdesc, D.fn_builtin(D.BUILTIN_GAUSSIAN, 3, 0.33, 0.075))
    u = L.explicit_euler(u0, dt, 107)
    ref = oracle.fv_euler(gdesc, D.flux(D.FLUX_BURGERS, D.NUMFLUX_UPWIND, []), u0, dt, 107)
    assert rel_err(u, ref) <= TOL
    assert abs(u.sum() - u0.sum()) / u0.sum() < 1e-14


def test_fv_interpolate_parity(gdt, ctx, oracle):
    import torch

    lib = gdt.capi.lib()
    for n, f in (([64], D.fn_builtin(D.BUILTIN_GAUSSIAN, 3, 0.33, 0.075)), ([20, 10], D.fn_builtin(D.BUILTIN_INDICATOR, 0, 0.25, 0.5)),
                 ([6, 5, 4], D.fn_builtin(D.BUILTIN_COS_PRODUCT, 3, 1.0, 2.0))):
        gdesc = D.grid_desc(0.0, 1.0, n, (1 << len(n)) - 1)
        space = make_space(gdt, ctx, gdesc, FV, 0)
        u = torch.empty(int(np.prod(n)), dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()
        gdt.capi.check(lib.gdtb_fv_interpolate(ctx._h, space._h, C.byref(f), C.c_void_p(u.data_ptr())))
        ref = oracle.fv_interpolate(gdesc, f)
        assert rel_err(u.cpu().numpy(), ref) <= TOL


def test_fv_large_grid_properties(gdt, ctx):
    """C4-sized properties (4096^2 periodic): conservation (sum of L(u) = 0) and linearity of the linear operator"""
    import torch

    n = [4096, 4096]
    gdesc = D.grid_desc(0.0, 1.0, n, 3)
    space = make_space(gdt, ctx, gdesc, FV, 0)
    L = gdt.make_advection_fv_operator(gdt.NumericalUpwindFlux(D.FLUX_LINEAR, [1.0, 0.5]), space)
    g = torch.Generator(device="cuda").manual_seed(SEED)
    u = torch.rand(4096 * 4096, dtype=torch.float64, device="cuda", generator=g)
    v = torch.rand(4096 * 4096, dtype=torch.float64, device="cuda", generator=g)
    Lu, Lv, Luv = torch.empty_like(u), torch.empty_like(u), torch.empty_like(u)
    torch.cuda.synchronize()
    L.apply_device(u.data_ptr(), Lu.data_ptr())
    L.apply_device(v.data_ptr(), Lv.data_ptr())
    w = 2.0 * u + v
    torch.cuda.synchronize()  # the library works on its own stream
    L.apply_device(w.data_ptr(), Luv.data_ptr())
    torch.cuda.synchronize()
    scale = Lu.abs().max().item()
    assert abs(Lu.sum().item()) <= 1e-9 * scale * 4096
    assert (Luv - (2.0 * Lu + Lv)).abs().max().item() <= 1e-12 * 3 * scale
    # x-upwind with a = (1, 0.5): L(u)_ij = ((u_ij - u_{i-1,j}) * 1 + (u_ij - u_{i,j-1}) * 0.5) / h
    U = u.view(4096, 4096)
    expect = ((U - torch.roll(U, 1, dims=1)) + 0.5 * (U - torch.roll(U, 1, dims=0))) * 4096.0
    assert (Lu.view(4096, 4096) - expect).abs().max().item() <= 1e-12 * scale


# ------------------------------------------------------------------------------------------------------------------
# error behaviour of the boundary (mirrors the reference's exceptions)
# ------------------------------------------------------------------------------------------------------------------
def test_error_conventions(gdt, ctx):
    lib = gdt.capi.lib()
    gdesc = D.grid_desc(0.0, 1.0, [4, 4])
    cg1, cg2 = make_space(gdt, ctx, gdesc, CG, 1), make_space(gdt, ctx, gdesc, CG, 2)
    pat1 = gdt.SparsityPattern(cg1, cg1, D.STENCIL_ELEMENT)
    with pytest.raises(gdt.capi.ShapesDoNotMatch):  # matrix-based.hh:73-80
        gdt.MatrixOperator(cg2, cg2, pat1)
    with pytest.raises(gdt.capi.WrongInputGiven):  # sparsity-pattern.hh:174-177
        gdt.SparsityPattern(cg1, cg1, 7)
    with pytest.raises(gdt.capi.FiniteElementError):  # lagrange.hh:107-136 (our limits are tighter)
        make_space(gdt, ctx, D.grid_desc(0.0, 1.0, [2, 2, 2]), CG, 4)
    op = gdt.MatrixOperator(cg1, cg1, pat1)
    bad = D.form(D.integrand(D.INT_IPDG_INNER_PENALTY, prefactor=1.0))
    assert lib.gdtb_matop_append_element(op._h, C.byref(bad)) == 3  # integrand_error
    with pytest.raises(gdt.capi.OperatorError):  # local/operators/advection-fv.hh:131-134
        gdt.make_advection_fv_operator(gdt.NumericalUpwindFlux(D.FLUX_LINEAR, [1.0, 0.0]), cg1)
    with pytest.raises(gdt.capi.SpaceError):
        cg1.mapper.global_indices(99)


# ------------------------------------------------------------------------------------------------------------------
# element-block partition (multi-GPU layout), exercised on one device
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("constant_kappa", [False, True], ids=["elem-kappa+mass", "const-kappa"])
@pytest.mark.parametrize("n,cuts", [([6, 5, 8], [0, 3, 8]), ([6, 5, 9], [0, 1, 4, 9]), ([7, 6], [0, 2, 6]), ([12], [0, 5, 12])])
def test_slab_owner_computes_rows(gdt, ctx, oracle, n, cuts, constant_kappa):
    """every slab produces its owned rows completely and without communication; the concatenation of all slabs is
    the global matrix / vector (constant kappa: the sum-factorised row kernel of the headline configuration, which is
    what `bench.py --gpus N` runs on every rank)"""
    kappa = rng_elem(n)
    forms = [laplace(0.75)] if constant_kappa else [laplace(D.fn_elem(kappa)), mass(0.5)]
    _slab_owner_computes_rows(gdt, ctx, oracle, n, cuts, forms)


@pytest.mark.parametrize("n,cuts", [([223, 2, 3], [0, 3]), ([223, 2, 3], [0, 1, 3]), ([300, 3, 4], [0, 2, 3, 4]), ([257, 1, 2], [0, 2])])
def test_q1_one_kappa_per_element_long_lines(gdt, ctx, oracle, n, cuts):
    """ONE Laplace integrand with one kappa per element on lattice lines of at least 224 vertices: the kernel variant
    with prefetched coefficients and work-item records (k_q1_items; items that end exactly at a line end, straddle a
    line break or a layer break), whole grid and slabs, fused right-hand side"""
    _slab_owner_computes_rows(gdt, ctx, oracle, n, cuts, [laplace(D.fn_elem(rng_elem(n, seed=3)))])


def _slab_owner_computes_rows(gdt, ctx, oracle, n, cuts, forms):
    lib = gdt.capi.lib()
    check = gdt.capi.check
    gdesc = D.grid_desc(-1.0, 1.0, n)
    src = D.fn_builtin(D.BUILTIN_COS_PRODUCT, 3, 2.0, 1.3)
    rp, ci = oracle.pattern(gdesc, (CG, 1))
    ref_v, ref_b = oracle.assemble(gdesc, CG, 1, rp, ci, forms, rhs_forms=[source(src), source(D.fn_const(0.25))])
    space = make_space(gdt, ctx, gdesc, CG, 1)
    got_v, got_b = np.full_like(ref_v, np.nan), np.full_like(ref_b, np.nan)
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        op_h, fun_h = C.c_void_p(), C.c_void_p()
        check(lib.gdtb_matop_create(ctx._h, space._h, space._h, None, C.byref(op_h)))
        check(lib.gdtb_vecfun_create(ctx._h, space._h, C.byref(fun_h)))
        check(lib.gdtb_matop_set_slab(op_h, lo, hi))
        check(lib.gdtb_vecfun_set_slab(fun_h, lo, hi))
        for f in forms:
            check(lib.gdtb_matop_append_element(op_h, C.byref(f)))
        for f in (source(src), source(D.fn_const(0.25))):
            check(lib.gdtb_vecfun_append_element(fun_h, C.byref(f)))
        check(lib.gdtb_assemble(op_h, fun_h, D.ASSEMBLE_OVERWRITE))
        rb, re_, vo = C.c_int64(), C.c_int64(), C.c_int64()
        check(lib.gdtb_matop_local_rows(op_h, C.byref(rb), C.byref(re_), C.byref(vo)))
        nnz = lib.gdtb_matop_local_nnz(op_h)
        assert vo.value == rp[rb.value] and nnz == rp[re_.value] - rp[rb.value]
        v, b = np.empty(nnz), np.empty(re_.value - rb.value)
        check(lib.gdtb_matop_values_download(op_h, gdt.capi.dptr(v)))
        check(lib.gdtb_vecfun_download(fun_h, gdt.capi.dptr(b)))
        got_v[vo.value : vo.value + nnz] = v
        got_b[rb.value : re_.value] = b
        lib.gdtb_matop_destroy(op_h)
        lib.gdtb_vecfun_destroy(fun_h)
    assert not np.isnan(got_v).any() and not np.isnan(got_b).any()
    assert rel_err(got_v, ref_v) <= TOL and rel_err(got_b, ref_b) <= TOL


@pytest.mark.parametrize("n,cuts", [([4, 3, 6], [0, 2, 6]), ([3, 4, 7], [0, 1, 3, 7]), ([5, 6], [0, 2, 6]), ([3, 2, 4], [0, 1, 2, 3, 4]),
                                    ([90, 2, 5], [0, 2, 3, 5]), ([131, 7], [0, 3, 7])])
def test_slab_owner_computes_rows_q2(gdt, ctx, oracle, n, cuts):
    """CG Q2 (BASELINE config 5 is sharded across the GPUs): the MCMG numbering groups DoFs by sub-entity kind, so a
    slab owns one contiguous row range per group; every owned row is complete without communication, the ranges of
    all slabs tile the global rows / CSR values exactly once and reproduce the global matrix (pattern-free operator:
    every CSR position is a closed form)"""
    from dune_gdt_b200 import parallel

    gdesc = D.grid_desc(-1.0, 1.0, n)
    forms = [laplace(0.75), mass(0.5)]
    if len(n) == 3 and n[0] in (3, 90):  # one integrand with one coefficient per element: the per-element factor tables
        forms = [laplace(D.fn_elem(rng_elem(n, seed=5)))]
    rp, ci = oracle.pattern(gdesc, (CG, 2))
    ref_v, _ = oracle.assemble(gdesc, CG, 2, rp, ci, forms)
    space = make_space(gdt, ctx, gdesc, CG, 2)
    got_v = np.full_like(ref_v, np.nan)
    rows_seen = np.zeros(rp.size - 1, dtype=int)
    world = len(cuts) - 1
    for rank in range(world):
        sa = parallel.SlabAssembly(space, rank, world, with_functional=False)
        # SlabAssembly cuts evenly; re-cut to the requested layers
        gdt.capi.check(gdt.capi.lib().gdtb_matop_set_slab(sa.op_h, cuts[rank], cuts[rank + 1]))
        sa = _refresh_ranges(gdt, sa)
        for f in forms:
            sa.append(f)
        values, _ = sa.assemble()
        assert len(sa.row_ranges) == 2 ** len(n)
        at = 0
        for rb, re_, vo, cnt in sa.row_ranges:
            assert vo == rp[rb] and cnt == rp[re_] - rp[rb]  # closed-form CSR offsets == the pattern builder's
            rows_seen[rb:re_] += 1
            assert np.isnan(got_v[vo:vo + cnt]).all()
            got_v[vo:vo + cnt] = values[at:at + cnt]
            at += cnt
        assert at == values.size
    assert (rows_seen == 1).all() and not np.isnan(got_v).any()
    assert rel_err(got_v, ref_v) <= TOL


def _refresh_ranges(gdt, sa):
    lib = gdt.capi.lib()
    n = C.c_int32()
    gdt.capi.check(lib.gdtb_matop_local_row_ranges(sa.op_h, 0, None, None, None, None, C.byref(n)))
    arrs = [np.zeros(n.value, dtype=np.int64) for _ in range(4)]
    ptr = [a.ctypes.data_as(C.POINTER(C.c_int64)) for a in arrs]
    gdt.capi.check(lib.gdtb_matop_local_row_ranges(sa.op_h, n.value, *ptr, C.byref(n)))
    sa.row_ranges = [tuple(int(x) for x in t) for t in zip(*arrs)]
    sa.nnz_local = int(lib.gdtb_matop_local_nnz(sa.op_h))
    return sa


def test_q2_pattern_free_operator_matches_pattern_based(gdt, ctx, oracle):
    gdesc = D.grid_desc(0.0, 1.0, [3, 4, 2])
    space = make_space(gdt, ctx, gdesc, CG, 2)
    lib, check = gdt.capi.lib(), gdt.capi.check
    op_h = C.c_void_p()
    check(lib.gdtb_matop_create(ctx._h, space._h, space._h, None, C.byref(op_h)))
    f = laplace(1.0)
    check(lib.gdtb_matop_append_element(op_h, C.byref(f)))
    check(lib.gdtb_assemble(op_h, None, D.ASSEMBLE_OVERWRITE))
    rp, ci = oracle.pattern(gdesc, (CG, 2))
    assert lib.gdtb_matop_local_nnz(op_h) == ci.size
    v = np.empty(ci.size)
    check(lib.gdtb_matop_values_download(op_h, gdt.capi.dptr(v)))
    ref_v, _ = oracle.assemble(gdesc, CG, 2, rp, ci, [f])
    assert rel_err(v, ref_v) <= TOL
    # the closed-form pattern is materialised on demand (mat-vec, constraints, solvers): same CSR as the oracle's
    import scipy.sparse as sp

    d_rp, d_ci = C.c_void_p(), C.c_void_p()
    check(lib.gdtb_matop_pattern_device(op_h, C.byref(d_rp), C.byref(d_ci)))
    x = np.random.default_rng(SEED).uniform(-1.0, 1.0, rp.size - 1)
    y = np.empty_like(x)
    check(lib.gdtb_matop_apply_host(op_h, gdt.capi.dptr(x), gdt.capi.dptr(y)))
    A = sp.csr_matrix((ref_v, ci, rp), shape=(rp.size - 1, rp.size - 1))
    assert rel_err(y, A @ x) <= TOL
    lib.gdtb_matop_destroy(op_h)
