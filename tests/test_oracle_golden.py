"""Pins the CPU restatement oracle against the reference's own known-answer values.

CPU only.  If these fail the oracle cannot be trusted and every parity claim is void.
"""
import numpy as np
import pytest

from dune_gdt_b200 import descriptors as D

CG, DG, FV = D.SPACE_CG, D.SPACE_DG, D.SPACE_FV


def laplace(kappa=1.0, **kw):
    return D.form(D.integrand(D.INT_LAPLACE, diffusion=kappa), **kw)


def mass(w=1.0, **kw):
    return D.form(D.integrand(D.INT_PRODUCT, diffusion=w), **kw)


def test_q2_stiffness_reference_golden(oracle):
    # dune/gdt/test/integrands/integrands_laplace.cc:115-135 (grid: integrands.hh:64-70)
    g = D.grid_desc([0, 0], [3, 1], [9, 2])
    rp, ci = oracle.pattern(g, (CG, 2))
    assert len(rp) - 1 == 95
    v, _ = oracle.assemble(g, CG, 2, rp, ci, [laplace()])
    assert v.min() == pytest.approx(-1.896296296296300, abs=1e-13)
    assert v.max() == pytest.approx(6.162962962962970, abs=1e-13)
    assert (v * v).sum() == pytest.approx(1704.099039780521, abs=5e-12)


def test_q2_mass_reference_golden(oracle):
    # dune/gdt/test/integrands/integrands_product.cc:107-126
    g = D.grid_desc([0, 0], [3, 1], [9, 2])
    rp, ci = oracle.pattern(g, (CG, 2))
    v, _ = oracle.assemble(g, CG, 2, rp, ci, [mass()])
    assert v.min() == pytest.approx(-0.002962962962963, abs=1e-13)
    assert v.max() == pytest.approx(0.047407407407407, abs=1e-13)
    assert (v * v).sum() == pytest.approx(0.066475994513031, abs=1e-13)


# SURVEY.md Appendix C addendum (derived 3D values, same construction as the reference's tests)
ADDENDUM = [
    # order, form, ndof, nnz, min, max, sumsq
    (1, "laplace", 150, 2548, -1.851851851851852e-01, 1.259259259259259e00, 8.193141289437587e01),
    (1, "mass", 150, 2548, 0.0, 2.469135802469136e-02, 4.107605547934767e-02),
    (2, "laplace", 855, 40953, -4.424691358024692e-01, 2.149135802469137e00, 8.992255959762240e02),
    (2, "mass", 855, 40953, -7.901234567901237e-04, 1.264197530864197e-02, 2.437453132144491e-02),
]


@pytest.mark.parametrize("order,kind,ndof,nnz,vmin,vmax,ssq", ADDENDUM)
def test_3d_known_answers(oracle, order, kind, ndof, nnz, vmin, vmax, ssq):
    g = D.grid_desc([0, 0, 0], [3, 1, 2], [9, 2, 4])
    rp, ci = oracle.pattern(g, (CG, order))
    assert len(rp) - 1 == ndof and len(ci) == nnz
    v, _ = oracle.assemble(g, CG, order, rp, ci, [laplace() if kind == "laplace" else mass()])
    scale = max(abs(vmin), abs(vmax))
    # the reference takes min/max over the dense n x n serialisation, i.e. including the structural zeros
    assert abs(min(v.min(), 0.0) - vmin) <= 1e-12 * scale
    assert abs(v.max() - vmax) <= 1e-12 * scale
    assert abs((v * v).sum() - ssq) <= 1e-11 * ssq


def test_2d_q1_laplace_known_answer(oracle):
    g = D.grid_desc([0, 0], [3, 1], [9, 2])
    rp, ci = oracle.pattern(g, (CG, 1))
    assert len(rp) - 1 == 30 and len(ci) == 196
    v, _ = oracle.assemble(g, CG, 1, rp, ci, [laplace()])
    assert v.min() == pytest.approx(-7.777777777777778e-01, abs=1e-13)
    assert v.max() == pytest.approx(2.888888888888889e00, abs=1e-13)
    assert (v * v).sum() == pytest.approx(1.322345679012346e02, abs=1e-11)


def test_closed_form_q1_stencils(oracle):
    # 3D Q1 Laplace interior row: centre 8h/3, faces 0, edges -h/6, corners -h/12
    n, h = 4, 0.25
    g = D.grid_desc(0.0, 1.0, [n, n, n])
    rp, ci = oracle.pattern(g, (CG, 1))
    v, _ = oracle.assemble(g, CG, 1, rp, ci, [laplace()])
    V = n + 1
    row = 2 + V * (2 + V * 2)
    cols, vals = ci[rp[row] : rp[row + 1]], v[rp[row] : rp[row + 1]]
    assert len(cols) == 27
    for c, a in zip(cols, vals):
        dx, dy, dz = c % V - 2, (c // V) % V - 2, c // (V * V) - 2
        nz = abs(dx) + abs(dy) + abs(dz)
        expect = {0: 8 * h / 3, 1: 0.0, 2: -h / 6, 3: -h / 12}[nz]
        assert a == pytest.approx(expect, abs=1e-15)
    assert abs(vals.sum()) < 1e-15
    # mass: 8h^3/27, 2h^3/27, h^3/54, h^3/216
    v, _ = oracle.assemble(g, CG, 1, rp, ci, [mass()])
    vals = v[rp[row] : rp[row + 1]]
    for c, a in zip(cols, vals):
        dx, dy, dz = c % V - 2, (c // V) % V - 2, c // (V * V) - 2
        nz = abs(dx) + abs(dy) + abs(dz)
        expect = {0: 8 * h**3 / 27, 1: 2 * h**3 / 27, 2: h**3 / 54, 3: h**3 / 216}[nz]
        assert a == pytest.approx(expect, rel=1e-13)


@pytest.mark.parametrize("dim,n", [(1, [7]), (2, [5, 3]), (3, [4, 3, 2])])
@pytest.mark.parametrize("order", [1, 2])
def test_cg_mapper_is_a_bijection_onto_lagrange_points(oracle, dim, n, order):
    # dune/gdt/test/spaces/h1_continuous_lagrange.hh:51-86: every global Lagrange point has exactly one global
    # index and the indices are consecutive from 0
    g = D.grid_desc(0.0, 1.0, n)
    size = oracle.space_size(g, CG, order)
    assert size == int(np.prod([order * k + 1 for k in n]))
    point_of = {}
    ne = int(np.prod(n))
    for e in range(ne):
        idx = [e % n[0], (e // n[0]) % (n[1] if dim > 1 else 1), e // (n[0] * (n[1] if dim > 1 else 1))]
        gi = oracle.global_indices(g, CG, order, e)
        for i, gidx in enumerate(gi):
            a = [i % (order + 1), (i // (order + 1)) % (order + 1), i // (order + 1) ** 2]
            point = tuple(order * idx[k] + a[k] for k in range(dim))
            assert point_of.setdefault(int(gidx), point) == point
    assert sorted(point_of) == list(range(size))
    assert len(set(point_of.values())) == size


def test_q1_global_index_is_vertex_index(oracle):
    g = D.grid_desc(0.0, 1.0, [3, 2, 2])
    for e in range(12):
        ex, ey, ez = e % 3, (e // 3) % 2, e // 6
        gi = oracle.global_indices(g, CG, 1, e)
        expect = [(ex + a) + 4 * ((ey + b) + 3 * (ez + c)) for c in (0, 1) for b in (0, 1) for a in (0, 1)]
        assert list(gi) == expect


@pytest.mark.parametrize("order,per_dir", [(1, lambda N: 3 * N + 1), (2, lambda N: 8 * N + 1)])
def test_pattern_size_formula(oracle, order, per_dir):
    # SURVEY Appendix C addendum: nnz = prod_i (N_i (k+1)^2 - (N_i - 1))
    for n in ([5], [4, 3], [3, 2, 2]):
        g = D.grid_desc(0.0, 1.0, n)
        rp, ci = oracle.pattern(g, (CG, order))
        assert len(ci) == int(np.prod([per_dir(k) for k in n]))
        for r in range(len(rp) - 1):
            row = ci[rp[r] : rp[r + 1]]
            assert np.all(np.diff(row) > 0)


def test_dg_patterns(oracle):
    # DG-Q1 on Yasp N^2 with element_and_intersection stencil: nnz = 16 (N^2 + 4N(N-1))
    N = 5
    g = D.grid_desc(-1.0, 1.0, [N, N])
    rp, ci = oracle.pattern(g, (DG, 1), stencil=D.STENCIL_ELEMENT_AND_INTERSECTION)
    assert len(rp) - 1 == 4 * N * N
    assert len(ci) == 16 * (N * N + 4 * N * (N - 1))
    rp, ci = oracle.pattern(g, (DG, 1), stencil=D.STENCIL_ELEMENT)
    assert len(ci) == 16 * N * N
    rp, ci = oracle.pattern(g, (DG, 1), stencil=D.STENCIL_INTERSECTION)
    assert len(ci) == 16 * 4 * N * (N - 1)


def test_gauss_rules_are_exact(oracle):
    for order in range(0, 12):
        x, w = oracle.gauss_rule(order)
        assert len(x) == order // 2 + 1
        for p in range(order + 1):
            assert np.dot(w, x**p) == pytest.approx(1.0 / (p + 1), rel=1e-14)


def test_rhs_constant_source(oracle):
    # sum of the load vector of f = 1 is the domain volume; every Q1 interior entry is h^d
    g = D.grid_desc([0, 0, 0], [1, 2, 3], [4, 4, 4])
    rp, ci = oracle.pattern(g, (CG, 1))
    f = D.form(D.integrand(D.INT_PRODUCT, diffusion=1.0, weight=D.fn_const(1.0)))
    _, b = oracle.assemble(g, CG, 1, rp, ci, rhs_forms=[f])
    assert b.sum() == pytest.approx(6.0, rel=1e-14)
    assert b[1 + 5 * (1 + 5 * 1)] == pytest.approx(0.25 * 0.5 * 0.75, rel=1e-14)


def test_heat_equation_example_converges_second_order(oracle):
    """examples/stationary-heat-equation.cc:62-127 with the oracle's assembly + a scipy solve: the L2 error against
    cos(pi x/2) cos(pi y/2) must fall by ~4 per refinement."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla

    errs = []
    for N in (8, 16, 32):
        g = D.grid_desc(-1.0, 1.0, [N, N])
        rp, ci = oracle.pattern(g, (CG, 1))
        src = D.fn_builtin(D.BUILTIN_COS_PRODUCT, 3, np.pi**2 / 2, np.pi / 2)
        rhs = D.form(D.integrand(D.INT_PRODUCT, diffusion=1.0, weight=src))
        v, b = oracle.assemble(g, CG, 1, rp, ci, [laplace()], rhs_forms=[rhs])
        A = sp.csr_matrix((v, ci, rp), shape=(len(b), len(b))).tolil()
        V = N + 1
        xs = np.linspace(-1, 1, V)
        X, Y = np.meshgrid(xs, xs, indexing="xy")
        boundary = ((np.abs(X) == 1) | (np.abs(Y) == 1)).reshape(-1)
        interior = np.where(~boundary)[0]
        A = A.tocsr()[interior][:, interior]
        u = np.zeros(len(b))
        u[interior] = spla.spsolve(A.tocsc(), b[interior])
        exact = (np.cos(np.pi / 2 * X) * np.cos(np.pi / 2 * Y)).reshape(-1)
        errs.append(np.sqrt(np.mean((u - exact) ** 2)))
    assert errs[0] / errs[1] == pytest.approx(4.0, rel=0.1)
    assert errs[1] / errs[2] == pytest.approx(4.0, rel=0.1)


# ---- SWIPDG: dune/gdt/test/stationary-heat-equation/stationary_heat_equation__ESV2007__table_1.mini:23-36 -------
def _q1_dg_h1_semi_error(N, u):
    """broken H1 semi-norm of u_h - cos(pi x/2)cos(pi y/2) on [-1,1]^2, 4x4 Gauss points per cell"""
    h = 2.0 / N
    gx, gw = np.polynomial.legendre.leggauss(4)
    gx, gw = (gx + 1) / 2, gw / 2
    err2 = 0.0
    for e in range(N * N):
        ex, ey = e % N, e // N
        c = u[4 * e : 4 * e + 4]
        for qx, wx in zip(gx, gw):
            for qy, wy in zip(gx, gw):
                dphix = np.array([-(1 - qy), (1 - qy), -qy, qy]) / h
                dphiy = np.array([-(1 - qx), -qx, (1 - qx), qx]) / h
                x, y = -1 + (ex + qx) * h, -1 + (ey + qy) * h
                dux = -np.pi / 2 * np.sin(np.pi / 2 * x) * np.cos(np.pi / 2 * y)
                duy = -np.pi / 2 * np.cos(np.pi / 2 * x) * np.sin(np.pi / 2 * y)
                err2 += wx * wy * h * h * ((c @ dphix - dux) ** 2 + (c @ dphiy - duy) ** 2)
    return np.sqrt(err2)


def swipdg_forms(sigma_inner=8.0, sigma_dirichlet=14.0, hI=D.HI_VOLUME):
    element = D.form(D.integrand(D.INT_LAPLACE, diffusion=1.0))
    coupling = D.form(
        [
            D.integrand(D.INT_IPDG_INNER_COUPLING, diffusion=1.0, weight=1.0, prefactor=1.0),
            D.integrand(D.INT_IPDG_INNER_PENALTY, weight=1.0, prefactor=sigma_inner, hI_kind=hI),
        ]
    )
    boundary = D.form(
        [
            D.integrand(D.INT_IPDG_BOUNDARY_PENALTY, weight=1.0, prefactor=sigma_dirichlet, hI_kind=hI),
            D.integrand(D.INT_IPDG_DIRICHLET_COUPLING, diffusion=1.0, prefactor=1.0),
        ]
    )
    force = D.fn_builtin(D.BUILTIN_COS_PRODUCT, 2, np.pi**2 / 2, np.pi / 2)
    rhs = D.form(D.integrand(D.INT_PRODUCT, diffusion=1.0, weight=force))
    return element, coupling, boundary, rhs


def test_swipdg_esv2007_h1_errors(oracle):
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla

    expected = [2.52e-01, 1.26e-01, 6.30e-02]  # norm.H_1_semi of the YaspGrid variant (.mini:31)
    for N, ref in zip((8, 16, 32), expected):
        g = D.grid_desc(-1.0, 1.0, [N, N])
        rp, ci = oracle.pattern(g, (DG, 1), stencil=D.STENCIL_ELEMENT_AND_INTERSECTION)
        el, co, bo, rhs = swipdg_forms()
        v, b = oracle.assemble(g, DG, 1, rp, ci, [el], [co], [bo], [rhs])
        A = sp.csr_matrix((v, ci, rp), shape=(len(b), len(b)))
        assert abs(A - A.T).max() < 1e-12  # symmetric interior penalty
        u = spla.spsolve(A.tocsc(), b)
        err = _q1_dg_h1_semi_error(N, u)
        assert err == pytest.approx(ref, rel=6e-3), (N, err)


# ---- FV: dune/gdt/test/linear-transport/linear_transport__1d__explicit__fv.{cc,mini} ----------------------------
@pytest.mark.parametrize("numflux", [D.NUMFLUX_UPWIND, D.NUMFLUX_LAX_FRIEDRICHS])
@pytest.mark.parametrize("N,expected", [(16, 1.77e-01), (32, 1.25e-01), (64, 8.84e-02)])
def test_fv_linear_transport_1d(oracle, N, expected, numflux):
    g = D.grid_desc([0.0], [1.0], [N], periodic=1)
    u0 = oracle.fv_interpolate(g, D.fn_builtin(D.BUILTIN_INDICATOR, 0, 0.25, 0.5))
    assert u0.sum() == N // 4 or u0.sum() == N // 4 + 1
    fl = D.flux(D.FLUX_LINEAR, numflux, [1.0])
    dt = 1.0 / N  # dt = h: the scheme is an exact shift (examples/mpi_2019_02...cc:286-288)
    u, steps, time = u0.copy(), 1, 0.0
    worst = 0.0
    while time < 1.0 + dt:  # examples/mpi_2019_02...cc:152
        un = oracle.fv_euler(g, fl, u, dt, 1)
        np.testing.assert_allclose(un, np.roll(u, 1), atol=1e-15)  # exact shift
        assert abs(un.sum() - u0.sum()) / u0.sum() <= 1e-15  # quantity.rel_mass_conserv_error = 0
        # L_infty(L_2): the reference compares the (in time linearly interpolated) discrete solution with the
        # travelling indicator; the gap peaks mid-step, where the front sits in the middle of a cell:
        # 2 fronts * h * (1/2)^2 = h/2
        mid, t_mid = 0.5 * (u + un), time + 0.5 * dt
        err2 = 0.0
        for sub in (0.25, 0.75):
            x = (np.arange(N) + sub) / N
            xi = np.fmod(x - t_mid + 10.0, 1.0)
            exact = ((0.25 <= xi) & (xi <= 0.5)).astype(float)
            err2 += ((mid - exact) ** 2).sum() * (0.5 / N)
        err = np.sqrt(err2)
        worst = max(worst, err)
        u, steps, time = un, steps + 1, time + dt
    assert steps == N + 2  # quantity.num_timesteps = [18 34 66]
    assert worst == pytest.approx(expected, rel=5e-3)  # norm.L_infty_L_2 = sqrt(h/2)


def test_fv_burgers_1d_mass_and_reference_loop(oracle):
    # dune/gdt/test/burgers/burgers__1d__explicit__fv.mini:3-15 (fixed dt, rel. mass conservation error 0)
    N, dt = 16, 0.0096815612792968738
    g = D.grid_desc([0.0], [1.0], [N], periodic=1)
    u0 = oracle.fv_interpolate(g, D.fn_builtin(D.BUILTIN_GAUSSIAN, 3, 0.33, 0.075))
    fl = D.flux(D.FLUX_BURGERS, D.NUMFLUX_UPWIND, [])
    u = oracle.fv_euler(g, fl, u0, dt, 105)
    assert abs(u.sum() - u0.sum()) / u0.sum() < 5e-15
    # independent numpy restatement of upwind Burgers (u >= 0: flux = u_left^2 / 2)
    v = u0.copy()
    for _ in range(105):
        left = np.roll(v, 1)
        fl_right = np.where((v + np.roll(v, -1)) / 2 > 0, 0.5 * v * v, 0.5 * np.roll(v, -1) ** 2)
        fl_left = np.where((left + v) / 2 > 0, 0.5 * left * left, 0.5 * v * v)
        v = v - dt * N * (fl_right - fl_left)
    np.testing.assert_allclose(u, v, rtol=1e-12, atol=1e-14)


# ---- SURVEY 8f "next" rows: constraints, norms, interpolation (the steps around the assembly in config 1) --------
@pytest.mark.parametrize("dim,n,order,expected", [
    (2, [5, 4], 1, 2 * (6 + 5) - 4),  # boundary vertices of a 6 x 5 vertex lattice
    (2, [5, 4], 2, 2 * (11 + 9) - 4),
    (3, [3, 4, 2], 1, 4 * 5 * 3 - 2 * 3 * 1),
    (3, [2, 2, 2], 2, 5**3 - 3**3),
    (1, [7], 3, 2),
])
def test_dirichlet_dofs_are_the_boundary_lattice_points(oracle, dim, n, order, expected):
    """tools/dirichlet-constraints.hh:85-110 with AllDirichletBoundaryInfo: every Lagrange point on the domain boundary,
    nothing else; the set is ascending and duplicate-free (std::set)"""
    g = D.grid_desc(0.0, 1.0, n)
    dofs = oracle.dirichlet_dofs(g, CG, order)
    assert len(dofs) == expected
    assert np.all(np.diff(dofs) > 0)
    # geometric cross-check through the interpolation of the coordinate functions
    on_boundary = np.zeros(oracle.space_size(g, CG, order), dtype=bool)
    for k in range(dim):
        p = [0.0] * 4
        p[1 + k] = 1.0
        xk = oracle.lagrange_interpolate(g, CG, order, D.fn_builtin(D.BUILTIN_AFFINE, 1, *p))
        on_boundary |= (np.abs(xk) < 1e-14) | (np.abs(xk - 1.0) < 1e-14)
    assert np.array_equal(np.where(on_boundary)[0], dofs)


def test_dirichlet_mask_and_non_lagrange_spaces(oracle):
    g = D.grid_desc(0.0, 1.0, [4, 3])
    left = oracle.dirichlet_dofs(g, CG, 1, boundary_mask=0b0001)
    assert np.array_equal(left, np.arange(4) * 5)  # vertices with ix == 0
    assert len(oracle.dirichlet_dofs(g, FV, 0)) == 0  # P0: the only local key sits in the element interior
    dg = oracle.dirichlet_dofs(g, DG, 1)
    assert len(dg) == 2 * (2 * 4 + 2 * 3) - 4  # two Q1 DoFs per boundary face; a corner element's corner DoF counts once
    gp = D.grid_desc(0.0, 1.0, [4, 3], periodic=1)
    assert len(oracle.dirichlet_dofs(gp, DG, 1)) == 2 * (2 * 4)  # no boundary intersections in the periodic direction


def test_heat_equation_example_with_constraints_and_norms(oracle):
    """examples/stationary-heat-equation.cc:87-127 step by step with the oracle: assemble, DirichletConstraints::apply,
    solve, H^1-semi and L^2 errors by BilinearForm::apply2 -- O(h) and O(h^2)"""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla

    exact = D.fn_builtin(D.BUILTIN_COS_PRODUCT, 3, 1.0, np.pi / 2)
    h1, l2 = [], []
    for N in (8, 16, 32):
        g = D.grid_desc(-1.0, 1.0, [N, N])
        rp, ci = oracle.pattern(g, (CG, 1))
        src = D.fn_builtin(D.BUILTIN_COS_PRODUCT, 3, np.pi**2 / 2, np.pi / 2)
        rhs = D.form(D.integrand(D.INT_PRODUCT, diffusion=1.0, weight=src))
        v, b = oracle.assemble(g, CG, 1, rp, ci, [laplace()], rhs_forms=[rhs])
        dofs = oracle.dirichlet_dofs(g, CG, 1)
        v, b = oracle.dirichlet_apply(rp, ci, v, b, dofs)
        A = sp.csr_matrix((v, ci, rp), shape=(len(b), len(b)))
        assert abs(A - A.T).max() == 0.0  # unit_col + unit_row keep the symmetry
        assert np.all(A.diagonal()[dofs] == 1.0) and np.all(b[dofs] == 0.0)
        u = spla.spsolve(A.tocsc(), b)
        assert np.all(u[dofs] == 0.0)
        assert np.allclose(oracle.csr_mv(rp, ci, v, u), b, atol=1e-12)
        h1.append(np.sqrt(oracle.bilinear_form_apply2(g, CG, 1, u, exact, laplace())))
        l2.append(np.sqrt(oracle.bilinear_form_apply2(g, CG, 1, u, exact, D.form(D.integrand(D.INT_PRODUCT)))))
    assert h1[0] / h1[1] == pytest.approx(2.0, rel=0.05) and h1[1] / h1[2] == pytest.approx(2.0, rel=0.05)
    assert l2[0] / l2[1] == pytest.approx(4.0, rel=0.05) and l2[1] / l2[2] == pytest.approx(4.0, rel=0.05)


def test_apply2_norms_of_known_functions(oracle):
    """BilinearForm::apply2 (operators/bilinear-form.hh:340-452): |x + 2y|_{H1}^2 = 5 |Omega|, ||1||_{L2}^2 = |Omega|;
    the interpolant of an affine function reproduces it, so the error norms vanish"""
    g = D.grid_desc([0.0, 0.0], [2.0, 1.0], [6, 5])
    f = D.fn_builtin(D.BUILTIN_AFFINE, 1, 0.5, 1.0, 2.0)
    assert oracle.bilinear_form_apply2(g, CG, 1, None, f, laplace()) == pytest.approx(10.0, rel=1e-13)
    one = D.fn_const(1.0)
    assert oracle.bilinear_form_apply2(g, CG, 1, None, one, D.form(D.integrand(D.INT_PRODUCT))) == pytest.approx(2.0, rel=1e-13)
    for order in (1, 2):
        u = oracle.lagrange_interpolate(g, CG, order, f)
        assert oracle.bilinear_form_apply2(g, CG, order, u, f, laplace()) < 1e-24
        assert oracle.bilinear_form_apply2(g, CG, order, u, None, laplace()) == pytest.approx(10.0, rel=1e-13)


def test_swipdg_esv2007_h1_errors_via_apply2(oracle):
    """the ESV2007 table again (.mini:31), the error norm now by the restated BilinearForm::apply2 with the
    exact solution of declared order 4 instead of the hand-written numpy quadrature above"""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla

    exact = D.fn_builtin(D.BUILTIN_COS_PRODUCT, 4, 1.0, np.pi / 2)
    for N, ref in zip((8, 16, 32), [2.52e-01, 1.26e-01, 6.30e-02]):
        g = D.grid_desc(-1.0, 1.0, [N, N])
        rp, ci = oracle.pattern(g, (DG, 1), stencil=D.STENCIL_ELEMENT_AND_INTERSECTION)
        el, co, bo, rhs = swipdg_forms()
        v, b = oracle.assemble(g, DG, 1, rp, ci, [el], [co], [bo], [rhs])
        u = spla.spsolve(sp.csr_matrix((v, ci, rp), shape=(len(b), len(b))).tocsc(), b)
        err = np.sqrt(oracle.bilinear_form_apply2(g, DG, 1, u, exact, laplace()))
        assert err == pytest.approx(_q1_dg_h1_semi_error(N, u), rel=1e-9)
        assert err == pytest.approx(ref, rel=6e-3), (N, err)


# ------------------------------------------------------------------------------------------------------------------
# pointwise integrand identities of the reference's own tests: polynomial bases
# {x, x^2 y} (ansatz) / {y, x y^3} (test) and the NON-symmetric diffusion tensor kappa(x) = x y [[x, y], [1, 2]]
# (dune/gdt/test/integrands/integrands.hh:75-104, integrands_laplace.cc:45-52, 71-91; integrands_product.cc:58-76).
# They pin the index convention of a full tensor: values[i][j] = (kappa grad phi_j) . grad psi_i, kappa row-major.
# ------------------------------------------------------------------------------------------------------------------
def _kat_points(oracle, order):
    x1, _ = oracle.gauss_rule(order)
    return [(a, b) for b in x1 for a in x1]


def _ulp_close(a, b, ulps=4):
    # EXPECT_DOUBLE_EQ: within 4 units in the last place
    return np.all(np.abs(a - b) <= ulps * np.spacing(np.maximum(np.abs(a), np.abs(b))))


def test_laplace_integrand_pointwise_reference_kat(oracle):
    # integrand order = kappa.order (3) + test.order (4) + ansatz.order (3) (laplace.hh:74-79)
    for x, y in _kat_points(oracle, 3 + 4 + 3):
        kappa = x * y * np.array([[x, y], [1.0, 2.0]])
        integrand = D.integrand(D.INT_LAPLACE, diffusion=D.fn_const(kappa, order=3))
        ansatz_v, ansatz_g = [x, x**2 * y], [[1.0, 0.0], [2.0 * x * y, x**2]]
        test_v, test_g = [y, x * y**3], [[0.0, 1.0], [y**3, 3.0 * x * y**2]]
        got = oracle.element_integrand_evaluate(integrand, 2, test_v, test_g, ansatz_v, ansatz_g, [x, y])
        expected = x * y * np.array([
            [1.0, 2.0 * (x * y + x**2)],
            [x * y**3 + 3.0 * x * y**2, 3.0 * x**2 * y**4 + 6.0 * (x**2 * y**3 + x**3 * y**2)],
        ])
        assert _ulp_close(got, expected), (x, y, got, expected)
        # the transposed convention (kappa^T, i.e. (kappa grad psi_i) . grad phi_j) must NOT reproduce the table
        wrong = D.integrand(D.INT_LAPLACE, diffusion=D.fn_const(kappa.T.copy(), order=3))
        bad = oracle.element_integrand_evaluate(wrong, 2, test_v, test_g, ansatz_v, ansatz_g, [x, y])
        assert not _ulp_close(bad, expected, ulps=64)


def test_product_integrand_pointwise_reference_kat(oracle):
    for x, y in _kat_points(oracle, 2 + 4 + 3):  # weight.order + test.order + ansatz.order (product.hh:89-100)
        integrand = D.integrand(D.INT_PRODUCT, diffusion=D.fn_const(x * y, order=2))
        got = oracle.element_integrand_evaluate(integrand, 2, [y, x * y**3], np.zeros((2, 2)), [x, x**2 * y],
                                                np.zeros((2, 2)), [x, y])
        expected = np.array([[(x * y) ** 2, (x * y) ** 3], [x**3 * y**4, x**4 * y**5]])
        assert _ulp_close(got, expected), (x, y, got, expected)
