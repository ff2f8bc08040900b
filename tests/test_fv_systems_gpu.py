"""FV advection operator for systems (m > 1) on the GPU against the oracle: the Euler equations with the
Vijayasundaram and Lax-Friedrichs fluxes (examples/mpi_2019_02_talk_on_hyperbolic_equations.cc:381-434,
test/inviscid-compressible-flow/), through the C ABI (gdtb_fv_space_create, gdtb_fvop_*)."""
import numpy as np
import pytest

from dune_gdt_b200 import descriptors as D
from helpers import TOL, rel_err

pytestmark = pytest.mark.gpu

GAMMA = 1.4


def random_states(oracle, d, ne, seed=11):
    rng = np.random.default_rng(seed)
    return np.stack([oracle.euler_conservative(GAMMA, rng.uniform(0.8, 1.5), rng.uniform(-0.4, 0.4, d), rng.uniform(0.6, 1.2))
                     for _ in range(ne)]).ravel()


CASES = [([16], 1), ([7], 1), ([9], 0), ([12, 9], 3), ([8, 5], 1), ([6, 7], 2), ([5, 4], 0), ([1, 6], 3), ([3, 1], 1)]


@pytest.mark.parametrize("n,periodic", CASES)
@pytest.mark.parametrize("numflux,params", [(D.NUMFLUX_VIJAYASUNDARAM, [GAMMA]), (D.NUMFLUX_LAX_FRIEDRICHS, [GAMMA, 0.35])])
def test_euler_apply_parity(gdt, ctx, oracle, n, periodic, numflux, params):
    d = len(n)
    lo, up = [0.0, -1.0][:d], [3.0, 1.0][:d]  # anisotropic cells
    gdesc = D.grid_desc(lo, up, n, periodic)
    grid = gdt.Grid(ctx, gdesc)
    space = gdt.make_finite_volume_space(grid, d + 2)
    ne = int(np.prod(n))
    assert space.mapper.size == ne * (d + 2)
    assert np.array_equal(space.mapper.global_indices(ne - 1), np.arange((ne - 1) * (d + 2), ne * (d + 2)))
    flux = D.flux(D.FLUX_EULER, numflux, params)
    cls = gdt.NumericalVijayasundaramFlux if numflux == D.NUMFLUX_VIJAYASUNDARAM else gdt.NumericalLaxFriedrichsFlux
    op = gdt.make_advection_fv_operator(cls(D.FLUX_EULER, params), space)
    u = random_states(oracle, d, ne)
    got = op.apply(u)
    ref = oracle.fvsys_apply(gdesc, flux, u)
    assert rel_err(got, ref) <= TOL
    assert np.array_equal(got, op.apply(u))  # bit-identical reruns


@pytest.mark.parametrize("N", [16, 32])
def test_euler_1d_shock_tube_time_loop_and_dt(gdt, ctx, oracle, N):
    """the reference's 1D EOC test setup (test/inviscid-compressible-flow/base.hh): dt estimate, step count, explicit
    Euler loop on the device == the oracle; mass is conserved"""
    from test_fv_systems_oracle import shock_tube_1d

    gdesc, u0 = shock_tube_1d(oracle, N)
    grid = gdt.Grid(ctx, gdesc)
    space = gdt.make_finite_volume_space(grid, 3)
    euler = gdt.EulerTools(1, GAMMA)
    op = gdt.make_advection_fv_operator(gdt.NumericalVijayasundaramFlux(*euler.flux()), space)
    flux = D.flux(D.FLUX_EULER, D.NUMFLUX_VIJAYASUNDARAM, [GAMMA])
    est = op.estimate_dt(u0)
    assert est == pytest.approx(oracle.fvsys_estimate_dt(gdesc, flux, u0), rel=1e-14)
    dt = 0.99 * est
    steps, t = 0, 0.0
    while t < 1.0 + dt:
        t += dt
        steps += 1
    assert steps + 1 == {16: 64, 32: 126}[N]  # inviscid_compressible_flow__euler_1d__explicit__fv.mini:12
    u = op.explicit_euler(u0, dt, steps)
    ref = oracle.fvsys_euler(gdesc, flux, u0, dt, steps)
    assert rel_err(u, ref) <= 1e-11  # 63 / 125 nonlinear steps of rounding-level differences
    assert abs(u.reshape(N, 3)[:, 0].sum() - u0.reshape(N, 3)[:, 0].sum()) <= 1e-13 * N


def test_euler_2d_driver_setup(gdt, ctx, oracle):
    """examples/mpi_2019_02_talk_on_hyperbolic_equations.cc:381-434 on a coarser grid: 2D periodic, the box initial
    values, dt from estimate_dt_for_hyperbolic_system, a few explicit Euler steps"""
    N = 24
    gdesc = D.grid_desc([-1.0, -1.0], [1.0, 1.0], [N, N], periodic=3)
    grid = gdt.Grid(ctx, gdesc)
    space = gdt.make_finite_volume_space(grid, 4)
    euler = gdt.EulerTools(2, GAMMA)
    c = -1.0 + (np.arange(N) + 0.5) * (2.0 / N)
    X, Y = np.meshgrid(c, c, indexing="xy")  # x fastest in the element numbering
    inside = ((X >= -0.5) & (X <= 0.0) & (Y >= -0.5) & (Y <= 0.0)).ravel()
    u0 = np.where(inside[:, None], euler.conservative(4.0, 0.0, 1.6)[None, :], euler.conservative(1.0, 0.0, 0.4)[None, :]).ravel()
    op = gdt.make_advection_fv_operator(gdt.NumericalVijayasundaramFlux(*euler.flux()), space)
    flux = D.flux(D.FLUX_EULER, D.NUMFLUX_VIJAYASUNDARAM, [GAMMA])
    dt = op.estimate_dt(u0)
    assert dt == pytest.approx(oracle.fvsys_estimate_dt(gdesc, flux, u0), rel=1e-14)
    u = op.explicit_euler(u0, dt, 10)
    ref = oracle.fvsys_euler(gdesc, flux, u0, dt, 10)
    assert rel_err(u, ref) <= 1e-12
    rho, v, p = euler.primitive(u.reshape(-1, 4))
    assert rho.min() > 0.9 and p.min() > 0.3
    assert abs(u.reshape(-1, 4)[:, 0].sum() - u0.reshape(-1, 4)[:, 0].sum()) <= 1e-12 * N * N


@pytest.mark.parametrize("n", [[16], [5], [10, 7], [4, 9], [1, 5]])
@pytest.mark.parametrize("kind", ["wall", "mirror", "wall-lower-mirror-upper"])
@pytest.mark.parametrize("numflux,params", [(D.NUMFLUX_VIJAYASUNDARAM, [GAMMA]), (D.NUMFLUX_LAX_FRIEDRICHS, [GAMMA, 0.35])])
def test_euler_impermeable_walls_parity(gdt, ctx, oracle, n, kind, numflux, params):
    """the two impermeable-wall treatments of test/inviscid-compressible-flow/base.hh:187-241 on every domain side"""
    d = len(n)
    lo, up = [0.0, -1.0][:d], [3.0, 1.0][:d]
    gdesc = D.grid_desc(lo, up, n, 0)
    all_sides = (1 << (2 * d)) - 1
    lower = sum(1 << (2 * k) for k in range(d))
    wall_mask, mirror_mask = {"wall": (all_sides, 0), "mirror": (0, all_sides),
                              "wall-lower-mirror-upper": (lower, all_sides & ~lower)}[kind]
    space = gdt.make_finite_volume_space(gdt.Grid(ctx, gdesc), d + 2)
    cls = gdt.NumericalVijayasundaramFlux if numflux == D.NUMFLUX_VIJAYASUNDARAM else gdt.NumericalLaxFriedrichsFlux
    op = gdt.make_advection_fv_operator(cls(D.FLUX_EULER, params), space)
    if wall_mask:
        op.append(D.fv_boundary(D.FVBND_EULER_IMPERMEABLE_WALL, wall_mask, 0.0, 0.0))
    if mirror_mask:
        op.append(D.fv_boundary(D.FVBND_EULER_INVISCID_MIRROR, mirror_mask, 0.0, 0.0))
    u = random_states(oracle, d, int(np.prod(n)), seed=23)
    got = op.apply(u)
    ref = oracle.fvsys_apply_walls(gdesc, D.flux(D.FLUX_EULER, numflux, params), u, wall_mask, mirror_mask)
    assert rel_err(got, ref) <= TOL


def test_euler_1d_wall_table_on_the_device(gdt, ctx, oracle):
    """inviscid_compressible_flow__euler_1d__explicit__fv.mini:26 (direct Euler treatment, 16 elements): the momentum
    deviation 3.50e-01 reproduced by the device time loop"""
    from test_fv_systems_oracle import shock_tube_1d

    N = 16
    gper, u0 = shock_tube_1d(oracle, N)
    walls = D.grid_desc([-1.0], [1.0], [N], periodic=0)
    euler = gdt.EulerTools(1, GAMMA)
    per_op = gdt.make_advection_fv_operator(gdt.NumericalVijayasundaramFlux(*euler.flux()),
                                            gdt.make_finite_volume_space(gdt.Grid(ctx, gper), 3))
    dt = 0.99 * per_op.estimate_dt(u0)
    op = gdt.make_advection_fv_operator(gdt.NumericalVijayasundaramFlux(*euler.flux()),
                                        gdt.make_finite_volume_space(gdt.Grid(ctx, walls), 3))
    op.append(D.fv_boundary(D.FVBND_EULER_IMPERMEABLE_WALL, 3, 0.0, 0.0))
    h, u, err = 2.0 / N, u0.copy(), 0.0
    for _ in range(63):
        u = op.explicit_euler(u, dt, 1)
        err = max(err, abs(u.reshape(N, 3)[:, 1].sum() * h))
    assert float(f"{err:.2e}") == 3.50e-01


def test_system_operator_error_conventions(gdt, ctx):
    grid = gdt.Grid(ctx, D.grid_desc([0.0], [1.0], [8], periodic=1))
    scalar, system = gdt.make_finite_volume_space(grid), gdt.make_finite_volume_space(grid, 3)
    with pytest.raises(gdt.capi.ShapesDoNotMatch):  # Euler needs m = d + 2
        gdt.make_advection_fv_operator(gdt.NumericalVijayasundaramFlux(D.FLUX_EULER, [GAMMA]), scalar)
    with pytest.raises(gdt.capi.ShapesDoNotMatch):
        gdt.make_advection_fv_operator(gdt.NumericalUpwindFlux(D.FLUX_LINEAR, [1.0]), system)
    with pytest.raises(gdt.capi.NotImplementedGdt):  # upwind.hh: m = 1 only
        gdt.make_advection_fv_operator(gdt.NumericalUpwindFlux(D.FLUX_EULER, [GAMMA]), system)
    with pytest.raises(gdt.capi.NotImplementedGdt):  # lax-friedrichs.hh:40-41: lambda must be provided for m > 1
        gdt.make_advection_fv_operator(gdt.NumericalLaxFriedrichsFlux(D.FLUX_EULER, [GAMMA]), system)
