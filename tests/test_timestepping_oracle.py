"""CPU checks of the oracle's restatement of SURVEY.md 8f row n3: estimate_dt_for_hyperbolic_system
(tools/hyperbolic.hh:38-86), ExplicitRungeKuttaTimeStepper (tools/timestepper/explicit-rungekutta.hh:63-270),
TimeStepperInterface::solve (tools/timestepper/interface.hh:191-263) and the FV boundary treatments
(local/operators/advection-fv.hh:188-457).  Pinned by the reference's own EOC tables: quantity.CFL = dt / explicit_fv_dt
and quantity.num_timesteps of linear_transport__1d__explicit__fv.mini:8-14 and burgers__1d__explicit__fv.mini:3-30."""
import numpy as np
import pytest

from dune_gdt_b200 import descriptors as D

BUTCHER = D.BUTCHER


# ---- dt estimate: the reference's tables print CFL = dt / estimate_dt_for_hyperbolic_system(...) -------------------
@pytest.mark.parametrize("N", [16, 32, 64])
def test_estimate_dt_linear_transport_cfl_table(oracle, N):
    """linear_transport__1d__explicit__fv.mini:14: quantity.CFL = [2 2 2] with dt = h (the exact-shift step)"""
    g = D.grid_desc([0.0], [1.0], [N], periodic=1)
    u0 = oracle.fv_interpolate(g, D.fn_builtin(D.BUILTIN_INDICATOR, 0, 0.25, 0.5))
    est = oracle.fv_estimate_dt(g, D.flux(D.FLUX_LINEAR, D.NUMFLUX_UPWIND, [1.0]), u0)
    assert (1.0 / N) / est == pytest.approx(2.0, rel=1e-14)


@pytest.mark.parametrize("numflux,fixed_dt,cfl,timesteps", [
    (D.NUMFLUX_UPWIND, 0.0096815612792968738, [2.31e-01, 4.80e-01], 107),         # burgers__1d__explicit__fv.mini:7-15
    (D.NUMFLUX_LAX_FRIEDRICHS, 0.009193328857421872, [2.20e-01, 4.56e-01], 112),  # :22-30
])
def test_estimate_dt_burgers_cfl_table(oracle, numflux, fixed_dt, cfl, timesteps):
    """quantity.CFL = (dt_factor * use_fixed_dt) / explicit_fv_dt with setup.dt_factor = 0.99 (test/burgers/base.hh:154-168);
    quantity.num_timesteps counts the time points of solve_instationary_system_explicit_euler
    (test/instationary-eocstudies/base.hh:429-446): `while (time < T_end + dt)`"""
    dt = 0.99 * fixed_dt
    for N, expected in zip((16, 32), cfl):
        g = D.grid_desc([0.0], [1.0], [N], periodic=1)
        u0 = oracle.fv_interpolate(g, D.fn_builtin(D.BUILTIN_GAUSSIAN, 3, 0.33, 0.075))
        est = oracle.fv_estimate_dt(g, D.flux(D.FLUX_BURGERS, numflux, []), u0)
        assert float(f"{dt / est:.2e}") == expected  # the table prints three significant digits
    time, points = 0.0, 1
    while time < 1.0 + dt:
        time += dt
        points += 1
    assert points == timesteps


def test_estimate_dt_defaults_and_degenerate_ranges(oracle):
    """hyperbolic.hh:47-48,62-64: default boundary range {max(), min()} (min() = smallest positive normal) and the
    1e-6 widening of a degenerate data range"""
    g = D.grid_desc(0.0, 1.0, [8, 4], periodic=3)
    fl = D.flux(D.FLUX_BURGERS, D.NUMFLUX_UPWIND, [])
    u = np.full(32, 2.0)
    # range [2, 2 + 2e-6]; Gauss-2 maximum point; perimeter / volume = 2 (h_x + h_y) / (h_x h_y)
    umax = 2.0 + (0.5 + 0.5 / np.sqrt(3.0)) * 2e-6
    pov = 2.0 * (1 / 8 + 1 / 4) / (1 / 8 * 1 / 4)
    assert oracle.fv_estimate_dt(g, fl, u) == pytest.approx(1.0 / (pov * umax), rel=1e-13)
    # all-negative state: data_maximum stays at numeric_limits<double>::min() > 0
    u = -np.linspace(1.0, 2.0, 32)
    est = oracle.fv_estimate_dt(g, fl, u)
    lo, hi = -2.0, np.finfo(np.float64).tiny
    pts = lo + (0.5 + np.array([-0.5, 0.5]) / np.sqrt(3.0)) * (hi - lo)
    assert est == pytest.approx(1.0 / (pov * np.abs(pts).max()), rel=1e-13)
    # an explicit boundary data range widens the state range
    est2 = oracle.fv_estimate_dt(g, fl, u, boundary_data_range=[-4.0, 1.0])
    pts = -4.0 + (0.5 + np.array([-0.5, 0.5]) / np.sqrt(3.0)) * 5.0
    assert est2 == pytest.approx(1.0 / (pov * np.abs(pts).max()), rel=1e-13)


# ---- Runge-Kutta ------------------------------------------------------------------------------------------------
def _op_matrix(oracle, g, fl, n):
    return np.array([oracle.fv_apply(g, fl, e) for e in np.eye(n)]).T


@pytest.mark.parametrize("method,order", [(D.RK_EULER, 1), (D.RK_SSP2, 2), (D.RK_SSP3, 3), (D.RK_CLASSIC4, 4)])
def test_rk_step_is_the_taylor_polynomial_for_a_linear_operator(oracle, method, order):
    """For linear L and u_t = r L u a p-stage, order-p method gives u_1 = sum_{k<=p} (r dt L)^k / k! u_0"""
    n = [7, 5]
    g = D.grid_desc(0.0, 1.0, n, periodic=3)
    fl = D.flux(D.FLUX_LINEAR, D.NUMFLUX_UPWIND, [1.0, -0.5])
    u0 = np.random.default_rng(20251017).uniform(-1, 1, 35)
    Lm = _op_matrix(oracle, g, fl, 35)
    dt, r = 0.013, -1.0
    u1, t1 = oracle.rk_step(g, fl, BUTCHER[method], u0, t=0.25, dt=dt, r=r)
    expect, term = u0.copy(), u0.copy()
    for k in range(1, order + 1):
        term = (r * dt / k) * (Lm @ term)
        expect += term
    np.testing.assert_allclose(u1, expect, rtol=0, atol=1e-14 * np.abs(u0).max() * 10)
    assert t1 == 0.25 + dt


def test_rk_euler_step_equals_the_examples_euler_loop(oracle):
    """r = -1, explicit_euler: u + k (r dt b_0) == u - L(u) dt (examples/mpi_2019_02...cc:154) bit for bit"""
    g = D.grid_desc([0.0], [1.0], [32], periodic=1)
    fl = D.flux(D.FLUX_BURGERS, D.NUMFLUX_UPWIND, [])
    u0 = oracle.fv_interpolate(g, D.fn_builtin(D.BUILTIN_GAUSSIAN, 3, 0.33, 0.075))
    a, _ = oracle.rk_step(g, fl, BUTCHER[D.RK_EULER], u0, 0.0, 0.004, r=-1.0)
    b = oracle.fv_euler(g, fl, u0, 0.004, 1)
    assert np.array_equal(a, b)


def test_rk_ssp_methods_are_convex_combinations_of_euler_steps(oracle):
    """Shu-Osher form: SSP2 = 1/2 u + 1/2 E(E(u)); SSP3 = 1/3 u + 2/3 E(3/4 u + 1/4 E(E(u))) -- also for nonlinear L"""
    g = D.grid_desc(0.0, 1.0, [24, 6], periodic=3)
    fl = D.flux(D.FLUX_BURGERS, D.NUMFLUX_LAX_FRIEDRICHS, [])
    u = np.random.default_rng(3).uniform(0.1, 1.0, 144)
    dt = 0.002
    E = lambda v: oracle.fv_euler(g, fl, v, dt, 1)  # noqa: E731
    ssp2, _ = oracle.rk_step(g, fl, BUTCHER[D.RK_SSP2], u, 0.0, dt, r=-1.0)
    np.testing.assert_allclose(ssp2, 0.5 * u + 0.5 * E(E(u)), rtol=1e-13)
    ssp3, _ = oracle.rk_step(g, fl, BUTCHER[D.RK_SSP3], u, 0.0, dt, r=-1.0)
    np.testing.assert_allclose(ssp3, u / 3 + 2 / 3 * E(0.75 * u + 0.25 * E(E(u))), rtol=1e-13)


def test_rk_solve_step_plan(oracle):
    """interface.hh:216-226: full steps of initial_dt, the last one cut to hit t_end; FloatCmp::lt ends the loop"""
    g = D.grid_desc([0.0], [1.0], [16], periodic=1)
    fl = D.flux(D.FLUX_LINEAR, D.NUMFLUX_UPWIND, [1.0])
    u0 = oracle.fv_interpolate(g, D.fn_builtin(D.BUILTIN_INDICATOR, 0, 0.25, 0.5))
    # dt = h divides T = 1: exactly 16 exact-shift steps, back to the start
    u, steps, t = oracle.rk_solve(g, fl, BUTCHER[D.RK_EULER], u0, 1.0, 1.0 / 16, r=-1.0)
    assert steps == 16 and t == 1.0
    np.testing.assert_allclose(u, u0, atol=1e-15)
    # dt = 0.3: 3 full steps + one of 0.1
    u, steps, t = oracle.rk_solve(g, fl, BUTCHER[D.RK_SSP2], u0, 1.0, 0.03, r=-1.0)
    assert steps == 34 and abs(t - 1.0) < 1e-14
    assert abs(u.sum() - u0.sum()) < 1e-13  # conservative


# ---- boundary treatments ------------------------------------------------------------------------------------------
def test_boundary_extrapolation_equals_a_ghost_cell(oracle):
    """...ByCustomExtrapolationOperator (local/operators/advection-fv.hh:418-443) on a 1D grid: the boundary cell sees
    the numerical flux against v = a u + b, scaled by |I| / |E| = 1 / h"""
    N = 10
    g = D.grid_desc([0.0], [2.0], [N])
    h = 2.0 / N
    u = np.random.default_rng(5).uniform(0.2, 1.0, N)
    for fl in (D.flux(D.FLUX_LINEAR, D.NUMFLUX_UPWIND, [1.5]), D.flux(D.FLUX_BURGERS, D.NUMFLUX_LAX_FRIEDRICHS, [])):
        plain = oracle.fv_apply(g, fl, u)
        bnd = [D.fv_boundary(D.FVBND_EXTRAPOLATION, 0b01, 0.0, 0.7), D.fv_boundary(D.FVBND_EXTRAPOLATION, 0b10, 1.0, 0.0)]
        out = oracle.fv_apply_bnd(g, fl, bnd, u)
        np.testing.assert_array_equal(out[1:-1], plain[1:-1])
        # extended periodic-free grid with explicit ghost cells carrying v
        ge = D.grid_desc([-h], [2.0 + h], [N + 2])
        ue = np.concatenate([[0.7], u, [u[-1]]])
        ext = oracle.fv_apply(ge, fl, ue)
        np.testing.assert_allclose(out, ext[1:-1], rtol=1e-13, atol=1e-15)


def test_boundary_numerical_flux_and_summation(oracle):
    """...ByCustomNumericalFluxOperator (local/operators/advection-fv.hh:281-296): g = a f(u).n + b on the selected
    sides only; treatments on the same side add up; outflow of the physical flux (a = 1) equals absorbing extrapolation"""
    g = D.grid_desc(0.0, 1.0, [6, 4])
    fl = D.flux(D.FLUX_LINEAR, D.NUMFLUX_UPWIND, [1.0, 0.5])
    u = np.random.default_rng(7).uniform(0.0, 1.0, 24)
    plain = oracle.fv_apply(g, fl, u)
    out = oracle.fv_apply_bnd(g, fl, [D.fv_boundary(D.FVBND_NUMERICAL_FLUX, 0b0010, 0.0, 3.0)], u).reshape(4, 6)
    ref = plain.reshape(4, 6).copy()
    ref[:, -1] += 3.0 * 6  # g |I| / |E| = 3 / h_x on the x+ side
    np.testing.assert_allclose(out, ref, rtol=1e-14)
    two = oracle.fv_apply_bnd(g, fl, [D.fv_boundary(D.FVBND_NUMERICAL_FLUX, 0b0010, 0.0, 1.0),
                                     D.fv_boundary(D.FVBND_NUMERICAL_FLUX, 0b0010, 0.0, 2.0)], u)
    np.testing.assert_allclose(two.reshape(4, 6), ref, rtol=1e-14)
    a = oracle.fv_apply_bnd(g, fl, [D.fv_boundary(D.FVBND_NUMERICAL_FLUX, 0b1111, 1.0, 0.0)], u)
    b = oracle.fv_apply_bnd(g, fl, [D.fv_boundary(D.FVBND_EXTRAPOLATION, 0b1111, 1.0, 0.0)], u)
    np.testing.assert_allclose(a, b, rtol=1e-14, atol=1e-15)
    # constant state + absorbing boundaries: nothing changes (div of a constant flux = 0)
    c = oracle.fv_apply_bnd(g, fl, [D.fv_boundary(D.FVBND_EXTRAPOLATION, 0b1111, 1.0, 0.0)], np.full(24, 0.3))
    np.testing.assert_allclose(c, 0.0, atol=1e-14)
