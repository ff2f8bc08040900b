"""CPU checks of the oracle's restatement of the FV path for SYSTEMS (m > 1): EulerTools (tools/euler.hh), the
Vijayasundaram / Lax-Friedrichs numerical fluxes (local/numerical-fluxes/vijayasundaram.hh:111-133, lax-friedrichs.hh:66-88),
LocalAdvectionFvCouplingOperator::apply for m components (local/operators/advection-fv.hh:127-153) and
estimate_dt_for_hyperbolic_system for m > 1 (tools/hyperbolic.hh:38-86).  Pinned by the reference's own EOC tables
test/inviscid-compressible-flow/inviscid_compressible_flow__euler_1d__explicit__fv.mini:8-42 (periodic boundaries,
impermeable walls by the direct Euler treatment and by the inviscid mirror treatment)."""
import numpy as np
import pytest

from dune_gdt_b200 import descriptors as D

GAMMA = 1.4


def shock_tube_1d(oracle, N):
    """test/inviscid-compressible-flow/base.hh:62-80: [-1, 1], 4 / 1.6 inside [-0.5, 0], 1 / 0.4 outside, v = 0"""
    g = D.grid_desc([-1.0], [1.0], [N], periodic=1)
    x = -1.0 + (np.arange(N) + 0.5) * (2.0 / N)
    hi = oracle.euler_conservative(GAMMA, 4.0, [0.0], 1.6)
    lo = oracle.euler_conservative(GAMMA, 1.0, [0.0], 0.4)
    u0 = np.where(((x >= -0.5) & (x <= 0.0))[:, None], hi[None, :], lo[None, :]).ravel()
    return g, u0


@pytest.mark.parametrize("N,time_points", [(16, 64), (32, 126), (64, 250)])
def test_euler_1d_periodic_table(oracle, N, time_points):
    """quantity.num_timesteps = [64 126 250], quantity.rel_mass_conserv_error = [0 0 0] (zero_tolerance 1e-15),
    quantity.CFL = 0.99: dt = 0.99 * estimate_dt_for_hyperbolic_system, explicit Euler up to T_end = 1
    (test/inviscid-compressible-flow/base.hh:358-381, test/instationary-eocstudies/base.hh:429-447)"""
    g, u0 = shock_tube_1d(oracle, N)
    fl = D.flux(D.FLUX_EULER, D.NUMFLUX_VIJAYASUNDARAM, [GAMMA])
    dt = 0.99 * oracle.fvsys_estimate_dt(g, fl, u0)
    steps, t = 0, 0.0
    while t < 1.0 + dt:
        t += dt
        steps += 1
    assert steps + 1 == time_points  # num_timesteps counts the time points (initial values included)
    u = oracle.fvsys_euler(g, fl, u0, dt, steps)
    m0, m1 = u0.reshape(N, 3).sum(0), u.reshape(N, 3).sum(0)
    assert abs(m1[0] - m0[0]) / m0[0] <= 1e-15 * 4  # a few ulps of the sum (the table's 0 is < 1e-15 per its tolerance)
    assert np.isfinite(u).all() and u.reshape(N, 3)[:, 0].min() > 0.9


@pytest.mark.parametrize("d", [1, 2])
def test_euler_jacobian_and_eigendecomposition_are_consistent(oracle, d):
    """flux_jacobian is the derivative of flux (central differences), and T diag(lambda) T^{-1} = sum_s n_s A_s,
    T T^{-1} = I for the formulas of tools/euler.hh:325-462 (Kroener's M T, (M T)^{-1})"""
    rng = np.random.default_rng(3)
    for _ in range(5):
        w = oracle.euler_conservative(GAMMA, rng.uniform(0.5, 2.0), rng.uniform(-0.6, 0.6, d), rng.uniform(0.4, 2.0))
        J = oracle.euler_jacobian(d, GAMMA, w)
        eps = 1e-6
        for c in range(d + 2):
            dw = np.zeros(d + 2)
            dw[c] = eps
            fd = (oracle.euler_flux(d, GAMMA, w + dw) - oracle.euler_flux(d, GAMMA, w - dw)) / (2 * eps)
            assert np.abs(J[:, :, c] - fd).max() <= 1e-8
        for n in ([1.0, 0.0][:d], [-1.0, 0.0][:d], [0.0, 1.0][:d] if d == 2 else [1.0], list(rng.normal(size=d))):
            n = np.asarray(n) / np.linalg.norm(n)
            ev, T, Ti = oracle.euler_eigen(d, GAMMA, w, n)
            P = sum(n[s] * J[s] for s in range(d))
            assert np.abs(T @ Ti - np.eye(d + 2)).max() <= 1e-13
            assert np.abs(T @ np.diag(ev) @ Ti - P).max() <= 1e-12 * max(1.0, np.abs(P).max())


def test_vijayasundaram_is_consistent_and_conservative(oracle):
    """g(w, w, n) = f(w) . n (consistency), and a constant state is a fixed point of the operator"""
    for d, n in ((1, [8]), (2, [6, 5])):
        g = D.grid_desc([0.0] * d, [1.0] * d, n, periodic=(1 << d) - 1)
        w = oracle.euler_conservative(GAMMA, 1.2, [0.3, -0.2][:d], 0.8)
        u = np.tile(w, int(np.prod(n)))
        fl = D.flux(D.FLUX_EULER, D.NUMFLUX_VIJAYASUNDARAM, [GAMMA])
        assert np.abs(oracle.fvsys_apply(g, fl, u)).max() <= 1e-13
        # conservation on a periodic grid: sum over cells of |E| L(u) vanishes per component
        rng = np.random.default_rng(5)
        ne = int(np.prod(n))
        states = np.stack([oracle.euler_conservative(GAMMA, rng.uniform(0.8, 1.5), rng.uniform(-0.3, 0.3, d), rng.uniform(0.6, 1.2))
                           for _ in range(ne)])
        for numflux, params in ((D.NUMFLUX_VIJAYASUNDARAM, [GAMMA]), (D.NUMFLUX_LAX_FRIEDRICHS, [GAMMA, 0.4])):
            L = oracle.fvsys_apply(g, D.flux(D.FLUX_EULER, numflux, params), states.ravel()).reshape(ne, d + 2)
            assert np.abs(L.sum(0)).max() <= 1e-12 * np.abs(L).max()


@pytest.mark.parametrize("treatment,wall_mask,mirror_mask,expected", [
    ("impermeable_walls_by_direct_euler_treatment", 3, 0, [3.50e-01, 4.20e-01, 4.57e-01]),   # .mini:26
    ("impermeable_walls_by_inviscid_mirror_treatment", 0, 3, [3.43e-01, 4.06e-01, 4.51e-01]),  # .mini:40
])
def test_euler_1d_impermeable_wall_tables(oracle, treatment, wall_mask, mirror_mask, expected):
    """quantity.rel_mass_conserv_error = max over the time points and components of |m_0 - m(t)| / (m_0 > 0 ? m_0 : 1)
    (test/instationary-eocstudies/base.hh:325-348; the momentum is what the walls change) and quantity.num_timesteps =
    [64 126 250], for both wall treatments of test/inviscid-compressible-flow/base.hh:187-241 -- to the table's 3 digits"""
    for N, time_points, exp in zip((16, 32, 64), (64, 126, 250), expected):
        g, u0 = shock_tube_1d(oracle, N)  # the dt estimate sees the same grid view as the periodic test
        walls = D.grid_desc([-1.0], [1.0], [N], periodic=0)
        fl = D.flux(D.FLUX_EULER, D.NUMFLUX_VIJAYASUNDARAM, [GAMMA])
        dt = 0.99 * oracle.fvsys_estimate_dt(g, fl, u0)
        steps, t = 0, 0.0
        while t < 1.0 + dt:
            t += dt
            steps += 1
        assert steps + 1 == time_points
        h = 2.0 / N
        m0 = u0.reshape(N, 3).sum(0) * h
        u, err = u0.copy(), np.zeros(3)
        for _ in range(steps):
            u = u - oracle.fvsys_apply_walls(walls, fl, u, wall_mask, mirror_mask) * dt
            m = u.reshape(N, 3).sum(0) * h
            err = np.maximum(err, np.abs(m0 - m) / np.where(m0 > 0, m0, 1.0))
        assert float(f"{err.max():.2e}") == exp
        assert err[0] <= 1e-14 and err[2] <= 1e-14  # mass and energy stay conserved: only the momentum sees the walls
