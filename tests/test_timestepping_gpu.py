"""Parity of SURVEY.md 8f row n3 through the C ABI against the oracle: FV boundary treatments
(local/operators/advection-fv.hh:188-457), estimate_dt_for_hyperbolic_system (tools/hyperbolic.hh:38-86),
ExplicitRungeKuttaTimeStepper::step / TimeStepperInterface::solve (tools/timestepper/explicit-rungekutta.hh:237-270,
interface.hh:191-263).  Step plans (counts, end time) are exact; values within 1e-12 relative."""
import numpy as np
import pytest

from dune_gdt_b200 import descriptors as D
from helpers import TOL, rel_err

pytestmark = pytest.mark.gpu

FV = D.SPACE_FV
SEED = 20251017


def fluxes(d):
    a = [1.0, -0.5, 0.75][:d]
    return [
        ("linear-upwind", D.flux(D.FLUX_LINEAR, D.NUMFLUX_UPWIND, a)),
        ("burgers-upwind", D.flux(D.FLUX_BURGERS, D.NUMFLUX_UPWIND, [])),
        ("linear-lf", D.flux(D.FLUX_LINEAR, D.NUMFLUX_LAX_FRIEDRICHS, a)),
        ("burgers-lf", D.flux(D.FLUX_BURGERS, D.NUMFLUX_LAX_FRIEDRICHS, [])),
    ]


def make_op(gdt, ctx, gdesc, fl, boundary=()):
    space = gdt.Space(gdt.Grid(ctx, gdesc), FV, 0)
    num = gdt.NumericalUpwindFlux(fl.kind, list(fl.p))
    num.desc.numflux = fl.numflux
    L = gdt.make_advection_fv_operator(num, space)
    for t in boundary:
        L.append(t)
    return L


def bnd_cases():
    E, NF = D.FVBND_EXTRAPOLATION, D.FVBND_NUMERICAL_FLUX
    out = []
    for n, per in (([17], 0), ([16], 0), ([12, 10], 0), ([9, 8], 0), ([12, 10], 1), ([10, 6], 2), ([6, 5, 4], 0), ([8, 5, 4], 2),
                   ([7, 4, 5], 5)):
        d = len(n)
        every = (1 << (2 * d)) - 1
        sets = [
            [D.fv_boundary(E, every, 1.0, 0.0)],                                          # absorbing everywhere
            [D.fv_boundary(E, 0b01, 0.0, 0.7), D.fv_boundary(NF, every & ~0b01, 1.0, 0.0)],  # inflow value left, outflow else
            [D.fv_boundary(E, every & 0b101010, -1.0, 0.2), D.fv_boundary(NF, every, 0.5, -0.3),
             D.fv_boundary(NF, 0b10, 0.0, 1.0)],                                          # mixed, two treatments on x+
        ]
        for name, fl in fluxes(d):
            for k, b in enumerate(sets):
                out.append((f"{name}-{'x'.join(map(str, n))}-p{per}-b{k}", n, per, fl, b))
    return out


@pytest.mark.parametrize("name,n,periodic,fl,boundary", bnd_cases(), ids=lambda v: v if isinstance(v, str) else None)
def test_fv_boundary_treatments_parity(gdt, ctx, oracle, name, n, periodic, fl, boundary):
    gdesc = D.grid_desc(0.0, 1.0, n, periodic)
    L = make_op(gdt, ctx, gdesc, fl, boundary)
    u = np.random.default_rng(SEED).uniform(-1.0, 1.0, int(np.prod(n)))
    out = L.apply(u)
    ref = oracle.fv_apply_bnd(gdesc, fl, boundary, u)
    assert rel_err(out, ref) <= TOL
    # periodic sides carry no boundary intersections: treatments there must not act
    if periodic:
        again = L.apply(u)
        assert np.array_equal(out, again)  # run-to-run bit-identical


def test_fv_boundary_error_conventions(gdt, ctx):
    L = make_op(gdt, ctx, D.grid_desc(0.0, 1.0, [8, 8]), D.flux(D.FLUX_LINEAR, D.NUMFLUX_UPWIND, [1.0, 0.0]))
    L.append(D.fv_boundary(D.FVBND_EXTRAPOLATION, 0b0011, 1.0, 0.0))
    with pytest.raises(gdt.capi.NotImplementedGdt):
        L.append(D.fv_boundary(D.FVBND_EXTRAPOLATION, 0b0001, 0.0, 1.0))
    with pytest.raises(gdt.capi.WrongInputGiven):
        L.append(D.fv_boundary(D.FVBND_EXTRAPOLATION, 0b110000, 1.0, 0.0))  # a 2D grid has no z sides
    with pytest.raises(gdt.capi.WrongInputGiven):
        L.append(D.fv_boundary(7, 0b0100, 1.0, 0.0))


@pytest.mark.parametrize("n,per", [([16], 1), ([64], 1), ([33, 20], 3), ([12, 9, 7], 7), ([40, 30], 0)])
def test_estimate_dt_parity(gdt, ctx, oracle, n, per):
    gdesc = D.grid_desc(0.0, 1.0, n, per)
    rng = np.random.default_rng(SEED)
    for name, fl in fluxes(len(n))[:2]:
        L = make_op(gdt, ctx, gdesc, fl)
        for u in (rng.uniform(-1.0, 2.0, int(np.prod(n))), -rng.uniform(1.0, 2.0, int(np.prod(n))), np.full(int(np.prod(n)), 0.5)):
            assert L.estimate_dt(u) == pytest.approx(oracle.fv_estimate_dt(gdesc, fl, u), rel=1e-14)
        u = rng.uniform(0.0, 1.0, int(np.prod(n)))
        assert L.estimate_dt(u, [-3.0, 2.5]) == pytest.approx(oracle.fv_estimate_dt(gdesc, fl, u, [-3.0, 2.5]), rel=1e-14)


def test_estimate_dt_reference_tables(gdt, ctx, oracle):
    """quantity.CFL of linear_transport__1d__explicit__fv.mini:14 and burgers__1d__explicit__fv.mini:15 (dt_factor 0.99)"""
    for N in (16, 32, 64):
        gdesc = D.grid_desc([0.0], [1.0], [N], periodic=1)
        u0 = oracle.fv_interpolate(gdesc, D.fn_builtin(D.BUILTIN_INDICATOR, 0, 0.25, 0.5))
        L = make_op(gdt, ctx, gdesc, D.flux(D.FLUX_LINEAR, D.NUMFLUX_UPWIND, [1.0]))
        assert (1.0 / N) / L.estimate_dt(u0) == pytest.approx(2.0, rel=1e-14)
    for N, cfl in ((16, 2.31e-01), (32, 4.80e-01)):
        gdesc = D.grid_desc([0.0], [1.0], [N], periodic=1)
        u0 = oracle.fv_interpolate(gdesc, D.fn_builtin(D.BUILTIN_GAUSSIAN, 3, 0.33, 0.075))
        L = make_op(gdt, ctx, gdesc, D.flux(D.FLUX_BURGERS, D.NUMFLUX_UPWIND, []))
        assert float(f"{0.99 * 0.0096815612792968738 / L.estimate_dt(u0):.2e}") == cfl


HEUN3 = ([[0.0, 0.0, 0.0], [1.0 / 3.0, 0.0, 0.0], [0.0, 2.0 / 3.0, 0.0]], [0.25, 0.0, 0.75], [0.0, 1.0 / 3.0, 2.0 / 3.0])
# a 6-stage array (more k-terms than one fused axpy pass takes): two Heun steps of half length chained
SIX = (np.tril(np.arange(36, dtype=float).reshape(6, 6) % 5 + 1.0, -1) / 20.0, [0.1, 0.2, 0.1, 0.25, 0.15, 0.2], [0.0, 0.05, 0.2, 0.35, 0.6, 0.85])


@pytest.mark.parametrize("method", [D.RK_EULER, D.RK_SSP2, D.RK_SSP3, D.RK_CLASSIC4, "heun3", "six"])
@pytest.mark.parametrize("n,per", [([64], 1), ([20, 12], 3), ([21, 8], 0), ([6, 5, 4], 7)])
def test_rk_step_parity(gdt, ctx, oracle, method, n, per):
    gdesc = D.grid_desc(0.0, 1.0, n, per)
    bnd = [] if per else [D.fv_boundary(D.FVBND_EXTRAPOLATION, 0b1111, 1.0, 0.0)]
    butcher = {"heun3": HEUN3, "six": SIX}.get(method) or D.BUTCHER[method]
    for name, fl in fluxes(len(n))[1:3]:
        L = make_op(gdt, ctx, gdesc, fl, bnd)
        u0 = np.random.default_rng(SEED).uniform(0.1, 1.0, int(np.prod(n)))
        if isinstance(method, str):
            ts = gdt.ExplicitRungeKuttaTimeStepper(L, u0, r=-1.0, t_0=0.5, method=D.RK_OTHER, A=butcher[0], b=butcher[1], c=butcher[2])
        else:
            ts = gdt.ExplicitRungeKuttaTimeStepper(L, u0, r=-1.0, t_0=0.5, method=method)
        dt = 0.2 * oracle.fv_estimate_dt(gdesc, fl, u0)
        assert ts.step(dt) == dt
        assert ts.step(dt, 0.5 * dt) == dt  # step returns the dt it was given, advances by min(dt, max_dt)
        ref, t = oracle.rk_step(gdesc, fl, butcher, u0, 0.5, dt, r=-1.0, boundary=bnd)
        ref, t = oracle.rk_step(gdesc, fl, butcher, ref, t, dt, 0.5 * dt, r=-1.0, boundary=bnd)
        assert ts.current_time() == t
        assert rel_err(ts.current_solution(), ref) <= TOL


@pytest.mark.parametrize("method,t_end,dt_factor", [
    (D.RK_EULER, 0.3, 1.0), (D.RK_EULER, 0.31, 0.93), (D.RK_SSP2, 0.25, 1.0), (D.RK_SSP3, 0.2, 0.8), (D.RK_CLASSIC4, 0.2, 1.0),
    (D.RK_EULER, 0.05, 1.0),
])
def test_rk_solve_parity(gdt, ctx, oracle, method, t_end, dt_factor):
    """the whole time loop on the device (graph replay of the full steps + the shortened last step) against the
    oracle's restatement of TimeStepperInterface::solve: same number of steps, same end time, same values"""
    n = [48, 20]
    gdesc = D.grid_desc(0.0, 1.0, n, 3)
    fl = D.flux(D.FLUX_BURGERS, D.NUMFLUX_UPWIND, [])
    u0 = oracle.fv_interpolate(gdesc, D.fn_builtin(D.BUILTIN_GAUSSIAN, 3, 0.33, 0.075)) + 0.1
    L = make_op(gdt, ctx, gdesc, fl)
    dt = dt_factor * L.estimate_dt(u0)
    ts = gdt.ExplicitRungeKuttaTimeStepper(L, u0, r=-1.0, method=method)
    next_dt = ts.solve(t_end, dt)
    ref, steps, t = oracle.rk_solve(gdesc, fl, D.BUTCHER[method], u0, t_end, dt, r=-1.0)
    assert next_dt == dt
    assert ts.num_steps == steps and steps >= 2
    assert ts.current_time() == t
    assert rel_err(ts.current_solution(), ref) <= TOL
    assert abs(ts.current_solution().sum() - u0.sum()) <= 1e-12 * u0.sum()
    # a second solve continues from the current time
    ts.solve(t_end + 3.5 * dt, dt)
    ref2, steps2, t2 = oracle.rk_solve(gdesc, fl, D.BUTCHER[method], ref, t_end + 3.5 * dt, dt, t0=t, r=-1.0)
    assert ts.num_steps == steps2 == 4 and ts.current_time() == t2
    assert rel_err(ts.current_solution(), ref2) <= TOL


def test_rk_linear_transport_table_through_the_stepper(gdt, ctx, oracle):
    """linear_transport__1d__explicit__fv.mini:8-14 driven by the time stepper: dt = h = 2 * estimate is an exact
    shift; 16 / 32 / 64 steps bring the indicator back, mass error 0"""
    for N in (16, 32, 64):
        gdesc = D.grid_desc([0.0], [1.0], [N], periodic=1)
        u0 = oracle.fv_interpolate(gdesc, D.fn_builtin(D.BUILTIN_INDICATOR, 0, 0.25, 0.5))
        L = make_op(gdt, ctx, gdesc, D.flux(D.FLUX_LINEAR, D.NUMFLUX_UPWIND, [1.0]))
        ts = gdt.ExplicitRungeKuttaTimeStepper(L, u0, r=-1.0)
        ts.solve(1.0, 2.0 * L.estimate_dt(u0))
        assert ts.num_steps == N and ts.current_time() == 1.0
        np.testing.assert_allclose(ts.current_solution(), u0, atol=1e-14)
        assert abs(ts.current_solution().sum() - u0.sum()) / u0.sum() <= 1e-15


def test_rk_error_conventions(gdt, ctx):
    L = make_op(gdt, ctx, D.grid_desc(0.0, 1.0, [8, 8], 3), D.flux(D.FLUX_LINEAR, D.NUMFLUX_UPWIND, [1.0, 0.0]))
    u0 = np.zeros(64)
    with pytest.raises(gdt.capi.WrongInputGiven):  # explicit-rungekutta.hh:216-222
        gdt.ExplicitRungeKuttaTimeStepper(L, u0, method=D.RK_OTHER, A=[[0.0, 0.5], [0.5, 0.0]], b=[0.5, 0.5], c=[0.0, 1.0])
    with pytest.raises(gdt.capi.NotImplementedGdt):  # :40-58: explicit_rungekutta_other without arrays
        gdt.ExplicitRungeKuttaTimeStepper(L, u0, method=D.RK_OTHER)
    with pytest.raises(gdt.capi.ShapesDoNotMatch):
        gdt.ExplicitRungeKuttaTimeStepper(L, np.zeros(63))
    ts = gdt.ExplicitRungeKuttaTimeStepper(L, u0)
    with pytest.raises(gdt.capi.WrongInputGiven):
        ts.solve(1.0, 0.0)


def test_rk_full_size_properties(gdt, ctx):
    """C4-sized (4096^2 periodic) SSP3 steps on the device: conservation and the discrete maximum principle of an
    SSP method under the CFL bound"""
    import ctypes as C

    import torch

    n = [4096, 4096]
    gdesc = D.grid_desc(0.0, 1.0, n, 3)
    L = make_op(gdt, ctx, gdesc, D.flux(D.FLUX_LINEAR, D.NUMFLUX_UPWIND, [1.0, 0.5]))
    g = torch.Generator(device="cuda").manual_seed(SEED)
    u = torch.rand(4096 * 4096, dtype=torch.float64, device="cuda", generator=g)
    mass0, lo, hi = u.sum().item(), u.min().item(), u.max().item()
    torch.cuda.synchronize()
    lib = gdt.capi.lib()
    dt = C.c_double()
    gdt.capi.check(lib.gdtb_fv_estimate_dt(L._h, C.c_void_p(u.data_ptr()), None, C.byref(dt)))
    assert dt.value == pytest.approx(1.0 / (4 * 4096 * 1.0), rel=1e-12)
    ts = C.c_void_p()
    gdt.capi.check(lib.gdtb_rk_create(L._h, D.RK_SSP3, 0, None, None, None, -1.0, 0.0, C.byref(ts)))
    steps = C.c_int64()
    gdt.capi.check(lib.gdtb_rk_solve(ts, C.c_void_p(u.data_ptr()), 12.5 * dt.value, dt.value, C.byref(steps), None))
    assert steps.value == 13
    torch.cuda.synchronize()
    assert abs(u.sum().item() - mass0) <= 1e-12 * mass0
    assert u.min().item() >= lo - 1e-14 and u.max().item() <= hi + 1e-14
    gdt.capi.check(lib.gdtb_rk_destroy(ts))
