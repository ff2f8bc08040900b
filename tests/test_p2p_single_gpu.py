"""Peer-memory FV time loop on ONE device (world size 1, periodic: the neighbour on both sides is the rank itself): the
in-kernel ghost hand-over, the step counters and the ping-pong buffers against the oracle's Euler loop."""
import numpy as np
import pytest

from dune_gdt_b200 import descriptors as D
from helpers import TOL, rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,periodic,fk,params,steps", [
    ([64, 48], 3, D.FLUX_LINEAR, [1.0, 0.5], 9), ([40, 33], 3, D.FLUX_BURGERS, [], 6), ([12, 10, 9], 7, D.FLUX_LINEAR, [1.0, -0.5, 0.25], 5),
    ([31, 20], 0, D.FLUX_BURGERS, [], 4), ([4096, 64], 3, D.FLUX_LINEAR, [1.0, 0.5], 3),
])
def test_peer_memory_loop_world_one(gdt, ctx, oracle, n, periodic, fk, params, steps):
    from dune_gdt_b200 import parallel

    grid = gdt.make_cube_grid(ctx, 0.0, 1.0, n, periodic=periodic)
    space = gdt.make_finite_volume_space(grid)
    loop = parallel.PeerMemoryFvTimeLoop(gdt.NumericalUpwindFlux(fk, params), space, 0, 1)
    u = np.random.default_rng(20251017).random(int(np.prod(n)))
    loop.set_initial_values(u)
    dt = 0.2 / max(n)
    loop.euler_steps(dt, steps)
    loop.check()
    ref = oracle.fv_euler(D.grid_desc(0.0, 1.0, n, periodic=periodic), D.flux(fk, D.NUMFLUX_UPWIND, params), u, dt, steps)
    got = loop.owned_view(loop.current()).cpu().numpy()
    assert rel_err(got, ref) <= TOL
    loop.close()


@pytest.mark.parametrize("method", [D.RK_SSP2, D.RK_SSP3, D.RK_CLASSIC4])
@pytest.mark.parametrize("n,periodic,fk,params", [([48, 40], 3, D.FLUX_BURGERS, []), ([12, 10, 9], 7, D.FLUX_LINEAR, [1.0, -0.5, 0.25]),
                                                  ([31, 20], 1, D.FLUX_LINEAR, [1.0, 0.5])])
def test_peer_memory_runge_kutta_world_one(gdt, ctx, oracle, method, n, periodic, fk, params):
    """Runge-Kutta on a slab (one slab: the periodic neighbour is the rank itself; [31, 20] is not periodic along the
    slab direction: no neighbours at all): stage-vector hand-over, counters and alternating stage buffers against the
    oracle's restatement of ExplicitRungeKuttaTimeStepper::solve"""
    from dune_gdt_b200 import parallel

    grid = gdt.make_cube_grid(ctx, 0.0, 1.0, n, periodic=periodic)
    space = gdt.make_finite_volume_space(grid)
    ts = parallel.PeerMemoryRkTimeStepper(gdt.NumericalUpwindFlux(fk, params), space, 0, 1, method)
    u = np.random.default_rng(20251017).uniform(0.1, 1.0, int(np.prod(n)))
    ts.set_initial_values(u)
    gd = D.grid_desc(0.0, 1.0, n, periodic=periodic)
    fl = D.flux(fk, D.NUMFLUX_UPWIND, params)
    dt = 0.3 * oracle.fv_estimate_dt(gd, fl, u)
    t_end = 6.5 * dt
    ts.solve(t_end, dt)
    ts.check()
    ref, steps, t = oracle.rk_solve(gd, fl, D.BUTCHER[method], u, t_end, dt, r=-1.0)
    assert ts.num_steps == steps == 7 and ts.current_time() == t
    assert rel_err(ts.owned_view().cpu().numpy(), ref) <= TOL
    ts.close()


def test_runge_kutta_on_a_slab_needs_two_stages(gdt, ctx):
    import ctypes as C

    grid = gdt.make_cube_grid(ctx, 0.0, 1.0, [8, 8], periodic=3)
    space = gdt.make_finite_volume_space(grid)
    L = gdt.make_advection_fv_operator(gdt.NumericalUpwindFlux(D.FLUX_LINEAR, [1.0, 0.0]), space)
    lib = gdt.capi.lib()
    gdt.capi.check(lib.gdtb_fvop_set_slab(L._h, 0, 8))
    ts = C.c_void_p()
    assert lib.gdtb_rk_create(L._h, D.RK_EULER, 0, None, None, None, -1.0, 0.0, C.byref(ts)) == 7  # NotImplemented
