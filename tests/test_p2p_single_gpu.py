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
