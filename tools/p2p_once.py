"""One slab (world size 1, periodic: the neighbour is the rank itself) of the 4096^2 FV problem through the
peer-memory paths, for an ncu launch list: fused Euler steps (k_fv_march<..., P2P>) and SSP3 steps on a slab
(k_rk_axpy, k_p2p_send_layers, k_fv_march<..., P2P>)."""
import os
import sys

sys.path.insert(0, os.getcwd())
import numpy as np
import torch

import dune_gdt_b200 as gdt
from dune_gdt_b200 import descriptors as D
from dune_gdt_b200 import parallel

ctx = gdt.Context(0)
n = 4096
grid = gdt.make_cube_grid(ctx, 0.0, 1.0, [n, n], periodic=3)
space = gdt.make_finite_volume_space(grid)
u0 = np.random.default_rng(20251017).random(n * n)
loop = parallel.PeerMemoryFvTimeLoop(gdt.NumericalUpwindFlux(D.FLUX_LINEAR, [1.0, 0.5]), space, 0, 1)
loop.set_initial_values(u0)
loop.euler_steps(0.25 / n, 6)
loop.check()
loop.close()
ts = parallel.PeerMemoryRkTimeStepper(gdt.NumericalUpwindFlux(D.FLUX_LINEAR, [1.0, 0.5]), space, 0, 1, D.RK_SSP3)
ts.set_initial_values(u0)
for _ in range(3):
    ts.step(0.25 / n)
ts.check()
ts.close()
print("ok")
