import os, sys
sys.path.insert(0, os.getcwd())
import torch
import dune_gdt_b200 as gdt
from dune_gdt_b200 import descriptors as D
ctx = gdt.Context(0)
n = 4096
grid = gdt.make_cube_grid(ctx, 0.0, 1.0, [n, n], periodic=3)
space = gdt.make_finite_volume_space(grid)
L = gdt.make_advection_fv_operator(gdt.NumericalUpwindFlux(D.FLUX_LINEAR, [1.0, 0.5]), space)
u = torch.rand(n * n, dtype=torch.float64, device="cuda"); v = torch.empty_like(u)
for _ in range(6): L.apply_device(u.data_ptr(), v.data_ptr())
torch.cuda.synchronize()
