import os, sys, ctypes as C
sys.path.insert(0, os.getcwd())
import torch
import dune_gdt_b200 as gdt
from dune_gdt_b200 import descriptors as D
ctx = gdt.Context(0)
lib = gdt.capi.lib()
def kt():
    ms, n = C.c_double(), C.c_int64()
    lib.gdtb_ctx_kernel_time(ctx._h, b"fv_apply", C.byref(ms), C.byref(n)); return ms.value, n.value
n = 4096
grid = gdt.make_cube_grid(ctx, 0.0, 1.0, [n, n], periodic=3)
space = gdt.make_finite_volume_space(grid)
for rows in [0, 8, 16, 32, 64, 128]:
    os.environ["GDTB_FV_ROWS"] = str(rows)
    for name, flux in (("linear", gdt.NumericalUpwindFlux(D.FLUX_LINEAR, [1.0, 0.5])), ("burgers", gdt.NumericalUpwindFlux(D.FLUX_BURGERS))):
        L = gdt.make_advection_fv_operator(flux, space)
        u = torch.rand(n * n, dtype=torch.float64, device="cuda"); v = torch.empty_like(u)
        for _ in range(5): L.apply_device(u.data_ptr(), v.data_ptr())
        lib.gdtb_ctx_enable_timing(ctx._h, 1); kt()
        for _ in range(50): L.apply_device(u.data_ptr(), v.data_ptr())
        ms, cnt = kt(); lib.gdtb_ctx_enable_timing(ctx._h, 0)
        per = ms / cnt
        print(f"rows={rows:4d} {name:8s} {per*1e3:8.1f} us  {16.0*n*n/per/1e6:8.1f} GB/s")
