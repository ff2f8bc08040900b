"""FV apply A/B: per-launch CUDA events (the library's timers) against one event pair around 100 back-to-back applies,
and a same-size device copy / fill for scale.  python tools/fv_ab.py"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import dune_gdt_b200 as gdt
from dune_gdt_b200 import descriptors as D

ctx = gdt.Context(0)
lib = gdt.capi.lib()
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx.set_stream(stream.cuda_stream)


def kt():
    ms, n = C.c_double(), C.c_int64()
    lib.gdtb_ctx_kernel_time(ctx._h, b"fv_apply", C.byref(ms), C.byref(n))
    return ms.value, n.value


def pair(fn, reps=100):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps * 1e3


n = 4096
grid = gdt.make_cube_grid(ctx, 0.0, 1.0, [n, n], periodic=3)
space = gdt.make_finite_volume_space(grid)
u = torch.rand(n * n, dtype=torch.float64, device="cuda")
v = torch.empty_like(u)
out = {"lib": os.environ.get("GDTB_LIB", "default"), "rows_env": os.environ.get("GDTB_FV_ROWS", ""), "tma": os.environ.get("GDTB_FV_TMA", "0")}
out["copy_us"] = pair(lambda: v.copy_(u))
out["fill_us"] = pair(lambda: v.fill_(1.0))
for name, flux in (("linear", gdt.NumericalUpwindFlux(D.FLUX_LINEAR, [1.0, 0.5])), ("burgers", gdt.NumericalUpwindFlux(D.FLUX_BURGERS))):
    L = gdt.make_advection_fv_operator(flux, space)
    out[name + "_pair_us"] = pair(lambda: L.apply_device(u.data_ptr(), v.data_ptr()))
    lib.gdtb_ctx_enable_timing(ctx._h, 1)
    kt()
    for _ in range(50):
        L.apply_device(u.data_ptr(), v.data_ptr())
    ms, cnt = kt()
    lib.gdtb_ctx_enable_timing(ctx._h, 0)
    out[name + "_events_us"] = ms / cnt * 1e3
    # ping-pong u -> v -> u (what a time loop does)
    out[name + "_pingpong_us"] = pair(lambda: (L.apply_device(u.data_ptr(), v.data_ptr()), L.apply_device(v.data_ptr(), u.data_ptr())), 50) / 2
print(json.dumps(out), flush=True)
