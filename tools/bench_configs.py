"""Times the non-headline configurations of SURVEY.md section 8 (C1, C3, C5) through the C ABI; device-resident, CUDA events.

  python tools/bench_configs.py [c1] [c3] [c5] [c2-qp] [c2-qpt] [c2-builtin] [c5-qp] [c5-qpt] [c5-builtin] [--n N]

The *-qp / *-qpt / *-builtin variants assemble C2 / C5 with a coefficient that varies inside every cell: one random
value (qp) or one full 3 x 3 tensor (qpt) per quadrature point in a device array, or an analytic kappa(x) sampled on the
device (builtin).  GDTB_NO_QP_GATHER=1 times the quadrature-faithful coloured-scatter kernels on the same input.
"""
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import dune_gdt_b200 as gdt
from dune_gdt_b200 import descriptors as D

lib = gdt.capi.lib()
check = gdt.capi.check
ctx = gdt.Context(0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx.set_stream(stream.cuda_stream)


def timed(fn, reps):
    reps = int(os.environ.get("GDTB_REPS", reps))
    fn()
    ctx.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    ctx.synchronize()
    return s.elapsed_time(e) / reps


def run(name, gdesc, kind, order, stencil, element=(), coupling=(), boundary=(), reps=3):
    grid = gdt.Grid(ctx, gdesc)
    space = gdt.Space(grid, kind, order)
    t0 = time.perf_counter()
    pat = gdt.SparsityPattern(space, space, stencil, D.PATTERN_AUTO)
    ctx.synchronize()
    t_pat = time.perf_counter() - t0
    op = gdt.MatrixOperator(space, space, pat)
    for f in element:
        check(lib.gdtb_matop_append_element(op._h, C.byref(f)))
    for f in coupling:
        check(lib.gdtb_matop_append_coupling(op._h, C.byref(f), D.FILTER_INNER_ONCE))
    for f in boundary:
        check(lib.gdtb_matop_append_boundary(op._h, C.byref(f), D.FILTER_ALL_BOUNDARY))
    ms = timed(lambda: check(lib.gdtb_assemble_async(op._h, None, D.ASSEMBLE_OVERWRITE)), reps)
    ne = int(np.prod([gdesc.n[k] for k in range(gdesc.dim)]))
    nnz, rows = pat.nnz, pat.rows
    alg = 8.0 * nnz + 8.0 * rows
    out = {"config": name, "plan": op.plan, "elements": ne, "rows": rows, "nnz": nnz, "pattern_build_s": t_pat,
           "ms_per_assembly": ms, "elements_per_s": ne / (ms * 1e-3), "algorithmic_GBps": alg / (ms * 1e-3) / 1e9}
    print(json.dumps(out), flush=True)


def qp_function(ne, nq, d, tensor, order):
    """device-resident coefficient samples (borrowed by the library: data_on_device = 1)"""
    g = torch.Generator(device="cuda").manual_seed(20251017)
    if tensor:
        a = 0.2 * torch.rand(ne, nq, d, d, dtype=torch.float64, device="cuda", generator=g)
        a += torch.eye(d, dtype=torch.float64, device="cuda")
    else:
        a = 0.5 + torch.rand(ne, nq, dtype=torch.float64, device="cuda", generator=g)
    f = D.Function()
    f.kind = D.FN_QP_TENSOR if tensor else D.FN_QP_SCALAR
    f.order = order
    f.qp_per_element = nq
    f.data_on_device = 1
    f.data = C.cast(a.data_ptr(), C.POINTER(C.c_double))
    f._keep = a
    return f


def variable_kappa(which, n, order):
    """(name suffix, Laplace form) for the *-qp / *-qpt / *-builtin variants; kappa has the declared order --kappa-order
    (default 0), so the rule has order kappa_order + 2 p: 2 points per direction for Q1 and 3 for Q2 by default"""
    out = []
    prefix = "c2" if order == 1 else "c5"
    kord = int(sys.argv[sys.argv.index("--kappa-order") + 1]) if "--kappa-order" in sys.argv else 0
    m = (kord + 2 * order) // 2 + 1
    if f"{prefix}-qp" in which:
        out.append(("scalar kappa per quadrature point", D.form(D.integrand(D.INT_LAPLACE, diffusion=qp_function(n**3, m**3, 3, False, kord)))))
    if f"{prefix}-qpt" in which:
        out.append(("3x3 kappa per quadrature point", D.form(D.integrand(D.INT_LAPLACE, diffusion=qp_function(n**3, m**3, 3, True, kord)))))
    if f"{prefix}-builtin" in which:
        out.append(("analytic kappa(x) = 1 + 0.5 |x|^2 sampled on the device",
                    D.form(D.integrand(D.INT_LAPLACE, diffusion=D.fn_builtin(D.BUILTIN_QUADRATIC, kord, 1.0, 0.5)))))
    return [(f"{label} ({m}^3 points per element)", form) for label, form in out]


def main():
    which = [a for a in sys.argv[1:] if not a.startswith("--") and not a.isdigit()] or ["c1", "c3", "c5"]
    n_override = int(sys.argv[sys.argv.index("--n") + 1]) if "--n" in sys.argv else None
    lap = D.form(D.integrand(D.INT_LAPLACE, diffusion=1.0))
    for order, n_default in ((1, 256), (2, 128)):
        n = n_override or n_default
        for label, form in variable_kappa(which, n, order):
            run(f"C{2 if order == 1 else 5} 3D Q{order} {n}^3, {label}", D.grid_desc(-1.0, 1.0, [n, n, n]), D.SPACE_CG, order,
                D.STENCIL_ELEMENT, element=[form], reps=3)
            torch.cuda.empty_cache()
    if "c2-elem" in which:
        n = n_override or 256
        g_ = torch.Generator(device="cuda").manual_seed(7)
        kap = 0.5 + torch.rand(n**3, dtype=torch.float64, device="cuda", generator=g_)
        f = D.Function()
        f.kind = D.FN_ELEM_SCALAR
        f.data_on_device = 1
        f.data = C.cast(kap.data_ptr(), C.POINTER(C.c_double))
        f._keep = kap
        run(f"C2 3D Q1 {n}^3, one kappa per element", D.grid_desc(-1.0, 1.0, [n, n, n]), D.SPACE_CG, 1, D.STENCIL_ELEMENT,
            element=[D.form(D.integrand(D.INT_LAPLACE, diffusion=f))], reps=10)
    if "c5-elem" in which:
        n = n_override or 128
        g_ = torch.Generator(device="cuda").manual_seed(7)
        kap = 0.5 + torch.rand(n**3, dtype=torch.float64, device="cuda", generator=g_)
        f = D.Function()
        f.kind = D.FN_ELEM_SCALAR
        f.data_on_device = 1
        f.data = C.cast(kap.data_ptr(), C.POINTER(C.c_double))
        f._keep = kap
        run(f"C5 3D Q2 {n}^3, one kappa per element", D.grid_desc(-1.0, 1.0, [n, n, n]), D.SPACE_CG, 2, D.STENCIL_ELEMENT,
            element=[D.form(D.integrand(D.INT_LAPLACE, diffusion=f))], reps=5)
    if "c2" in which:
        n = n_override or 256
        run(f"C2 3D Q1 {n}^3", D.grid_desc(-1.0, 1.0, [n, n, n]), D.SPACE_CG, 1, D.STENCIL_ELEMENT, element=[lap], reps=10)
    if "c1" in which:
        n = n_override or 128
        run(f"C1 2D Q1 {n}^2", D.grid_desc(-1.0, 1.0, [n, n]), D.SPACE_CG, 1, D.STENCIL_ELEMENT, element=[lap], reps=20)
    if "c3" in which:
        n = n_override or 2048
        inner = D.form([D.integrand(D.INT_IPDG_INNER_COUPLING, prefactor=1.0, diffusion=1.0, weight=1.0),
                        D.integrand(D.INT_IPDG_INNER_PENALTY, prefactor=8.0, weight=1.0, hI_kind=D.HI_VOLUME)])
        bnd = D.form([D.integrand(D.INT_IPDG_DIRICHLET_COUPLING, prefactor=1.0, diffusion=1.0),
                      D.integrand(D.INT_IPDG_BOUNDARY_PENALTY, prefactor=14.0, weight=1.0, hI_kind=D.HI_VOLUME)])
        run(f"C3 2D SWIPDG DG-Q1 {n}^2", D.grid_desc(-1.0, 1.0, [n, n]), D.SPACE_DG, 1,
            D.STENCIL_ELEMENT_AND_INTERSECTION, element=[lap], coupling=[inner], boundary=[bnd], reps=5)
    if "c3-elem" in which:
        n = n_override or 2048
        g_ = torch.Generator(device="cuda").manual_seed(9)
        kap = 0.5 + torch.rand(n * n, dtype=torch.float64, device="cuda", generator=g_)

        def fe():
            f = D.Function()
            f.kind = D.FN_ELEM_SCALAR
            f.data_on_device = 1
            f.data = C.cast(kap.data_ptr(), C.POINTER(C.c_double))
            f._keep = kap
            return f

        lap_e = D.form(D.integrand(D.INT_LAPLACE, diffusion=fe()))
        inner = D.form([D.integrand(D.INT_IPDG_INNER_COUPLING, prefactor=1.0, diffusion=fe(), weight=fe()),
                        D.integrand(D.INT_IPDG_INNER_PENALTY, prefactor=8.0, weight=fe(), hI_kind=D.HI_VOLUME)])
        bnd = D.form([D.integrand(D.INT_IPDG_DIRICHLET_COUPLING, prefactor=1.0, diffusion=fe()),
                      D.integrand(D.INT_IPDG_BOUNDARY_PENALTY, prefactor=14.0, weight=fe(), hI_kind=D.HI_VOLUME)])
        run(f"C3 2D SWIPDG DG-Q1 {n}^2, kappa = omega = one value per element", D.grid_desc(-1.0, 1.0, [n, n]), D.SPACE_DG, 1,
            D.STENCIL_ELEMENT_AND_INTERSECTION, element=[lap_e], coupling=[inner], boundary=[bnd], reps=5)
    if "c5" in which:
        n = n_override or 128
        run(f"C5 3D Q2 {n}^3", D.grid_desc(-1.0, 1.0, [n, n, n]), D.SPACE_CG, 2, D.STENCIL_ELEMENT, element=[lap], reps=2)


if __name__ == "__main__":
    main()
