"""Times the non-headline configurations of SURVEY.md section 8 (C1, C3, C5) through the C ABI; device-resident, CUDA events.

  python tools/bench_configs.py [c1] [c3] [c5] [--n N]
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


def main():
    which = [a for a in sys.argv[1:] if not a.startswith("--")] or ["c1", "c3", "c5"]
    n_override = int(sys.argv[sys.argv.index("--n") + 1]) if "--n" in sys.argv else None
    lap = D.form(D.integrand(D.INT_LAPLACE, diffusion=1.0))
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
    if "c5" in which:
        n = n_override or 128
        run(f"C5 3D Q2 {n}^3", D.grid_desc(-1.0, 1.0, [n, n, n]), D.SPACE_CG, 2, D.STENCIL_ELEMENT, element=[lap], reps=2)


if __name__ == "__main__":
    main()
