#!/usr/bin/env python
"""bench.py -- elements assembled / s for the 3D Q1 Laplace + RHS assembly on a 256^3 YaspGrid cube (BASELINE.json
configs[1]), plus the roofline of the dominant kernel and the CPU baseline.

  python bench.py --gpus N --steps K --warmup W            # product arm (one rank per GPU under torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port) on host cores

A step = one pass of the hot path: matrix + right-hand side of the whole (rank-local) grid in one fused walk.
Weak scaling: every rank owns a 256 x 256 x 256 slab of a 256 x 256 x (256 N) grid (owner-computes-rows, the ghost
element layer is recomputed, no data-path collective -- the reference assembles its overlap redundantly as well).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NX = 256  # elements per direction and rank
METRIC = "elements assembled/sec (3D Q1 Laplace + RHS, 256^3 per GPU, FP64)"
UNIT = "elements/s"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md recipe)"""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.samples = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.time(), line.strip()))

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self, t0, t1):
        rows = [s for t, s in self.samples if t0 <= t <= t1 + 0.06] or [s for _, s in self.samples[-3:]]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            parts = [p.strip() for p in r.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def forms():
    from dune_gdt_b200 import descriptors as D

    lap = D.form(D.integrand(D.INT_LAPLACE, diffusion=1.0))
    # 3D analogue of examples/stationary-heat-equation.cc:68-70, declared order 3
    src = D.fn_builtin(D.BUILTIN_COS_PRODUCT, 3, 0.75 * np.pi**2, 0.5 * np.pi)
    rhs = D.form(D.integrand(D.INT_PRODUCT, diffusion=1.0, weight=src))
    return lap, rhs


def kernel_time(gdt, ctx, family):
    ms, n = C.c_double(), C.c_int64()
    gdt.capi.check(gdt.capi.lib().gdtb_ctx_kernel_time(ctx._h, family.encode(), C.byref(ms), C.byref(n)))
    return ms.value, n.value


# ------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's assembly loops, threaded like walk(use_tbb = true)
# ------------------------------------------------------------------------------------------------------------------
def cpu_assemble_rate(n_side, threads, repeats=1):
    import oracle
    from dune_gdt_b200 import descriptors as D

    h = 2.0 / NX
    g = D.grid_desc(-1.0, -1.0 + n_side * h, [n_side] * 3)
    rp, ci = oracle.pattern(g, (D.SPACE_CG, 1))  # setup, not timed (neither is the GPU pattern)
    lap, rhs = forms()
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        oracle.assemble(g, D.SPACE_CG, 1, rp, ci, [lap], rhs_forms=[rhs], num_threads=threads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return n_side**3 / best, best


def cpu_sample_side(threads, budget_s):
    """pick the sample grid so that one assembly costs about budget_s of wall time"""
    rate, _ = cpu_assemble_rate(32, threads)
    side = int(round((rate * budget_s) ** (1.0 / 3.0)))
    return int(min(max(side, 32), 192))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle

    oracle.build()
    threads = os.cpu_count() or 1
    total_steps = args.steps + args.warmup
    side = cpu_sample_side(threads, budget_s=min(20.0, 100.0 / max(total_steps, 1)))
    for _ in range(args.warmup):
        cpu_assemble_rate(side, threads)
    times = []
    for _ in range(args.steps):
        _, dt = cpu_assemble_rate(side, threads)
        times.append(dt)
    total = float(sum(times))
    value = side**3 * args.steps / total
    sample = f"{side}^3 elements of the 256^3 grid (same h, same forms) per step, matrix + RHS walk, pattern build untimed"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "3D Q1 Laplace + RHS assembly, 256^3 YaspGrid cube [-1,1]^3 (BASELINE.json configs[1])",
                   "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------------------
# product arm
# ------------------------------------------------------------------------------------------------------------------
def fv_extra(gdt, ctx, torch, hbm_gbs, peak_src):
    """C4: explicit first-order FV upwind apply + fused Euler step on a 4096^2 periodic grid (reported alongside)"""
    from dune_gdt_b200 import descriptors as D

    n = 4096
    grid = gdt.make_cube_grid(ctx, 0.0, 1.0, [n, n], periodic=3)
    space = gdt.make_finite_volume_space(grid)
    out = {}
    for name, flux in (("linear_transport", gdt.NumericalUpwindFlux(D.FLUX_LINEAR, [1.0, 0.5])),
                       ("burgers", gdt.NumericalUpwindFlux(D.FLUX_BURGERS))):
        L = gdt.make_advection_fv_operator(flux, space)
        u = torch.rand(n * n, dtype=torch.float64, device="cuda")
        v = torch.empty_like(u)
        for _ in range(5):
            L.apply_device(u.data_ptr(), v.data_ptr())
        kernel_time(gdt, ctx, "fv_apply")
        steps = 50
        for _ in range(steps):
            L.apply_device(u.data_ptr(), v.data_ptr())
        ms, cnt = kernel_time(gdt, ctx, "fv_apply")
        per = ms / max(cnt, 1)
        bytes_per = 16.0 * n * n
        out[name] = {"cells_per_s": n * n / (per * 1e-3), "ms_per_apply": per,
                     "roofline": {"bound": "hbm", "achieved": bytes_per / (per * 1e-3) / 1e9, "peak": hbm_gbs,
                                  "unit": "GB/s", "frac": bytes_per / (per * 1e-3) / 1e9 / hbm_gbs,
                                  "peak_source": peak_src}}
    # scale: a device copy of the same 134 MB vector (read + write = the apply's 16 B/cell) on the same box -- at this
    # size launch ramp-up and tail keep a plain copy below the 2 GB copy MEASURED_PEAKS.json quotes
    a = torch.rand(n * n, dtype=torch.float64, device="cuda")
    b = torch.empty_like(a)
    for _ in range(5):
        b.copy_(a)
    s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s_.record()
    for _ in range(50):
        b.copy_(a)
    e_.record()
    torch.cuda.synchronize()
    copy_ms = s_.elapsed_time(e_) / 50
    out["same_size_device_copy"] = {"ms": copy_ms, "GBps": 16.0 * n * n / (copy_ms * 1e-3) / 1e9,
                                    "apply_over_copy": out["linear_transport"]["ms_per_apply"] / copy_ms}
    del a, b
    # systems (m = 4): the 2D Euler equations with the Vijayasundaram flux (examples/mpi_2019_02...cc:381-434), 2048^2 cells
    ns = 2048
    sgrid = gdt.make_cube_grid(ctx, -1.0, 1.0, [ns, ns], periodic=3)
    sspace = gdt.make_finite_volume_space(sgrid, 4)
    euler = gdt.EulerTools(2, 1.4)
    Ls = gdt.make_advection_fv_operator(gdt.NumericalVijayasundaramFlux(*euler.flux()), sspace)
    w = torch.empty(ns * ns, 4, dtype=torch.float64, device="cuda")
    w[:, 0] = 1.0 + 0.5 * torch.rand(ns * ns, dtype=torch.float64, device="cuda")
    w[:, 1:3] = 0.2 * (torch.rand(ns * ns, 2, dtype=torch.float64, device="cuda") - 0.5)
    w[:, 3] = 2.0 + torch.rand(ns * ns, dtype=torch.float64, device="cuda")
    wo = torch.empty_like(w)
    for _ in range(3):
        Ls.apply_device(w.data_ptr(), wo.data_ptr())
    kernel_time(gdt, ctx, "fv_apply")
    for _ in range(20):
        Ls.apply_device(w.data_ptr(), wo.data_ptr())
    ms, cnt = kernel_time(gdt, ctx, "fv_apply")
    per = ms / max(cnt, 1)
    out["euler_2d_vijayasundaram_2048^2"] = {
        "cells_per_s": ns * ns / (per * 1e-3), "ms_per_apply": per,
        "roofline": {"bound": "hbm", "achieved": 64.0 * ns * ns / (per * 1e-3) / 1e9, "peak": hbm_gbs, "unit": "GB/s",
                     "frac": 64.0 * ns * ns / (per * 1e-3) / 1e9 / hbm_gbs, "peak_source": peak_src},
        "note": "64 B/cell (4 components in, 4 out); every face flux (eigendecomposition of a 4 x 4 jacobian) is evaluated "
                "from both sides: FP64-pipe bound, not HBM bound"}
    del w, wo
    # the caller of the apply: one SSP3 Runge-Kutta step (tools/timestepper/explicit-rungekutta.hh:237-270), stages fused
    # into the applies (9 vector passes) against separate axpy passes (18)
    lib, check = gdt.capi.lib(), gdt.capi.check
    L = gdt.make_advection_fv_operator(gdt.NumericalUpwindFlux(D.FLUX_LINEAR, [1.0, 0.5]), space)
    u = torch.rand(n * n, dtype=torch.float64, device="cuda")
    for label, env in (("fused", None), ("separate_axpy", "GDTB_RK_NO_FUSE")):
        if env:
            os.environ[env] = "1"
        ts = C.c_void_p()
        check(lib.gdtb_rk_create(L._h, D.RK_SSP3, 0, None, None, None, -1.0, 0.0, C.byref(ts)))
        dt = 0.1 / n
        for _ in range(3):
            check(lib.gdtb_rk_step(ts, C.c_void_p(u.data_ptr()), dt, dt, None))
        ctx.synchronize()
        s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        s_.record()
        for _ in range(reps):
            check(lib.gdtb_rk_step(ts, C.c_void_p(u.data_ptr()), dt, dt, None))
        e_.record()
        ctx.synchronize()
        torch.cuda.synchronize()
        out.setdefault("ssp3_step", {})[label] = {"ms_per_step": s_.elapsed_time(e_) / reps}
        lib.gdtb_rk_destroy(ts)
        if env:
            del os.environ[env]
    return out


def demangle(name):
    """c++filt if the box has it; the mangled symbol otherwise (it is what ncu launch lists show either way)"""
    if not name:
        return name
    try:
        out = subprocess.run(["c++filt", name], capture_output=True, text=True, timeout=5).stdout.strip()
        return out or name
    except (OSError, subprocess.TimeoutExpired):
        return name


def kernel_name(gdt, ctx, family):
    return demangle(gdt.capi.lib().gdtb_ctx_kernel_name(ctx._h, family.encode()).decode())


def qp_function(torch, ne, nq, order):
    """one coefficient value per quadrature point in a device array the library borrows (data_on_device = 1)"""
    from dune_gdt_b200 import descriptors as D

    g = torch.Generator(device="cuda").manual_seed(20251017)
    a = 0.5 + torch.rand(ne, nq, dtype=torch.float64, device="cuda", generator=g)
    f = D.Function()
    f.kind, f.order, f.qp_per_element, f.data_on_device = D.FN_QP_SCALAR, order, nq, 1
    f.data = C.cast(a.data_ptr(), C.POINTER(C.c_double))
    f._keep = a
    return f


def assembly_extra(gdt, ctx, torch, hbm_gbs, peak_src):
    """The other single-GPU configurations, device resident, each with the roofline of ITS dominant kernel:
      C5 3D Q2 Laplace 128^3 and C3 2D SWIPDG DG-Q1 2048^2 (constant coefficients; the launches write the matrix only:
      algorithmic bytes = 8 nnz), C2 with one kappa value per ELEMENT (the non-specialised Q1 number) and C2 / C5 with one
      kappa value per QUADRATURE POINT (HBM: 8 nnz written + 8 n_qp read per element; also given against the FP64 peak),
      and the sort-and-unique pattern build for C2.  The parity tests cover every one of these paths."""
    from dune_gdt_b200 import descriptors as D

    lib, check = gdt.capi.lib(), gdt.capi.check
    out = {}
    fp64_peak = 34.1  # TFLOP/s, measured on this pool's B200 (profiles/r02_fp64_peak.json, tools/microbench/fp64_peak.cu)

    def timed(op_h, family, reps):
        for _ in range(3):
            check(lib.gdtb_assemble_async(op_h, None, D.ASSEMBLE_OVERWRITE))
        ctx.synchronize()
        kernel_time(gdt, ctx, family)
        for _ in range(reps):
            check(lib.gdtb_assemble_async(op_h, None, D.ASSEMBLE_OVERWRITE))
        ctx.synchronize()
        ms, cnt = kernel_time(gdt, ctx, family)
        return ms / max(cnt, 1)

    def entry(name, elements, rows, nnz, per_ms, plan, family, read_bytes=0.0, flops=None, note=None):
        alg = 8.0 * nnz + read_bytes  # the launch writes every CSR value once (no vector), reads only the coefficients
        ach = alg / (per_ms * 1e-3) / 1e9
        out[name] = {"plan": plan, "kernel": kernel_name(gdt, ctx, family), "elements": elements, "rows": rows, "nnz": nnz,
                     "ms_per_assembly": per_ms, "elements_per_s": elements / (per_ms * 1e-3),
                     "roofline": {"bound": "hbm", "achieved": ach, "peak": hbm_gbs, "unit": "GB/s", "frac": ach / hbm_gbs,
                                  "algorithmic_bytes_per_launch": alg, "bytes_per_element": alg / elements,
                                  "peak_source": peak_src}}
        if flops is not None:
            tf = flops / (per_ms * 1e-3) / 1e12
            out[name]["fp64"] = {"flops_per_launch": flops, "achieved_tflops": tf, "peak_tflops": fp64_peak,
                                 "frac": tf / fp64_peak, "peak_source": "measured DFMA loop (profiles/r02_fp64_peak.json)"}
        if note:
            out[name]["note"] = note

    lap = D.form(D.integrand(D.INT_LAPLACE, diffusion=1.0))

    def cg_case(name, n, order, form, family, reps, read_bytes=0.0, flops=None, note=None):
        grid = gdt.make_cube_grid(ctx, -1.0, 1.0, [n, n, n])
        space = gdt.make_continuous_lagrange_space(grid, order)
        op_h = C.c_void_p()
        check(lib.gdtb_matop_create(ctx._h, space._h, space._h, None, C.byref(op_h)))  # closed-form CSR positions
        check(lib.gdtb_matop_append_element(op_h, C.byref(form)))
        per = timed(op_h, family, reps)
        entry(name, n**3, space.mapper.size, int(lib.gdtb_matop_local_nnz(op_h)), per, lib.gdtb_matop_plan(op_h).decode(),
              family, read_bytes, flops, note)
        lib.gdtb_matop_destroy(op_h)
        del space, grid
        torch.cuda.empty_cache()

    # C5, constant kappa: the sum-factorised Q2 gather
    cg_case("c5_q2_laplace_128^3", 128, 2, lap, "q2_gather", 20)
    # C2 with one kappa value per element: what "assembly" costs when kappa is data (no constant-coefficient identities)
    kap = torch.rand(256**3, dtype=torch.float64, device="cuda") + 0.5
    fe = D.Function()
    fe.kind, fe.data_on_device, fe.data = D.FN_ELEM_SCALAR, 1, C.cast(kap.data_ptr(), C.POINTER(C.c_double))
    cg_case("c2_q1_laplace_256^3_kappa_per_element", 256, 1, D.form(D.integrand(D.INT_LAPLACE, diffusion=fe)), "q1_gather", 20,
            read_bytes=8.0 * 256**3)
    del kap
    # the same for C5 (one integrand, one kappa per element): per-element 1D factor tables in the Q2 gather
    kap5 = torch.rand(128**3, dtype=torch.float64, device="cuda") + 0.5
    fe5 = D.Function()
    fe5.kind, fe5.data_on_device, fe5.data = D.FN_ELEM_SCALAR, 1, C.cast(kap5.data_ptr(), C.POINTER(C.c_double))
    cg_case("c5_q2_laplace_128^3_kappa_per_element", 128, 2, D.form(D.integrand(D.INT_LAPLACE, diffusion=fe5)), "q2_gather", 10,
            read_bytes=8.0 * 128**3)
    del kap5
    # kappa per quadrature point (declared order 0: 2^3 points for Q1, 3^3 for Q2): the quadrature loop itself, sum-factorised
    f1 = qp_function(torch, 256**3, 8, 0)
    cg_case("c2_q1_laplace_256^3_kappa_per_qp", 256, 1, D.form(D.integrand(D.INT_LAPLACE, diffusion=f1)), "q1_gather", 5,
            read_bytes=8.0 * 8 * 256**3, flops=2.0 * 8 * 3 * 8 * 8 * 256**3,
            note="flops: dense B^T D B count without symmetry, 2 * n_qp * d * n^2 per element (SURVEY.md 8d)")
    del f1
    f2 = qp_function(torch, 128**3, 27, 0)
    cg_case("c5_q2_laplace_128^3_kappa_per_qp", 128, 2, D.form(D.integrand(D.INT_LAPLACE, diffusion=f2)), "q2_gather", 3,
            read_bytes=8.0 * 27 * 128**3, flops=2.0 * 27 * 3 * 27 * 27 * 128**3,
            note="flops: dense B^T D B count, 118 098 per element (SURVEY.md 8d); the kernel executes the sum-factorised "
                 "form (about 1 300 flops per local-matrix row)")
    del f2
    torch.cuda.empty_cache()

    # C3: SWIPDG as in examples/adaptive_elliptic_swipdg.cc:230-251 / test ESV2007.hh:108-112 on a Yasp grid
    n = 2048
    grid = gdt.make_cube_grid(ctx, -1.0, 1.0, [n, n])
    space = gdt.make_discontinuous_lagrange_space(grid, 1)
    pat = gdt.make_sparsity_pattern(space, space, gdt.Stencil.element_and_intersection)
    op = gdt.MatrixOperator(space, space, pat)
    inner = D.form([D.integrand(D.INT_IPDG_INNER_COUPLING, prefactor=1.0, diffusion=1.0, weight=1.0),
                    D.integrand(D.INT_IPDG_INNER_PENALTY, prefactor=8.0, weight=1.0, hI_kind=D.HI_VOLUME)])
    bnd = D.form([D.integrand(D.INT_IPDG_DIRICHLET_COUPLING, prefactor=1.0, diffusion=1.0),
                  D.integrand(D.INT_IPDG_BOUNDARY_PENALTY, prefactor=14.0, weight=1.0, hI_kind=D.HI_VOLUME)])
    check(lib.gdtb_matop_append_element(op._h, C.byref(lap)))
    check(lib.gdtb_matop_append_coupling(op._h, C.byref(inner), D.FILTER_INNER_ONCE))
    check(lib.gdtb_matop_append_boundary(op._h, C.byref(bnd), D.FILTER_ALL_BOUNDARY))
    per = timed(op._h, "dg_gather", 20)
    entry("c3_swipdg_dg_q1_2048^2", n * n, pat.rows, pat.nnz, per, op.plan, "dg_gather")
    # the same operator with kappa = omega = one value per element (SWIPDG's weight is the diffusion)
    kap3 = torch.rand(n * n, dtype=torch.float64, device="cuda") + 0.5

    def fe3():
        f = D.Function()
        f.kind, f.data_on_device, f.data = D.FN_ELEM_SCALAR, 1, C.cast(kap3.data_ptr(), C.POINTER(C.c_double))
        return f

    op2 = gdt.MatrixOperator(space, space, pat)
    lap_e = D.form(D.integrand(D.INT_LAPLACE, diffusion=fe3()))
    inner_e = D.form([D.integrand(D.INT_IPDG_INNER_COUPLING, prefactor=1.0, diffusion=fe3(), weight=fe3()),
                      D.integrand(D.INT_IPDG_INNER_PENALTY, prefactor=8.0, weight=fe3(), hI_kind=D.HI_VOLUME)])
    bnd_e = D.form([D.integrand(D.INT_IPDG_DIRICHLET_COUPLING, prefactor=1.0, diffusion=fe3()),
                    D.integrand(D.INT_IPDG_BOUNDARY_PENALTY, prefactor=14.0, weight=fe3(), hI_kind=D.HI_VOLUME)])
    check(lib.gdtb_matop_append_element(op2._h, C.byref(lap_e)))
    check(lib.gdtb_matop_append_coupling(op2._h, C.byref(inner_e), D.FILTER_INNER_ONCE))
    check(lib.gdtb_matop_append_boundary(op2._h, C.byref(bnd_e), D.FILTER_ALL_BOUNDARY))
    per = timed(op2._h, "dg_gather", 10)
    entry("c3_swipdg_dg_q1_2048^2_kappa_per_element", n * n, pat.rows, pat.nnz, per, op2.plan, "dg_gather", read_bytes=8.0 * n * n)
    del op2, kap3
    del op, pat, space, grid
    torch.cuda.empty_cache()

    # K1: the sort-and-unique pattern builder on C2 (setup, not part of a step): emit keys, radix sort, unique, row pointer
    grid = gdt.make_cube_grid(ctx, -1.0, 1.0, [256, 256, 256])
    space = gdt.make_continuous_lagrange_space(grid, 1)
    times = {}
    for label, method in (("sort_unique", D.PATTERN_SORT_UNIQUE), ("closed_form", D.PATTERN_STRUCTURED)):
        ctx.synchronize()
        t0 = time.perf_counter()
        pat = gdt.SparsityPattern(space, space, D.STENCIL_ELEMENT, method)
        ctx.synchronize()
        times[label] = time.perf_counter() - t0
        nnz = pat.nnz
        del pat
        torch.cuda.empty_cache()
    out["c2_pattern_build_256^3"] = {"nnz": nnz, "sort_unique_s": times["sort_unique"], "closed_form_s": times["closed_form"],
                                    "note": "includes allocation; CUB radix sort / select inside the sort-and-unique builder"}
    return out


class Env:
    """one process per GPU: torch.distributed (NCCL) plumbing + the library context on a torch stream"""

    def __init__(self):
        import torch
        import torch.distributed as dist

        import dune_gdt_b200 as gdt

        self.torch, self.dist, self.gdt = torch, dist, gdt
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device -- the product has no CPU path (use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
        self.ctx = gdt.Context(self.local_rank)
        # the library launches on the stream bench.py times with torch CUDA events (a non-default torch stream: handle 0
        # would mean "use the library's own stream")
        self.stream = torch.cuda.Stream()
        torch.cuda.set_stream(self.stream)
        self.ctx.set_stream(self.stream.cuda_stream)

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, value):
        t = self.torch.tensor([value], dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.item()

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def sharded_workload(env, workload, steps, warmup):
    """The other multi-GPU rows of SURVEY.md 8e (one rank per GPU), device resident, CUDA events, max over ranks; returns
    the JSON record (every rank computes it, rank 0 prints it).
      c5           3D Q2 Laplace 128^3 SHARDED across the ranks (BASELINE.json configs[4], strong scaling): z-slabs,
                   owner-computes-rows per sub-entity group, no data-path collective
      c4 / c4-weak explicit FV upwind Euler steps on 4096^2 periodic, y-slabs, one ghost row per side over NCCL per step
                   (interior overlapped with the exchange)
      c4-p2p / c4-weak-p2p   the same with the ghost rows handed over INSIDE the kernel (NVLink peer stores)
      c3 / c3-weak 2D SWIPDG DG-Q1 2048^2 (BASELINE.json configs[2]), y-slabs of element-owned rows, no collective
      c2-halo      the headline workload with the interface-row halo partition (own elements only + one message per
                   slab face + add) instead of the ghost-layer recompute, weak scaling
      c2-halo-p2p  the same with the interface rows handed over inside the gather kernel (peer stores + counters)"""
    torch, gdt = env.torch, env.gdt
    from dune_gdt_b200 import descriptors as D
    from dune_gdt_b200 import parallel

    rank, world, ctx = env.rank, env.world, env.ctx
    warmup = max(warmup, 3)

    def timed(step, units, metric, unit, scaling, config):
        for _ in range(warmup):
            step()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        env.barrier()
        start.record()
        for _ in range(steps):
            step()
        end.record()
        env.barrier()
        dt = env.max_over_ranks(start.elapsed_time(end) * 1e-3)
        return {"workload": workload, "metric": metric, "value": units * steps / dt, "unit": unit, "n_gpus": world,
                "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True,
                "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config}

    if workload == "c5":
        n = 128
        grid = gdt.make_cube_grid(ctx, -1.0, 1.0, [n, n, n])
        space = gdt.make_continuous_lagrange_space(grid, 2)
        slab = parallel.SlabAssembly(space, rank, world, with_functional=False)
        slab.append(D.form(D.integrand(D.INT_LAPLACE, diffusion=1.0)))
        return timed(slab.assemble_device, n**3, "elements assembled/sec (3D Q2 Laplace, 128^3 sharded, FP64)", UNIT, "strong",
                     {"workload": "3D Q2 Laplace assembly, 128^3 YaspGrid cube sharded across the GPUs (BASELINE.json configs[4])",
                      "partition": f"z-slabs x{world}, owner-computes-rows per sub-entity group, no collective",
                      "nnz_this_rank": slab.nnz_local})
    if workload in ("c3", "c3-weak"):
        # BASELINE.json configs[2] at throughput size: SWIPDG DG-Q1 (element + inner coupling + Dirichlet boundary forms),
        # y-slabs of element rows; rows are element-owned, the neighbour across a slab face enters through its index only
        n = 2048
        ny = n * world if workload == "c3-weak" else n
        grid = gdt.make_cube_grid(ctx, [-1.0, -1.0], [1.0, -1.0 + ny * (2.0 / n)], [n, ny])
        space = gdt.make_discontinuous_lagrange_space(grid, 1)
        slab = parallel.SlabAssembly(space, rank, world, with_functional=False)
        slab.append(D.form(D.integrand(D.INT_LAPLACE, diffusion=1.0)))
        slab.append_coupling(D.form([D.integrand(D.INT_IPDG_INNER_COUPLING, prefactor=1.0, diffusion=1.0, weight=1.0),
                                     D.integrand(D.INT_IPDG_INNER_PENALTY, prefactor=8.0, weight=1.0, hI_kind=D.HI_VOLUME)]))
        slab.append_boundary(D.form([D.integrand(D.INT_IPDG_DIRICHLET_COUPLING, prefactor=1.0, diffusion=1.0),
                                     D.integrand(D.INT_IPDG_BOUNDARY_PENALTY, prefactor=14.0, weight=1.0, hI_kind=D.HI_VOLUME)]))
        return timed(slab.assemble_device, n * ny, "elements assembled/sec (2D SWIPDG DG-Q1, 2048^2, FP64)", UNIT,
                     "weak" if workload == "c3-weak" else "strong",
                     {"workload": f"2D SWIPDG DG-Q1 assembly, {n} x {ny} YaspGrid (BASELINE.json configs[2] at throughput size)",
                      "partition": f"y-slabs x{world}, element-owned rows, no collective", "nnz_this_rank": slab.nnz_local})
    if workload in ("c4-p2p", "c4-weak-p2p"):
        n = 4096
        ny = n * world if workload == "c4-weak-p2p" else n
        grid = gdt.make_cube_grid(ctx, [0.0, 0.0], [1.0, ny / n], [n, ny], periodic=3)
        space = gdt.make_finite_volume_space(grid)
        loop = parallel.PeerMemoryFvTimeLoop(gdt.NumericalUpwindFlux(D.FLUX_LINEAR, [1.0, 0.5]), space, rank, world)
        loop.set_initial_values(np.random.default_rng(20251017).random(n * ny))
        rec = timed(lambda: loop.euler_steps(0.25 / n, 1), n * ny,
                    "cells updated/sec (explicit FV upwind Euler step, 4096^2 periodic, FP64)", "cells/s",
                    "weak" if workload == "c4-weak-p2p" else "strong",
                    {"workload": f"FV linear advection, {n} x {ny} periodic YaspGrid (BASELINE.json configs[3]), fused apply + Euler update",
                     "partition": f"y-slabs x{world}, ghost rows handed over INSIDE the kernel (NVLink peer stores + step counters), "
                                  "one launch per step, no host-launched collective"})
        loop.check()
        loop.close()
        return rec
    if workload in ("c4", "c4-weak"):
        n = 4096
        ny = n * world if workload == "c4-weak" else n
        grid = gdt.make_cube_grid(ctx, [0.0, 0.0], [1.0, ny / n], [n, ny], periodic=3)
        space = gdt.make_finite_volume_space(grid)
        L = parallel.make_distributed_advection_fv_operator(gdt.NumericalUpwindFlux(D.FLUX_LINEAR, [1.0, 0.5]), space, rank, world)
        a = torch.rand(L.local_size, dtype=torch.float64, device="cuda")
        b = torch.empty_like(a)
        state = [a, b]

        def step():
            L.euler_step(state[0], state[1], 0.25 / n)
            state.reverse()

        return timed(step, n * ny, "cells updated/sec (explicit FV upwind Euler step, 4096^2 periodic, FP64)", "cells/s",
                     "weak" if workload == "c4-weak" else "strong",
                     {"workload": f"FV linear advection, {n} x {ny} periodic YaspGrid (BASELINE.json configs[3]), fused apply + Euler update",
                      "partition": f"y-slabs x{world}, one ghost row per side per step over NCCL send/recv, interior overlapped"})
    if workload in ("c2-halo", "c2-halo-p2p"):
        h = 2.0 / NX
        grid = gdt.make_cube_grid(ctx, [-1.0, -1.0, -1.0], [1.0, 1.0, -1.0 + NX * world * h], [NX, NX, NX * world])
        space = gdt.make_continuous_lagrange_space(grid, 1)
        p2p = workload == "c2-halo-p2p"
        halo = parallel.HaloSlabAssembly(space, rank, world, p2p=p2p)
        lap, rhs = forms()
        halo.append(lap)
        halo.append_rhs(rhs)
        how = ("handed over INSIDE the gather kernel (NVLink peer stores + counters, no host-launched collective)" if p2p
               else "one NCCL message per slab face + add kernel, stream-ordered")
        rec = timed(halo.assemble_device, NX**3 * world, METRIC, UNIT, "weak",
                    {"workload": "3D Q1 Laplace + RHS assembly, 256^3 per GPU (BASELINE.json configs[1])",
                     "partition": f"z-slabs x{world}, own elements only + interface-row halo (one layer of rows per slab face) {how}",
                     "halo_bytes_per_face": 8 * (halo.mat_layout[2] + halo.vec_layout[2])})
        halo.check()
        halo.close()
        return rec
    raise SystemExit(f"unknown workload {workload}")


def run_sharded_workload(args):
    env = Env()
    rec = sharded_workload(env, args.workload, args.steps, args.warmup)
    if env.rank == 0:
        print(json.dumps(rec))
    env.close()


def neighbour_self_check(env, op_h, fun_h, space, rows_local, nnz_local):
    """N > 1: every rank ALSO assembles the first vertex layer of the slab above it (from its own top element layer as the
    ghost layer below) and sends it up; the owner compares it with its own first layer, which it produced from ITS ghost
    layer: both are complete rows of the same global matrix / vector, so they must be bit-identical.  Returns the number
    of differing entries over all ranks (0 = the partition is consistent)."""
    torch, dist, gdt = env.torch, env.dist, env.gdt
    from dune_gdt_b200 import descriptors as D

    lib, check = gdt.capi.lib(), gdt.capi.check
    rank, world, ctx = env.rank, env.world, env.ctx
    layer_rows = (NX + 1) ** 2
    layer_nnz = 3 * (3 * NX + 1) ** 2  # an interior vertex layer: 3 z-planes of (3 N + 1)^2 entries
    bad = 0
    probe_v = probe_b = None
    if rank + 1 < world:
        ph, pf = C.c_void_p(), C.c_void_p()
        check(lib.gdtb_matop_create(ctx._h, space._h, space._h, None, C.byref(ph)))
        check(lib.gdtb_vecfun_create(ctx._h, space._h, C.byref(pf)))
        lo = (rank + 1) * NX
        check(lib.gdtb_matop_set_slab(ph, lo, lo + 1))
        check(lib.gdtb_vecfun_set_slab(pf, lo, lo + 1))
        lap, rhs = forms()
        check(lib.gdtb_matop_append_element(ph, C.byref(lap)))
        check(lib.gdtb_vecfun_append_element(pf, C.byref(rhs)))
        check(lib.gdtb_assemble(ph, pf, D.ASSEMBLE_OVERWRITE))
        assert lib.gdtb_matop_local_nnz(ph) == layer_nnz
        pv, pb = C.c_void_p(), C.c_void_p()
        check(lib.gdtb_matop_values_device(ph, C.byref(pv)))
        check(lib.gdtb_vecfun_device(pf, C.byref(pb)))
        probe_v = as_tensor(torch, pv.value, layer_nnz).clone()
        probe_b = as_tensor(torch, pb.value, layer_rows).clone()
        lib.gdtb_matop_destroy(ph)
        lib.gdtb_vecfun_destroy(pf)
    ops = []
    recv_v = recv_b = None
    if rank + 1 < world:
        ops += [dist.P2POp(dist.isend, probe_v, rank + 1), dist.P2POp(dist.isend, probe_b, rank + 1)]
    if rank > 0:
        recv_v = torch.empty(layer_nnz, dtype=torch.float64, device="cuda")
        recv_b = torch.empty(layer_rows, dtype=torch.float64, device="cuda")
        ops += [dist.P2POp(dist.irecv, recv_v, rank - 1), dist.P2POp(dist.irecv, recv_b, rank - 1)]
    if ops:
        for r in dist.batch_isend_irecv(ops):
            r.wait()
    if rank > 0:
        pv, pb = C.c_void_p(), C.c_void_p()
        check(lib.gdtb_matop_values_device(op_h, C.byref(pv)))
        check(lib.gdtb_vecfun_device(fun_h, C.byref(pb)))
        mine_v = as_tensor(torch, pv.value, nnz_local)[:layer_nnz]
        mine_b = as_tensor(torch, pb.value, rows_local)[:layer_rows]
        torch.cuda.synchronize()
        bad = int((mine_v != recv_v).sum().item() + (mine_b != recv_b).sum().item())
    t = torch.tensor([bad], dtype=torch.int64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return int(t.item())


def as_tensor(torch, ptr, n):
    class _Arr:
        __cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 3}

    return torch.as_tensor(_Arr(), device="cuda")


def oracle_checksum(threads):
    """the first 1024 CSR values + the first 1024 right-hand-side entries of the C2 system from the CPU oracle: the
    values belong to vertices of the line iy = iz = 0, the right-hand-side entries to the vertex lines iy <= 3 of the
    plane iz = 0 -- they only see the elements ey <= 3, ez = 0, so a 256 x 5 x 2 sub-grid with the same h (and the same
    257 vertices per line) reproduces them exactly, row for row"""
    import oracle
    from dune_gdt_b200 import descriptors as D

    h = 2.0 / NX
    g = D.grid_desc([-1.0, -1.0, -1.0], [1.0, -1.0 + 5 * h, -1.0 + 2 * h], [NX, 5, 2])
    rp, ci = oracle.pattern(g, (D.SPACE_CG, 1))
    lap, rhs = forms()
    v, b = oracle.assemble(g, D.SPACE_CG, 1, rp, ci, [lap], rhs_forms=[rhs], num_threads=threads)
    # rows of the vertices (ix, 0, 0): 8 entries at the two ends, 12 inside -- identical in the full grid
    n_rows = int(np.searchsorted(rp, 1024, side="right"))
    assert n_rows <= NX + 1
    return float(v[:1024].sum() + b[:1024].sum()), v[:1024], b[:1024]


def run_product(args):
    env = Env()
    torch, dist, gdt = env.torch, env.dist, env.gdt
    from dune_gdt_b200 import descriptors as D

    rank, world, local_rank, ctx = env.rank, env.world, env.local_rank, env.ctx
    lib = gdt.capi.lib()
    check = gdt.capi.check
    barrier = env.barrier

    # global grid: 256 x 256 x (256 * world), h = 2/256 everywhere; this rank owns element layers [rank*256, (rank+1)*256)
    h = 2.0 / NX
    grid = gdt.make_cube_grid(ctx, [-1.0, -1.0, -1.0], [1.0, 1.0, -1.0 + NX * world * h], [NX, NX, NX * world])
    space = gdt.make_continuous_lagrange_space(grid, 1)
    op_h, fun_h = C.c_void_p(), C.c_void_p()
    check(lib.gdtb_matop_create(ctx._h, space._h, space._h, None, C.byref(op_h)))  # closed-form Q1 stencil
    check(lib.gdtb_vecfun_create(ctx._h, space._h, C.byref(fun_h)))
    check(lib.gdtb_matop_set_slab(op_h, rank * NX, (rank + 1) * NX))
    check(lib.gdtb_vecfun_set_slab(fun_h, rank * NX, (rank + 1) * NX))
    lap, rhs = forms()
    check(lib.gdtb_matop_append_element(op_h, C.byref(lap)))
    check(lib.gdtb_vecfun_append_element(fun_h, C.byref(rhs)))
    plan = lib.gdtb_matop_plan(op_h).decode()
    nnz_local = lib.gdtb_matop_local_nnz(op_h)
    rb, re_, vo = C.c_int64(), C.c_int64(), C.c_int64()
    check(lib.gdtb_matop_local_rows(op_h, C.byref(rb), C.byref(re_), C.byref(vo)))
    rows_local = re_.value - rb.value
    elements_local = NX**3

    # ---- device-resident throughput -----------------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        check(lib.gdtb_assemble_async(op_h, fun_h, D.ASSEMBLE_OVERWRITE))
    ctx.synchronize()
    check(lib.gdtb_ctx_enable_timing(ctx._h, 1))
    kernel_time(gdt, ctx, "q1_gather")
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = ctx.launch_count
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.time()
    start.record()
    for _ in range(args.steps):
        check(lib.gdtb_assemble_async(op_h, fun_h, D.ASSEMBLE_OVERWRITE))
    end.record()
    barrier()
    t1 = time.time()
    ms_total = start.elapsed_time(end)
    launches = ctx.launch_count - launches0
    kern_ms, kern_n = kernel_time(gdt, ctx, "q1_gather")
    check(lib.gdtb_ctx_enable_timing(ctx._h, 0))
    ms_total = env.max_over_ranks(ms_total)
    value = elements_local * world * args.steps / (ms_total * 1e-3)
    clocks = None
    if sampler:
        # keep the GPU under the same load a little longer if the timed region was too short to be sampled
        if t1 - t0 < 0.3:
            t0 = time.time()
            while time.time() - t0 < 0.4:
                for _ in range(20):
                    check(lib.gdtb_assemble_async(op_h, fun_h, D.ASSEMBLE_OVERWRITE))
                ctx.synchronize()
            t1 = time.time()
        sampler.stop()
        clocks = sampler.summary(t0, t1)

    # ---- roofline of the dominant kernel -------------------------------------------------------------------
    hbm_gbs, peak_src = measured_peaks()
    alg_bytes = 8.0 * nnz_local + 8.0 * rows_local  # every CSR value and RHS entry written once, no array inputs
    per_launch_ms = kern_ms / max(kern_n, 1)
    achieved = alg_bytes / (per_launch_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    for name in ("r02_q1_gather_traffic.json", "q1_gather_traffic.json"):
        prof = os.path.join(ROOT, "profiles", name)
        if os.path.exists(prof):
            with open(prof) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
            traffic_src = f"profiles/{name} (ncu --set full: dram__bytes_read.sum + dram__bytes_write.sum of one launch)"
            break
    roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_gbs, "unit": "GB/s", "frac": achieved / hbm_gbs,
                "traffic": traffic, "traffic_source": traffic_src, "kernel": kernel_name(gdt, ctx, "q1_gather"),
                "algorithmic_bytes_per_launch": alg_bytes,
                "bytes_per_element": alg_bytes / elements_local, "ms_per_launch": per_launch_ms,
                "peak_source": peak_src,
                "note": "the kernel only writes (no array inputs): the copy-derived peak (half reads, half writes) is "
                        "not an upper bound for it; a pure fill reaches about 7.4 TB/s on this box (tools/copy_bandwidth.py)"}

    # ---- N > 1: the partition is consistent across the ranks (bit-identical interface rows) ------------------
    self_check = None
    if world > 1:
        differing = neighbour_self_check(env, op_h, fun_h, space, rows_local, nnz_local)
        self_check = {"interface_layers_compared": world - 1, "differing_entries": differing, "ok": differing == 0,
                      "what": "every rank re-assembles the first vertex layer of the slab above it and the owner compares "
                              "it bit for bit with its own (matrix rows + right-hand side)"}

    # ---- end to end through the C ABI with host buffers ----------------------------------------------------
    e2e_steps = max(1, min(args.steps, 5))
    values_host = torch.empty(nnz_local, dtype=torch.float64).pin_memory()
    rhs_host = torch.empty(rows_local, dtype=torch.float64).pin_memory()
    vp = C.cast(values_host.data_ptr(), C.POINTER(C.c_double))
    bp = C.cast(rhs_host.data_ptr(), C.POINTER(C.c_double))

    def e2e_step():
        # the call a user of the reference-facing API makes: append the local forms (host descriptors -> device),
        # one grid walk, results in host memory
        check(lib.gdtb_matop_clear_forms(op_h))
        check(lib.gdtb_vecfun_clear_forms(fun_h))
        check(lib.gdtb_matop_append_element(op_h, C.byref(lap)))
        check(lib.gdtb_vecfun_append_element(fun_h, C.byref(rhs)))
        check(lib.gdtb_assemble_host(op_h, fun_h, vp, bp))

    e2e_step()
    barrier()
    te = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = env.max_over_ranks(time.perf_counter() - te)
    e2e_value = elements_local * world * e2e_steps / e2e_s
    checksum = float(values_host[:1024].sum() + rhs_host[:1024].sum())
    head_values = values_host[:1024].numpy().copy()
    head_rhs = rhs_host[:1024].numpy().copy()
    d2h = 8 * (nnz_local + rows_local)
    e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 2 * C.sizeof(D.Form),
           "d2h_bytes_per_step": d2h, "steps": e2e_steps,
           "ms_per_step": 1e3 * e2e_s / e2e_steps, "result_checksum": checksum,
           "d2h_gb_per_s_per_rank": d2h / (e2e_s / e2e_steps) / 1e9,
           "note": "bound by the device-to-host copy of the 3.77 GB result per rank into pinned host memory (the assembly "
                   "itself is 0.6 ms of the step); with N ranks on one host the copies share the host's memory / PCIe "
                   "root bandwidth"}
    del values_host, rhs_host
    lib.gdtb_matop_destroy(op_h)
    lib.gdtb_vecfun_destroy(fun_h)
    torch.cuda.empty_cache()

    extra = None
    cpu = None
    if world == 1:
        check(lib.gdtb_ctx_enable_timing(ctx._h, 1))
        extra = {"fv_apply_4096x4096_upwind": fv_extra(gdt, ctx, torch, hbm_gbs, peak_src)}
        extra.update(assembly_extra(gdt, ctx, torch, hbm_gbs, peak_src))
        check(lib.gdtb_ctx_enable_timing(ctx._h, 0))
        if not args.no_cpu_baseline:
            import oracle

            oracle.build()
            threads = os.cpu_count() or 1
            # the oracle as the CHECKER of what the timed steps produced: first 1024 values + first 1024 rhs entries
            ref_sum, ref_v, ref_b = oracle_checksum(threads)
            scale = max(float(np.abs(ref_v).max()), float(np.abs(ref_b).max()))
            err = max(float(np.abs(head_values - ref_v).max()), float(np.abs(head_rhs - ref_b).max())) / scale
            e2e["result_checksum_check"] = {"oracle_checksum": ref_sum, "max_rel_err_first_1024": err, "ok": err <= 1e-12}
            side = cpu_sample_side(threads, budget_s=8.0)
            rate, dt = cpu_assemble_rate(side, threads, repeats=2)
            cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"{side}^3 elements of the 256^3 grid (same h and forms), best of 2 walks of {dt:.2f} s, "
                             f"oracle port with {threads} threads + row-striped locks, pattern build untimed"}
    else:
        # the multi-GPU rows that DO communicate, so that the scaling record carries them next to the headline
        extra = {"multi_gpu": {}}
        sub_steps = max(10, min(args.steps, 50))
        for wl in ("c5", "c3", "c2-halo-p2p", "c2-halo", "c4-weak-p2p", "c4-weak"):
            try:
                extra["multi_gpu"][wl] = sharded_workload(env, wl, sub_steps, args.warmup)
            except Exception as exc:  # a failed side workload must not take the headline line with it
                extra["multi_gpu"][wl] = {"error": f"{type(exc).__name__}: {exc}"}
            torch.cuda.empty_cache()

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": "3D Q1 Laplace + RHS assembly, 256^3 YaspGrid cube [-1,1]^3 per GPU (BASELINE.json configs[1]); "
                            "kappa = I, f = (3 pi^2/4) prod cos(pi x_i/2) declared order 3",
                "grid": [NX, NX, NX * world], "elements_per_gpu": elements_local, "nnz_per_gpu": nnz_local,
                "rows_per_gpu": rows_local, "plan": plan, "partition": f"z-slabs x{world}, owner-computes-rows",
                "l2": "no flush needed: each step streams 3.77 GB of output per GPU through a 126 MB L2",
            },
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
        }
        if self_check:
            line["self_check"] = self_check
        if extra:
            line["extra"] = extra
        print(json.dumps(line))
    env.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="product", choices=["product", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="c2", choices=["c2", "c5", "c3", "c3-weak", "c4", "c4-weak", "c4-p2p", "c4-weak-p2p", "c2-halo", "c2-halo-p2p"],
                    help="c2 (default, the headline line); c5 / c4 / c2-halo: the other multi-GPU rows, see run_sharded_workload")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.steps == 200:
            args.steps, args.warmup = 5, 1
        run_reference(args)
    elif args.workload != "c2":
        if args.steps == 200:
            args.steps = 50
        run_sharded_workload(args)
    else:
        run_product(args)


if __name__ == "__main__":
    main()
