"""Multi-GPU path on real devices: world_size 2 over NCCL (skipped on a single-GPU box; run with `gpurun --gpus 2`)."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    import dune_gdt_b200 as gdt
    import oracle
    from dune_gdt_b200 import descriptors as D
    from dune_gdt_b200 import parallel

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    try:
        ctx = gdt.Context(rank)
        rng = np.random.default_rng(20251017)
        # ---- FV: distributed apply and a few Euler steps against the oracle on the whole grid ------------------
        for n, periodic, fk, params in (([96, 64], 3, D.FLUX_LINEAR, [1.0, 0.5]), ([64, 50], 0, D.FLUX_BURGERS, []),
                                        ([24, 20, 18], 7, D.FLUX_LINEAR, [1.0, -0.5, 0.25])):
            u = rng.random(int(np.prod(n)))
            grid = gdt.make_cube_grid(ctx, 0.0, 1.0, n, periodic=periodic)
            space = gdt.make_finite_volume_space(grid)
            L = parallel.make_distributed_advection_fv_operator(gdt.NumericalUpwindFlux(fk, params), space, rank, world)
            gd = D.grid_desc(0.0, 1.0, n, periodic=periodic)
            fl = D.flux(fk, D.NUMFLUX_UPWIND, params)
            ref = oracle.fv_apply(gd, fl, u)
            src = torch.from_numpy(L.scatter_from_global(u)).cuda()
            dst = torch.zeros_like(src)
            L.apply(src, dst)
            torch.cuda.synchronize()
            own = L.owned_view(dst).cpu().numpy()
            err = np.abs(own - ref[L.begin * L.plane:L.end * L.plane]).max() / np.abs(ref).max()
            assert err <= 1e-12, ("apply", n, err)
            dt, steps = 0.2 / max(n), 5
            ref_e = oracle.fv_euler(gd, fl, u, dt, steps)
            a, b = src, dst
            for _ in range(steps):
                L.euler_step(a, b, dt)
                a, b = b, a
            torch.cuda.synchronize()
            own = L.owned_view(a).cpu().numpy()
            err = np.abs(own - ref_e[L.begin * L.plane:L.end * L.plane]).max() / np.abs(ref_e).max()
            assert err <= 1e-12, ("euler", n, err)
        # ---- FV: the same Euler loops with the peer-memory ghost exchange (no collective per step) --------------
        for n, periodic, fk, params in (([96, 64], 3, D.FLUX_LINEAR, [1.0, 0.5]), ([64, 50], 0, D.FLUX_BURGERS, []),
                                        ([24, 20, 18], 7, D.FLUX_LINEAR, [1.0, -0.5, 0.25]), ([33, 40], 2, D.FLUX_BURGERS, [])):
            u = rng.random(int(np.prod(n)))
            grid = gdt.make_cube_grid(ctx, 0.0, 1.0, n, periodic=periodic)
            space = gdt.make_finite_volume_space(grid)
            loop = parallel.PeerMemoryFvTimeLoop(gdt.NumericalUpwindFlux(fk, params), space, rank, world)
            loop.set_initial_values(u)
            dt, steps = 0.2 / max(n), 7
            loop.euler_steps(dt, steps)
            loop.check()
            ref_e = oracle.fv_euler(D.grid_desc(0.0, 1.0, n, periodic=periodic), D.flux(fk, D.NUMFLUX_UPWIND, params), u, dt, steps)
            own = loop.owned_view(loop.current()).cpu().numpy()
            err = np.abs(own - ref_e[loop.begin * loop.plane:loop.end * loop.plane]).max() / np.abs(ref_e).max()
            assert err <= 1e-12, ("p2p euler", n, err)
            loop.close()
        # ---- FV: Runge-Kutta on slabs, stage vectors handed over by peer stores ------------------------------
        for n, periodic, fk, params, method in (([64, 48], 3, D.FLUX_BURGERS, [], D.RK_SSP3),
                                                ([40, 33], 0, D.FLUX_LINEAR, [1.0, 0.5], D.RK_SSP2),
                                                ([12, 10, 16], 7, D.FLUX_LINEAR, [1.0, -0.5, 0.25], D.RK_CLASSIC4)):
            u = rng.uniform(0.1, 1.0, int(np.prod(n)))
            grid = gdt.make_cube_grid(ctx, 0.0, 1.0, n, periodic=periodic)
            space = gdt.make_finite_volume_space(grid)
            ts = parallel.PeerMemoryRkTimeStepper(gdt.NumericalUpwindFlux(fk, params), space, rank, world, method)
            ts.set_initial_values(u)
            gd = D.grid_desc(0.0, 1.0, n, periodic=periodic)
            fl = D.flux(fk, D.NUMFLUX_UPWIND, params)
            dt = 0.3 * oracle.fv_estimate_dt(gd, fl, u)
            ts.solve(5.5 * dt, dt)
            ts.check()
            ref, steps, t = oracle.rk_solve(gd, fl, D.BUTCHER[method], u, 5.5 * dt, dt, r=-1.0)
            assert ts.num_steps == steps and ts.current_time() == t
            own = ts.owned_view().cpu().numpy()
            err = np.abs(own - ref[ts.begin * ts.plane:ts.end * ts.plane]).max() / np.abs(ref).max()
            assert err <= 1e-12, ("p2p rk", n, err)
            ts.close()
        # ---- assembly: concatenated slab results == oracle on the whole grid ----------------------------------
        n = [10, 9, 11]
        grid = gdt.make_cube_grid(ctx, -1.0, 1.0, n)
        space = gdt.make_continuous_lagrange_space(grid, 1)
        slab = parallel.SlabAssembly(space, rank, world)
        kap = rng.uniform(0.5, 2.0, int(np.prod(n)))
        lap = D.form(D.integrand(D.INT_LAPLACE, diffusion=D.fn_elem(kap)))
        src_f = D.fn_builtin(D.BUILTIN_COS_PRODUCT, 3, 0.75 * np.pi**2, 0.5 * np.pi)
        rhs = D.form(D.integrand(D.INT_PRODUCT, diffusion=1.0, weight=src_f))
        slab.append(lap)
        slab.append_rhs(rhs)
        values, vector = slab.assemble()
        gd = D.grid_desc(-1.0, 1.0, n)
        rp, ci = oracle.pattern(gd, (D.SPACE_CG, 1))
        v_ref, b_ref = oracle.assemble(gd, D.SPACE_CG, 1, rp, ci, [lap], rhs_forms=[rhs])
        assert slab.value_offset == rp[slab.row_begin] and slab.nnz_local == rp[slab.row_end] - rp[slab.row_begin]
        seg = v_ref[slab.value_offset:slab.value_offset + slab.nnz_local]
        assert np.abs(values - seg).max() / np.abs(v_ref).max() <= 1e-12
        assert np.abs(vector - b_ref[slab.row_begin:slab.row_end]).max() / np.abs(b_ref).max() <= 1e-12
        nnz = torch.tensor([slab.nnz_local], device="cuda")
        dist.all_reduce(nnz)
        assert int(nnz.item()) == len(v_ref)
        # ---- the other partition: own elements only + interface-row halo over NCCL (SURVEY.md 8e scheme 2) ----
        halo = parallel.HaloSlabAssembly(space, rank, world)
        halo.append(lap)
        halo.append_rhs(rhs)
        hv, hb = halo.assemble_device()
        hv, hb = hv.cpu().numpy(), hb.cpu().numpy()
        assert halo.value_offset == slab.value_offset and halo.nnz_owned == slab.nnz_local
        assert np.abs(hv[:halo.nnz_owned] - seg).max() / np.abs(v_ref).max() <= 1e-12
        assert np.abs(hb[:halo.rows_owned] - b_ref[slab.row_begin:slab.row_end]).max() / np.abs(b_ref).max() <= 1e-12
        # ---- ... and with the interface rows handed over INSIDE the gather kernel (peer stores + counters): several
        # back-to-back assemblies without any host synchronisation exercise both buffer parities and the acknowledgements
        for forms in ([lap], [D.form(D.integrand(D.INT_LAPLACE, diffusion=0.75))]):
            p2p = parallel.HaloSlabAssembly(space, rank, world, p2p=True)
            p2p.append(forms[0])
            p2p.append_rhs(rhs)
            ref_v, _ = oracle.assemble(gd, D.SPACE_CG, 1, rp, ci, forms)
            for _ in range(5):
                pv, pb = p2p.assemble_device()
            p2p.check()
            pv, pb = pv.cpu().numpy(), pb.cpu().numpy()
            ref_seg = ref_v[slab.value_offset:slab.value_offset + slab.nnz_local]
            assert np.abs(pv[:p2p.nnz_owned] - ref_seg).max() / np.abs(ref_v).max() <= 1e-12
            assert np.abs(pb[:p2p.rows_owned] - b_ref[slab.row_begin:slab.row_end]).max() / np.abs(b_ref).max() <= 1e-12
            p2p.close()
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_two_gpu_fv_halo_and_slab_assembly(tmp_path, oracle, gdt):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    world = 2
    mp.spawn(_worker, args=(world, free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))
