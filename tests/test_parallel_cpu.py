"""Host-side logic of the multi-GPU path on CPU: world_size 2 and 3 `gloo` process groups.

The slab partition, the neighbour topology and the ghost-layer exchange are checked against the oracle: every rank runs
the oracle's FV apply on its ghosted slab (as a small non-periodic grid) and the owned layers must reproduce the
oracle's result on the whole grid.  The CUDA kernels themselves are covered by the `-m gpu` tests."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, periodic, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist

    import oracle
    from dune_gdt_b200 import descriptors as D
    from dune_gdt_b200 import parallel

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        d = len(n)
        plane = int(np.prod(n[:-1])) if d > 1 else 1
        rng = np.random.default_rng(20251017)
        u = rng.random(int(np.prod(n)))
        begin, end = parallel.slab_layers(n[-1], rank, world)
        per_last = bool(periodic & (1 << (d - 1))) and n[-1] > 1
        local = np.zeros((end - begin + 2) * plane)
        local[plane:-plane] = u[begin * plane:end * plane]
        t = torch.from_numpy(local)
        for r in parallel.exchange_ghost_layers(t, plane, rank, world, per_last):
            r.wait()
        # expected ghost content
        lower, upper = parallel.neighbours(rank, world, per_last)
        if lower is not None:
            jb = (begin - 1) % n[-1]
            assert np.array_equal(local[:plane], u[jb * plane:(jb + 1) * plane])
        if upper is not None:
            ja = end % n[-1]
            assert np.array_equal(local[-plane:], u[ja * plane:(ja + 1) * plane])
        # oracle on the ghosted slab == oracle on the whole grid (owned layers)
        lo, up = [0.0] * d, [1.0] * d
        h_last = 1.0 / n[-1]
        fl = D.flux(D.FLUX_LINEAR, D.NUMFLUX_UPWIND, [1.0, 0.5, -0.75][:d])
        full = oracle.fv_apply(D.grid_desc(lo, up, list(n), periodic=periodic), fl, u)
        has_lo, has_hi = lower is not None, upper is not None
        j0, j1 = begin - (1 if has_lo else 0), end + (1 if has_hi else 0)
        sub_n = list(n[:-1]) + [j1 - j0]
        sub_lo, sub_up = list(lo), list(up)
        sub_lo[-1], sub_up[-1] = j0 * h_last, j1 * h_last
        sub_u = local[(0 if has_lo else plane):(len(local) if has_hi else len(local) - plane)]
        sub = oracle.fv_apply(D.grid_desc(sub_lo, sub_up, sub_n, periodic=periodic & ~(1 << (d - 1))), fl, sub_u)
        own = sub[(plane if has_lo else 0):(plane if has_lo else 0) + (end - begin) * plane]
        ref = full[begin * plane:end * plane]
        err = np.abs(own - ref).max() / np.abs(full).max()
        assert err <= 1e-12, err
        # all ranks together cover every layer exactly once
        counts = torch.zeros(n[-1], dtype=torch.int64)
        counts[begin:end] += 1
        dist.all_reduce(counts)
        assert bool((counts == 1).all())
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,periodic", [(2, [16, 12], 3), (2, [16, 12], 0), (3, [6, 5, 10], 7), (3, [40], 1),
                                              (2, [8, 2], 3)])
def test_slab_partition_and_ghost_exchange_gloo(tmp_path, oracle, world, n, periodic):
    import torch.multiprocessing as mp

    port = free_port()
    mp.spawn(_worker, args=(world, port, n, periodic, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


def test_slab_layers_properties():
    sys.path.insert(0, ROOT)
    from dune_gdt_b200 import capi, parallel

    for n_last in (1, 7, 256, 257):
        for world in (1, 2, 3, 8):
            if n_last < world:
                with pytest.raises(capi.WrongInputGiven):
                    parallel.slab_layers(n_last, 0, world)
                continue
            cuts = [parallel.slab_layers(n_last, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n_last
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
            sizes = [e - b for b, e in cuts]
            assert max(sizes) - min(sizes) <= 1 and min(sizes) >= 1
    assert parallel.neighbours(0, 4, False) == (None, 1)
    assert parallel.neighbours(3, 4, False) == (2, None)
    assert parallel.neighbours(0, 4, True) == (3, 1)
    assert parallel.neighbours(0, 1, True) == (0, 0)


def _halo_worker(rank, world, port, n, out_dir):
    """interface-row halo of the element-partitioned assembly (SURVEY.md 8e scheme 2): every rank assembles its OWN
    element layers only (oracle with kappa = 0 outside the slab), holds the rows of the vertex layers [begin, end],
    sends the partial interface layer up and adds what arrives from below; owned rows == the global matrix / vector"""
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    import oracle
    from dune_gdt_b200 import descriptors as D
    from dune_gdt_b200 import parallel

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        d = len(n)
        g = D.grid_desc(-1.0, 1.0, list(n))
        plane_e = int(np.prod(n[:-1])) if d > 1 else 1
        plane_v = int(np.prod([k + 1 for k in n[:-1]])) if d > 1 else 1
        kap = np.random.default_rng(20251017).uniform(0.5, 2.0, int(np.prod(n)))
        rp, ci = oracle.pattern(g, (D.SPACE_CG, 1))
        f_src = np.random.default_rng(7).uniform(-1.0, 1.0, kap.size)
        lap = lambda k: D.form(D.integrand(D.INT_LAPLACE, diffusion=D.fn_elem(k)))  # noqa: E731
        rhs = lambda w: D.form(D.integrand(D.INT_PRODUCT, diffusion=1.0, weight=D.fn_elem(w)))  # noqa: E731
        v_ref, b_ref = oracle.assemble(g, D.SPACE_CG, 1, rp, ci, [lap(kap)], rhs_forms=[rhs(f_src)])
        begin, end = parallel.slab_layers(n[-1], rank, world)
        mask = np.zeros_like(kap)
        mask[begin * plane_e:end * plane_e] = 1.0
        v_part, b_part = oracle.assemble(g, D.SPACE_CG, 1, rp, ci, [lap(kap * mask)], rhs_forms=[rhs(f_src * mask)])
        r0, r1 = begin * plane_v, (end + 1) * plane_v  # rows of the vertex layers [begin, end]
        values = torch.from_numpy(v_part[rp[r0]:rp[r1]].copy())
        vector = torch.from_numpy(b_part[r0:r1].copy())
        has_lo, has_hi = rank > 0, rank < world - 1
        # interface layers are interior along the last direction: the one received and the one sent have the same shape
        layer_nnz = int(rp[r0 + plane_v] - rp[r0]) if has_lo else int(rp[r1] - rp[r1 - plane_v])
        if has_lo and has_hi:
            assert layer_nnz == rp[r1] - rp[r1 - plane_v]
        parallel.exchange_interface_rows(values, 0 if has_lo else -1, values.numel() - layer_nnz if has_hi else -1,
                                         layer_nnz, rank, world)
        parallel.exchange_interface_rows(vector, 0 if has_lo else -1, vector.numel() - plane_v if has_hi else -1,
                                         plane_v, rank, world)
        own_rows = (end - begin + (0 if has_hi else 1)) * plane_v
        own_nnz = int(rp[r0 + own_rows] - rp[r0])
        err_v = np.abs(values.numpy()[:own_nnz] - v_ref[rp[r0]:rp[r0] + own_nnz]).max() / np.abs(v_ref).max()
        err_b = np.abs(vector.numpy()[:own_rows] - b_ref[r0:r0 + own_rows]).max() / np.abs(b_ref).max()
        assert err_v <= 1e-12 and err_b <= 1e-12, (err_v, err_b)
        total = torch.tensor([own_nnz, own_rows])
        dist.all_reduce(total)
        assert total.tolist() == [len(v_ref), len(b_ref)]
        open(os.path.join(out_dir, f"halo{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, [5, 4, 6]), (3, [6, 7]), (3, [4, 3, 9])])
def test_interface_row_halo_gloo(tmp_path, oracle, world, n):
    import torch.multiprocessing as mp

    mp.spawn(_halo_worker, args=(world, free_port(), n, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"halo{r}").exists() for r in range(world))
