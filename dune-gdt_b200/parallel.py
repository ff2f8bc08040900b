"""One process per GPU: element-slab partition of the structured grid and the ghost-layer exchange of the FV path.

Reference: the only halo site of this path is the stage-vector exchange of the explicit Runge-Kutta stepper
(dune/gdt/tools/timestepper/explicit-rungekutta.hh:252-257, a DUNE `communicate()` with a data handle on an overlapping
YaspGrid).  Here the grid is cut into slabs of element layers along the LAST direction; a rank's FV vector is
[ghost layer below | owned layers | ghost layer above] and before every operator apply the first / last owned layer
travels to the neighbours' ghost layers (periodic wrap: rank 0 <-> rank N-1) with grouped point-to-point messages
(`torch.distributed` P2P: NCCL send/recv over NVLink on GPUs, gloo on CPU for the host-logic tests).

Assembly has two partitions.  `SlabAssembly`: rows are owned by the slab that owns the vertex layer and
the one element layer below a slab is recomputed locally (what the reference does with YaspGrid's overlap): no exchange.
`HaloSlabAssembly`: every rank walks only its own elements, the partial sums of the interface rows (one vertex layer of
CSR values + right-hand-side entries per slab face) travel to the owner with NCCL send / recv and are added there.
"""
import ctypes as C

import numpy as np

from . import capi
from . import descriptors as D


def slab_layers(n_last, rank, world):
    """element layers [begin, end) of `rank`: contiguous, sizes differ by at most one (lower ranks get the extra layer)"""
    if world < 1 or not 0 <= rank < world:
        raise capi.WrongInputGiven("slab_layers: need 0 <= rank < world")
    if n_last < world:
        raise capi.WrongInputGiven(f"cannot cut {n_last} element layers into {world} non-empty slabs")
    base, extra = divmod(n_last, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def neighbours(rank, world, periodic_last):
    """(rank owning the layer below my first layer, rank owning the layer above my last layer); None at a domain
    boundary without periodicity.  With world == 1 and a periodic direction both neighbours are the rank itself."""
    lower = rank - 1 if rank > 0 else (world - 1 if periodic_last else None)
    upper = rank + 1 if rank < world - 1 else (0 if periodic_last else None)
    return lower, upper


def exchange_ghost_layers(u_local, plane, rank, world, periodic_last, group=None):
    """fills the ghost layers of `u_local` (1D torch tensor, layout [ghost | owned | ghost], `plane` cells per layer)
    with the neighbours' boundary layers.  Returns the list of outstanding requests (call .wait() on each)."""
    import torch.distributed as dist

    lower, upper = neighbours(rank, world, periodic_last)
    first_owned = u_local[plane:2 * plane]
    last_owned = u_local[-2 * plane:-plane]
    ghost_lo = u_local[:plane]
    ghost_hi = u_local[-plane:]
    if world == 1:
        if periodic_last:
            ghost_lo.copy_(last_owned)
            ghost_hi.copy_(first_owned)
        return []
    ops = []
    # tags are not supported by NCCL P2P; message order per pair is fixed instead: "upward" messages first
    if upper is not None:
        ops.append(dist.P2POp(dist.isend, last_owned, upper, group))
    if lower is not None:
        ops.append(dist.P2POp(dist.irecv, ghost_lo, lower, group))
    if lower is not None:
        ops.append(dist.P2POp(dist.isend, first_owned, lower, group))
    if upper is not None:
        ops.append(dist.P2POp(dist.irecv, ghost_hi, upper, group))
    if world == 2 and periodic_last:
        # both neighbours are the same rank: keep the two message pairs apart (send/recv order must match on both sides)
        reqs = dist.batch_isend_irecv(ops[:2])
        for r in reqs:
            r.wait()
        return dist.batch_isend_irecv(ops[2:])
    return dist.batch_isend_irecv(ops) if ops else []


class DistributedAdvectionFvOperator:
    """AdvectionFvOperator (dune/gdt/operators/advection-fv.hh:44-141) on the slab of this rank.

    `apply` / `euler_step` work on device vectors in slab layout (torch float64 CUDA tensors of `local_size` entries).
    The interior layers are computed while the ghost layers are in flight; the first and last owned layer follow."""

    def __init__(self, numerical_flux, space, rank, world, group=None):
        from .api import AdvectionFvOperator

        self.space = space
        self.rank, self.world, self.group = rank, world, group
        g = space.grid.desc
        self.dim = int(g.dim)
        self.n_last = int(g.n[self.dim - 1])
        self.periodic_last = bool(g.periodic & (1 << (self.dim - 1))) and self.n_last > 1
        self.begin, self.end = slab_layers(self.n_last, rank, world)
        self.op = AdvectionFvOperator(numerical_flux, space)
        capi.check(capi.lib().gdtb_fvop_set_slab(self.op._h, self.begin, self.end))
        self.plane = int(capi.lib().gdtb_fvop_ghost_layer_size(self.op._h))
        self.owned = (self.end - self.begin) * self.plane
        self.local_size = self.owned + 2 * self.plane
        self._comm_stream = None
        self._compute_stream = None

    # ---- layout helpers -------------------------------------------------------------------------------------
    def scatter_from_global(self, u_global):
        """owned part of a global host vector in slab layout (ghost layers zero)"""
        out = np.zeros(self.local_size)
        out[self.plane:self.plane + self.owned] = np.asarray(u_global)[self.begin * self.plane:self.end * self.plane]
        return out

    def owned_view(self, u_local):
        return u_local[self.plane:self.plane + self.owned]

    # ---- operator ---------------------------------------------------------------------------------------------
    def _step(self, src, dst, euler, dt, lo, hi):
        capi.check(capi.lib().gdtb_fvop_step_async(self.op._h, C.c_void_p(src.data_ptr()), C.c_void_p(dst.data_ptr()),
                                                   int(euler), float(dt), int(lo), int(hi)))

    def _run(self, src, dst, euler, dt):
        import torch

        ctx = self.space.grid.ctx
        caller = torch.cuda.current_stream()
        if self._comm_stream is None:
            # the library launches on an explicit (non-default) stream: handle 0 would select its own stream
            self._compute_stream = torch.cuda.Stream()
            self._comm_stream = torch.cuda.Stream()
        compute, comm = self._compute_stream, self._comm_stream
        # the context is shared with the caller's other operators: its stream is re-pointed for this call only
        previous = getattr(ctx, "stream_handle", None)
        ctx.set_stream(compute.cuda_stream)
        compute.wait_stream(caller)  # src is complete on the caller's stream
        comm.wait_stream(caller)
        comm.wait_stream(compute)
        with torch.cuda.stream(comm):
            reqs = exchange_ghost_layers(src, self.plane, self.rank, self.world, self.periodic_last, self.group)
            for r in reqs:
                r.wait()  # stream-ordered: makes `comm` wait for the transfers, does not block the host
        lo, hi = self.begin, self.end
        if hi - lo > 2:
            self._step(src, dst, euler, dt, lo + 1, hi - 1)  # interior: independent of the ghost layers
            compute.wait_stream(comm)
            self._step(src, dst, euler, dt, lo, lo + 1)
            self._step(src, dst, euler, dt, hi - 1, hi)
        else:
            compute.wait_stream(comm)
            self._step(src, dst, euler, dt, lo, hi)
        caller.wait_stream(compute)
        src.record_stream(compute)
        dst.record_stream(compute)
        src.record_stream(comm)
        ctx.set_stream(previous)

    def apply(self, src, dst):
        """dst(owned) = L(src); fills src's ghost layers first"""
        self._run(src, dst, 0, 0.0)

    def euler_step(self, src, dst, dt):
        """dst(owned) = src - dt L(src) (examples/mpi_2019_02_talk_on_hyperbolic_equations.cc:152-157)"""
        self._run(src, dst, 1, dt)


def make_distributed_advection_fv_operator(numerical_flux, space, rank, world, group=None):
    return DistributedAdvectionFvOperator(numerical_flux, space, rank, world, group)


class PeerMemoryFvTimeLoop:
    """The explicit Euler loop of examples/mpi_2019_02_talk_on_hyperbolic_equations.cc:141-159 on slabs WITHOUT a
    host-launched collective per step: the apply kernel stores its first / last owned layer straight into the neighbour
    GPUs' ghost layers (NVLink peer stores into buffers opened through CUDA IPC) and raises a step counter there; the
    blocks that read a ghost layer wait for the neighbour's counter of the previous step (`gdtb_fvop_p2p_*`).  One
    kernel launch per step and rank; NCCL is only used to pass the IPC handles around and for the initial ghost fill."""

    def __init__(self, numerical_flux, space, rank, world, group=None):
        import torch
        import torch.distributed as dist

        from .api import AdvectionFvOperator

        lib = capi.lib()
        self.space, self.rank, self.world, self.group = space, rank, world, group
        g = space.grid.desc
        self.dim = int(g.dim)
        self.n_last = int(g.n[self.dim - 1])
        self.periodic_last = bool(g.periodic & (1 << (self.dim - 1))) and self.n_last > 1
        self.begin, self.end = slab_layers(self.n_last, rank, world)
        self.op = AdvectionFvOperator(numerical_flux, space)
        capi.check(lib.gdtb_fvop_set_slab(self.op._h, self.begin, self.end))
        self.plane = int(lib.gdtb_fvop_ghost_layer_size(self.op._h))
        self.owned = (self.end - self.begin) * self.plane
        self.local_size = self.owned + 2 * self.plane
        handles = (C.c_ubyte * (3 * 64))()
        p0, p1 = C.c_void_p(), C.c_void_p()
        capi.check(lib.gdtb_fvop_p2p_alloc(self.op._h, C.byref(p0), C.byref(p1), handles))
        dev = torch.device("cuda", space.grid.ctx.device)
        self.u = [_as_tensor(p0.value, self.local_size, dev), _as_tensor(p1.value, self.local_size, dev)]
        lower, upper = neighbours(rank, world, self.periodic_last)
        mine = torch.tensor(list(bytes(handles)), dtype=torch.uint8, device=dev)
        if world > 1:
            every = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(every, mine, group=group)
        else:
            every = [mine]

        def arg(nb):
            if nb is None:
                return None, 0, 0
            if nb == rank:
                return None, 0, 1
            b, e = slab_layers(self.n_last, nb, world)
            raw = bytes(every[nb].cpu().numpy().tobytes())
            return (C.c_ubyte * len(raw)).from_buffer_copy(raw), e - b, 0

        lo_h, lo_layers, lo_self = arg(lower)
        hi_h, hi_layers, hi_self = arg(upper)
        self._keep = (lo_h, hi_h)
        capi.check(lib.gdtb_fvop_p2p_connect(self.op._h, lo_h, lo_layers, lo_self, hi_h, hi_layers, hi_self))
        if world > 1:
            dist.barrier(group=group)  # every rank has opened its neighbours' buffers

    def set_initial_values(self, u_global):
        """owned layers of a global host vector into u[0]; the ghost layers come from the neighbours once (NCCL)"""
        import torch

        loc = np.zeros(self.local_size)
        loc[self.plane:self.plane + self.owned] = np.asarray(u_global)[self.begin * self.plane:self.end * self.plane]
        self.space.grid.ctx.synchronize()
        self.u[0].copy_(torch.from_numpy(loc))
        for r in exchange_ghost_layers(self.u[0], self.plane, self.rank, self.world, self.periodic_last, self.group):
            r.wait()
        torch.cuda.synchronize()
        if self.world > 1:
            import torch.distributed as dist

            dist.barrier(group=self.group)

    def euler_steps(self, dt, n_steps):
        """n_steps times u <- u - dt L(u): one kernel launch per step, the ghost exchange happens inside the kernel"""
        lib = capi.lib()
        for _ in range(n_steps):
            capi.check(lib.gdtb_fvop_p2p_step(self.op._h, 1, float(dt)))

    def apply_steps(self, n_steps):
        lib = capi.lib()
        for _ in range(n_steps):
            capi.check(lib.gdtb_fvop_p2p_step(self.op._h, 0, 0.0))

    def check(self):
        capi.check(capi.lib().gdtb_fvop_p2p_check(self.op._h))

    def current(self):
        """device tensor (slab layout) that holds the current solution"""
        p, s = C.c_void_p(), C.c_int64()
        capi.check(capi.lib().gdtb_fvop_p2p_current(self.op._h, C.byref(p), C.byref(s)))
        return self.u[s.value & 1]

    def owned_view(self, u_local):
        return u_local[self.plane:self.plane + self.owned]

    def close(self):
        """collective: nobody may free its buffers while a neighbour can still store into them"""
        import torch

        self.space.grid.ctx.synchronize()
        torch.cuda.synchronize()
        if self.world > 1:
            import torch.distributed as dist

            dist.barrier(group=self.group)
        self.u = None
        self.op = None


class PeerMemoryRkTimeStepper:
    """ExplicitRungeKuttaTimeStepper (tools/timestepper/explicit-rungekutta.hh:158-270) on slabs: the reference exchanges
    every stage vector with a DataHandle communicate() (:252-257); here the stepper-owned solution / stage vectors are
    opened by the neighbour processes through CUDA IPC, a small kernel hands the boundary layers over after every stage
    vector and every operator apply waits inside the kernel for its source (`gdtb_rk_p2p_*`).  >= 2 stages."""

    def __init__(self, numerical_flux, space, rank, world, method, r=-1.0, t_0=0.0, group=None):
        import torch
        import torch.distributed as dist

        from .api import AdvectionFvOperator

        lib = capi.lib()
        self.space, self.rank, self.world, self.group = space, rank, world, group
        g = space.grid.desc
        self.dim = int(g.dim)
        self.n_last = int(g.n[self.dim - 1])
        self.periodic_last = bool(g.periodic & (1 << (self.dim - 1))) and self.n_last > 1
        self.begin, self.end = slab_layers(self.n_last, rank, world)
        self.op = AdvectionFvOperator(numerical_flux, space)
        capi.check(lib.gdtb_fvop_set_slab(self.op._h, self.begin, self.end))
        self.plane = int(lib.gdtb_fvop_ghost_layer_size(self.op._h))
        self.owned = (self.end - self.begin) * self.plane
        self.local_size = self.owned + 2 * self.plane
        self._h = C.c_void_p()
        capi.check(lib.gdtb_rk_create(self.op._h, int(method), 0, None, None, None, float(r), float(t_0), C.byref(self._h)))
        handles = (C.c_ubyte * (4 * 64))()
        p0 = C.c_void_p()
        capi.check(lib.gdtb_rk_p2p_handles(self._h, C.byref(p0), handles))
        dev = torch.device("cuda", space.grid.ctx.device)
        self._ptr = p0.value
        self.u = _as_tensor(p0.value, self.local_size, dev)
        lower, upper = neighbours(rank, world, self.periodic_last)
        mine = torch.tensor(list(bytes(handles)), dtype=torch.uint8, device=dev)
        if world > 1:
            every = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(every, mine, group=group)
        else:
            every = [mine]

        def arg(nb):
            if nb is None:
                return None, 0, 0
            if nb == rank:
                return None, 0, 1
            b, e = slab_layers(self.n_last, nb, world)
            raw = bytes(every[nb].cpu().numpy().tobytes())
            return (C.c_ubyte * len(raw)).from_buffer_copy(raw), e - b, 0

        lo_h, lo_layers, lo_self = arg(lower)
        hi_h, hi_layers, hi_self = arg(upper)
        self._keep = (lo_h, hi_h)
        capi.check(lib.gdtb_rk_p2p_connect(self._h, lo_h, lo_layers, lo_self, hi_h, hi_layers, hi_self))
        if world > 1:
            dist.barrier(group=group)

    def set_initial_values(self, u_global):
        import torch

        loc = np.zeros(self.local_size)
        loc[self.plane:self.plane + self.owned] = np.asarray(u_global)[self.begin * self.plane:self.end * self.plane]
        self.space.grid.ctx.synchronize()
        self.u.copy_(torch.from_numpy(loc))
        for r in exchange_ghost_layers(self.u, self.plane, self.rank, self.world, self.periodic_last, self.group):
            r.wait()
        torch.cuda.synchronize()
        if self.world > 1:
            import torch.distributed as dist

            dist.barrier(group=self.group)

    def step(self, dt, max_dt=None):
        ret = C.c_double()
        capi.check(capi.lib().gdtb_rk_step(self._h, C.c_void_p(self._ptr), float(dt), float(dt if max_dt is None else max_dt),
                                           C.byref(ret)))
        return ret.value

    def solve(self, t_end, initial_dt):
        n, nxt = C.c_int64(), C.c_double()
        capi.check(capi.lib().gdtb_rk_solve(self._h, C.c_void_p(self._ptr), float(t_end), float(initial_dt), C.byref(n),
                                            C.byref(nxt)))
        self.num_steps = n.value
        return nxt.value

    def current_time(self):
        return capi.lib().gdtb_rk_current_time(self._h)

    def check(self):
        capi.check(capi.lib().gdtb_rk_p2p_check(self._h))

    def owned_view(self):
        return self.u[self.plane:self.plane + self.owned]

    def close(self):
        """collective: nobody may free its buffers while a neighbour can still store into them"""
        import torch

        self.space.grid.ctx.synchronize()
        torch.cuda.synchronize()
        if self.world > 1:
            import torch.distributed as dist

            dist.barrier(group=self.group)
        self.u = None
        if self._h.value:
            capi.lib().gdtb_rk_destroy(self._h)
            self._h = C.c_void_p()
        self.op = None


class SlabAssembly:
    """Matrix operator + functional of one rank: owner-computes-rows on the rank's element slab, no communication.
    The global CSR matrix is the concatenation of the ranks' value arrays in rank order (`value_offset` is the
    global position of the first local value, `row_begin/row_end` the global row range)."""

    def __init__(self, space, rank, world, with_functional=True):
        lib = capi.lib()
        g = space.grid.desc
        d = int(g.dim)
        self.begin, self.end = slab_layers(int(g.n[d - 1]), rank, world)
        self.space = space
        self.op_h, self.fun_h = C.c_void_p(), C.c_void_p()
        ctx = space.grid.ctx
        capi.check(lib.gdtb_matop_create(ctx._h, space._h, space._h, None, C.byref(self.op_h)))
        capi.check(lib.gdtb_matop_set_slab(self.op_h, self.begin, self.end))
        if with_functional:
            capi.check(lib.gdtb_vecfun_create(ctx._h, space._h, C.byref(self.fun_h)))
            capi.check(lib.gdtb_vecfun_set_slab(self.fun_h, self.begin, self.end))
        rb, re_, vo = C.c_int64(), C.c_int64(), C.c_int64()
        capi.check(lib.gdtb_matop_local_rows(self.op_h, C.byref(rb), C.byref(re_), C.byref(vo)))
        self.row_begin, self.row_end, self.value_offset = rb.value, re_.value, vo.value
        self.nnz_local = int(lib.gdtb_matop_local_nnz(self.op_h))
        # CG Q2: one (row range, global value offset) per sub-entity group of the MCMG numbering; Q1: a single range
        n = C.c_int32()
        capi.check(lib.gdtb_matop_local_row_ranges(self.op_h, 0, None, None, None, None, C.byref(n)))
        rbs, res, vos, cnt = (np.zeros(n.value, dtype=np.int64) for _ in range(4))
        i64p = lambda a: a.ctypes.data_as(C.POINTER(C.c_int64))  # noqa: E731
        capi.check(lib.gdtb_matop_local_row_ranges(self.op_h, n.value, i64p(rbs), i64p(res), i64p(vos), i64p(cnt), C.byref(n)))
        # (global row begin, global row end, global CSR offset of the first value, number of values) per range
        self.row_ranges = [tuple(int(x) for x in t) for t in zip(rbs, res, vos, cnt)]

    def append(self, form):
        capi.check(capi.lib().gdtb_matop_append_element(self.op_h, C.byref(form)))

    def append_coupling(self, form, filter=D.FILTER_INNER_ONCE):
        """DG spaces: inner-intersection forms (the neighbour across a slab face enters through its index, geometry and
        coefficients only -- no exchange)"""
        capi.check(capi.lib().gdtb_matop_append_coupling(self.op_h, C.byref(form), filter))

    def append_boundary(self, form):
        capi.check(capi.lib().gdtb_matop_append_boundary(self.op_h, C.byref(form), D.FILTER_ALL_BOUNDARY))

    def append_rhs(self, form):
        capi.check(capi.lib().gdtb_vecfun_append_element(self.fun_h, C.byref(form)))

    def assemble(self):
        values = np.empty(self.nnz_local)
        if not self.fun_h.value:
            capi.check(capi.lib().gdtb_assemble_host(self.op_h, None, capi.dptr(values), None))
            return values, None
        vector = np.empty(sum(re_ - rb for rb, re_, _, _ in self.row_ranges))  # the owned rows, range after range
        capi.check(capi.lib().gdtb_assemble_host(self.op_h, self.fun_h, capi.dptr(values), capi.dptr(vector)))
        return values, vector

    def assemble_device(self):
        """enqueue the assembly on the context's stream (values stay on the device)"""
        capi.check(capi.lib().gdtb_assemble_async(self.op_h, self.fun_h if self.fun_h.value else None, D.ASSEMBLE_OVERWRITE))

    def scatter_into_global(self, values_local, values_global):
        """place this rank's value segments (stored back to back) at their global CSR positions"""
        at = 0
        for _, _, offset, count in self.row_ranges:
            values_global[offset:offset + count] = values_local[at:at + count]
            at += count

    def __del__(self):
        if getattr(self, "op_h", None) and self.op_h.value:
            capi.lib().gdtb_matop_destroy(self.op_h)
            if self.fun_h.value:
                capi.lib().gdtb_vecfun_destroy(self.fun_h)


def exchange_interface_rows(local, recv_offset, send_offset, count, rank, world, add=None, group=None,
                            stream_ordered=False):
    """interface-row halo of the element-partitioned assembly: the last `count` entries at `send_offset` of `local`
    (the partial sums of the interface layer owned by rank + 1) are sent up, the layer arriving from rank - 1 is added
    to the `count` entries at `recv_offset`.  `local` is a 1D torch tensor (CUDA for NCCL, CPU for gloo); `add(y, x)`
    performs y += x (default: torch).  `stream_ordered`: the producer of `local`, the transfer and `add` all run on
    torch's current stream (the library context was given that stream): nothing synchronises with the host.  Returns the
    received buffer (None on rank 0)."""
    import torch
    import torch.distributed as dist

    ops, recv = [], None
    if send_offset >= 0 and rank < world - 1:
        ops.append(dist.P2POp(dist.isend, local[send_offset:send_offset + count], rank + 1, group))
    if recv_offset >= 0 and rank > 0:
        recv = torch.empty(count, dtype=local.dtype, device=local.device)
        ops.append(dist.P2POp(dist.irecv, recv, rank - 1, group))
    if ops:
        for r in dist.batch_isend_irecv(ops):
            r.wait()  # NCCL: orders torch's current stream after the transfer (no host synchronisation); gloo: blocks
    if local.is_cuda and not stream_ordered:
        # the add may run on another stream than torch's current one
        torch.cuda.current_stream(local.device).synchronize()
    if recv is not None:
        target = local[recv_offset:recv_offset + count]
        if add is None:
            target += recv
        else:
            add(target, recv)
    return recv


class HaloSlabAssembly:
    """Element-partitioned assembly with an interface-row halo (CG Q1): each rank walks its own element layers only;
    the rows of the interface layer it shares with the rank above are partial and are completed on their owner by
    one NCCL message per slab face.  After `assemble_device()` the owned rows of this rank are the first
    `nnz_owned` values / `rows_owned` vector entries of the device buffers (same layout as `SlabAssembly`)."""

    def __init__(self, space, rank, world, group=None, p2p=False):
        """p2p: the interface rows travel INSIDE the gather kernel (NVLink peer stores into the neighbour's receive buffer
        + counters, gdtb_halo_p2p_*) instead of one NCCL message per slab face followed by an add kernel"""
        lib = capi.lib()
        g = space.grid.desc
        d = int(g.dim)
        self.rank, self.world, self.group, self.p2p = rank, world, group, bool(p2p) and world > 1
        self.begin, self.end = slab_layers(int(g.n[d - 1]), rank, world)
        self.space = space
        self.op_h, self.fun_h = C.c_void_p(), C.c_void_p()
        ctx = space.grid.ctx
        self.ctx = ctx
        capi.check(lib.gdtb_matop_create(ctx._h, space._h, space._h, None, C.byref(self.op_h)))
        capi.check(lib.gdtb_vecfun_create(ctx._h, space._h, C.byref(self.fun_h)))
        capi.check(lib.gdtb_matop_set_slab_halo(self.op_h, self.begin, self.end))
        capi.check(lib.gdtb_vecfun_set_slab_halo(self.fun_h, self.begin, self.end))
        if self.p2p:
            import torch
            import torch.distributed as dist

            handles = (C.c_ubyte * (2 * 64))()
            capi.check(lib.gdtb_halo_p2p_alloc(self.op_h, handles))
            dev = torch.device("cuda", ctx.device)
            mine = torch.tensor(list(bytes(handles)), dtype=torch.uint8, device=dev)
            every = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(every, mine, group=group)

            def raw(nb):
                if nb < 0 or nb >= world:
                    return None
                b = bytes(every[nb].cpu().numpy().tobytes())
                return (C.c_ubyte * len(b)).from_buffer_copy(b)

            lo_h, hi_h = raw(rank - 1), raw(rank + 1)
            lo_layers = 0
            if rank > 0:
                b, e = slab_layers(int(g.n[d - 1]), rank - 1, world)
                lo_layers = e - b
            self._keep = (lo_h, hi_h)
            capi.check(lib.gdtb_halo_p2p_connect(self.op_h, lo_h, lo_layers, hi_h))
            dist.barrier(group=group)  # every rank has opened its neighbours' buffers
        rb, re_, vo = C.c_int64(), C.c_int64(), C.c_int64()
        capi.check(lib.gdtb_matop_local_rows(self.op_h, C.byref(rb), C.byref(re_), C.byref(vo)))
        self.row_begin, self.value_offset = rb.value, vo.value
        self.nnz_local = int(lib.gdtb_matop_local_nnz(self.op_h))
        self.rows_local = re_.value - rb.value
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        capi.check(lib.gdtb_matop_halo_layout(self.op_h, C.byref(a), C.byref(b), C.byref(c)))
        self.mat_layout = (a.value, b.value, c.value)
        capi.check(lib.gdtb_vecfun_halo_layout(self.fun_h, C.byref(a), C.byref(b), C.byref(c)))
        self.vec_layout = (a.value, b.value, c.value)
        # the interface layer at the top belongs to the rank above: not part of the owned rows
        top = rank < world - 1
        self.nnz_owned = self.nnz_local - (self.mat_layout[2] if top else 0)
        self.rows_owned = self.rows_local - (self.vec_layout[2] if top else 0)

    def append(self, form):
        capi.check(capi.lib().gdtb_matop_append_element(self.op_h, C.byref(form)))

    def append_rhs(self, form):
        capi.check(capi.lib().gdtb_vecfun_append_element(self.fun_h, C.byref(form)))

    def _device_views(self):
        import torch

        lib = capi.lib()
        pv, pb = C.c_void_p(), C.c_void_p()
        capi.check(lib.gdtb_matop_values_device(self.op_h, C.byref(pv)))
        capi.check(lib.gdtb_vecfun_device(self.fun_h, C.byref(pb)))
        dev = torch.device("cuda", self.ctx.device)
        values = _as_tensor(pv.value, self.nnz_local, dev)
        vector = _as_tensor(pb.value, self.rows_local, dev)
        return values, vector

    def assemble_device(self):
        """one walk over the own elements + the interface-row exchange; returns (values, vector) device tensors that
        alias the library's buffers (owned rows first)"""
        import torch

        lib = capi.lib()
        if self.p2p:  # the exchange happens inside the kernel: nothing follows the walk
            capi.check(lib.gdtb_assemble_async(self.op_h, self.fun_h, D.ASSEMBLE_OVERWRITE))
            return self._device_views()
        # stream-ordered when the library launches on torch's current stream (ctx.set_stream(torch stream)): the walk,
        # the NCCL transfer and the add kernels queue up behind each other without a host round trip
        ordered = self.ctx.stream_handle is not None and self.ctx.stream_handle == torch.cuda.current_stream().cuda_stream
        capi.check(lib.gdtb_assemble_async(self.op_h, self.fun_h, D.ASSEMBLE_OVERWRITE))
        if not ordered:
            self.ctx.synchronize()  # the exchange runs on torch's stream
        values, vector = self._device_views()

        def add(y, x):
            capi.check(lib.gdtb_vector_add(self.ctx._h, C.c_void_p(y.data_ptr()), C.c_void_p(x.data_ptr()), y.numel()))

        keep = [exchange_interface_rows(values, *self.mat_layout, self.rank, self.world, add, self.group, ordered),
                exchange_interface_rows(vector, *self.vec_layout, self.rank, self.world, add, self.group, ordered)]
        for t in keep:
            if t is not None and ordered:
                t.record_stream(torch.cuda.current_stream())  # the add kernels read the receive buffers asynchronously
        if not ordered:
            torch.cuda.synchronize()  # received layers are complete before the add kernels read them
            self.ctx.synchronize()
        del keep
        return values, vector

    def check(self):
        """peer-memory mode: synchronise and report a timed-out wait"""
        if self.p2p:
            capi.check(capi.lib().gdtb_halo_p2p_check(self.op_h))

    def close(self):
        """collective (peer-memory mode): nobody frees its receive buffers while a neighbour can still store into them"""
        if self.p2p:
            import torch
            import torch.distributed as dist

            self.ctx.synchronize()
            torch.cuda.synchronize()
            dist.barrier(group=self.group)
        if self.op_h.value:
            capi.lib().gdtb_matop_destroy(self.op_h)
            capi.lib().gdtb_vecfun_destroy(self.fun_h)
            self.op_h = C.c_void_p()

    def __del__(self):
        if getattr(self, "op_h", None) and self.op_h.value:
            capi.lib().gdtb_matop_destroy(self.op_h)
            capi.lib().gdtb_vecfun_destroy(self.fun_h)


def _as_tensor(ptr, n, device):
    """torch view of `n` doubles of library-owned device memory (no copy)"""
    import torch

    class _Arr:
        __cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 3}

    return torch.as_tensor(_Arr(), device=device)
