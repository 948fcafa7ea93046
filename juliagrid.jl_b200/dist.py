"""Scenario sharding across GPUs (one process per GPU) and the single all-gather of converged states.

The reference has no distributed code; contingencies / Monte-Carlo draws are independent user-loop iterations
(SURVEY.md §3.4-3.5). Every rank runs the same deterministic setup (no broadcast), solves its contiguous block of
scenarios, and one all-gather over NCCL (NVLink/NVSwitch) — gloo on CPU in the tests — assembles the global result.
`torch.distributed` is plumbing only: no collective sits on the numerical path.
"""
from __future__ import annotations

import numpy as np


def shard_bounds(total: int, rank: int, world: int):
    """Contiguous block [lo, hi) of `total` scenarios owned by `rank`; sizes differ by at most one."""
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard(items, rank: int, world: int):
    lo, hi = shard_bounds(len(items), rank, world)
    return items[lo:hi]


def allgather_rows(local, total_rows: int | None = None, group=None):
    """All-gather row blocks of a 2-D (or 1-D) tensor whose per-rank row counts follow shard_bounds().
    `local` is a torch tensor (CUDA for NCCL, CPU for gloo). Returns the concatenation in rank order."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if total_rows is None:
        counts = [torch.zeros(1, dtype=torch.int64, device=local.device) for _ in range(world)]
        dist.all_gather(counts, torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device), group=group)
        sizes = [int(c.item()) for c in counts]
    else:
        sizes = [shard_bounds(total_rows, r, world)[1] - shard_bounds(total_rows, r, world)[0] for r in range(world)]
    assert sizes[rank] == local.shape[0]
    if len(set(sizes)) == 1:
        out = torch.empty((world * sizes[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    pad = max(sizes)
    buf = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    buf[: local.shape[0]] = local
    out = torch.empty((world * pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, buf, group=group)
    return torch.cat([out[r * pad: r * pad + sizes[r]] for r in range(world)], dim=0)


def gather_batch_result(vm, va, iterations, status, total_rows=None, group=None):
    """The one collective of the sweep: converged states [Vm | Va] (FP64), iteration counts and status bytes."""
    import torch
    state = allgather_rows(torch.cat([vm, va], dim=1), total_rows, group)
    n = vm.shape[1]
    meta = allgather_rows(torch.stack([iterations.to(torch.int32), status.to(torch.int32)], dim=1), total_rows, group)
    return state[:, :n], state[:, n:], meta[:, 0], meta[:, 1]


class PendingGather:
    """Handle of an all-gather in flight (equal block sizes on every rank): the collective runs on the backend's own
    stream, so the next batch can be solved while the converged states of this one travel over NVLink. `wait()` makes
    the current stream wait for it and returns (vm, va, iterations, status) of all ranks."""

    def __init__(self, works, state, meta, n, keep=()):
        self._works, self._state, self._meta, self._n, self._keep = works, state, meta, n, keep

    def wait(self):
        for w in self._works:
            w.wait()
        self._works, self._keep = [], ()
        s, m, n = self._state, self._meta, self._n
        return s[:, :n], s[:, n:], m[:, 0], m[:, 1]


def gather_batch_result_async(vm, va, iterations, status, group=None) -> PendingGather:
    """gather_batch_result without blocking the current stream; every rank must hold the same number of rows."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    src = torch.cat([vm, va], dim=1)                 # private copy: the caller may overwrite vm / va right away
    meta_src = torch.stack([iterations.to(torch.int32), status.to(torch.int32)], dim=1)
    state = torch.empty((world * src.shape[0], src.shape[1]), dtype=src.dtype, device=src.device)
    meta = torch.empty((world * meta_src.shape[0], 2), dtype=torch.int32, device=src.device)
    works = [dist.all_gather_into_tensor(state, src, group=group, async_op=True),
             dist.all_gather_into_tensor(meta, meta_src, group=group, async_op=True)]
    return PendingGather(works, state, meta, vm.shape[1], keep=(src, meta_src))


# ---- the library's own collective (jgb_comm_init / jgb_allgather_states, include/jgb200.h) ---------------------------
def comm_init(ctx, rank: int | None = None, world: int | None = None, group=None):
    """jgb_comm_init on every rank of the job. Rank 0 draws the NCCL unique id (jgb_comm_unique_id) and
    torch.distributed — any backend, it only carries 128 bytes of host data — hands it to the others. With world == 1
    no process group is needed."""
    import ctypes as C
    import torch.distributed as dist
    have_pg = dist.is_available() and dist.is_initialized()
    if rank is None:
        rank = dist.get_rank(group) if have_pg else 0
    if world is None:
        world = dist.get_world_size(group) if have_pg else 1
    ident = (C.c_uint8 * 128)()
    if rank == 0:
        rc = ctx.lib.jgb_comm_unique_id(ident)
        if rc != 0:
            raise RuntimeError("jgb_comm_unique_id failed: " + ctx.lib.jgb_last_error(None).decode())
    if world > 1:
        box = [bytes(ident)]
        dist.broadcast_object_list(box, src=0, group=group)
        ident = (C.c_uint8 * 128).from_buffer_copy(box[0])
    ctx.check(ctx.lib.jgb_comm_init(ctx.handle, rank, world, ident))
    return rank, world


def allgather_states(ctx, vm, va, iterations, status, out=None):
    """jgb_allgather_states on torch CUDA tensors: vm / va [rows][n] float64, iterations int32, status int8 (any may be
    None). Returns (vm_all, va_all, iterations_all, status_all) of shape [world * rows, ...]; the collective is in flight
    on the library's private stream — call comm_wait(ctx) (or the next allgather_states) before touching the buffers."""
    import ctypes as C
    import torch
    r, w = C.c_int32(0), C.c_int32(0)
    ctx.check(ctx.lib.jgb_comm_size(ctx.handle, C.byref(r), C.byref(w)))
    world = w.value
    first = next(t for t in (vm, va, iterations, status) if t is not None)
    rows = first.shape[0]
    n = vm.shape[1] if vm is not None else (va.shape[1] if va is not None else 1)
    if out is None:
        out = tuple(None if t is None else torch.empty((world * rows,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
                    for t in (vm, va, iterations, status))
    P = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
    ctx.check(ctx.lib.jgb_allgather_states(ctx.handle, rows, n, P(vm), P(va), P(iterations), P(status), P(out[0]),
                                           P(out[1]), P(out[2]), P(out[3])))
    return out


def comm_wait(ctx, host_blocking: bool = True):
    ctx.check(ctx.lib.jgb_comm_wait(ctx.handle, 1 if host_blocking else 0))
