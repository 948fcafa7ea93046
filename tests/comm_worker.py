"""Worker of tests/test_comm_gpu.py: rank r of a WORLD_SIZE-rank job on GPU r. Exchanges the NCCL unique id over a gloo
process group, all-gathers a rank-stamped block of states through jgb_allgather_states and checks every block."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jgb200  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(rank)
    ctx = jgb200.Context(rank)
    jgb200.dist.comm_init(ctx)
    rows, n = 96, 257
    dev = torch.device("cuda", rank)
    vm = torch.full((rows, n), float(rank + 1), dtype=torch.float64, device=dev) + torch.arange(n, device=dev) * 1e-3
    va = -vm
    it = torch.full((rows,), 10 + rank, dtype=torch.int32, device=dev)
    st = torch.full((rows,), rank, dtype=torch.int8, device=dev)
    torch.cuda.synchronize()
    for rep in range(2):                               # the second call waits for the first one
        vm_all, va_all, it_all, st_all = jgb200.dist.allgather_states(ctx, vm, va, it, st)
    jgb200.dist.comm_wait(ctx)
    for r in range(world):
        blk = slice(r * rows, (r + 1) * rows)
        assert torch.equal(vm_all[blk], torch.full((rows, n), float(r + 1), dtype=torch.float64, device=dev)
                           + torch.arange(n, device=dev) * 1e-3)
        assert torch.equal(va_all[blk], -vm_all[blk])
        assert bool((it_all[blk] == 10 + r).all()) and bool((st_all[blk] == r).all())
    # partial calls: states only
    out = jgb200.dist.allgather_states(ctx, vm, None, None, None)
    jgb200.dist.comm_wait(ctx, host_blocking=False)
    ctx.synchronize()
    assert out[1] is None and float(out[0][(world - 1) * rows, 0]) == float(world)
    print(f"rank {rank} of {world}: all-gather ok", flush=True)
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
