"""jgb_comm_init / jgb_allgather_states: the library's own NCCL all-gather of converged states."""
import os
import subprocess
import sys

import pytest

import jgb200
from conftest import ROOT

pytestmark = pytest.mark.gpu


def _run(world):
    port = 29500 + os.getpid() % 2000
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "comm_worker.py")], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "all-gather ok" in o


def test_allgather_single_rank():
    """One rank: exercises the run-time NCCL binding and the private-stream ordering on any GPU box."""
    _run(1)


def test_allgather_two_ranks():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    _run(2)


def test_comm_errors(ctx):
    import ctypes as C
    lib = ctx.lib
    fresh = jgb200.Context(0)
    r, w = C.c_int32(7), C.c_int32(7)
    assert lib.jgb_comm_size(fresh.handle, C.byref(r), C.byref(w)) == 0 and (r.value, w.value) == (-1, 0)
    assert lib.jgb_allgather_states(fresh.handle, 1, 1, None, None, None, None, None, None, None, None) == -1
    ident = (C.c_uint8 * 128)()
    assert lib.jgb_comm_init(fresh.handle, 3, 2, ident) == -1            # rank >= nranks
    assert lib.jgb_comm_wait(fresh.handle, 1) == 0                        # nothing pending: a no-op
    fresh.close()
