"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: ray sharding and the gradient reducer."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import wisp_b200 as W
from wisp_b200 import parallel as P


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    big = torch.nn.Parameter(torch.zeros(1 << 20)); s1 = torch.nn.Parameter(torch.zeros(7, 3)); s2 = torch.nn.Parameter(torch.zeros(5))
    big.grad = torch.full_like(big, float(rank + 1)); s1.grad = torch.full_like(s1, float(10 * (rank + 1))); s2.grad = torch.arange(5.0) * (rank + 1)
    red = P.GradientReducer([big, s1, s2])
    assert len(red.big) == 1 and len(red.small) == 2
    red.reduce()
    ok = bool(torch.allclose(big.grad, torch.full_like(big, 1.5)) and torch.allclose(s1.grad, torch.full_like(s1, 15.0))
              and torch.allclose(s2.grad, torch.arange(5.0) * 1.5))
    S, R = P.sync_sample_counts(100 * (rank + 1), 10, torch.device("cpu"))
    ok = ok and (S, R) == (300, 20)
    b, e = P.shard_range(1024 * 1024 + 3, rank, world)
    t = torch.tensor([b, e], dtype=torch.int64)
    gathered = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gathered, t)
    ok = ok and int(gathered[0][0]) == 0 and int(gathered[0][1]) == int(gathered[1][0]) and int(gathered[1][1]) == 1024 * 1024 + 3
    # the trainer step's gradient exchange (MultiviewStep._all_reduce): sum over ranks of the grid buffers, and ONE small collective for
    # decoder gradients + loss; the 1/world of the mean is in the loss gradient already (inv_count uses the global ray count)
    from types import SimpleNamespace
    from wisp_b200.trainers import MultiviewStep
    st = SimpleNamespace(g_grid=[torch.full((1000, 2), float(rank + 1)), torch.full((10,), 2.0 * (rank + 1))], g_dens=torch.full((37,), float(rank)),
                         g_col=torch.arange(11.0) * (rank + 1), g_rest=[torch.full((3, 3), 5.0 + rank)], group=None)
    loss = torch.tensor([0.25 * (rank + 1)])
    MultiviewStep._all_reduce(st, loss)
    ok = ok and bool(torch.allclose(st.g_grid[0], torch.full((1000, 2), 3.0)) and torch.allclose(st.g_grid[1], torch.full((10,), 6.0))
                     and torch.allclose(st.g_dens, torch.full((37,), 1.0)) and torch.allclose(st.g_col, torch.arange(11.0) * 3)
                     and torch.allclose(st.g_rest[0], torch.full((3, 3), 11.0)) and abs(float(loss) - 0.75) < 1e-6)
    out[rank] = ok
    dist.destroy_process_group()


def test_gradient_reducer_and_sharding_world2():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert all(out[r] for r in range(world))


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 1024, 1048576):
        for w in (1, 2, 3, 8):
            spans = [P.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(e - b for b, e in spans) - min(e - b for b, e in spans) <= 1


def test_install_is_noop_without_wisp():
    """Without an importable wisp nothing is patched (the patch body itself is exercised against the real classes in
    tests/test_host_cpu.py::test_install_patches_the_real_wisp_classes)."""
    import sys
    from wisp_b200 import install
    if "wisp" not in sys.modules:
        assert install.install() is False and not install._ORIG
