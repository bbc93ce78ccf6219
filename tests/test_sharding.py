"""Multi-rank path on CPU: world_size-2 gloo processes shard a batch, "solve" their shard with the
host-emulation build of the kernel source (tests/emu; no GPU here) and gather the results."""
import os
import sys
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions():
    from boundmpc_b200.sharding import shard_range
    for total in (0, 1, 7, 64, 65536, 65537):
        for world in (1, 2, 4, 8):
            r = [shard_range(total, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, total, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from boundmpc_b200.sharding import shard_range, gather_results
    from tests.emu import emu
    from tests.util import load
    S = load("seq_exp2.npz")
    idx = np.arange(total) % len(S["x0"])
    lo, hi = shard_range(total, rank, world)
    r = emu.solve(S["x0"][idx[lo:hi]], S["p"][idx[lo:hi]], tol=1e-9)
    out = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in r.items()}
    g = gather_results(out, total, rank, world)
    q.put((rank, g["x"].numpy(), g["iters"].numpy(), g["status"].numpy(), g["f"].numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [6, 7])
def test_two_rank_gather_matches_single_rank(total):
    from tests.emu import emu
    from tests.util import load
    S = load("seq_exp2.npz")
    idx = np.arange(total) % len(S["x0"])
    ref = emu.solve(S["x0"][idx], S["p"][idx], tol=1e-9)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + total
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p_ in procs:
        p_.start()
    got = [q.get(timeout=300) for _ in range(2)]
    for p_ in procs:
        p_.join(timeout=60)
        assert p_.exitcode == 0
    for rank, x, iters, status, f in got:
        # same instance -> bitwise the same solution on any rank / shard size
        assert np.array_equal(x, ref["x"])
        assert np.array_equal(iters, ref["iters"]) and np.array_equal(status, ref["status"])
        assert np.array_equal(f, ref["f"])
