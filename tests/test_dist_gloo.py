"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: track sharding and the algebra of the
per-iteration all-reduce (sum of per-shard reduced camera systems == the unsharded system), exercised
through the oracle's LM with vggsfm_b200.dist.HostAllReduce.  The CUDA path uses the same shard ranges and
the same reduction points through vgg_allreduce_fn (vggsfm_b200/dist.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ba_oracle as bo
from tests.helpers import ba_case
from vggsfm_b200.dist import HostAllReduce, shard_range


def test_shard_range_covers_and_aligns():
    for N in (4096, 2048, 100, 17, 16):
        for world in (1, 2, 4, 8):
            spans = [shard_range(N, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == N
            for (a, b), (c, d) in zip(spans[:-1], spans[1:]):
                assert b == c and a <= b
            for a, b in spans:
                assert (b - a) % 16 == 0 or b == N      # only the tail shard may be ragged


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    c = ba_case(6, 64, "SIMPLE_RADIAL", bo.INTR_SHARED, seed=1)
    lo, hi = shard_range(64, rank, world)
    opt = bo.LMOptions()
    opt.max_num_iterations = 6
    tr = []
    p, i, x, summ = bo.lm_solve(c["poses"], c["intr"], c["points"][lo:hi], c["uv"][:, lo:hi], c["mask"][:, lo:hi],
                                c["model"], c["mode"], options=opt, trace=tr, allreduce=HostAllReduce())
    q.put((rank, p, i, x, summ, [t.get("candidate_cost") for t in tr]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_lm_equals_unsharded(world):
    """world 2: even shards; world 3: the 16-aligned split of 64 tracks is 32 / 32 / 0 -- a rank without tracks still
    takes part in every reduction and ends with the same cameras."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    c = ba_case(6, 64, "SIMPLE_RADIAL", bo.INTR_SHARED, seed=1)
    opt = bo.LMOptions()
    opt.max_num_iterations = 6
    tr = []
    p0, i0, x0, s0 = bo.lm_solve(c["poses"], c["intr"], c["points"], c["uv"], c["mask"], c["model"], c["mode"],
                                 options=opt, trace=tr)
    for rank, p, i, x, summ, costs in res:
        lo, hi = shard_range(64, rank, world)
        assert summ["iterations"] == s0["iterations"] and summ["successful"] == s0["successful"]
        assert np.allclose(costs, [t.get("candidate_cost") for t in tr], rtol=1e-9)
        assert np.abs(p - p0).max() < 1e-9 and np.abs(i - i0).max() < 1e-8
        assert hi == lo or np.abs(x - x0[lo:hi]).max() < 1e-9      # world 3 leaves the last rank an EMPTY shard: still in step
    # every rank holds identical cameras
    for r in res[1:]:
        assert np.array_equal(res[0][1], r[1])
