"""Multi-rank host logic on CPU (gloo, world_size 2): row sharding and the
"all-reduce fp32 partial sums, then cast" composition of the row-sharded backward
(reference mechanism: lower_to_communication.cpp:266-284 + post_communication.cpp:288-313).
The kernels themselves need a GPU; here the per-shard partial sums come from the oracle."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_rows_partitions_exactly():
    from fuser_b200.parallel import shard_rows

    for rows in (0, 1, 7, 8, 32768, 32771):
        for world in (1, 2, 3, 4, 8):
            blocks = [shard_rows(rows, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == rows
            for (a0, a1), (b0, b1) in zip(blocks, blocks[1:]):
                assert a1 == b0 and a1 - a0 >= b1 - b0 >= 0
            assert max(b - a for a, b in blocks) - min(b - a for a, b in blocks) <= 1
    with pytest.raises(ValueError):
        shard_rows(8, 2, 2)


def _worker(rank, world, port, ok):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fuser_b200.parallel import local_rows
        from oracle import aten_oracle as AO

        torch.manual_seed(0)  # same data on every rank
        m, n, dt = 64, 256, torch.bfloat16
        x = torch.randn(m, n).to(dt)
        dy = torch.randn(m, n).to(dt)
        w = torch.randn(n).to(dt)
        _, mean, invstd = AO.layer_norm_fwd(x, None, None, 1e-5)
        # single-device expectation (cast once, after the full sum)
        dx_ref, dw_ref, db_ref = AO.layer_norm_bwd(dy, x, mean, invstd, w)
        # this rank's shard
        xs, dys, ms, iss = (local_rows(t, world, rank) for t in (x, dy, mean, invstd))
        xhat = (xs.float() - ms) * iss
        sums = torch.stack([(dys.float() * xhat).sum(0), dys.float().sum(0)])  # fp32 [2, N] partials
        dist.all_reduce(sums)
        dw, db = sums[0].to(dt), sums[1].to(dt)  # cast AFTER the reduce
        dx_local = AO.layer_norm_bwd(dys, xs, ms, iss, w)[0]
        a, b = (rank * m) // world, ((rank + 1) * m) // world
        AO.assert_close(dx_local, dx_ref[a:b], n, "sharded dx")
        AO.assert_close(dw, dw_ref, m, "all-reduced dw")
        AO.assert_close(db, db_ref, m, "all-reduced db")
        # the reduced result is identical on every rank
        gathered = [torch.empty_like(sums) for _ in range(world)]
        dist.all_gather(gathered, sums)
        assert all(torch.equal(g, gathered[0]) for g in gathered)
        ok[rank] = 1
    finally:
        dist.destroy_process_group()


def test_row_sharded_backward_composition_gloo_world2():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    ctx = mp.get_context("spawn")
    ok = ctx.Array("i", [0] * world)
    procs = [ctx.Process(target=_worker, args=(r, world, port, ok)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert list(ok) == [1] * world
