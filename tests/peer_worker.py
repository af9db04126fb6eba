"""Worker of tests/test_peer_gpu.py::test_peer_two_gpus (run under torchrun, one rank per GPU):
row-sharded LayerNorm / RMSNorm backward with the all-reduce of dgamma / dbeta fused into the kernel
over peer memory, checked against (a) the NCCL path on the same shards — bit-exact, both add the
ranks' fp32 sums in rank order... NCCL's order is its own, so within 1 fp32 ulp then cast — and
(b) the ATen oracle on the unsharded problem."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from fuser_b200 import ops, parallel  # noqa: E402
from oracle import aten_oracle as AO  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    deferred = len(sys.argv) > 1 and sys.argv[1] == "deferred"
    px = parallel.PeerExchange(max_cols=12288, device=dev, deferred=deferred)
    nccl = parallel.ColumnSumReducer(device=dev)
    for (m, n, dt) in [(4099, 4096, torch.bfloat16), (1000, 12288, torch.bfloat16), (777, 2048, torch.float32),
                       (4099, 4096, torch.bfloat16)]:
        g = torch.Generator().manual_seed(m + n)
        x = torch.randn(m, n, generator=g).to(dt)
        dy = torch.randn(m, n, generator=g).to(dt)
        w = torch.randn(n, generator=g).to(dt)
        _, mean, invstd = AO.layer_norm_fwd(x, None, None, 1e-5)
        _, rstd = AO.rms_norm_fwd(x, w, 1e-5)
        a, b = parallel.shard_rows(m, world, rank)
        xs, dys, ms, iss, rs, wd = (t.to(dev) for t in (x[a:b], dy[a:b], mean[a:b], invstd[a:b], rstd[a:b], w))
        for it in range(3):  # consecutive launches alternate the exchange slots
            dx, h = parallel.layer_norm_bwd_sharded(dys, xs, ms, iss, wd, px)
            dw, db = h.result()
            dx2, h2 = parallel.layer_norm_bwd_sharded(dys, xs, ms, iss, wd, nccl)
            dw2, db2 = h2.result()
            torch.cuda.synchronize()
            assert torch.equal(dx, dx2)
            dxe, dwe, dbe = AO.layer_norm_bwd(dy, x, mean, invstd, w)
            AO.assert_close(dx, dxe[a:b], n, "peer ln dx")
            AO.assert_close(dw, dwe, m, "peer ln dw")
            AO.assert_close(db, dbe, m, "peer ln db")
            # the fused reduce and NCCL add the same fp32 partial sums: at most a rounding apart
            assert (dw.float() - dw2.float()).abs().max() <= 2 * torch.finfo(dt).eps * dwe.float().abs().max()
            assert (db.float() - db2.float()).abs().max() <= 2 * torch.finfo(dt).eps * dbe.float().abs().max()
            # every rank holds the same bits
            gathered = [torch.empty_like(dw) for _ in range(world)]
            dist.all_gather(gathered, dw)
            assert all(torch.equal(gathered[0], t) for t in gathered), "ranks disagree on dgamma"
            dx, h = parallel.rms_norm_bwd_sharded(dys, xs, rs, wd, px, centered=True)
            if deferred:  # two publishes in flight, combined in order
                dx_b, h_b = parallel.rms_norm_bwd_sharded(dys, xs, rs, wd, px, centered=True)
                (dw_b,) = h_b.result()
                assert torch.equal(h.result()[0], dw_b) and torch.equal(dx, dx_b)
            (dw,) = h.result()
            torch.cuda.synchronize()
            dxe, dwe = AO.rms_norm_bwd(dy, x, rstd, w)
            AO.assert_close(dx, dxe[a:b], n, "peer rms dx")
            AO.assert_close(dw, dwe, m, "peer rms dw")
    dist.barrier()
    px.close()
    if rank == 0:
        print("peer worker ok", "deferred" if deferred else "fused")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
