#!/usr/bin/env python
"""Per-CTA timeline of the row-sharded backward with the fused peer all-reduce, on every rank (run under torchrun).

    torchrun --nproc-per-node 8 scripts/timeline_peer.py [c4|c2]

Each rank launches its shard of the backward a few times and instruments the last launch
(b200norm_debug_set_timeline); rank 0 prints, per rank, min / median / max over the CTAs of every mark in
microseconds relative to that rank's first CTA entry: first tile, row loop done, partial row written, grid barrier
passed, column sums finished and pushed to the peers, peers' sums collected, exit."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from fuser_b200 import _lib as L  # noqa: E402
from fuser_b200 import ops, parallel  # noqa: E402

SLOTS = 12
NAMES = {2: "first tile", 3: "loop done", 4: "partial row", 5: "grid barrier", 8: "finished+pushed", 9: "collected", 6: "exit"}


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "c4"
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    M, N, ln = (32768, 12288, True) if which == "c4" else (16384 * world, 4096, False)
    a, b = parallel.shard_rows(M, world, rank)
    m = b - a
    g = torch.Generator(device="cuda").manual_seed(rank)
    sets = []
    for _ in range(3):
        x = torch.randn(m, N, device=dev, generator=g).to(torch.bfloat16)
        dy = torch.randn(m, N, device=dev, generator=g).to(torch.bfloat16)
        sets.append((x, dy))
    w = torch.randn(N, device=dev, generator=g).to(torch.bfloat16)
    px = parallel.PeerExchange(max_cols=N, device=dev, deferred=False)
    saved = []
    for x, _ in sets:
        if ln:
            saved.append(ops.layer_norm_fwd(x, None, None, 1e-5)[1:])
        else:
            saved.append((ops.rms_norm_fwd(x, w, 1e-5, stat_is_rms=True)[1],))

    def launch(i):
        x, dy = sets[i % 3]
        if ln:
            return parallel.layer_norm_bwd_sharded(dy, x, saved[i % 3][0], saved[i % 3][1], w, px)
        return parallel.rms_norm_bwd_sharded(dy, x, saved[i % 3][0], w, px, stat_is_rms=True)

    for i in range(5):
        launch(i)
    torch.cuda.synchronize()
    dist.barrier()
    buf = torch.zeros(2048 * SLOTS, dtype=torch.int64, device=dev)
    for i in range(3):
        launch(i)
    L.set_timeline(buf.data_ptr())
    launch(3)
    L.set_timeline(None)
    torch.cuda.synchronize()
    t = buf.cpu().view(-1, SLOTS)
    nct = int((t[:, 0] != 0).sum())
    t = t[:nct]
    t0 = int(t[:, 0].min())
    row = [float(nct)]
    for s in (2, 3, 4, 5, 8, 9, 6):
        col = (t[:, s][t[:, s] != 0] - t0).double() / 1e3
        row += [float(col.min()), float(col.median()), float(col.max())] if col.numel() else [0.0, 0.0, 0.0]
    allrows = [torch.zeros(len(row), dtype=torch.float64, device=dev) for _ in range(world)]
    dist.all_gather(allrows, torch.tensor(row, dtype=torch.float64, device=dev))
    if rank == 0:
        print(f"== {which}: rows per rank {m}, N {N}, {world} ranks, fused peer all-reduce; us after the rank's first CTA entry "
              f"(min / median / max over the CTAs)")
        for r, rr in enumerate(allrows):
            rr = rr.tolist()
            parts = []
            for i, s in enumerate((2, 3, 4, 5, 8, 9, 6)):
                lo, med, hi = rr[1 + 3 * i: 4 + 3 * i]
                parts.append(f"{NAMES[s]} {lo:.1f}/{med:.1f}/{hi:.1f}")
            print(f" rank {r} ({int(rr[0])} CTAs): " + "  ".join(parts), flush=True)
    dist.barrier()
    px.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
