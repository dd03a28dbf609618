#!/usr/bin/env python
"""Host <-> device link probe: what the box's host side can feed N GPUs at once.

The end-to-end dynamics path (unwrap -> MSD -> ACF from a host-resident store) is bound by the
host link, not by the kernels; this script measures the ceiling that path can reach at N ranks:
per-rank and aggregate GB/s for pinned H2D, D2H and both directions at once, all ranks copying
simultaneously (barrier-aligned), plus the page-locking rate of fresh pinned memory.

    python scripts/hostlink_probe.py                  # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29511 scripts/hostlink_probe.py

Rank 0 prints one JSON line.
"""
import json
import os
import subprocess
import sys
import time

import torch
import torch.distributed as dist


def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=20).stdout.strip()
    except Exception as exc:  # pragma: no cover
        return f"{type(exc).__name__}: {exc}"


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    gib = float(os.environ.get("PROBE_GIB", "2"))
    n = int(gib * (1 << 30)) // 4

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    barrier()
    t0 = time.perf_counter()
    h_src = torch.empty(n, dtype=torch.float32, pin_memory=True)
    t_pin = time.perf_counter() - t0
    h_src.fill_(1.0)
    h_dst = torch.empty(n, dtype=torch.float32, pin_memory=True)
    h_dst.fill_(0.0)
    d_a = torch.empty(n, dtype=torch.float32, device=dev)
    d_b = torch.ones(n, dtype=torch.float32, device=dev)
    s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
    nbytes = n * 4

    def timed(fn, reps=4):
        fn()
        barrier()
        t = time.perf_counter()
        for _ in range(reps):
            fn()
        barrier()
        return (time.perf_counter() - t) / reps

    def up():
        with torch.cuda.stream(s_up):
            d_a.copy_(h_src, non_blocking=True)

    def down():
        with torch.cuda.stream(s_dn):
            h_dst.copy_(d_b, non_blocking=True)

    def both():
        up()
        down()

    res = {}
    for name, fn, factor in (("h2d", up, 1), ("d2h", down, 1), ("duplex", both, 2)):
        t = timed(fn)
        res[name] = factor * nbytes / t * 1e-9   # GB/s of this rank while all ranks copy

    vals = torch.tensor([res["h2d"], res["d2h"], res["duplex"], gib / t_pin], dtype=torch.float64,
                        device=dev)
    if world > 1:
        allv = [torch.zeros_like(vals) for _ in range(world)]
        dist.all_gather(allv, vals)
    else:
        allv = [vals]
    if rank == 0:
        table = torch.stack(allv).cpu()
        line = {
            "n_ranks": world, "gib_per_copy": gib,
            "per_rank_GBps": {k: [round(float(x), 1) for x in table[:, i]]
                              for i, k in enumerate(["h2d", "d2h", "duplex"])},
            "aggregate_GBps": {k: round(float(table[:, i].sum()), 1)
                               for i, k in enumerate(["h2d", "d2h", "duplex"])},
            "pin_rate_GiBps_per_rank": [round(float(x), 2) for x in table[:, 3]],
            "host": {
                "nproc": os.cpu_count(),
                "mem": sh("free -g | sed -n 2p"),
                "shm": sh("df -h /dev/shm | tail -1"),
                "numa": sh("lscpu | grep -i -E 'numa|model name|socket'"),
                "topo": sh("nvidia-smi topo -m | head -12"),
                "pcie": sh("nvidia-smi --query-gpu=index,pcie.link.gen.current,"
                           "pcie.link.width.current --format=csv,noheader"),
            },
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    sys.exit(main())
