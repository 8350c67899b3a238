#!/usr/bin/env python3
"""PCIe / host-memory floor of the e2e number at N GPUs: every rank copies the bench workload's bytes (pinned H2D of the
GAF, D2H of the PAF) alone and concurrently on two streams -- what g2p_convert_host can at best overlap -- all ranks at
the same time, max over ranks.  One JSON line on rank 0.

    python tools/pcie_floor.py                                          # 1 GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_floor.py
"""
import json, os, subprocess, time
import torch
import torch.distributed as dist

IN, OUT = 1_373_582_239, 3_057_789_728   # bytes per GPU of the short-read workload (10 M records)
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
h_in = torch.empty(IN, dtype=torch.uint8).pin_memory(); h_out = torch.empty(OUT, dtype=torch.uint8).pin_memory()
d_in = torch.empty(IN, dtype=torch.uint8, device="cuda"); d_out = torch.empty(OUT, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def t(f, n=5):
    f(); barrier(); t0 = time.perf_counter()
    for _ in range(n):
        f()
    torch.cuda.synchronize()
    v = torch.tensor([(time.perf_counter() - t0) / n * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
    return float(v.item())


def h2d():
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)


def both():
    h2d(); d2h()


a, b, c = t(h2d), t(d2h), t(both)
if rank == 0:
    link = subprocess.run(["nvidia-smi", "--query-gpu=index,pcie.link.gen.current,pcie.link.width.current", "--format=csv,noheader"],
                          stdout=subprocess.PIPE, text=True).stdout.strip().split("\n")
    print(json.dumps({"n_gpus": world, "bytes_per_gpu": {"h2d": IN, "d2h": OUT},
                      "h2d_ms": a, "d2h_ms": b, "both_ms": c,
                      "h2d_GBps_total": IN * world / a / 1e6, "d2h_GBps_total": OUT * world / b / 1e6,
                      "both_GBps_total": (IN + OUT) * world / c / 1e6,
                      "e2e_floor_records_per_s": 10_000_000 * world / (c / 1e3),
                      "pcie_links": link, "host_cores": os.cpu_count()}))
if world > 1:
    dist.destroy_process_group()
