#!/usr/bin/env python3
"""bench.py — gaf2paf throughput on B200 (BASELINE.json metric: GAF records/s and input GB/s
vs the HBM roofline), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload short|asm] [--records R]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference's CPU gaf2paf on the host cores

A step is one pass of the whole device pipeline (line index -> size pass -> scan -> emit
pass; kernels k_rec / k_long / k_convert_list) over one synthetic batch.  `value` is measured with the batch already resident in HBM
(CUDA events around exactly K steps, max over ranks); `e2e` is the same metric through the
host-buffer C-ABI call (pinned host input -> H2D -> pipeline -> D2H), i.e. what the
`gaf2paf` executable does per chunk.  Multi-GPU runs shard by records: every rank converts
its own newline-aligned shard (weak scaling, no collective on the data path).
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOADS = {
    # BASELINE.json configs[2]: synthetic 10M-record short-read GAF on 1 B200
    "short": {"preset": "short", "records": 10_000_000, "desc": "configs[2]: synthetic short-read GAF, 1-5 node steps, short cg CIGARs"},
    # BASELINE.json configs[3] at a record count whose PAF fits one GPU next to the input
    # shapes of configs[0] / configs[1] (stable-interval and node-coordinate assembly alignments), parity-test sized
    "stable": {"preset": "stable", "records": 300_000, "desc": "configs[0] shape: stable-interval steps (>contig:start-end), ~2 kB records"},
    "medium": {"preset": "medium", "records": 100_000, "desc": "configs[1] shape: node-coordinate records of a few hundred steps, ~12 kB"},
    "asm": {"preset": "asm", "records": 4000, "desc": "configs[3] shape: assembly-scale records, 5k-15k steps, 4000 of the 100k records (PAF of all would not fit next to the input)"},
}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_baseline(H, preset_name, seed, sample_records, procs):
    """Time the CPU gaf2paf (oracle/_ref reference build when present, else the port) on a bounded
    sample of the same workload: `procs` independent processes on disjoint record ranges."""
    binary, kind = H.oracle_path()
    p = H.preset(preset_name, seed=seed)
    lengths = H.gen_lengths(p)
    per = max(1, sample_records // procs)
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as td:
        lp = os.path.join(td, "l.tsv")
        open(lp, "wb").write(lengths)
        files, nbytes = [], 0
        for i in range(procs):
            g = H.gen_records(p, i * per, per)
            fp = os.path.join(td, "s%d.gaf" % i)
            open(fp, "wb").write(g)
            files.append(fp)
            nbytes += len(g)
        t0 = time.perf_counter()
        ps = [subprocess.Popen([binary, fp, "-l", lp], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for fp in files]
        rcs = [q.wait() for q in ps]
        dt = time.perf_counter() - t0
    if any(rcs):
        raise RuntimeError("cpu baseline failed: rc %r" % rcs)
    return {"value": per * procs / dt, "unit": "records/s", "cores": procs, "kind": kind,
            "sample": "%d records (%d B) of the same workload, %d process(es), incl. lengths-table load" % (per * procs, nbytes, procs),
            "seconds": dt, "input_MBps": nbytes / dt / 1e6}


def run_reference(a):
    """--impl reference: the reference's own CPU implementation on the host cores."""
    import helpers as H
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    wl = WORKLOADS[a.workload]
    procs = min(os.cpu_count() or 1, 64)
    # per-step sample sized for ~2-4 s of work per process
    per_proc = {"short": 60000, "stable": 4000, "medium": 600}.get(a.workload, 12)
    vals = []
    for s in range(a.warmup + a.steps):
        r = cpu_baseline(H, wl["preset"], 1000 + s, per_proc * procs, procs)
        if s >= a.warmup:
            vals.append(r)
    tot_rec = sum(float(r["sample"].split()[0]) for r in vals)
    tot_s = sum(r["seconds"] for r in vals)
    v = tot_rec / tot_s
    line = {
        "impl": "reference", "metric": "gaf2paf_records_per_s", "value": v, "unit": "records/s", "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1000.0 * tot_s / max(1, len(vals)), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": wl["desc"], "preset": wl["preset"], "records_per_step": per_proc * procs},
        "cpu_baseline": {"value": v, "unit": "records/s", "cores": procs, "kind": vals[-1]["kind"], "sample": vals[-1]["sample"]},
        "e2e": {"value": v, "unit": "records/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "input_MBps": statistics.mean(r["input_MBps"] for r in vals), "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="short", choices=sorted(WORKLOADS))
    ap.add_argument("--records", type=int, default=0, help="records per GPU (default: the workload's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl == "b200" else a.warmup
    if a.impl == "reference":
        return run_reference(a)

    import torch
    import torch.distributed as dist
    import cactus_gfa_tools_b200 as g2p
    import helpers as H

    if not torch.cuda.is_available():
        sys.exit("bench.py: no CUDA device; this path has no CPU fallback (use --impl reference for the CPU baseline)")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on STDOUT when the first communicator is created (measured on the
        # GPU box with NCCL_DEBUG=VERSION in the environment).  stdout carries the JSON line only: file
        # descriptor 1 points at stderr while the process group and its first collective come up.
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    wl = WORKLOADS[a.workload]
    nrec = a.records or wl["records"]
    p = H.preset(wl["preset"], seed=1)
    lengths = H.gen_lengths(p)
    cv = g2p.Converter(local)
    assert cv.load_lengths(lengths)

    # this rank's shard: records [rank*nrec, (rank+1)*nrec), generated on the host, copied to pinned memory
    threads = max(1, (os.cpu_count() or 8) // max(1, world))
    addr, nbytes = H.gen_records_raw(p, rank * nrec, nrec, threads=min(64, threads))
    pinned = g2p.lib.g2p_host_alloc(nbytes + 16)
    ctypes.memmove(pinned, addr, nbytes)
    H.gen_free(addr)
    d_in = torch.empty(nbytes + 16, dtype=torch.uint8, device="cuda")
    assert g2p.lib.g2p_copy_to_device(d_in.data_ptr(), pinned, nbytes) == 0
    torch.cuda.synchronize()
    stream = torch.cuda.current_stream().cuda_stream

    # ---- device-resident timing (value)
    res = None
    for _ in range(a.warmup):
        d_out, res = cv.convert_device(d_in.data_ptr(), nbytes, stream)
    assert g2p.exit_code(res) == 0, "synthetic workload must convert cleanly"
    out_bytes = res.out_bytes
    n_records = res.n_records
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    emit_ms, size_ms, index_ms, dev_ms, launches = [], [], [], [], 0
    e0.record()
    for _ in range(a.steps):
        d_out, res = cv.convert_device(d_in.data_ptr(), nbytes, stream)
        emit_ms.append(res.emit_ms); size_ms.append(res.size_ms); index_ms.append(res.index_ms); dev_ms.append(res.device_ms)
        launches += res.gpu_launches
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    t_ms = e0.elapsed_time(e1)
    t = torch.tensor([t_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t_ms = float(t.item())

    # ---- end to end through the host-buffer C-ABI call (pinned input, H2D + D2H inside)
    e2e = None
    if not a.no_e2e:
        for _ in range(2):
            cv.convert_host_raw(pinned, nbytes)
        barrier()
        w0 = time.perf_counter()
        for _ in range(a.steps):
            o_addr, r2 = cv.convert_host_raw(pinned, nbytes)
        torch.cuda.synchronize()
        w = torch.tensor([time.perf_counter() - w0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(w, op=dist.ReduceOp.MAX)
        w_s = float(w.item())
        e2e = {"value": n_records * world * a.steps / w_s, "unit": "records/s", "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": int(r2.out_bytes),
               "ms_per_step": 1000.0 * w_s / a.steps, "input_GBps": nbytes * world * a.steps / w_s / 1e9}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peak, peak_src = peaks()
    value = n_records * world * a.steps / (t_ms / 1000.0)
    # dominant kernel: the emit pass (reads the GAF, writes the PAF) or the size pass (reads the GAF);
    # k_rec (or k_short with G2P_SIZE_KERNEL=short) converts short records, k_long the ones it delegates (res.n_long)
    em, sz = statistics.mean(emit_ms), statistics.mean(size_ms)
    short_kernel = "k_short<8>" if os.environ.get("G2P_SIZE_KERNEL") == "short" else "k_rec"
    kname = "k_long" if res.n_long * 2 > n_records else short_kernel
    if em >= sz:   # k_emit_lines: reads descriptors + GAF text, writes the PAF
        dom, dom_ms, dom_bytes, dom_key = "k_emit_lines (emit pass)", em, nbytes + out_bytes, "k_emit_lines"
    else:          # size pass: parses the GAF, writes sizes + line descriptors
        dom, dom_ms, dom_bytes, dom_key = kname + " (size pass)", sz, nbytes, kname.split("<")[0] + "_size"
    # DRAM traffic of that kernel from the committed ncu --set full capture (bytes per record of the
    # captured launch, scaled to this launch's record count); null when no capture is committed
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
        per_rec = tj.get(a.workload, {}).get(dom_key)
        if per_rec:
            traffic = per_rec * n_records
    except Exception:
        pass
    achieved = dom_bytes / (dom_ms / 1000.0) / 1e9
    line = {
        "metric": "gaf2paf_records_per_s", "value": value, "unit": "records/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": t_ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64",
        "data": "synthetic",
        "config": {"workload": wl["desc"], "preset": wl["preset"], "records_per_gpu": int(n_records), "gaf_bytes_per_gpu": nbytes,
                   "paf_bytes_per_gpu": int(out_bytes), "table_entries": int(cv.table_entries), "l2": "inputs_larger_than_l2",
                   "sharding": "newline-aligned record ranges, one per GPU, no collective"},
        "input_GBps": nbytes * world * a.steps / (t_ms / 1000.0) / 1e9,
        "pipeline_in_plus_out_GBps": (nbytes + out_bytes) * world * a.steps / (t_ms / 1000.0) / 1e9,
        "pipeline_frac_of_hbm_peak": (nbytes + out_bytes) * a.steps / (t_ms / 1000.0) / 1e9 / peak,
        "kernel_ms": {"index": statistics.mean(index_ms), "size": sz, "emit": em, "device_pipeline": statistics.mean(dev_ms)},
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "algorithmic_bytes_per_launch": dom_bytes, "peak_source": peak_src},
        "records_by_kernel": {short_kernel.split("<")[0]: int(n_records - res.n_long), "k_long": int(res.n_long - res.n_delegated), "general": int(res.n_delegated)},
        "gpu_launches": launches,
        "clocks": clocks,
    }
    if e2e:
        line["e2e"] = e2e
    if not a.no_cpu_baseline and world == 1:
        try:
            line["cpu_baseline"] = cpu_baseline(H, wl["preset"], 1, {"short": 600000, "stable": 40000, "medium": 6000}.get(a.workload, 40), 1)
        except Exception as ex:   # the baseline must not take the GPU number down with it
            line["cpu_baseline"] = {"error": str(ex)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
