#!/usr/bin/env python3
"""gaffilter on the device: time g2p_filter_host / device on a synthetic PAF (or GAF) and compare with the reference
executable when it is present.   python tools/filter_bench.py [--records N] [--gaf] [--no-ref] [--reps K]"""
import argparse, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cactus_gfa_tools_b200 as g2p, helpers as H

ap = argparse.ArgumentParser()
ap.add_argument("--records", type=int, default=600000)
ap.add_argument("--gaf", action="store_true")
ap.add_argument("--queries", type=int, default=0, help="number of query sequences (default: records / 3)")
ap.add_argument("--no-ref", action="store_true")
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
text, lengths = H.gen_filter_case(5, n_records=a.records, n_queries=a.queries or max(1, a.records // 3), paf=not a.gaf)
cv = g2p.Converter(0)
par = g2p.Converter.filter_params(paf=not a.gaf, ratio=2)
for _ in range(a.reps):
    out, res = cv.filter_host(text, par)
t0 = time.perf_counter(); out, res = cv.filter_host(text, par); wall = time.perf_counter() - t0
msg = "gaffilter %s-r 2: %d lines (%d B) -> %d kept; device %.3f ms (%d launches), host call %.1f ms" % (
    "" if a.gaf else "-p ", res.n_loaded, len(text), res.n_loaded - res.n_filtered, res.device_ms, res.gpu_launches, wall * 1e3)
ref = os.path.join(H.REF_BIN, "gaffilter")
if not a.no_ref and os.path.exists(ref):
    open("/dev/shm/f.txt", "wb").write(text)
    t0 = time.perf_counter()
    try:   # (the reference is quadratic in the alignments per query sequence: bounded, and skipped with --no-ref for contig-shaped inputs)
        r = subprocess.run([ref, "/dev/shm/f.txt"] + ([] if a.gaf else ["-p"]) + ["-r", "2"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, timeout=120)
        msg += "; reference %.2f s; identical: %s" % (time.perf_counter() - t0, out == r.stdout)
    except subprocess.TimeoutExpired:
        msg += "; reference: no result within 120 s"
    os.unlink("/dev/shm/f.txt")
print(msg)
