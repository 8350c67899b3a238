#!/usr/bin/env python3
"""Timing of the gaf2unstable path (BASELINE config 2 shape) on one GPU next to the reference binary.
Not part of bench.py's contract (config 2 is a parity configuration); run by hand:
    python tools/bench_unstable.py [records]"""
import os, sys, time, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cactus_gfa_tools_b200 as g2p
import helpers as H

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
t0 = time.time()
rgfa, gaf = H.gen_rgfa_case(7, n_contigs=24, n_records=n, aligned=True)
print("generated %d records (%d B GAF, %d B rGFA) in %.1f s" % (n, len(gaf), len(rgfa), time.time() - t0))
cv = g2p.Converter(0)
ok, code, msg = cv.load_rgfa(rgfa)
assert ok
for _ in range(3):
    out, res, warns = cv.unstable_host(gaf)
t0 = time.perf_counter()
out, res, warns = cv.unstable_host(gaf)
wall = time.perf_counter() - t0
print("B200 gaf2unstable: device %.3f ms (size %.3f, emit %.3f), host call %.1f ms, %d records -> %d B, %.1f M records/s device, in+out %.1f GB/s"
      % (res.device_ms, res.size_ms, res.emit_ms, wall * 1e3, res.n_records, len(out), res.n_records / res.device_ms / 1e3,
         (len(gaf) + len(out)) / res.device_ms / 1e6))
sample = gaf[:gaf.rfind(b"\n", 0, len(gaf) // 4) + 1]
t0 = time.perf_counter()
rc, ref, err = H.run_gaf2unstable_ref(sample, rgfa)
dt = time.perf_counter() - t0
print("reference gaf2unstable: %.2f s for %d records (incl. rGFA load) -> %.0f records/s" % (dt, sample.count(b"\n"), sample.count(b"\n") / dt))
assert out[:len(ref)] == ref
print("parity on the sample: OK")
