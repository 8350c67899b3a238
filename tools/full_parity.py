#!/usr/bin/env python3
"""Full-size parity: the whole bench workload converted on the GPU, byte-compared with the reference
binary run in parallel on newline-aligned shards of the same input (one process per host core).
    python tools/full_parity.py short 10000000
    python tools/full_parity.py asm 4000"""
import hashlib, os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cactus_gfa_tools_b200 as g2p
import helpers as H

preset, n = sys.argv[1], int(sys.argv[2])
p = H.preset(preset, seed=1)
lengths = H.gen_lengths(p)
gaf = H.gen_records(p, 0, n)
cv = g2p.Converter(0)
assert cv.load_lengths(lengths)
t0 = time.perf_counter()
out, res = cv.convert_host(gaf)
t_gpu = time.perf_counter() - t0
assert g2p.exit_code(res) == 0
procs = min(os.cpu_count() or 1, 48)
binary, kind = H.oracle_path()
ranges = g2p.shard_ranges(gaf, procs)
t0 = time.perf_counter()
with tempfile.TemporaryDirectory(dir="/dev/shm") as td:
    lp = os.path.join(td, "l.tsv"); open(lp, "wb").write(lengths)
    ps = []
    for i, (a, b) in enumerate(ranges):
        fp = os.path.join(td, "s%d.gaf" % i); open(fp, "wb").write(gaf[a:b])
        ps.append(subprocess.Popen([binary, fp, "-l", lp], stdout=open(os.path.join(td, "o%d.paf" % i), "wb")))
    assert all(q.wait() == 0 for q in ps)
    h_ref, n_ref = hashlib.md5(), 0
    for i in range(procs):
        d = open(os.path.join(td, "o%d.paf" % i), "rb").read(); h_ref.update(d); n_ref += len(d)
t_ref = time.perf_counter() - t0
same = n_ref == len(out) and hashlib.md5(out).hexdigest() == h_ref.hexdigest()
print("%s %d records: GAF %d B -> PAF %d B; GPU host call %.2f s; %s on %d processes %.1f s; md5 %s -> %s"
      % (preset, n, len(gaf), len(out), t_gpu, kind, procs, t_ref, h_ref.hexdigest(), "IDENTICAL" if same else "DIFFERENT"))
sys.exit(0 if same else 1)
