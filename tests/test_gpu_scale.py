"""GPU parity at scale and on the multi-GPU executable path (pytest -m gpu, B200 box).

* whole-output md5 of every bench workload shape against the reference binary, which runs in parallel on
  newline-aligned shards of the same input (one process per host core) -- the full output, not a slice;
* the drop-in executable with G2P_GPUS=2: one input sharded over two GPUs by newline-aligned byte ranges,
  stdout in input order (reference analogue: the in-order loop gaf2paf_main.cpp:342-374), including a failing
  record in a chunk converted by the second GPU;
* the double-buffered pinned results of the host-buffer call (include/g2p.h).
"""
import hashlib
import os
import subprocess
import tempfile

import pytest

import helpers as H

pytestmark = pytest.mark.gpu


def ref_md5_parallel(g2p, gaf, lengths):
    """md5 + size of the reference's output for `gaf`, computed on shards (its records are independent)."""
    binary, kind = H.oracle_path()
    procs = max(1, min(os.cpu_count() or 1, 48))
    ranges = [r for r in g2p.shard_ranges(gaf, procs) if r[1] > r[0]]
    d = "/dev/shm" if os.path.isdir("/dev/shm") else None
    with tempfile.TemporaryDirectory(dir=d) as td:
        lp = os.path.join(td, "l.tsv")
        open(lp, "wb").write(lengths)
        ps = []
        for i, (a, b) in enumerate(ranges):
            fp = os.path.join(td, "s%d.gaf" % i)
            open(fp, "wb").write(gaf[a:b])
            ps.append(subprocess.Popen([binary, fp, "-l", lp], stdout=open(os.path.join(td, "o%d.paf" % i), "wb"), stderr=subprocess.DEVNULL))
        assert all(q.wait() == 0 for q in ps)
        h, n = hashlib.md5(), 0
        for i in range(len(ranges)):
            with open(os.path.join(td, "o%d.paf" % i), "rb") as f:
                while True:
                    blk = f.read(1 << 24)
                    if not blk:
                        break
                    h.update(blk)
                    n += len(blk)
    return h.hexdigest(), n, kind


@pytest.mark.parametrize("name,count,over", [
    ("short", 3_000_000, {}),            # BASELINE configs[2] shape (the bench runs 10 M of these)
    ("tagged", 600_000, {}),             # 250-500 B records: k_long's descriptor blocks
    ("mixed", 1_200_000, {"mix_every": 16400}),   # configs[4] shape: ~73 assembly-scale records among the short ones
    ("stable", 60_000, {}),              # configs[0] shape
    ("medium", 40_000, {}),              # configs[1] shape
    ("asm", 600, {}),                    # configs[3] shape
])
def test_whole_output_md5_matches_reference(g2p, name, count, over):
    p = H.preset(name, seed=5, **over)
    lengths = H.gen_lengths(p)
    gaf = H.gen_records(p, 0, count)
    cv = g2p.Converter(0)
    try:
        assert cv.load_lengths(lengths)
        out, res = cv.convert_host(gaf)
        assert g2p.exit_code(res) == 0 and res.n_records == gaf.count(b"\n")
        got = hashlib.md5(out).hexdigest()
        n_got = len(out)
        del out
    finally:
        cv.close()
    ref, n_ref, kind = ref_md5_parallel(g2p, gaf, lengths)
    assert n_got == n_ref, "PAF size differs from the %s oracle" % kind
    assert got == ref, "PAF md5 differs from the %s oracle" % kind


def test_default_dispatch_sends_250_to_1000_byte_records_to_k_fuse(g2p):
    p = H.preset("tagged", seed=15)
    lengths = H.gen_lengths(p)
    gaf = H.gen_records(p, 0, 50000)
    cv = g2p.Converter(0)
    try:
        assert cv.load_lengths(lengths)
        out, res = cv.convert_host(gaf)
        assert res.n_fused == res.n_records == 50000 and res.n_long == 0
        assert out == H.run_gaf2paf_cpu(gaf, lengths)[1]
    finally:
        cv.close()


def n_gpus():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
        return sum(1 for ln in out.splitlines() if ln.startswith("GPU "))
    except Exception:
        return 0


@pytest.mark.parametrize("gpus", [1, 2])
def test_cli_sharded_over_gpus_matches_reference(g2p, gpus):
    if n_gpus() < gpus:
        pytest.skip("needs %d GPUs" % gpus)
    exe = os.path.join(g2p.BIN_DIR, "gaf2paf")
    binary, kind = H.oracle_path()
    p = H.preset("mixed", seed=9, mix_every=4000)
    lengths = H.gen_lengths(p)
    gaf = H.gen_records(p, 0, 40000)        # 10 assembly-scale records among 40 k short ones, ~8 MB
    env = dict(os.environ, G2P_GPUS=str(gpus), G2P_CHUNK_MB="1", G2P_IO_MIN_BYTES="262144", G2P_IO_THREADS="4")
    with tempfile.TemporaryDirectory() as td:
        lp, gp, op = os.path.join(td, "l.tsv"), os.path.join(td, "in.gaf"), os.path.join(td, "out.paf")
        open(lp, "wb").write(lengths)
        open(gp, "wb").write(gaf)
        rrc, rout, rerr = H.run_tool(binary, [gp, "-l", lp])
        assert rrc == 0
        # stdout a pipe
        pr = subprocess.run([exe, "-l", lp, gp], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env)
        assert pr.returncode == 0 and pr.stdout == rout
        # stdout a regular file (parallel pwrite), input from stdin (sequential read + carry)
        with open(op, "wb") as f:
            pr = subprocess.run([exe, "-l", lp, "-"], input=gaf, stdout=f, stderr=subprocess.PIPE, env=env)
        assert pr.returncode == 0 and open(op, "rb").read() == rout
        # a failing record in a late chunk (the second GPU's when gpus == 2): output up to it, message, exit 1
        lines = gaf.split(b"\n")
        k = len(lines) * 3 // 4
        f6 = lines[k].split(b"\t")
        f6[5] = b">nosuchnode" + f6[5]
        lines[k] = b"\t".join(f6)
        bad = b"\n".join(lines)
        open(gp, "wb").write(bad)
        rrc, rout2, rerr2 = H.run_tool(binary, [gp, "-l", lp])
        pr = subprocess.run([exe, "-l", lp, gp], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env)
        assert rrc == 1 and pr.returncode == 1
        assert pr.stdout == rout2 and pr.stderr.decode("latin-1") == rerr2


def test_host_results_are_double_buffered(g2p):
    """A pinned result stays valid until the next-but-one *_host call on the same context."""
    import ctypes
    p = H.preset("short", seed=13)
    lengths = H.gen_lengths(p)
    g1, g2_, g3 = (H.gen_records(p, i * 20000, 20000) for i in range(3))
    cv = g2p.Converter(0)
    try:
        assert cv.load_lengths(lengths)
        a1, r1 = cv.convert_host_raw(*g2p._buf_ptr(g1)[:2])
        ref1 = ctypes.string_at(a1, r1.out_bytes)
        a2, r2 = cv.convert_host_raw(*g2p._buf_ptr(g2_)[:2])
        assert ctypes.string_at(a1, r1.out_bytes) == ref1, "result 1 must survive call 2"
        ref2 = ctypes.string_at(a2, r2.out_bytes)
        a3, r3 = cv.convert_host_raw(*g2p._buf_ptr(g3)[:2])
        assert ctypes.string_at(a2, r2.out_bytes) == ref2, "result 2 must survive call 3"
        rc, ref, err, kind = H.run_gaf2paf_cpu(g1 + g2_ + g3, lengths)
        assert ref == ref1 + ref2 + ctypes.string_at(a3, r3.out_bytes)
    finally:
        cv.close()
