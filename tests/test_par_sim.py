"""k_par_* (g2p_par.cuh: the token-parallel conversion of the records k_rec does not take) under the SIMT emulator against
the reference executable, and the same inputs with G2P_PAR=0 so that k_long (one warp per record), which otherwise only
sees what k_par rejects, stays covered."""
import os
import re
import subprocess
import tempfile

import pytest

import helpers as H

SIMT = os.path.join(H.BUILD, "g2p_simt")
REF = os.path.join(H.REF_BIN, "gaf2paf")
PORT = os.path.join(H.ORACLE_BIN, "gaf2paf_oracle")
CHECK = REF if os.path.exists(REF) else PORT


def _simt(lp, data, **env):
    e = dict(os.environ, G2P_FUSE="0", G2P_SIMT_STATS="1", **env)
    p = subprocess.run([SIMT, "-l", lp, "-"], input=data, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=e)
    rc = p.returncode if p.returncode >= 0 else 128 - p.returncode
    m = re.search(r"(\d+) to k_par \((\d+) steps, (\d+) ops\), (\d+) to k_long, (\d+) to the general kernel", p.stderr.decode("latin-1"))
    return rc, p.stdout, tuple(int(x) for x in m.groups()) if m else None


@pytest.mark.parametrize("name,count,over", [
    ("stable", 400, {}),
    ("medium", 200, {}),
    ("asm", 2, {"steps_lo": 700, "steps_hi": 1300}),
    ("mixed", 2500, {"mix_every": 400, "n_nodes": 20000}),
    ("tagged", 600, {}),
    # few long steps with 5..6-digit numbers; many tiny steps and ops
    ("stable", 600, {"node_len_lo": 2000, "node_len_hi": 300000, "mrun_lo": 500, "mrun_hi": 150000, "steps_lo": 2, "steps_hi": 9, "max_runs": 30}),
    ("medium", 150, {"node_len_lo": 5, "node_len_hi": 40, "mrun_lo": 1, "mrun_hi": 12, "indel_lo": 1, "indel_hi": 3}),
])
def test_par_matches_reference(name, count, over):
    p = H.preset(name, seed=23, **over)
    lengths = H.gen_lengths(p)
    gaf = H.gen_records(p, 0, count, threads=4)
    with tempfile.TemporaryDirectory() as td:
        lp = os.path.join(td, "l.tsv")
        open(lp, "wb").write(lengths)
        rc, ref, _ = H.run_tool(CHECK, ["-", "-l", lp], gaf)
        assert rc == 0 and ref.count(b"\n") > 0
        rc1, out1, st1 = _simt(lp, gaf)
        assert rc1 == 0 and out1 == ref
        assert st1 and st1[0] > 0 and st1[1] > 0 and st1[3] == 0, st1      # k_par took what k_rec left, nothing went on to k_long
        rc0, out0, st0 = _simt(lp, gaf, G2P_PAR="0")
        assert rc0 == 0 and out0 == ref
        assert st0 and st0[1] == 0 and st0[3] == st0[0], st0               # ... and k_long takes it all when told to


def _edit_cg(line, f):
    i = line.rfind(b"cg:Z:")
    j = line.find(b"\t", i)
    cg = line[i + 5: j if j > 0 else len(line)]
    ops = re.findall(rb"\d+[A-Z=]", cg)
    return line[:i] + b"cg:Z:" + b"".join(f(ops)) + (line[j:] if j > 0 else b"")


def _col(line, k, f):
    c = line.split(b"\t")
    c[k] = f(c[k])
    return b"\t".join(c)


def test_par_rejects_go_down_the_chain():
    """Records k_par must not convert (non-canonical text, reference errors) among records it does convert: the
    canonical ones keep their descriptors, the others are converted by k_long / the general kernel, and a record that
    aborts the reference stops the output at the same byte."""
    pm = H.preset("medium", seed=7)
    lengths = H.gen_lengths(pm)
    med = H.gen_records(pm, 0, 40, threads=1).split(b"\n")[:-1]
    plus = [l for l in med if l.split(b"\t")[4] == b"+"]
    minus = [l for l in med if l.split(b"\t")[4] == b"-"]
    assert plus and minus
    variants = {
        "leading-zero-op": _edit_cg(plus[0], lambda o: o[:-1] + [b"0" + o[-1]]),            # valid, not canonical: k_long rejects it too
        "leading-zero-first-op": _edit_cg(minus[0], lambda o: [b"00" + o[0]] + o[1:]),
        "plus-sign-number": _col(plus[1], 7, lambda c: b"+" + c),                            # strtol accepts it
        "mapq-255": _col(minus[1], 11, lambda c: b"255"),
        "mapq-300": _col(plus[2], 11, lambda c: b"300"),
        "star-block": _col(plus[3], 10, lambda c: b"*"),
        "empty-tag-field": plus[4] + b"\t",
        "extra-tags": minus[2] + b"\tzz:Z:" + b"x" * 300 + b"\tyy:i:5",
        "long-qname": _col(plus[5], 0, lambda c: c + b"_" * 500),
        "path-end-short": _col(plus[6], 8, lambda c: str(int(c) - 3).encode()),
        "path-start-late": _col(minus[3], 7, lambda c: str(int(c) + 2).encode()),
    }
    with tempfile.TemporaryDirectory() as td:
        lp = os.path.join(td, "l.tsv")
        open(lp, "wb").write(lengths)
        body = []
        for k, v in variants.items():
            body += [med[len(body) % len(med)], v]
        data = b"\n".join(body + med[:5]) + b"\n"
        rc, ref, _ = H.run_tool(CHECK, ["-", "-l", lp], data)
        for env in ({}, {"G2P_PAR": "0"}, {"G2P_SIMT_REVERSE": "1"}):
            rc1, out1, st = _simt(lp, data, **env)
            assert rc1 == rc, env
            assert (out1 == ref) if rc == 0 else ref.startswith(out1), env
        # each variant alone after a few good records (some abort the reference: same exit code, same flushed prefix)
        aborting = {
            "bad-letter-last": _edit_cg(plus[0], lambda o: o[:-1] + [o[-1][:-1] + b"Q"]),
            "bad-letter-first": _edit_cg(minus[0], lambda o: [o[0][:-1] + b"Q"] + o[1:]),
            "unknown-node": plus[1].replace(b">", b">nosuchnode_", 1),
            "cigar-too-short": _edit_cg(plus[2], lambda o: o[: len(o) // 2]),
            "no-cg": b"\t".join(c for c in plus[3].split(b"\t") if not c.startswith(b"cg:Z:")),
        }
        for k, v in list(variants.items()) + list(aborting.items()):
            data = b"\n".join(med[:3] + [v] + med[3:6]) + b"\n"
            rc, ref, _ = H.run_tool(CHECK, ["-", "-l", lp], data)
            rc1, out1, st = _simt(lp, data)
            assert rc1 == rc, k
            assert (out1 == ref) if rc == 0 else ref.startswith(out1), k


def test_par_descriptor_room():
    """No room for the descriptor runs: k_par leaves the block alone and k_long (with its own overflow path) takes it."""
    p = H.preset("medium", seed=3)
    lengths = H.gen_lengths(p)
    gaf = H.gen_records(p, 0, 60, threads=2)
    with tempfile.TemporaryDirectory() as td:
        lp = os.path.join(td, "l.tsv")
        open(lp, "wb").write(lengths)
        rc, ref, _ = H.run_tool(CHECK, ["-", "-l", lp], gaf)
        rc1, out1, st = _simt(lp, gaf, G2P_SIMT_DESC_CAP="1024")
        assert rc == 0 and rc1 == 0 and out1 == ref and st[1] == 0 and st[3] == st[0]


@pytest.mark.parametrize("preset,pat", [("medium", rb"s"), ("stable", rb"ctg")])
def test_par_long_node_names(preset, pat):
    """Names longer than 16 bytes (hashed keys verified against the arena) and of exactly 16 / 17 bytes."""
    p = H.preset(preset, seed=9)
    lengths = H.gen_lengths(p)
    gaf = H.gen_records(p, 0, 120, threads=2)

    def rename(m):
        n = int(m.group(2))
        pre = (b"chromosome_with_a_long_name_", b"exactly16b_", b"seventeen_b_", b"n")[n % 4]
        return m.group(1) + pre + m.group(2)
    gaf2 = re.sub(rb"([<>])" + pat + rb"(\d+)", rename, gaf)
    len2 = re.sub(rb"(^|\n)" + pat + rb"(\d+)", rename, lengths)
    assert gaf2 != gaf
    with tempfile.TemporaryDirectory() as td:
        lp = os.path.join(td, "l.tsv")
        open(lp, "wb").write(len2)
        rc, ref, _ = H.run_tool(CHECK, ["-", "-l", lp], gaf2)
        assert rc == 0 and ref.count(b"\n") > 0
        rc1, out1, st1 = _simt(lp, gaf2)
        assert rc1 == 0 and out1 == ref and st1[0] > 0 and st1[3] == 0, st1
