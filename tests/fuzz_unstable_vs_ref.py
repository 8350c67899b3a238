#!/usr/bin/env python3
"""Mutation fuzzer for gaf2unstable: a converter binary (default: the kernels under the SIMT emulator, build/g2u_simt)
vs the reference gaf2unstable on single malformed or unusual records between two good ones.  Compares the exit code
and stdout (when the reference dies it loses the bytes still in its stdio buffer: then ours must start with its).

    python tests/fuzz_unstable_vs_ref.py [--n 1500] [--seed 1] [--bin build/g2u_simt]

Development tool (needs oracle/_ref).  Tolerated: a zero-length or inverted stable interval, where the reference
indexes an empty vector (undefined behaviour: it usually dies with SIGSEGV, this build reports an assertion abort)."""
import argparse
import os
import random
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H
from fuzz_vs_ref import mutate, WEIRD

REF = os.path.join(ROOT, "oracle", "_ref", "gaf2unstable")
WEIRD_IVL = [":0-0", ":5-5", ":10-5", ":-1-5", ":5", ":", ":5-", ":a-b", ":0-99999999999", ":1-2:3", ""]


def run(binary, gp, data):
    try:
        p = subprocess.run([binary, "-g", gp, "-"], input=data, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=20)
    except subprocess.TimeoutExpired:
        return 999, b""
    rc = p.returncode
    return (128 - rc if rc < 0 else rc), p.stdout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1500)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--bin", default=os.path.join(ROOT, "build", "g2u_simt"))
    a = ap.parse_args()
    rnd = random.Random(a.seed)
    rgfa, gaf = H.gen_rgfa_case(a.seed, n_records=120, aligned=(a.seed % 2 == 0))
    seeds = [l.decode("latin-1") for l in gaf.split(b"\n") if l and not l.startswith(b"*")]
    good = [s for s in seeds if s.split("\t")[5] != "*"][:2]
    bad = tol = 0
    with tempfile.TemporaryDirectory() as td:
        gp = os.path.join(td, "g.gfa")
        open(gp, "wb").write(rgfa)
        for it in range(a.n):
            line = rnd.choice(seeds)
            r = rnd.random()
            if r < 0.3:   # interval surgery on one step
                cols = line.split("\t")
                toks = re.findall(r"[<>][^<>]*", cols[5]) or [cols[5]]
                k = rnd.randrange(len(toks))
                toks[k] = toks[k].split(":")[0] + rnd.choice(WEIRD_IVL)
                cols[5] = "".join(toks)
                line = "\t".join(cols)
            else:
                for _ in range(rnd.randrange(1, 3)):
                    line = mutate(rnd, line.replace(" ", "\x01").replace("\t", " ")).replace("\x01", " ")
            data = (good[0] + "\n" + line + "\n" + good[1] + "\n").encode("latin-1")
            ref = run(REF, gp, data)
            got = run(a.bin, gp, data)
            same = ref[0] == got[0] and (ref[1] == got[1] if ref[0] == 0 else got[1].startswith(ref[1]) or ref[1].startswith(got[1]))
            if not same:
                if ref[0] == 139 and got[0] == 134:
                    tol += 1
                    continue
                bad += 1
                if bad <= 12:
                    print("MISMATCH #%d: %r" % (it, line))
                    print("   ref rc=%d out=%r" % (ref[0], ref[1][-200:]))
                    print("   got rc=%d out=%r" % (got[0], got[1][-200:]))
    print("fuzz gaf2unstable: %d cases, %d mismatches, %d tolerated (reference SIGSEGV on an empty node range)" % (a.n, bad, tol))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
