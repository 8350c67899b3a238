#!/usr/bin/env python3
"""Generate the committed golden vectors by running the UNMODIFIED reference binaries
(oracle/_ref/gaf2paf, oracle/_ref/gaf2unstable; built by oracle/build_ref.sh from
/root/reference).  The reference ships no golden files for this path (SURVEY.md §4,
§8c), so these known-answer vectors pin the byte-level behaviour instead.

    python tests/golden/make_golden.py        # rewrites tests/golden/*.json

Only needed in the build container; the JSON files are what travels to the GPU box.
"""
import json
import os
import random
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref")

KAT_LENGTHS = "chrA\t1000\t6\t60\t61\nchrB\t500\nchrC\t300\na\t10\nb\t5\nc\t10\nn2\t20\nn3\t50\nz\t0\n"


def T(s):
    return s.replace(" ", "\t")


# SURVEY.md Appendix A (K1-K23) + extra edge cases found while reading the reference.
KAT = [
    ("K1-two-interval-steps-fwd", T("q1 200 5 132 + >chrA:100-200>chrB:0-50 150 10 138 120 130 60 tp:A:P cm:i:5 cg:Z:50M2I40M3D35M")),
    ("K2-second-step-reversed-mapq255", T("q1 200 5 132 + >chrA:100-200<chrB:0-50 150 10 138 120 130 255 tp:A:P cg:Z:50M2I40M3D35M")),
    ("K3-minus-strand-record", T("q1 200 5 132 - >chrA:100-200<chrB:0-50 150 10 138 120 130 0 cg:Z:50M2I40M3D35M")),
    ("K4-whole-contig-minus-rc-before-tp", T("q1 200 5 132 - chrC 300 10 138 120 130 60 rc:Z:foo tp:A:S cg:Z:50M2I40M3D35M")),
    ("K5-node-steps-reverse-middle", T("q1 200 5 132 + >chrA:0-100<n2>n3 170 10 138 120 130 60 cg:Z:50M2I40M3D35M")),
    ("K6-insertion-at-boundary-goes-to-next-step", T("q 100 0 28 + >a>b>c 25 0 25 25 28 60 cg:Z:10M3I5M10M")),
    ("K7-leading-I-kept-trailing-I-dropped", T("q 100 0 31 + >a>b>c 25 0 25 25 31 60 cg:Z:2I10M5M10M4I")),
    ("K8-one-op-spans-three-steps", T("q 100 0 25 + >a<b>c 25 0 25 25 25 60 cg:Z:25M")),
    ("K9-N-S-H-P-X-ops", T("q 100 0 24 + >a>b>c 25 0 25 20 25 60 cg:Z:2S2H8M2N3P5M7M3X")),
    ("K10-minus-strand-clipped-both-ends-mapq300", T("q 100 10 27 - >a<b>c 25 3 20 17 17 300 cg:Z:17M")),
    ("K11-deletion-only-step-suppressed", T("q 100 0 20 + >a>b>c 25 0 25 18 25 60 cg:Z:8=2X5D10=")),
    ("K12-mismatch-only-step-suppressed-query-still-advances", T("q 100 0 25 + >a>b>c 25 0 25 18 25 60 cg:Z:10=5X10=")),
    ("K13-gi-zero-blocklen", T("q1 200 5 132 + chrC 300 10 138 120 0 60 cg:Z:50M2I40M3D35M")),
    ("K14-gi-exponent", T("q1 200 5 132 + chrC 300 10 138 1300000 1 60 cg:Z:50M2I40M3D35M")),
    ("K15-gi-rounds-to-0.001", T("q1 200 5 132 + chrC 300 10 138 1 2000 60 cg:Z:50M2I40M3D35M")),
    ("K16-gi-rounds-to-0", T("q1 200 5 132 + chrC 300 10 138 1 3000 60 cg:Z:50M2I40M3D35M")),
    ("K17-star-matches-and-blocklen", T("q1 200 5 132 + chrC 300 10 138 * * 60 cg:Z:50M2I40M3D35M")),
    ("K18-cigar-longer-than-path-truncated", T("q 100 0 15 + >a>b 15 0 15 15 15 60 cg:Z:20M")),
    ("K19-cigar-too-short-abort", T("q 100 0 15 + >a>b 15 0 15 15 15 60 cg:Z:14M")),
    ("K20-unknown-name-exit1", T("q 100 0 15 + >zzz:0-15 15 0 15 15 15 60 cg:Z:15M")),
    ("K21-no-cg-exit1", T("q 100 0 15 + >a>b 15 0 15 15 15 60 tp:A:P")),
    ("K22-empty-column-abort", T("q  0 15 + >a>b 15 0 15 15 15 60 cg:Z:15M")),
    ("K23-star-line-skipped", T("* >s43 97 12 0 6 92")),
    # extra
    ("X1-zero-length-middle-step", T("q 100 0 23 + >a>z>c 20 0 20 20 23 60 cg:Z:10M3I10M")),
    ("X2-minus-strand-node-steps-with-indels", T("q 100 3 30 - >a<b>c 25 2 24 20 27 13 tp:A:P cg:Z:4M2I6M3D5M1I4M")),
    ("X3-gi-negative", T("q1 200 5 132 + chrC 300 10 138 * 7 60 cg:Z:50M2I40M3D35M")),
    ("X4-gi-big-ratio", T("q1 200 5 132 + chrC 300 10 138 1234565 1000 60 cg:Z:50M2I40M3D35M")),
    ("X5-gi-large", T("q1 200 5 132 + chrC 300 10 138 9223372036854775807 1 60 cg:Z:50M2I40M3D35M")),
    ("X6-star-qlen-and-mapq", T("q1 * 5 132 + chrC 300 10 138 120 130 * cg:Z:50M2I40M3D35M")),
    ("X7-unknown-name-second-step-partial-output", T("q 100 0 25 + >a>nope>c 25 0 25 25 25 60 cg:Z:25M")),
    ("X8-minus-unknown-names-last-reported", T("q 100 0 25 - >nope1>a>nope2 25 0 25 25 25 60 cg:Z:25M")),
    ("X9-duplicate-tag-abort", T("q 100 0 15 + >a>b 15 0 15 15 15 60 tp:A:P tp:A:S cg:Z:15M")),
    ("X10-short-tag-abort", T("q 100 0 15 + >a>b 15 0 15 15 15 60 ab:1 cg:Z:15M")),
    ("X11-bad-strand-abort", T("q 100 0 15 x >a>b 15 0 15 15 15 60 cg:Z:15M")),
    ("X12-star-strand-abort", T("q 100 0 15 * >a>b 15 0 15 15 15 60 cg:Z:15M")),
    ("X13-range-without-dash-abort", T("q 100 0 15 + >chrA:5 15 0 15 15 15 60 cg:Z:15M")),
    ("X14-star-path", T("q 100 0 15 + * 15 0 15 15 15 60 cg:Z:15M")),
    ("X15-trailing-tab-and-empty-tag", T("q 100 0 15 + >a>b 15 0 15 15 15 60 tp:A:P  cg:Z:15M ")),
    ("X16-long-tag-names", T("q 100 0 15 + >a>b 15 0 15 15 15 60 tpx:A:P cgg:Z:1M cg:Z:15M rc:Z:id=x|y:z")),
    ("X17-leading-zeros-and-zero-op", T("q 0100 00 15 + >a>b 15 0 15 15 15 060 cg:Z:007M0I0M08M")),
    ("X18-junk-after-ints", T("q 100x 0 15abc + >a:0-10zz>b 15 0 15 15 15 60 cg:Z:15M")),
    ("X19-too-few-columns-abort", T("q 100 0 15 + >a>b 15 0 15")),
    ("X20-clip-negative-abort", T("q 100 0 15 + >a>b 15 0 16 15 15 60 cg:Z:16M")),
    ("X21-reversed-step-multi-op", T("q 100 0 30 + <c<b<a 25 1 24 20 28 60 cg:Z:3M2I4M1D2M3I10M2D1M")),
    ("X22-minus-whole-contig-stable", T("q1 200 5 132 - chrC 300 10 138 120 130 60 cg:Z:50M2I40M3D35M")),
    ("X23-long-name-over-16-bytes", T("q 100 0 15 + >averyveryverylongname_12345:0-15 15 0 15 15 15 60 cg:Z:15M")),
    ("X24-empty-cg-plus", T("q 100 0 15 + >a>b 15 0 15 15 15 60 cg:Z:")),
    ("X25-cg-bad-op-abort", T("q 100 0 15 + >a>b 15 0 15 15 15 60 cg:Z:15Q")),
    ("X26-cg-trailing-digits-abort", T("q 100 0 15 + >a>b 15 0 15 15 15 60 cg:Z:15M3")),
]
KAT_LENGTHS_FULL = KAT_LENGTHS + "averyveryverylongname_12345\t77\n"

UNSTABLE_RGFA = "\n".join(
    [
        T("S s1 %s LN:i:100 SN:Z:chrA SO:i:0 SR:i:0"),
        T("S s2 %s LN:i:20 SN:Z:chrA SO:i:100 SR:i:0"),
        T("S s3 %s LN:i:50 SN:Z:chrA SO:i:120 SR:i:0"),
        T("S s4 %s LN:i:30 SN:Z:HG.ctg1 SO:i:500 SR:i:1"),
        T("L s1 + s2 + 0M SR:i:0 L1:i:100 L2:i:20"),
        T("L s2 + s3 + 0M SR:i:0 L1:i:20 L2:i:50"),
        T("L s1 + s4 + 0M SR:i:1 L1:i:100 L2:i:30"),
        T("L s4 + s3 + 0M SR:i:1 L1:i:30 L2:i:50"),
    ]
) + "\n"

UNSTABLE_GAF = [
    T("q1 200 5 132 + >chrA:0-120>chrA:120-170 170 10 138 120 130 60 tp:A:P cm:i:5 cg:Z:50M2I40M3D35M"),
    T("q2 200 0 60 + >chrA:0-100>HG.ctg1:500-530>chrA:120-170 180 90 150 60 60 255 tp:A:S cg:Z:60M"),
    T("q3 200 0 60 - <chrA:120-170<HG.ctg1:500-530<chrA:0-100 180 30 90 60 60 7 cg:Z:60M ds:Z:foo"),
    T("q4 200 5 132 + chrA 170 10 138 120 130 60 tp:A:P cg:Z:50M2I40M3D35M"),
    T("q5 200 5 33 - chrA 170 105 133 28 28 60 tp:A:P cg:Z:28M"),
    T("q6 200 0 60 + * * * * * * 255 tp:A:P"),
    T("* >s43 97 12 0 6 92"),
]


def run(cmd, stdin_bytes):
    p = subprocess.run(cmd, input=stdin_bytes, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    rc = p.returncode
    if rc < 0:
        rc = 128 - rc   # killed by signal N -> shell convention 128+N (SIGABRT -> 134)
    return rc, p.stdout.decode("latin-1"), p.stderr.decode("latin-1")


def rgfa_text():
    rnd = random.Random(7)
    seqs = ["".join(rnd.choice("ACGT") for _ in range(n)) for n in (100, 20, 50, 30)]
    parts = UNSTABLE_RGFA.split("%s")
    out = parts[0]
    for s, p in zip(seqs, parts[1:]):
        out += s + p
    return out


def main():
    if not os.path.exists(os.path.join(REF, "gaf2paf")):
        sys.exit("oracle/_ref/gaf2paf missing: run oracle/build_ref.sh first")
    with tempfile.TemporaryDirectory() as td:
        lp = os.path.join(td, "kat.len.tsv")
        with open(lp, "w") as f:
            f.write(KAT_LENGTHS_FULL)
        vectors = []
        for name, line in KAT:
            rc, out, err = run([os.path.join(REF, "gaf2paf"), "-", "-l", lp], (line + "\n").encode("latin-1"))
            vectors.append({"name": name, "in": line, "rc": rc, "out": out, "err": err})
        # all rc==0 vectors in one stream as well (record order, multi-record input)
        ok_lines = [v["in"] for v in vectors if v["rc"] == 0]
        rc, out, err = run([os.path.join(REF, "gaf2paf"), "-", "-l", lp], ("\n".join(ok_lines) + "\n").encode("latin-1"))
        assert rc == 0
        doc = {"lengths": KAT_LENGTHS_FULL, "vectors": vectors, "stream": {"in": ok_lines, "out": out}}
        with open(os.path.join(HERE, "gaf2paf_kat.json"), "w") as f:
            json.dump(doc, f, indent=1)
        print("gaf2paf_kat.json: %d vectors" % len(vectors))

        # gaf2unstable
        gp = os.path.join(td, "g.gfa")
        with open(gp, "w") as f:
            f.write(rgfa_text())
        nl = os.path.join(td, "nl.tsv")
        uv = []
        for line in UNSTABLE_GAF:
            rc, out, err = run([os.path.join(REF, "gaf2unstable"), "-", "-g", gp, "-o", nl], (line + "\n").encode("latin-1"))
            uv.append({"in": line, "rc": rc, "out": out, "err": err})
        rc, out, err = run([os.path.join(REF, "gaf2unstable"), "-", "-g", gp, "-o", nl], ("\n".join(UNSTABLE_GAF) + "\n").encode("latin-1"))
        assert rc == 0, err
        node_lengths = open(nl).read()
        rc2, paf, err2 = run([os.path.join(REF, "gaf2paf"), "-", "-l", nl], out.encode("latin-1"))
        doc = {"rgfa": rgfa_text(), "vectors": uv, "stream": {"in": UNSTABLE_GAF, "out": out},
               "node_lengths": node_lengths, "paf": {"rc": rc2, "out": paf}}
        with open(os.path.join(HERE, "gaf2unstable_kat.json"), "w") as f:
            json.dump(doc, f, indent=1)
        print("gaf2unstable_kat.json: %d vectors" % len(uv))


if __name__ == "__main__":
    main()
