"""Dependency-free restatement of the reference's only acceptance property for the gaf2paf path,
`check_cigar` of /root/reference/test/verify_matches.py:40-92 (used by test/gaf2paf.t:26-67): for every PAF line, the
query / target lengths in columns 2 / 7 equal the FASTA sequence lengths, and walking the cg:Z CIGAR over
query[qs:qe] and target[ts:te] (target reverse-complemented and the op list reversed on '-' lines) every M / = run
is an exact match, with both sequences consumed exactly.  TEST INFRASTRUCTURE ONLY."""
import re

_COMP = str.maketrans("ACGTNacgtn", "TGCANtgcan")


def revcomp(s):
    return s[::-1].translate(_COMP)


def check_cigar(paf_line, fa, min_identity=1.0):
    toks = paf_line.rstrip("\n").split("\t")
    cigar = toks[-1]
    assert cigar[:4] == "cg:Z", paf_line
    qs, qe, ts, te = int(toks[2]), int(toks[3]), int(toks[7]), int(toks[8])
    qname, tname = toks[0], toks[5]
    assert qname in fa, "query %s not in the sequences" % qname
    assert tname in fa, "target %s not in the sequences" % tname
    q = fa[qname][qs:qe]
    assert len(q) == qe - qs and len(fa[qname]) == int(toks[1]), paf_line
    t = fa[tname][ts:te]
    assert len(t) == te - ts and len(fa[tname]) == int(toks[6]), paf_line
    assert toks[4] in ("+", "-")
    ops = re.findall("([0-9]+)(=|X|M|D|I)", cigar[5:])
    if toks[4] == "-":
        t = revcomp(t)
        ops = ops[::-1]
    qp = tp = 0
    for ln, op in ops:
        ln = int(ln)
        if op in ("M", "="):
            a, b = q[qp:qp + ln].upper(), t[tp:tp + ln].upper()
            assert len(a) == len(b) == ln, paf_line
            same = sum(1 for x, y in zip(a, b) if x == y or (min_identity < 1 and "N" in (x, y)))
            iden = same / float(ln) if ln else 1.0
            assert not ((min_identity == 1 and iden < 1) or (ln > 100 and iden < min_identity)), \
                "identity %.4f in %d%s at query %d target %d of\n%s" % (iden, ln, op, qp, tp, paf_line)
        if op != "I":
            tp += ln
        if op != "D":
            qp += ln
    assert qp == qe - qs and tp == te - ts, paf_line


def check_paf(paf_text, fa, min_identity=1.0):
    """Every line of a PAF (bytes or str); returns the number of lines checked."""
    if isinstance(paf_text, bytes):
        paf_text = paf_text.decode("latin-1")
    n = 0
    for line in paf_text.split("\n"):
        if line:
            check_cigar(line, fa, min_identity)
            n += 1
    return n
