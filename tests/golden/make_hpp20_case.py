#!/usr/bin/env python3
"""Generates tests/golden/hpp20_case.json.gz from the reference's own test sequences
(/root/reference/test/hpp-20-2M/*.fa.gz; run in the build container, the GPU box has no /root/reference).

The reference's test/gaf2paf.t builds its inputs with minigraph / gfatools / samtools, none of which exist here
(SURVEY.md §4), so BASELINE configs[0] / configs[1] are substituted by a GAF and an rGFA SYNTHESISED from slices of the
same FASTA records with the structure minigraph gives them:

  graph "hpp"      rank-0 chain of CHM13 nodes (both scaffolds), rank-1 bubble nodes cut from HG003 / HG004
  graph "hg38rev"  rank-0 chain of hg38.chr20.reversed nodes, rank-1 bubbles cut from CHM13
  stable GAFs      queries CHM13 (self), hg38.chr20 against hg38rev ('-' strand, and '+' strand over all-'<' paths),
                   paths of node-aligned stable intervals `>contig:a-b`, skipped nodes (I), bubble detours (D),
                   whole-contig paths and '*' lines; CIGARs of = / M runs that are exact matches by construction

so that the reference's acceptance property (tests/check_cigar.py) must hold on the PAF of both `gaf2paf` and
`gaf2unstable | gaf2paf` (the three scenarios of test/gaf2paf.t:31-67: forward, reverse strand, another assembly)."""
import gzip
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/test/hpp-20-2M"
_COMP = str.maketrans("ACGTNacgtn", "TGCANtgcan")


def revcomp(s):
    return s[::-1].translate(_COMP)


def load(path):
    d, name = {}, None
    for line in gzip.open(path, "rt"):
        line = line.rstrip()
        if line.startswith(">"):
            name = line[1:].split()[0]
            d[name] = []
        else:
            d[name].append(line)
    return {k: "".join(v) for k, v in d.items()}


def build_graph(rnd, chains, bubble_src, prefix):
    """chains: [(contig name, sequence)] -> rank-0 nodes; bubble_src: [(contig name, sequence)] -> rank-1 nodes.
    Returns (rgfa text, contigs {name: [(node, offset, length)]}, bubbles [(node, contig, offset, length, left, right)], node_seq)."""
    S, L, contigs, node_seq, bubbles = [], [], {}, {}, []
    nid = 1
    for name, seq in chains:
        nodes, off = [], 0
        while off < len(seq):
            ln = min(len(seq) - off, rnd.randrange(2000, 12000))
            if len(seq) - off - ln < 1500:
                ln = len(seq) - off
            node = "%s%d" % (prefix, nid)
            nid += 1
            nodes.append((node, off, ln))
            node_seq[node] = seq[off:off + ln]
            S.append("S\t%s\t%s\tLN:i:%d\tSN:Z:%s\tSO:i:%d\tSR:i:0" % (node, node_seq[node], ln, name, off))
            if len(nodes) > 1:
                L.append("L\t%s\t+\t%s\t+\t0M\tSR:i:0\tL1:i:%d\tL2:i:%d" % (nodes[-2][0], node, nodes[-2][2], ln))
            off += ln
        contigs[name] = nodes
    for bname, bseq in bubble_src:
        off = rnd.randrange(0, 3000)
        while off + 500 < len(bseq) and len(bubbles) < 40:
            ln = rnd.randrange(300, 4000)
            ln = min(ln, len(bseq) - off)
            cname = rnd.choice(list(contigs))
            i = rnd.randrange(0, len(contigs[cname]) - 1)
            node = "%s%d" % (prefix, nid)
            nid += 1
            node_seq[node] = bseq[off:off + ln]
            S.append("S\t%s\t%s\tLN:i:%d\tSN:Z:%s\tSO:i:%d\tSR:i:1" % (node, node_seq[node], ln, bname, off))
            left, right = contigs[cname][i][0], contigs[cname][i + 1][0]
            L.append("L\t%s\t+\t%s\t+\t0M\tSR:i:1\tL1:i:%d\tL2:i:%d" % (left, node, contigs[cname][i][2], ln))
            L.append("L\t%s\t+\t%s\t+\t0M\tSR:i:1\tL1:i:%d\tL2:i:%d" % (node, right, ln, contigs[cname][i + 1][2]))
            bubbles.append((node, bname, off, ln, cname, i))
            off += ln + rnd.randrange(1000, 6000)
    return "\n".join(S + L) + "\n", contigs, bubbles, node_seq


def make_records(rnd, contigs, bubbles, chain_name, qname, qlen, mode, count, tag):
    """Records of query `qname` against paths over the rank-0 chain `chain_name`.
    mode 'fwd'  : query == chain sequence,        '+' strand, '>' steps
    mode 'minus': query == revcomp(chain),        '-' strand, '>' steps
    mode 'rev'  : query == revcomp(chain),        '+' strand, '<' steps in reverse order"""
    nodes = contigs[chain_name]
    by_gap = {}
    for b in bubbles:
        if b[4] == chain_name:
            by_gap.setdefault(b[5], []).append(b)
    recs = []
    for r in range(count):
        n = rnd.randrange(1, min(18, len(nodes)) + 1)
        i0 = rnd.randrange(0, len(nodes) - n + 1)
        steps, ops = [], []   # steps: (contig, a, b); ops: [op, len] along the path, forwards
        qcov = 0
        k = i0
        while k < i0 + n:
            node, off, ln = nodes[k]
            inner = i0 < k < i0 + n - 1
            if inner and rnd.random() < 0.12:      # the path skips this node: its bases are an insertion in the query
                ops.append(["I", ln])
                k += 1
                continue
            steps.append((chain_name, off, off + ln))
            ops.append(["=", ln])
            if k + 1 < i0 + n and k in by_gap and rnd.random() < 0.5:   # detour through a bubble: deletion
                b = rnd.choice(by_gap[k])
                steps.append((b[1], b[2], b[2] + b[3]))
                ops.append(["D", b[3]])
            k += 1
        first_len = steps[0][2] - steps[0][1]
        last_len = steps[-1][2] - steps[-1][1]
        ps = rnd.randrange(0, first_len)
        ec = rnd.randrange(0, last_len) if len(steps) > 1 else rnd.randrange(0, first_len - ps)
        # clip the first / last '=' run (first and last steps are always rank-0 steps of the chain)
        ops[0][1] -= ps
        ops[-1][1] -= ec
        assert ops[0][0] == "=" and ops[-1][0] == "=" and ops[0][1] > 0 and ops[-1][1] > 0
        merged = []
        for op, ln in ops:
            if merged and merged[-1][0] == op:
                merged[-1][1] += ln
            else:
                merged.append([op, ln])
        # long exact runs are split into =, and some are written as M, like a real aligner's output
        cg = []
        for op, ln in merged:
            if op == "=" and ln > 50 and rnd.random() < 0.3:
                a = rnd.randrange(1, ln)
                cg += [("M", a), ("=", ln - a)]
            else:
                cg.append((op if op != "=" or rnd.random() < 0.7 else "M", ln))
        total = sum(b - a for _, a, b in steps)
        pe = total - ec
        c_lo = nodes[i0][1] + ps                       # chain coordinates covered by the alignment
        c_hi = nodes[i0 + n - 1][1] + nodes[i0 + n - 1][2] - ec
        nm = sum(ln for op, ln in cg if op in "=M")
        nb = sum(ln for op, ln in cg)
        if mode == "fwd":
            qs, qe, strand = c_lo, c_hi, "+"
            path = "".join(">%s:%d-%d" % s for s in steps)
            cgs = "".join("%d%s" % (ln, op) for op, ln in cg)
            pps, ppe = ps, pe
        elif mode == "minus":
            qs, qe, strand = qlen - c_hi, qlen - c_lo, "-"
            path = "".join(">%s:%d-%d" % s for s in steps)
            cgs = "".join("%d%s" % (ln, op) for op, ln in cg)
            pps, ppe = ps, pe
        else:
            qs, qe, strand = qlen - c_hi, qlen - c_lo, "+"
            path = "".join("<%s:%d-%d" % s for s in reversed(steps))
            cgs = "".join("%d%s" % (ln, op) for op, ln in reversed(cg))
            pps, ppe = ec, total - ps
        assert qe - qs == sum(ln for op, ln in cg if op in "=MI")
        tags = ["tp:A:%s" % rnd.choice("PS"), "cm:i:%d" % rnd.randrange(5, 900), "s1:i:%d" % rnd.randrange(100, 90000),
                "s2:i:%d" % rnd.randrange(0, 5000), "dv:f:0.%04d" % rnd.randrange(0, 300)]
        rnd.shuffle(tags)
        tags = tags[:rnd.randrange(0, 6)]
        pos = rnd.randrange(0, len(tags) + 1)
        cols = [qname, str(qlen), str(qs), str(qe), strand, path, str(total), str(pps), str(ppe), str(nm), str(nb),
                str(rnd.choice([0, 3, 60, 255]))] + tags[:pos] + ["cg:Z:" + cgs] + tags[pos:]
        recs.append("\t".join(cols))
    return recs


def whole_contig_records(rnd, contigs, chain_name, seq_len, count):
    """`q ... + contig plen ps pe` with a bare stable name as the path (minigraph writes these for alignments that stay
    on one stable sequence); the query is the chain itself."""
    recs = []
    for _ in range(count):
        a = rnd.randrange(0, seq_len - 2000)
        b = rnd.randrange(a + 500, min(seq_len, a + 40000))
        recs.append("\t".join([chain_name, str(seq_len), str(a), str(b), "+", chain_name, str(seq_len), str(a), str(b), str(b - a), str(b - a), "60",
                               "tp:A:P", "cg:Z:%d=" % (b - a)]))
    return recs


def main():
    rnd = random.Random(20)
    chm, hg3, hg4 = load(SRC + "/CHM13.fa.gz"), load(SRC + "/HG003.fa.gz"), load(SRC + "/HG004.fa.gz")
    hg38 = load(SRC + "/hg38.fa.gz")
    fa = {
        "CHM13.CHM13_Super-Scaffold_117": chm["CHM13.CHM13_Super-Scaffold_117"][1000000:1110000],
        "CHM13.CHM13_Super-Scaffold_100037": chm["CHM13.CHM13_Super-Scaffold_100037"][100000:130000],
        "HG003.HG003_h1tg000030l": hg3["HG003.HG003_h1tg000030l"][1000000:1035000],
        "HG004.HG004_h2tg000013l": hg4["HG004.HG004_h2tg000013l"][1000000:1035000],
        "hg38.chr20": hg38["hg38.chr20"][1000000:1070000],
    }
    fa["hg38.chr20.reversed"] = revcomp(fa["hg38.chr20"])    # what hg38-rev.fa.gz holds for the full record (checked)
    c117, c100 = "CHM13.CHM13_Super-Scaffold_117", "CHM13.CHM13_Super-Scaffold_100037"
    g1, contigs1, bub1, nseq1 = build_graph(rnd, [(c117, fa[c117]), (c100, fa[c100])],
                                            [("HG003.HG003_h1tg000030l", fa["HG003.HG003_h1tg000030l"]), ("HG004.HG004_h2tg000013l", fa["HG004.HG004_h2tg000013l"])], "s")
    g2, contigs2, bub2, nseq2 = build_graph(rnd, [("hg38.chr20.reversed", fa["hg38.chr20.reversed"])], [(c100, fa[c100])], "s")
    star = "*\t>s3\t97\t12\t0\t6\t92"
    # scenario 1 (gaf2paf.t:31-40): CHM13 aligned back to the graph built from it -- forward
    gaf1 = make_records(rnd, contigs1, bub1, c117, c117, len(fa[c117]), "fwd", 60, "a") + [star] + \
        make_records(rnd, contigs1, bub1, c100, c100, len(fa[c100]), "fwd", 25, "b") + whole_contig_records(rnd, contigs1, c117, len(fa[c117]), 6)
    # scenarios 2 / 3 (gaf2paf.t:44-66): hg38 against the graph whose rank-0 chain is its reverse complement
    L38 = len(fa["hg38.chr20"])
    gaf2 = make_records(rnd, contigs2, bub2, "hg38.chr20.reversed", "hg38.chr20", L38, "minus", 50, "c") + [star] + \
        make_records(rnd, contigs2, bub2, "hg38.chr20.reversed", "hg38.chr20", L38, "rev", 40, "d") + \
        make_records(rnd, contigs2, bub2, "hg38.chr20.reversed", "hg38.chr20.reversed", L38, "fwd", 15, "e")
    case = {
        "note": "generated by tests/golden/make_hpp20_case.py from slices of /root/reference/test/hpp-20-2M/*.fa.gz",
        "fasta": fa,
        "fai": "".join("%s\t%d\t0\t60\t61\n" % (k, len(v)) for k, v in fa.items()),
        "graphs": {
            "hpp": {"rgfa": g1, "gaf": "\n".join(gaf1) + "\n"},
            "hg38rev": {"rgfa": g2, "gaf": "\n".join(gaf2) + "\n"},
        },
    }
    out = os.path.join(HERE, "hpp20_case.json.gz")
    with gzip.GzipFile(out, "wb", mtime=0) as f:
        f.write(json.dumps(case, sort_keys=True).encode())
    print("%s: %d bytes; %d + %d GAF records; %d + %d nodes" % (out, os.path.getsize(out), len(gaf1), len(gaf2), len(nseq1), len(nseq2)))


if __name__ == "__main__":
    sys.exit(main())
