"""BASELINE configs[0] / configs[1] substitutes: a stable-coordinate GAF and an rGFA synthesised
from the reference's own test sequences (test/hpp-20-2M/*.fa.gz; fixture tests/golden/hpp20_case.json.gz, generator
tests/golden/make_hpp20_case.py), checked with the reference's own acceptance property -- check_cigar of
test/verify_matches.py:40-92, restated in tests/check_cigar.py -- for the three scenarios of test/gaf2paf.t:31-67:

    gaf2paf CHM13.gaf -l all.fa.fai                                      -> verify against the FASTA sequences
    gaf2unstable CHM13.gaf -g graph.gfa -o L | gaf2paf - -l L            -> verify against the node sequences

on the forward graph ("hpp") and on the reverse-strand graph ("hg38rev").  The CPU tests pin the fixture, the property
restatement and the emulated kernels to the reference binary; the GPU tests run the product through the C-ABI."""
import gzip
import json
import os
import subprocess
import tempfile

import pytest

import helpers as H
from check_cigar import check_paf


@pytest.fixture(scope="module")
def case():
    with gzip.open(os.path.join(H.GOLDEN, "hpp20_case.json.gz"), "rt") as f:
        return json.load(f)


def node_sequences(rgfa):
    d = {}
    for line in rgfa.split("\n"):
        if line.startswith("S\t"):
            t = line.split("\t")
            d[t[1]] = t[2]
    return d


def ref_two_stage(gaf, rgfa):
    rc, ugaf, err, nl = H.run_gaf2unstable_ref(gaf, rgfa, want_lengths=True)
    assert rc == 0, err
    rc2, paf, err2, kind = H.run_gaf2paf_cpu(ugaf, nl)
    assert rc2 == 0, err2
    return ugaf, nl, paf


@pytest.mark.parametrize("graph", ["hpp", "hg38rev"])
def test_reference_output_satisfies_the_property(case, graph):
    """Pins the fixture and tests/check_cigar.py: the reference's own output passes its own acceptance test."""
    g = case["graphs"][graph]
    gaf, rgfa, fai = g["gaf"].encode(), g["rgfa"].encode(), case["fai"].encode()
    rc, paf, err, kind = H.run_gaf2paf_cpu(gaf, fai)
    assert rc == 0, err
    n_records = sum(1 for ln in g["gaf"].split("\n") if ln and not ln.startswith("*"))
    assert check_paf(paf, case["fasta"]) >= n_records
    ugaf, nl, upaf = ref_two_stage(gaf, rgfa)
    fa = dict(case["fasta"])
    fa.update(node_sequences(g["rgfa"]))
    assert check_paf(upaf, fa) >= n_records
    # the property is not vacuous: one substituted base in a query is caught (node-space PAF: the targets are the
    # node sequences, separate entries from the query even when the query is the graph's own backbone)
    bad = dict(fa)
    line = upaf.split(b"\n")[0].decode().split("\t")
    q, pos = line[0], int(line[2]) + 5
    bad[q] = bad[q][:pos] + ("A" if bad[q][pos].upper() != "A" else "C") + bad[q][pos + 1:]
    with pytest.raises(AssertionError):
        check_paf(upaf, bad)


@pytest.mark.parametrize("graph", ["hpp", "hg38rev"])
def test_emulated_kernels_match_reference(case, graph):
    """The product kernels under the SIMT emulator on the stable GAF (contig names longer than 16 bytes, interval
    steps) and on the node-space GAF the reference's gaf2unstable makes of it."""
    g = case["graphs"][graph]
    gaf, rgfa, fai = g["gaf"].encode(), g["rgfa"].encode(), case["fai"].encode()
    simt = os.path.join(H.BUILD, "g2p_simt")
    ugaf, nl, upaf = ref_two_stage(gaf, rgfa)
    rc, paf, err, kind = H.run_gaf2paf_cpu(gaf, fai)
    for text, table, want in ((gaf, fai, paf), (ugaf, nl, upaf)):
        with tempfile.TemporaryDirectory() as td:
            lp = os.path.join(td, "l.tsv")
            open(lp, "wb").write(table)
            p = subprocess.run([simt, "-l", lp, "-"], input=text, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        assert p.returncode == 0 and p.stdout == want


@pytest.mark.gpu
@pytest.mark.parametrize("graph", ["hpp", "hg38rev"])
def test_gpu_gaf2paf_and_two_stage(case, graph, g2p):
    g = case["graphs"][graph]
    gaf, rgfa, fai = g["gaf"].encode(), g["rgfa"].encode(), case["fai"].encode()
    cv = g2p.Converter(0)
    try:
        # gaf2paf CHM13.gaf -l all.fa.fai
        assert cv.load_lengths(fai)
        paf, res = cv.convert_host(gaf)
        assert g2p.exit_code(res) == 0
        rc, ref, err, kind = H.run_gaf2paf_cpu(gaf, fai)
        assert rc == 0 and paf == ref
        assert check_paf(paf, case["fasta"]) > 0
        # gaf2unstable -g graph.gfa -o L | gaf2paf - -l L
        ok, code, msg = cv.load_rgfa(rgfa)
        assert ok, msg
        ugaf, ures, warns = cv.unstable_host(gaf)
        assert g2p.exit_code(ures) == 0
        nl = cv.node_lengths()
        rugaf, rnl, rupaf = ref_two_stage(gaf, rgfa)
        assert ugaf == rugaf and nl == rnl
        assert cv.load_lengths(nl)
        upaf, res2 = cv.convert_host(ugaf)
        assert g2p.exit_code(res2) == 0 and upaf == rupaf
        fa = dict(case["fasta"])
        fa.update(node_sequences(g["rgfa"]))
        assert check_paf(upaf, fa) > 0
    finally:
        cv.close()
