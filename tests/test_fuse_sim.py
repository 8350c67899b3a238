"""k_fuse (the one-pass kernel, csrc/g2p_fuse.cuh) under the SIMT emulator (build/g2p_simt runs the product's own
kernels on the CPU): byte parity with the reference on every workload shape it takes, on every alignment of a
record against a tile boundary, on the grow-and-rerun path of its output buffer, and on inputs it must hand to the
general pipeline (long records, non-canonical records, errors)."""
import os
import subprocess
import tempfile

import pytest

import helpers as H

SIMT = os.path.join(H.BUILD, "g2p_simt")
TILE = 32768


def simt(gaf, lengths, env=None):
    e = dict(os.environ, G2P_SIMT_STATS="1", G2P_FUSE="2")   # always try k_fuse first (the default does when it pays)
    e.update(env or {})
    with tempfile.TemporaryDirectory() as td:
        lp = os.path.join(td, "l.tsv")
        open(lp, "wb").write(lengths)
        p = subprocess.run([SIMT, "-l", lp, "-"], input=gaf, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=e)
    return p.returncode, p.stdout, p.stderr.decode("latin-1")


@pytest.mark.parametrize("name,count,over", [
    ("short", 12000, {"pct_star": 2}),
    ("short_eqx", 6000, {}),
    ("tagged", 5000, {}),
    ("short", 4000, {"pct_minus": 100, "pct_rev": 100}),
    ("short", 4000, {"pct_minus": 0, "pct_rev": 0, "steps_hi": 8}),
])
def test_fused_kernel_matches_reference(name, count, over):
    p = H.preset(name, seed=17, **over)
    lengths = H.gen_lengths(p)
    gaf = H.gen_records(p, 0, count, threads=2)
    rc, out, err = simt(gaf, lengths)
    rrc, ref, rerr, kind = H.run_gaf2paf_cpu(gaf, lengths)
    assert rc == rrc == 0
    assert "k_fuse converted" in err, err
    assert out == ref


def test_every_alignment_of_a_tile_boundary():
    """The record that straddles the first tile boundary starts 0..N bytes before it, ends exactly at it, begins
    exactly at it..., with and without a final newline."""
    p = H.preset("short", seed=23)
    lengths = H.gen_lengths(p)
    body = H.gen_records(p, 0, 520, threads=1)      # ~70 kB: three tiles
    lines = body.split(b"\n")[:-1]
    first = lines[0].split(b"\t")
    for shift in range(0, 170, 1):
        f = list(first)
        f[0] = b"r" + b"x" * shift
        gaf = b"\n".join([b"\t".join(f)] + lines[1:]) + (b"\n" if shift % 2 else b"")
        rc, out, err = simt(gaf, lengths)
        rrc, ref, rerr, kind = H.run_gaf2paf_cpu(gaf, lengths)
        assert rc == rrc == 0 and "k_fuse converted" in err
        assert out == ref, "shift %d" % shift


@pytest.mark.parametrize("cfg", ["1", "2", "3", "4", "5", "6", "7"])
def test_other_configurations(cfg):
    """The tuning alternatives (24 KiB tiles; 16 KiB tiles with three CTAs per SM) and the dense-input configurations."""
    p = H.preset("short", seed=19, pct_star=1)
    lengths = H.gen_lengths(p)
    gaf = H.gen_records(p, 0, 4000, threads=1)
    rc, out, err = simt(gaf, lengths, {"G2P_FUSE_CFG": cfg})
    rrc, ref, rerr, kind = H.run_gaf2paf_cpu(gaf, lengths)
    assert rc == rrc == 0 and "k_fuse converted" in err and out == ref


def test_output_buffer_grows_and_the_kernel_runs_again():
    p = H.preset("short", seed=29)
    lengths = H.gen_lengths(p)
    gaf = H.gen_records(p, 0, 3000, threads=1)
    rc, out, err = simt(gaf, lengths, {"G2P_FUSE_OUT_CAP": "70000"})
    rrc, ref, rerr, kind = H.run_gaf2paf_cpu(gaf, lengths)
    assert rc == rrc == 0 and "k_fuse converted" in err and out == ref


@pytest.mark.parametrize("kind", ["long", "noncanonical", "name", "abort", "blank"])
def test_falls_back_to_the_general_pipeline(kind):
    p = H.preset("short", seed=31)
    lengths = H.gen_lengths(p)
    gaf = H.gen_records(p, 0, 900, threads=1)
    lines = gaf.split(b"\n")[:-1]
    k = 700
    f = lines[k].split(b"\t")
    if kind == "long":
        pm = H.preset("medium", seed=31)
        lines[k] = H.gen_records(pm, 0, 1, threads=1).rstrip(b"\n")
    elif kind == "noncanonical":
        f[1] = b"0" + f[1]              # leading zero in the query length: printed as a number by the reference
        lines[k] = b"\t".join(f)
    elif kind == "name":
        f[5] = b">nosuchnode" + f[5]
        lines[k] = b"\t".join(f)
    elif kind == "abort":
        f[2] = b""
        lines[k] = b"\t".join(f)
    else:
        lines[k] = b""
    gaf = b"\n".join(lines) + b"\n"
    rc, out, err = simt(gaf, lengths)
    rrc, ref, rerr, _ = H.run_gaf2paf_cpu(gaf, lengths)
    assert "fell back to the general pipeline" in err
    assert rc == rrc
    if rc != 134:
        assert out == ref
    if rc == 1:
        assert err.splitlines()[-1] + "\n" == rerr


def test_default_dispatch_picks_the_kernel_that_pays():
    """G2P_FUSE unset: records of <= 240 bytes stay with the two-pass pipeline (k_rec), 250-500 byte records go to k_fuse."""
    for name, want in (("short", "to k_long"), ("tagged", "k_fuse converted")):
        p = H.preset(name, seed=41)
        lengths = H.gen_lengths(p)
        gaf = H.gen_records(p, 0, 2500, threads=1)
        e = dict(os.environ, G2P_SIMT_STATS="1")
        e.pop("G2P_FUSE", None)
        with tempfile.TemporaryDirectory() as td:
            lp = os.path.join(td, "l.tsv")
            open(lp, "wb").write(lengths)
            pr = subprocess.run([SIMT, "-l", lp, "-"], input=gaf, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=e)
        rrc, ref, rerr, kind = H.run_gaf2paf_cpu(gaf, lengths)
        assert pr.returncode == 0 and pr.stdout == ref
        assert want in pr.stderr.decode(), pr.stderr.decode()


def test_mutation_fuzz_through_k_fuse():
    """tests/fuzz_vs_ref.py with k_fuse forced first: every mutated record is either converted by it identically or handed on."""
    env = dict(os.environ, G2P_FUSE="2")
    pr = subprocess.run([os.sys.executable, os.path.join(H.ROOT, "tests", "fuzz_vs_ref.py"), "--n", "400", "--seed", "3", "--bin", SIMT],
                        stdout=subprocess.PIPE, stderr=subprocess.STDOUT, env=env)
    assert pr.returncode == 0 and b" 0 mismatches" in pr.stdout, pr.stdout[-500:]


def test_general_pipeline_alone_still_matches():
    """G2P_FUSE=0: the two-pass pipeline (k_rec / k_long / k_convert_list / k_emit_lines) in its default configuration,
    including an unterminated last line."""
    p = H.preset("short", seed=37, pct_star=1)
    lengths = H.gen_lengths(p)
    gaf = H.gen_records(p, 0, 5000, threads=1)[:-1]
    rc, out, err = simt(gaf, lengths, {"G2P_FUSE": "0"})
    rrc, ref, rerr, kind = H.run_gaf2paf_cpu(gaf, lengths)
    assert rc == rrc == 0 and "k_fuse" not in err and out == ref
