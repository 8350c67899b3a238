"""Shared test / bench helpers: synthetic generator binding, oracle runners, golden vectors.

Everything here is test infrastructure.  The oracle (oracle/bin/*, oracle/_ref/*) is only
ever used as the checker or as the timed CPU baseline — never by the product path."""
import ctypes
import json
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
ORACLE_BIN = os.path.join(ROOT, "oracle", "bin")
REF_BIN = os.path.join(ROOT, "oracle", "_ref")
BUILD = os.path.join(ROOT, "build")


class GenParams(ctypes.Structure):
    _fields_ = [("seed", ctypes.c_uint64)] + [(k, ctypes.c_uint32) for k in (
        "n_nodes", "node_len_lo", "node_len_hi", "steps_lo", "steps_hi", "mrun_lo", "mrun_hi", "indel_lo", "indel_hi",
        "max_runs", "pct_rev", "pct_minus", "use_eqx", "stable", "qlen_min", "pct_star")]


_gen = None


def gen_lib():
    global _gen
    if _gen is None:
        path = os.path.join(BUILD, "libgafgen.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-C", ROOT, "build/libgafgen.so"], stdout=subprocess.DEVNULL)
        _gen = ctypes.CDLL(path)
        _gen.gafgen_preset.argtypes = [ctypes.c_char_p, ctypes.POINTER(GenParams)]
        _gen.gafgen_preset.restype = None
        _gen.gafgen_lengths.argtypes = [ctypes.POINTER(GenParams), ctypes.c_void_p, ctypes.c_size_t]
        _gen.gafgen_lengths.restype = ctypes.c_size_t
        _gen.gafgen_records.argtypes = [ctypes.POINTER(GenParams), ctypes.c_uint64, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
        _gen.gafgen_records.restype = ctypes.c_size_t
        _gen.gafgen_records_alloc.argtypes = [ctypes.POINTER(GenParams), ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int, ctypes.POINTER(ctypes.c_size_t)]
        _gen.gafgen_records_alloc.restype = ctypes.c_void_p
        _gen.gafgen_free.argtypes = [ctypes.c_void_p]
        _gen.gafgen_free.restype = None
    return _gen


def preset(name, seed=1, **over):
    p = GenParams()
    gen_lib().gafgen_preset(name.encode(), ctypes.byref(p))
    if p.n_nodes == 0:
        raise ValueError("unknown preset " + name)
    p.seed = seed
    for k, v in over.items():
        setattr(p, k, v)
    return p


def gen_lengths(p):
    n = gen_lib().gafgen_lengths(ctypes.byref(p), None, 0)
    buf = ctypes.create_string_buffer(n)
    gen_lib().gafgen_lengths(ctypes.byref(p), buf, n)
    return buf.raw


def gen_records_raw(p, first, count, threads=None):
    """-> (address, size) of a malloc'd buffer; release with gen_free(address)."""
    threads = threads or min(32, os.cpu_count() or 1)
    n = ctypes.c_size_t()
    addr = gen_lib().gafgen_records_alloc(ctypes.byref(p), first, count, threads, ctypes.byref(n))
    if not addr:
        raise MemoryError("gafgen_records_alloc")
    return addr, n.value


def gen_free(addr):
    gen_lib().gafgen_free(addr)


def gen_records(p, first, count, threads=None):
    addr, n = gen_records_raw(p, first, count, threads)
    try:
        return ctypes.string_at(addr, n)
    finally:
        gen_free(addr)


def gen_records_into(p, first, count, addr, cap, threads=None):
    """Generate straight into a caller buffer (e.g. pinned memory); returns the byte count
    (nothing written if it exceeds cap)."""
    threads = threads or min(32, os.cpu_count() or 1)
    return gen_lib().gafgen_records(ctypes.byref(p), first, count, addr, cap, threads)


def oracle_path(kind="auto", tool="gaf2paf"):
    """Path of the checker binary: the unmodified reference build when present
    (kind 'reference'), else the restatement (kind 'port')."""
    ref = os.path.join(REF_BIN, tool)
    port = os.path.join(ORACLE_BIN, tool + "_oracle")
    if kind in ("auto", "reference") and os.path.exists(ref):
        return ref, "reference"
    if kind == "reference":
        raise FileNotFoundError(ref)
    if not os.path.exists(port):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    return port, "port"


def run_tool(binary, args, stdin_bytes=None):
    p = subprocess.run([binary] + list(args), input=stdin_bytes, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    rc = p.returncode
    if rc < 0:
        rc = 128 - rc
    return rc, p.stdout, p.stderr.decode("latin-1")


def run_gaf2paf_cpu(gaf, lengths, kind="auto"):
    """Run the CPU checker on in-memory inputs -> (rc, stdout bytes, stderr text, kind)."""
    binary, k = oracle_path(kind)
    with tempfile.TemporaryDirectory() as td:
        lp = os.path.join(td, "l.tsv")
        with open(lp, "wb") as f:
            f.write(lengths)
        rc, out, err = run_tool(binary, ["-", "-l", lp], gaf)
    return rc, out, err, k


def golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)
