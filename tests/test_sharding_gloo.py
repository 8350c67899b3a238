"""Multi-GPU host logic on the CPU: two ranks (torch.distributed, gloo) each convert their
newline-aligned shard and rank 0 concatenates the outputs in rank order -- exactly what
bench.py / the gaf2paf executable do with one GPU per rank (no collective on the
data path, only the gather of results).  The conversion itself runs the product kernels under the
SIMT emulator (build/g2p_simt), so no GPU is needed."""
import os
import subprocess
import sys
import tempfile

import helpers as H

ROOT = H.ROOT
SIMT = os.path.join(H.BUILD, "g2p_simt")

WORKER = r'''
import os, subprocess, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import torch.distributed as dist
import cactus_gfa_tools_b200 as g2p
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
gaf = open({gaf!r}, "rb").read()
a, b = g2p.shard_ranges(gaf, world)[rank]
p = subprocess.run([{simt!r}, "-l", {lengths!r}, "-"], input=gaf[a:b], stdout=subprocess.PIPE)
assert p.returncode == 0
parts = [None] * world
dist.all_gather_object(parts, (rank, a, b, p.stdout))
if rank == 0:
    parts.sort()
    assert parts[0][1] == 0 and parts[-1][2] == len(gaf) and all(parts[i][2] == parts[i + 1][1] for i in range(world - 1))
    open({out!r}, "wb").write(b"".join(x[3] for x in parts))
dist.barrier()
dist.destroy_process_group()
'''


def test_two_rank_sharding_concatenates_to_the_unsharded_output():
    p = H.preset("short", seed=101, pct_star=1)
    lengths = H.gen_lengths(p)
    pm = H.preset("medium", seed=101)
    gaf = H.gen_records(p, 0, 1500, threads=2) + H.gen_records(pm, 0, 20, threads=2) + H.gen_records(p, 1500, 700, threads=2)
    with tempfile.TemporaryDirectory() as td:
        gp, lp, op = os.path.join(td, "in.gaf"), os.path.join(td, "l.tsv"), os.path.join(td, "out.paf")
        open(gp, "wb").write(gaf)
        open(lp, "wb").write(lengths)
        script = os.path.join(td, "worker.py")
        open(script, "w").write(WORKER.format(root=ROOT, gaf=gp, lengths=lp, simt=SIMT, out=op))
        env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", WORLD_SIZE="2")
        procs = [subprocess.Popen([sys.executable, script], env=dict(env, RANK=str(r))) for r in range(2)]
        assert [q.wait(timeout=300) for q in procs] == [0, 0]
        rc, ref, err, kind = H.run_gaf2paf_cpu(gaf, lengths)
        assert rc == 0 and open(op, "rb").read() == ref
