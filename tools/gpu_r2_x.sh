#!/bin/bash
# round-2 run X (8 GPUs, final build): configs[4] -- the mixed workload as ONE logical input sharded over 8 GPUs
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 280 $TR bench.py --gpus 8 --workload mixed --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2x_bench_mixed.json 2> gpurun_out/r2x_bench_mixed.err; echo "bench mixed N=8 rc $?"
head -c 600 gpurun_out/r2x_bench_mixed.json; echo
