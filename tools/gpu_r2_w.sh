#!/bin/bash
# round-2 run W (2 GPUs, final build): the whole GPU suite with 2 devices visible (incl. the sharded-executable tests),
# and the mixed workload (k_par on its assembly records) as one logical input over 2 GPUs
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2w_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2w_pytest.log; tail -3 gpurun_out/r2w_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for w in mixed short; do
  timeout 900 $TR bench.py --gpus 2 --workload $w --steps 10 --warmup 3 > gpurun_out/r2w_bench_$w.json 2> gpurun_out/r2w_bench_$w.err; echo "bench $w N=2 rc $?"
  head -c 400 gpurun_out/r2w_bench_$w.json; echo
done
