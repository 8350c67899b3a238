#!/bin/bash
# round-2 run Z3: flat reversed-piece copy in write_line: long-record workloads + short
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
Q="--no-cli --no-cpu-baseline --no-e2e --steps 10 --warmup 3"
for w in stable medium asm short mixed; do
  timeout 400 python bench.py --workload $w $Q > gpurun_out/r2z3_${w}.json 2> gpurun_out/r2z3_${w}.err
  echo "$w rc $? $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2z3_${w}.json | head -1) $(grep -o '"kernel_ms": {[^}]*}' gpurun_out/r2z3_${w}.json | head -1)"
done
