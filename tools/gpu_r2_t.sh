#!/bin/bash
# round-2 run T: staged k_unstable on the GPU (parity tests, bench with and without staging)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "unstable or hpp20 or rgfa" > gpurun_out/r2t_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2t_pytest.log
tail -3 gpurun_out/r2t_pytest.log
Q="--no-cli --no-cpu-baseline --steps 10 --warmup 3"
for v in 1 0; do
  G2P_UNSTABLE_STAGED=$v timeout 400 python bench.py --workload unstable $Q > gpurun_out/r2t_unstable_$v.json 2> gpurun_out/r2t_unstable_$v.err
  echo "staged=$v rc $? $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2t_unstable_$v.json | head -1) $(grep -o '"kernel_ms": {[^}]*}' gpurun_out/r2t_unstable_$v.json | head -1) $(grep -o '"e2e": {"value": [0-9.]*' gpurun_out/r2t_unstable_$v.json | head -1)"
done
