#!/bin/bash
# round-2 run Y: gaf2unstable with the warp-phased per-record code (parity tests, bench for both kernel variants, one full capture)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "unstable or hpp20 or rgfa or filter" > gpurun_out/r2y_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2y_pytest.log
tail -3 gpurun_out/r2y_pytest.log
Q="--no-cli --no-cpu-baseline --steps 10 --warmup 3"
for v in 0 1; do
  G2P_UNSTABLE_STAGED=$v timeout 400 python bench.py --workload unstable $Q > gpurun_out/r2y_unstable_$v.json 2> gpurun_out/r2y_unstable_$v.err
  echo "staged=$v rc $? $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2y_unstable_$v.json | head -1) $(grep -o '"kernel_ms": {[^}]*}' gpurun_out/r2y_unstable_$v.json | head -1) $(grep -o '"e2e": {"value": [0-9.]*' gpurun_out/r2y_unstable_$v.json | head -1)"
done
S="python bench.py --workload unstable --steps 1 --warmup 2 --no-cli --no-e2e --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_unstable" -s 2 -c 2 -f -o gpurun_out/r2y_k_unstable $S > gpurun_out/r2y_ncu_full.log 2>&1; echo "full rc $?"
cp cactus-gfa-tools_b200/csrc/g2u_core.cuh gpurun_out/r2y_g2u_core.cuh
