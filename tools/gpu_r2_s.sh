#!/bin/bash
# round-2 run S: GPU suite + long-record workloads after the k_par / k_emit_lines changes; launch lists
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2s_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2s_pytest.log
tail -3 gpurun_out/r2s_pytest.log
Q="--no-cli --no-cpu-baseline --no-e2e --steps 6 --warmup 3"
for w in stable medium asm mixed short; do
  timeout 400 python bench.py --workload $w $Q > gpurun_out/r2s_${w}.json 2> gpurun_out/r2s_${w}.err
  echo "$w rc $? $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2s_${w}.json | head -1) $(grep -o '"kernel_ms": {[^}]*}' gpurun_out/r2s_${w}.json | head -1)"
done
for w in asm stable; do
  S="python bench.py --workload $w --steps 1 --warmup 3 --no-cli --no-e2e --no-cpu-baseline"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2s_launches_$w.csv $S > gpurun_out/r2s_ncu_$w.log 2>&1
  echo "ncu list $w rc $?"
done
