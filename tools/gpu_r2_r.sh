#!/bin/bash
# round-2 run R: k_par with the tile map; per-kernel times (ncu launch lists) on asm / stable / medium
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
Q="--no-cli --no-cpu-baseline --no-e2e --steps 6 --warmup 3"
for w in stable medium asm mixed; do
  timeout 400 python bench.py --workload $w $Q > gpurun_out/r2r_${w}.json 2> gpurun_out/r2r_${w}.err
  echo "$w rc $? $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2r_${w}.json | head -1) $(grep -o '"kernel_ms": {[^}]*}' gpurun_out/r2r_${w}.json | head -1)"
done
for w in asm stable medium; do
  S="python bench.py --workload $w --steps 1 --warmup 3 --no-cli --no-e2e --no-cpu-baseline"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2r_launches_$w.csv $S > gpurun_out/r2r_ncu_$w.log 2>&1
  echo "ncu list $w rc $?"
done
