#!/bin/bash
# round-2 run U: where does the gaf2unstable stage spend its time?  launch list + full capture of the size pass
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
S="python bench.py --workload unstable --steps 1 --warmup 2 --no-cli --no-e2e --no-cpu-baseline"
$S > gpurun_out/r2u_plain.log 2>&1; echo "plain rc $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2u_launches_unstable.csv $S > gpurun_out/r2u_ncu_list.log 2>&1; echo "list rc $?"
G2P_UNSTABLE_STAGED=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2u_launches_unstable_unstaged.csv $S > gpurun_out/r2u_ncu_list0.log 2>&1; echo "list0 rc $?"
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_unstable_staged" -s 2 -c 1 -f -o gpurun_out/r2u_k_unstable $S > gpurun_out/r2u_ncu_full.log 2>&1; echo "full rc $?"
cp cactus-gfa-tools_b200/csrc/g2u_core.cuh gpurun_out/r2u_g2u_core.cuh
