#!/bin/bash
# round-2 run Q: --set full captures of the k_par kernels on 1000 asm records
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
S="python bench.py --workload asm --records 1000 --steps 1 --warmup 1 --no-cli --no-e2e --no-cpu-baseline"
$S > gpurun_out/r2q_plain.log 2>&1; echo "plain rc $?"
for k in k_par_count k_par_fill k_par_lines k_par_steps; do
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$k" -s 1 -c 1 -f -o gpurun_out/r2q_$k $S > gpurun_out/r2q_ncu_$k.log 2>&1
  echo "ncu $k rc $?"
done
cp cactus-gfa-tools_b200/csrc/g2p_par.cuh gpurun_out/r2q_g2p_par.cuh
