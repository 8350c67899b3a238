#!/usr/bin/env bash
# ncu captures of the short-read workload with k_rec: launch list + --set full of k_rec and k_emit_lines.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
S="python bench.py --records 1000000 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
$S > gpurun_out/plain_rec.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_rec.csv $S > gpurun_out/ncu_list_rec.log 2>&1
echo "list rc=$?"
$S > gpurun_out/plain_rec2.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:k_rec|k_emit_lines" -s 6 -c 2 -f -o gpurun_out/prof_rec $S > gpurun_out/ncu_full_rec.log 2>&1
echo "full rc=$?"
ls -la gpurun_out | tail -8
