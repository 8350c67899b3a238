#!/usr/bin/env bash
# ncu captures (launch list + --set full) of the short-read and assembly workloads.  Each capture
# runs only after the same command exited 0 without ncu.  Outputs -> gpurun_out/.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
S="python bench.py --records 1000000 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
A="python bench.py --workload asm --records 1000 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
echo "== short: launch list"
$S > gpurun_out/plain_short.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_short.csv $S > gpurun_out/ncu_list_short.log 2>&1
echo "rc=$?"
echo "== short: full (k_short size pass + k_emit_lines of the timed step)"
$S > gpurun_out/plain_short2.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:k_short|k_emit_lines" -s 6 -c 2 -f -o gpurun_out/prof_short $S > gpurun_out/ncu_full_short.log 2>&1
echo "rc=$?"
if [ "${SKIP_ASM:-0}" != "1" ]; then
echo "== asm: launch list"
$A > gpurun_out/plain_asm.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_asm.csv $A > gpurun_out/ncu_list_asm.log 2>&1
echo "rc=$?"
echo "== asm: full (k_long size + emit of the timed step)"
$A > gpurun_out/plain_asm2.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:k_long|k_emit_lines" -s 6 -c 2 -f -o gpurun_out/prof_asm $A > gpurun_out/ncu_full_asm.log 2>&1
echo "rc=$?"
fi
ls -la gpurun_out | tail -20
